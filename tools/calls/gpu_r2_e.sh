#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.log
echo "== tail diag 2"; timeout 900 python tools/gpu_tail_diag2.py 2>&1 | tail -40 | tee gpurun_out/r02_tail_diag2.log
