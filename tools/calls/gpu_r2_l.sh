#!/bin/bash
# round 2, call L: tests, sanitizer on the new shared-memory paths, long-sequence numbers
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu.log
echo "== long sequences"; timeout 900 python tools/gpu_longseq.py 2>&1 | tail -20 | tee gpurun_out/r02_longseq.log
echo "== sanitizer"; bash tools/gpu_sanitize.sh
