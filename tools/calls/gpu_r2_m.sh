#!/bin/bash
# round 2, call M: tests + sanitizer after the atomic tail flag
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | tee gpurun_out/r02_pytest_gpu.log
echo "== sanitizer"; bash tools/gpu_sanitize.sh
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
echo "== bench exact"; timeout 600 python bench.py $B 2>&1 | tail -1 | cut -c1-200
echo "== bench fast"; timeout 600 python bench.py --mode fast $B 2>&1 | tail -1 | cut -c1-200
