#!/bin/bash
# round 2, call K: shared-memory multiplier table for long periods, per-sample hybrid guard -- tests and measurements
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/r02_pytest_gpu.log
echo "== long sequences"; timeout 900 python tools/gpu_longseq.py 2>&1 | tail -20 | tee gpurun_out/r02_longseq.log
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
echo "== bench hybrid j0"; timeout 600 python bench.py --mode hybrid --jitter 0 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_hybrid_j0.json | cut -c1-300
echo "== bench hybrid_host j0"; timeout 600 python bench.py --mode hybrid_host --jitter 0 --steps 3 --warmup 1 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_hybrid_host_j0.json | cut -c1-300
