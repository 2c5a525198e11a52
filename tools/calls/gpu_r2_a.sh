#!/bin/bash
# round 2, call A: whole GPU suite, headline bench with sub-records, hybrid bench
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.log
echo "== bench exact"; timeout 900 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_exact.err | tail -1 | tee gpurun_out/r02_bench_exact.json
tail -5 gpurun_out/bench_exact.err
echo "== bench hybrid jitter 0"; timeout 600 python bench.py --mode hybrid --jitter 0 --no-subrecords --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02_bench_hybrid_j0.json
echo "== bench exact jitter 0"; timeout 600 python bench.py --mode exact --jitter 0 --no-subrecords --no-cpu-baseline --no-reference-cuda 2>&1 | tail -1 | tee gpurun_out/r02_bench_exact_j0.json
echo "== bench fast jitter 0"; timeout 600 python bench.py --mode fast --jitter 0 --no-subrecords --no-cpu-baseline --no-reference-cuda 2>&1 | tail -1 | tee gpurun_out/r02_bench_fast_j0.json
echo "== bench host"; timeout 600 python bench.py --mode host --steps 2 --warmup 1 --no-subrecords --no-cpu-baseline --no-reference-cuda 2>&1 | tail -1 | tee gpurun_out/r02_bench_host.json
