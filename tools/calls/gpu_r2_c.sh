#!/bin/bash
# round 2, call C: GPU suite (incl. assisted march, reference lyap_interactive programs), assist report, host/hybrid benches
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -25 | tee gpurun_out/r02_pytest_gpu.log
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
echo "== bench host"; timeout 600 python bench.py --mode host --steps 3 --warmup 1 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_host.json
echo "== bench hybrid jitter 0"; timeout 600 python bench.py --mode hybrid --jitter 0 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_hybrid_j0.json
echo "== bench hybrid_host jitter 0"; timeout 600 python bench.py --mode hybrid_host --jitter 0 --steps 3 --warmup 1 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_hybrid_host_j0.json
echo "== bench host jitter 0"; timeout 600 python bench.py --mode host --jitter 0 --steps 3 --warmup 1 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_host_j0.json
echo "== assist report"; timeout 900 python tools/gpu_assist_report.py 2>&1 | tail -30
