#!/bin/bash
# round 2, call F (re-entry): full GPU suite + every bench line of the round, saved under gpurun_out/
mkdir -p gpurun_out
echo "== pytest"; timeout 1800 python -m pytest tests -m gpu -q -x --durations=15 2>&1 | tail -40 | tee gpurun_out/r02_pytest_gpu.log
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
echo "== bench exact (headline, all sub-records)"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/r02_bench_exact.json | cut -c1-600
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r02_bench_reference.json | cut -c1-300
echo "== bench fast"; timeout 600 python bench.py --mode fast $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_fast.json | cut -c1-400
echo "== bench host"; timeout 600 python bench.py --mode host --steps 3 --warmup 1 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_host.json | cut -c1-400
echo "== bench hybrid j0"; timeout 600 python bench.py --mode hybrid --jitter 0 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_hybrid_j0.json | cut -c1-400
echo "== bench exact j0"; timeout 600 python bench.py --mode exact --jitter 0 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_exact_j0.json | cut -c1-400
echo "== bench hybrid_host j0"; timeout 600 python bench.py --mode hybrid_host --jitter 0 --steps 3 --warmup 1 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_hybrid_host_j0.json | cut -c1-400
echo "== bench host j0"; timeout 600 python bench.py --mode host --jitter 0 --steps 3 --warmup 1 $B 2>&1 | tail -1 | tee gpurun_out/r02_bench_host_j0.json | cut -c1-400
echo "== bake512 fast"; timeout 600 python bench.py --workload bake512 --mode fast 2>&1 | tail -1 | tee gpurun_out/r02_bench_bake512_fast.json | cut -c1-400
echo "== bake512 exact"; timeout 600 python bench.py --workload bake512 --mode exact 2>&1 | tail -1 | tee gpurun_out/r02_bench_bake512_exact.json | cut -c1-400
echo "== bake512 host"; timeout 600 python bench.py --workload bake512 --mode host --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/r02_bench_bake512_host.json | cut -c1-400
echo "== tail diag 2"; timeout 900 python tools/gpu_tail_diag2.py 2>&1 | tail -40 | tee gpurun_out/r02_tail_diag2.log
echo "== assist report"; timeout 900 python tools/gpu_assist_report.py 2>&1 | tail -30 | tee gpurun_out/r02_assist_report.log
