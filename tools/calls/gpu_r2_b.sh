#!/bin/bash
# round 2, call B: hybrid diagnosis, GPU suite, sanitizer logs, ncu launch list + full captures
mkdir -p gpurun_out
echo "== hybrid diag"; timeout 600 python tools/gpu_diag_hybrid.py 2>&1 | tail -20 | tee gpurun_out/r02_hybrid_diag.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee gpurun_out/r02_pytest_gpu.log
echo "== bench hybrid jitter 0"; timeout 600 python bench.py --mode hybrid --jitter 0 --no-subrecords --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/r02_bench_hybrid_j0.json
echo "== sanitizer"; bash tools/gpu_sanitize.sh
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_exact.csv python bench.py --steps 2 --warmup 1 $B > gpurun_out/ncu_launches.log 2>&1
echo "== ncu full render exact"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_render_exact python bench.py --steps 1 --warmup 1 $B > gpurun_out/ncu_render_exact.log 2>&1
echo "== ncu full render host"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_render_host python bench.py --mode host --steps 1 --warmup 1 $B > gpurun_out/ncu_render_host.log 2>&1
echo "== ncu full march hybrid"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:march_fast2_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_march_hybrid python bench.py --mode hybrid --jitter 0 --steps 1 --warmup 1 $B > gpurun_out/ncu_march_hybrid.log 2>&1
ls -la gpurun_out | tail -20
