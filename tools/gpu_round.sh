#!/bin/bash
# One gpurun call: smoke, bench (both arms), ncu launch list + full captures.  Output under gpurun_out/.
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench exact"; timeout 600 python bench.py 2>&1 | tail -2 | tee gpurun_out/bench_exact.json
echo "== bench fast"; timeout 600 python bench.py --mode fast --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fast.json
echo "== bench reference"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
echo "== bench bake512 fast"; timeout 600 python bench.py --workload bake512 --mode fast 2>&1 | tail -1 | tee gpurun_out/bench_bake_fast.json
echo "== bench bake512 exact"; timeout 600 python bench.py --workload bake512 --mode exact 2>&1 | tail -1 | tee gpurun_out/bench_bake_exact.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
echo "== ncu full render exact"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/prof_render_exact python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_render_exact.log 2>&1
echo "== ncu full render fast"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_fast2_kernel -s 1 -c 1 -f -o gpurun_out/prof_render_fast python bench.py --mode fast --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_render_fast.log 2>&1
echo "== ncu full bake fast"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bake_kernel -s 1 -c 1 -f -o gpurun_out/prof_bake_fast python bench.py --workload bake512 --mode fast --steps 1 --warmup 1 > gpurun_out/ncu_bake_fast.log 2>&1
ls -la gpurun_out
