#!/bin/bash
# round 2, call Q: ncu --set full of the fast frame kernel and the fast bake kernel (final fold form)
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_fast2_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_render_fast python bench.py --mode fast --steps 1 --warmup 1 $B > gpurun_out/ncu_render_fast.log 2>&1
tail -2 gpurun_out/ncu_render_fast.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:bake_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_bake_fast python bench.py --workload bake512 --mode fast --steps 1 --warmup 1 $B > gpurun_out/ncu_bake_fast.log 2>&1
tail -2 gpurun_out/ncu_bake_fast.log
ls -la gpurun_out
