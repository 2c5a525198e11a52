#!/bin/bash
mkdir -p gpurun_out
B="--no-subrecords --no-cpu-baseline --no-reference-cuda"
echo "== bake512 host"; timeout 600 python bench.py --workload bake512 --mode host --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-400
echo "== ncu bake host"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bake_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_bake_host python bench.py --workload bake512 --mode host --steps 1 --warmup 1 > gpurun_out/ncu_bake_host.log 2>&1
echo "== ncu render host small"
cat > /tmp/rh.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, lyapunov3d_b200 as lp
prm, cam, lights, n, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, n); seq = lp.scene_convert_sequence(s)
lp.scene_cam_recalculate(cam, 640, 360, 1)
for _ in range(2):
    lp.render(cam, prm, seq, lights, n, 640, 360, mode="host")
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 1 -c 1 -f -o gpurun_out/r02_prof_render_host_small python /tmp/rh.py > gpurun_out/ncu_render_host_small.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
