#!/bin/bash
# compute-sanitizer over small configurations of every kernel family (memcheck + racecheck).
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, lyapunov3d_b200 as lp
from lyapunov3d_b200.structs import clone
prm, cam, lights, n, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, n)
prm.accum = 120
for sq in ("BCABA", "A6B6C6", "A9B9C9D9"):
    seq = lp.scene_convert_sequence(sq)
    c = clone(cam); lp.scene_cam_recalculate(c, 37, 21, 1)
    for mode in ("exact", "fast", "host"):
        lp.render(c, prm, seq, lights, n, 37, 21, mode=mode)
        lp.render(c, prm, seq, lights, n, 37, 21, mode=mode, tile=8, rank=1, world=3, compact=True)
        lp.bake(prm, seq, 13, 7, 5, mode=mode)
        lp.bake(prm, seq, 13, 7, 5, mode=mode, dtype="f16", z0=1, z1=4)
        lp.exponent_points(torch.rand(100, 3, device="cuda") * 4, prm, seq, mode=mode)
torch.cuda.synchronize(); print("sanitizer workload done")
PY
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py 2>&1 | tail -4
done
