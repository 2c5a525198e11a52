#!/bin/bash
# compute-sanitizer over small configurations of every kernel family (memcheck + racecheck + initcheck on the
# library's own scratch).  Logs land in gpurun_out/r02_sanitizer_<tool>.log; copy them to profiles/.
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch, lyapunov3d_b200 as lp
from lyapunov3d_b200 import api
from lyapunov3d_b200.structs import clone
prm, cam, lights, n, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, n)
prm.accum = 80
launches = 0
# (sequence, seq_table): register-table periods 5, 21, 40; generic path with the shared-memory
# multiplier table forced (2) and with the run-length loop (0)
for sq, table in (("BCABA", 1), ("A9B9C9D9", 1), ("A9A8B9B9", 2), ("A9A8B9B9", 0), ("ABCABBCACABBACBCAABCABCCBACBABCABACBBCA", 2)):
    api.set_option("seq_table", table)
    seq = lp.scene_convert_sequence(sq)
    c = clone(cam); lp.scene_cam_recalculate(c, 37, 21, 1)
    for mode in ("exact", "fast", "host"):
        lp.render(c, prm, seq, lights, n, 37, 21, mode=mode)
        lp.render(c, prm, seq, lights, n, 37, 21, mode=mode, tile=8, rank=1, world=3, compact=True)
        lp.bake(prm, seq, 13, 7, 5, mode=mode)
        lp.bake(prm, seq, 13, 7, 5, mode=mode, dtype="f16", z0=1, z1=4)
        lp.exponent_points(torch.rand(100, 3, device="cuda") * 4, prm, seq, mode=mode)
        launches += 5
    p0 = clone(prm); p0.jitter = 0.0
    for mode in ("hybrid", "hybrid_host"):
        for batch in (0, 1, 32):
            api.set_option("hybrid_guard_batch", batch)
            lp.render(c, p0, seq, lights, n, 37, 21, mode=mode)
            lp.render(c, p0, seq, lights, n, 37, 21, mode=mode, tile=8, rank=2, world=3, compact=True)
            launches += 4
        api.set_option("hybrid_guard_batch", 0)
    # element-aligned (odd offset) fast-mode volume: scalar-store fallback
    big = torch.zeros(13 * 7 * 5 + 3, device="cuda")
    lp.bake(prm, seq, 13, 7, 5, mode="fast", out=big[1:1 + 13 * 7 * 5].view(5, 7, 13))
    pts = lp.render(c, prm, seq, lights, n, 37, 21, mode="exact")[1]
    lp.shade_points(pts, c, lights, n, mode="exact"); lp.shade_points(pts, c, lights, n, mode="host")
    api.ray_probe([[0, 0], [5, 7], [36, 20]], c, prm, mode="exact")
    api.normalize_vectors(torch.rand(50, 3, device="cuda"), mode="host")
    launches += 6
torch.cuda.synchronize(); print("sanitizer workload done: %d library calls" % launches)
PY
for tool in memcheck racecheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py > gpurun_out/r02_sanitizer_$tool.log 2>&1
  echo "exit code $?" >> gpurun_out/r02_sanitizer_$tool.log
  tail -5 gpurun_out/r02_sanitizer_$tool.log
done
