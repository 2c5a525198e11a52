import os, sys
sys.path.insert(0, os.getcwd())
import torch, lyapunov3d_b200 as lp
from lyapunov3d_b200 import api
pk = api.probe_peaks()
prm, cam, lights, n, s, _ = lp.params_init()
buf = torch.empty((512, 512, 512), dtype=torch.float32, device="cuda")
for mode in ("fast", "exact"):
    for sq in ("A", "AB", "ABC", "BCABA", "A6B6C6", "A9B9C9D9"):
        seq = lp.scene_convert_sequence(sq)
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); lp.bake(prm, seq, 512, mode=mode, out=buf); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        git = 512 ** 3 * 1026 / best / 1e6
        peak = (pk["ffma_lane_ops_per_s"] / 3.965 if mode == "fast" else pk["mufu_lane_ops_per_s"] * 1026 / 1008) / 1e9
        print(mode, sq, "P=%d" % api.plan_period(seq, 18, 1008), "ms %.3f Giter/s %.1f frac %.3f" % (best, git, git / peak), flush=True)
