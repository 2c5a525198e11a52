import os, sys
sys.path.insert(0, os.getcwd())
import torch, lyapunov3d_b200 as lp
from lyapunov3d_b200 import api
from lyapunov3d_b200.structs import clone
prm, cam, lights, n, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, n); seq = lp.scene_convert_sequence(s)
c = clone(cam); lp.scene_cam_recalculate(c, 1920, 1080, 1)
for mode in ("fast",):
    for wps in (12, 16):
        api.set_option("render_warps_per_sm", wps)
        best = 1e9
        for _ in range(3):
            torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); r = lp.render(c, prm, seq, lights, n, 1920, 1080, mode=mode); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        print(mode, "warps/SM", wps, "ms %.2f" % best, "Giter/s %.1f" % (int(r[2].item()) * 1026 / best / 1e6), flush=True)
