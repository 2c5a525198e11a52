"""Where does hybrid mode's time go?  Kernel split via CUDA events and knob sensitivity."""
import sys, os, time
sys.path.insert(0, os.getcwd())
import torch, lyapunov3d_b200 as lp
from lyapunov3d_b200 import api
from lyapunov3d_b200.structs import clone
prm, cam, lights, n, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, n)
seq = lp.scene_convert_sequence(s)
prm.jitter = 0.0
w, h = 960, 540
lp.scene_cam_recalculate(cam, w, h, 1)
dl = api.upload_lights(lights)
def run(mode, reps=3):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter()
        r = lp.render(cam, prm, seq, dl, n, w, h, mode=mode)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best * 1e3, int(r[2].item())
for mode in ("exact", "fast", "hybrid"):
    print(mode, "%.1f ms, %d evals" % run(mode), flush=True)
for pct in (0, 10, 100, 1000):
    api.set_option("hybrid_guard_percent", pct)
    print("guard %d%%: %.1f ms" % (pct, run("hybrid")[0]), flush=True)
api.set_option("hybrid_guard_percent", 100)
for b in (1, 2, 4, 8, 16, 32):
    api.set_option("hybrid_guard_batch", b)
    print("batch %d: %.1f ms" % (b, run("hybrid")[0]), flush=True)
api.set_option("hybrid_guard_batch", 0)
for wps in (8, 12, 16):
    api.set_option("render_warps_per_sm", wps)
    print("warps/SM %d: %.1f ms" % (wps, run("hybrid")[0]), flush=True)
