import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, lyapunov3d_b200 as lp
from lyapunov3d_b200.structs import clone
prm, cam, lights, n, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, n); seq = lp.scene_convert_sequence(s)
c = clone(cam); lp.scene_cam_recalculate(c, 1920, 1080, 1)
host = torch.zeros((1080, 1920, 4), dtype=torch.uint8).pin_memory().numpy()
for i in range(4):
    t = time.perf_counter(); lp.render_host(c, prm, seq, lights, n, 1920, 1080, mode="exact", want_points=False, rgba=host); print("call", i, (time.perf_counter() - t) * 1e3, "ms", flush=True)
os.environ["LYAP_TRACE"] = "1"
for i in range(2):
    lp.render_host(c, prm, seq, lights, n, 1920, 1080, mode="exact", want_points=False, rgba=host)
del os.environ["LYAP_TRACE"]
# device path on the torch stream for comparison
for i in range(3):
    torch.cuda.synchronize(); t = time.perf_counter(); lp.render(c, prm, seq, lights, n, 1920, 1080, mode="exact"); torch.cuda.synchronize(); print("device call", i, (time.perf_counter() - t) * 1e3, "ms", flush=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for i in range(3):
    flush.zero_(); torch.cuda.synchronize()
    t = time.perf_counter(); lp.render_host(c, prm, seq, lights, n, 1920, 1080, mode="exact", want_points=False, rgba=host); print("after torch activity: call", i, (time.perf_counter() - t) * 1e3, "ms", flush=True)
