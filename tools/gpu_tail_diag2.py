"""Tail compaction on the two-rays-per-lane kernels: fast and hybrid, shard sizes, on/off."""
import sys, os, time, json
sys.path.insert(0, os.getcwd())
import torch, lyapunov3d_b200 as lp
from lyapunov3d_b200 import api
from lyapunov3d_b200.structs import clone
prm, cam, lights, n, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, n)
seq = lp.scene_convert_sequence(s)
w, h = 1920, 1080
lp.scene_cam_recalculate(cam, w, h, 1)
dl = api.upload_lights(lights)
rgba = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda"); pts = torch.zeros((h, w, 36), dtype=torch.uint8, device="cuda")
p0 = clone(prm); p0.jitter = 0.0
def run(mode, world, p, reps=2):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter()
        lp.render(cam, p, seq, dl, n, w, h, mode=mode, tile=8, rank=0, world=world, rgba=rgba, points=pts)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best * 1e3
for mode, p in (("exact", prm), ("fast", prm), ("hybrid", p0), ("exact", p0)):
    for world in (1, 4, 8):
        for tc in (0, 1):
            api.set_option("tail_compaction", tc)
            print(json.dumps({"mode": mode, "jitter": p.jitter, "world": world, "tail_compaction": tc, "ms": round(run(mode, world, p), 2)}), flush=True)
api.set_option("tail_compaction", 1)
