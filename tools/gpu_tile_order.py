"""Chord-descending tile order against image order: the default 1080p frame, whole and as one rank's shard
(world 2 / 4 / 8 emulated on one GPU), CUDA-event times of the launch (order kernel + sort included)."""
import sys, os, json
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
def run(mode, p, world, rank, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        lp.render(cam, p, seq, dl, n, w, h, mode=mode, tile=8, rank=rank, world=world, rgba=rgba, points=pts)
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for mode, p in (("exact", prm), ("fast", prm), ("hybrid", p0)):
    for world in (8, 4, 2, 1):
        rec = {"mode": mode, "world": world}
        for order in (0, 1):
            api.set_option("tile_order", order)
            rec["order%d_ms_rank0_rankLast" % order] = [round(run(mode, p, world, r), 2) for r in sorted({0, world - 1})]
        print(json.dumps(rec), flush=True)
api.set_option("tile_order", 1)
