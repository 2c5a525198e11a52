"""One rank's shard of the default 1080p frame (world 4 / 8 emulated on one GPU): persistent warps per SM."""
import sys, os, time, json
sys.path.insert(0, os.getcwd())
import torch, lyapunov3d_b200 as lp
from lyapunov3d_b200 import api
prm, cam, lights, n, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, n)
seq = lp.scene_convert_sequence(s)
w, h = 1920, 1080
lp.scene_cam_recalculate(cam, w, h, 1)
dl = api.upload_lights(lights)
rgba = torch.zeros((h, w, 4), dtype=torch.uint8, device="cuda"); pts = torch.zeros((h, w, 36), dtype=torch.uint8, device="cuda")
def run(mode, world, rank, reps=3):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        lp.render(cam, prm, seq, dl, n, w, h, mode=mode, tile=8, rank=rank, world=world, rgba=rgba, points=pts)
        e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
from lyapunov3d_b200.structs import clone
p0 = clone(prm); p0.jitter = 0.0
which = sys.argv[1:] or ["exact", "fast", "hybrid"]
for mode, warps in (("exact", (0, 8, 12, 16)), ("fast", (0, 8, 12)), ("hybrid", (0, 8, 12))):
    if mode not in which:
        continue
    if mode == "hybrid":
        prm = p0
    for world in (8, 4, 2, 1):
        for wp in warps:
            api.set_option("render_warps_per_sm", wp)
            ms = [round(run(mode, world, r), 2) for r in (0, world - 1)]
            print(json.dumps({"mode": mode, "world": world, "warps_per_sm": wp, "ms_rank0_rankLast": ms}), flush=True)
api.set_option("render_warps_per_sm", 0)
