"""Tail behaviour of render_kernel: shard size x persistent warps x tail compaction (exact and host)."""
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
def run(mode, world, reps=2):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter()
        lp.render(cam, prm, seq, dl, n, w, h, mode=mode, tile=8, rank=0, world=world, rgba=rgba, points=pts)
        torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best * 1e3
out = []
for mode, worlds, warps in (("exact", (1, 2, 4, 8), (8, 12, 16)), ("host", (1, 8), (8, 16, 24)), ("fast", (1, 8), (0,))):
    for world in worlds:
        for wp in warps:
            for tc in (0, 1):
                api.set_option("render_warps_per_sm", wp); api.set_option("tail_compaction", tc)
                ms = run(mode, world, reps=2 if mode != "host" else 1)
                rec = {"mode": mode, "world": world, "warps_per_sm": wp, "tail_compaction": tc, "ms": round(ms, 2)}
                out.append(rec); print(json.dumps(rec), flush=True)
api.set_option("render_warps_per_sm", 0); api.set_option("tail_compaction", 1)
json.dump(out, open("gpurun_out/r02_tail_diag.json", "w"), indent=1)
