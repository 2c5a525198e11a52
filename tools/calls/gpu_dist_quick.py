#!/usr/bin/env python
"""Short multi-GPU timing (torchrun): the two tile-sharded frames through peer memory."""
import json, os, sys, time
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import lyapunov3d_b200 as lp
from lyapunov3d_b200 import api, dist as ld
from lyapunov3d_b200.structs import clone
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
prm, cam, lights, n, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, n); seq = lp.scene_convert_sequence(s)
d_lights = api.upload_lights(lights, dev)
out = {"world": world}
def run(name, c, p, sq, w, h, mode, per, reps):
    rgba, pts = ld.PeerBuffer(w * h * 4), ld.PeerBuffer(w * h * 36)
    best = 1e9
    for _ in range(reps):
        ev = torch.zeros(1, dtype=torch.int64, device=dev)
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        ld.render_frame_sharded_peer(rgba, pts, c, p, sq, d_lights, n, w, h, mode=mode, evals=ev)
        best = min(best, time.perf_counter() - t0)
    dist.all_reduce(ev)
    out[name] = {"s": best, "giter_s": int(ev.item()) * per / best / 1e9, "fps": 1 / best}
    rgba.close(); pts.close()
c2 = clone(cam); lp.scene_cam_recalculate(c2, 1920, 1080, 1)
p3 = clone(prm); p3.settle, p3.accum = 72, 4032; s3 = lp.scene_convert_sequence("A6B6C6")
c3 = clone(cam); lp.scene_cam_recalculate(c3, 3840, 2160, 1)
run("frame1080_tiles_exact_peer", c2, prm, seq, 1920, 1080, "exact", 1026, 3)
run("frame1080_tiles_fast_peer", c2, prm, seq, 1920, 1080, "fast", 1026, 3)
run("frame4k_long_exact_peer", c3, p3, s3, 3840, 2160, "exact", 4104, 2)
run("frame4k_long_fast_peer", c3, p3, s3, 3840, 2160, "fast", 4104, 2)
if rank == 0: print("DIST_QUICK " + json.dumps(out), flush=True)
dist.destroy_process_group()
