#!/usr/bin/env python
"""Run under torchrun on N GPUs: sharded frame / bake / animation == single-GPU results, bit for bit.
Also times the BASELINE multi-GPU configs (3: 4K long frame tile-sharded, 4: 512^3 z-slabs, 5: orbit frames)."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lyapunov3d_b200 as lp  # noqa: E402
from lyapunov3d_b200 import api, dist as ld  # noqa: E402
from lyapunov3d_b200.structs import clone  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
timing = "--time" in sys.argv
prm, cam, lights, n, s, _ = lp.params_init()
lp.scene_lights_recalculate(lights, n)
seq = lp.scene_convert_sequence(s)
out = {"world": world}


def timed(fn, reps=2):
    best = 1e30
    for _ in range(reps):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        dist.barrier()
        best = min(best, time.perf_counter() - t0)
    return r, best


# ---- correctness: small sharded frame and bake equal the single-GPU result
w, h = 200, 120
c = clone(cam)
lp.scene_cam_recalculate(c, w, h, 1)
for mode in ("exact", "host"):
    rgba, pts, ev = ld.render_frame_sharded(c, prm, seq, lights, n, w, h, mode=mode)
    tot = ev.clone()
    dist.all_reduce(tot)
    if rank == 0:
        r1, p1, e1 = lp.render(c, prm, seq, lights, n, w, h, mode=mode)
        assert torch.equal(rgba, r1) and torch.equal(pts, p1) and int(tot.item()) == int(e1.item()), mode
vol = ld.bake_sharded(prm, seq, 48, 40, 36 + world - 1, mode="fast")
if rank == 0:
    v1 = lp.bake(prm, seq, 48, 40, 36 + world - 1, mode="fast")
    assert torch.equal(vol.view(torch.int32), v1.view(torch.int32))
# ---- peer-memory variants: kernels store straight into rank 0's buffers, no gather
dev = torch.device("cuda", local)
d_lights = api.upload_lights(lights, dev)
p_rgba, p_pts = ld.PeerBuffer(w * h * 4), ld.PeerBuffer(w * h * 36)
ev = torch.zeros(1, dtype=torch.int64, device=dev)
ld.render_frame_sharded_peer(p_rgba, p_pts, c, prm, seq, d_lights, n, w, h, mode="exact", evals=ev)
if rank == 0:
    r1, p1, e1 = lp.render(c, prm, seq, lights, n, w, h, mode="exact")
    assert torch.equal(p_rgba.view((h, w, 4), "|u1").tensor(), r1) and torch.equal(p_pts.view((h, w, 36), "|u1").tensor(), p1)
p_rgba.close(); p_pts.close()
nzp = 36 + world - 1
p_vol = ld.PeerBuffer(48 * 40 * nzp * 4)
ld.bake_sharded_peer(p_vol, prm, seq, 48, 40, nzp, mode="fast")
if rank == 0:
    v1 = lp.bake(prm, seq, 48, 40, nzp, mode="fast")
    assert torch.equal(p_vol.view((nzp, 40, 48), "<f4").tensor().view(torch.int32), v1.view(torch.int32))
    print("peer-memory sharding ok on", world, "GPUs", flush=True)
p_vol.close()

frames = ld.render_animation_sharded(2 * world + 1, 64, 36, prm, cam, seq, lights, n, mode="exact")
if rank == 0:
    for f, fr in enumerate(frames):
        cf = clone(cam)
        lp.campath_frame(f, 2 * world + 1, cf)
        lp.scene_cam_recalculate(cf, 64, 36, 1)
        assert torch.equal(fr, lp.render(cf, prm, seq, lights, n, 64, 36, mode="exact")[0]), f
    print("dist correctness ok on", world, "GPUs", flush=True)

if timing:
    iters = prm.settle + prm.accum
    # config 4: 512^3 bake by z-slabs
    for mode in ("fast", "exact"):
        _, t = timed(lambda: ld.bake_sharded(prm, seq, 512, mode=mode))
        out[f"bake512_{mode}"] = {"s": t, "giter_s": 512 ** 3 * iters / t / 1e9}
    p_vol = ld.PeerBuffer(512 ** 3 * 4)
    for mode in ("fast", "exact"):
        _, t = timed(lambda: ld.bake_sharded_peer(p_vol, prm, seq, 512, 512, 512, mode=mode), reps=3)
        out[f"bake512_{mode}_peer"] = {"s": t, "giter_s": 512 ** 3 * iters / t / 1e9}
    p_vol.close()
    # config 2 frame tile-sharded (strong scaling of one 1080p frame)
    c2 = clone(cam)
    lp.scene_cam_recalculate(c2, 1920, 1080, 1)
    for mode in ("exact", "fast"):
        (rgba, pts, ev), t = timed(lambda: ld.render_frame_sharded(c2, prm, seq, lights, n, 1920, 1080, mode=mode, want_points=False))
        tot = ev.clone()
        dist.all_reduce(tot)
        out[f"frame1080_tiles_{mode}"] = {"s": t, "giter_s": int(tot.item()) * iters / t / 1e9, "fps": 1 / t}
    # config 3: 4K, A6B6C6, 72 + 4032 iterations, tile-sharded
    p3 = clone(prm)
    p3.settle, p3.accum = 72, 4032
    s3 = lp.scene_convert_sequence("A6B6C6")
    c3 = clone(cam)
    lp.scene_cam_recalculate(c3, 3840, 2160, 1)
    for mode in ("exact", "fast"):
        (rgba, pts, ev), t = timed(lambda: ld.render_frame_sharded(c3, p3, s3, lights, n, 3840, 2160, mode=mode, want_points=False), reps=1)
        tot = ev.clone()
        dist.all_reduce(tot)
        out[f"frame4k_long_{mode}"] = {"s": t, "giter_s": int(tot.item()) * 4104 / t / 1e9, "fps": 1 / t, "evals": int(tot.item())}
        if rank == 0 and mode == "exact":
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            api.write_png(os.path.join(ROOT, "gpurun_out", "frame4k_long_exact.png"), rgba.cpu().numpy())
    # the same two frames through peer memory (no gather, no scatter)
    for name, (cc, pp, ss, ww, hh, per) in {"frame1080_tiles": (c2, prm, seq, 1920, 1080, iters), "frame4k_long": (c3, p3, s3, 3840, 2160, 4104)}.items():
        p_rgba, p_pts = ld.PeerBuffer(ww * hh * 4), ld.PeerBuffer(ww * hh * 36)
        for mode in ("exact", "fast"):
            ev = torch.zeros(1, dtype=torch.int64, device=dev)
            _, t = timed(lambda: ld.render_frame_sharded_peer(p_rgba, p_pts, cc, pp, ss, d_lights, n, ww, hh, mode=mode, evals=ev), reps=1)
            dist.all_reduce(ev)
            out[f"{name}_{mode}_peer"] = {"s": t, "giter_s": int(ev.item()) * per / t / 1e9, "fps": 1 / t}
        p_rgba.close(); p_pts.close()
    # config 5: 120-frame 1080p orbit, frames dealt to ranks
    n_frames = 120 if world >= 4 else 8 * world
    count = [0]
    _, t = timed(lambda: ld.render_animation_sharded(n_frames, 1920, 1080, prm, cam, seq, lights, n, mode="exact",
                                                     on_frame=lambda f, fr: count.__setitem__(0, count[0] + 1)), reps=1)
    out["orbit1080_exact"] = {"frames": n_frames, "s": t, "fps": n_frames / t}
    if rank == 0:
        print("DIST_TIMING " + json.dumps(out), flush=True)
dist.destroy_process_group()
