"""Multi-GPU headless front end (one process per GPU; run under torchrun for N > 1).

    python -m lyapunov3d_b200.cli frame  --width 3840 --height 2160 --seq A6B6C6 --settle 72 --accum 4032 --out out/
    python -m lyapunov3d_b200.cli bake   --n 512 --mode fast --out exps.raw
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 -m lyapunov3d_b200.cli orbit --frames 120 --out out/

frame: one frame, 8x8 tiles dealt round-robin to the ranks, written by every rank's kernel straight
       into rank 0's buffers (peer memory), saved as PNG (+ P3 PPM / raw LyapPoint dump on request)
bake:  z-slabs of the volume, same mechanism, saved as the reference's exps.raw
orbit: the scale.pl camera path, whole frames dealt to the ranks, one PNG per frame

The single-GPU C++ programs are lyapunov3d_b200/bin/lyap_render and lyap_calculate.
"""
import argparse
import os
import time

import numpy as np


def _scene(args):
    import lyapunov3d_b200 as lp
    prm, cam, lights, n_lights, seq_s, _ = lp.params_init()
    seq_s = args.seq or seq_s
    for k in ("settle", "accum"):
        if getattr(args, k) is not None:
            setattr(prm, k, getattr(args, k))
    for k in ("d", "depth", "jitter", "refine"):
        if getattr(args, k, None) is not None:
            setattr(prm, k, getattr(args, k))
    lp.scene_lights_recalculate(lights, n_lights)
    return prm, cam, lights, n_lights, seq_s, lp.scene_convert_sequence(seq_s)


def main(argv=None):
    ap = argparse.ArgumentParser(prog="lyapunov3d_b200.cli")
    ap.add_argument("what", choices=["frame", "bake", "orbit"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--n", type=int, default=512, help="voxels per edge (bake)")
    ap.add_argument("--frames", type=int, default=120, help="frames of the orbit")
    ap.add_argument("--seq", default=None)
    ap.add_argument("--settle", type=int, default=None)
    ap.add_argument("--accum", type=int, default=None)
    ap.add_argument("--d", type=float, default=None)
    ap.add_argument("--depth", type=float, default=None)
    ap.add_argument("--jitter", type=float, default=None)
    ap.add_argument("--refine", type=float, default=None)
    ap.add_argument("--mode", default=None, choices=["exact", "fast", "host"])
    ap.add_argument("--f16", action="store_true")
    ap.add_argument("--ppm", action="store_true")
    ap.add_argument("--points", action="store_true")
    ap.add_argument("--out", default=".")
    args = ap.parse_args(argv)

    import torch
    import torch.distributed as dist

    import lyapunov3d_b200 as lp
    from lyapunov3d_b200 import api, dist as ld
    from lyapunov3d_b200.structs import clone

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    prm, cam, lights, n_lights, seq_s, seq = _scene(args)
    iters = prm.settle + prm.accum
    d_lights = api.upload_lights(lights, dev)
    stamp = int(time.time())

    if args.what == "bake":
        mode = args.mode or "fast"
        n = args.n
        vol = ld.PeerBuffer(n ** 3 * (2 if args.f16 else 4))
        t0 = time.perf_counter()
        ld.bake_sharded_peer(vol, prm, seq, n, n, n, mode=mode, f16=args.f16)
        dt = time.perf_counter() - t0
        if rank == 0:
            host = vol.view((n, n, n), "<f2" if args.f16 else "<f4").tensor().cpu().numpy()
            api.write_raw(args.out if args.out.endswith(".raw") else os.path.join(args.out, "exps.raw"), host)
            print(f"baked {n}^3 x {iters} iterations on {world} GPU(s) in {dt * 1e3:.2f} ms ({n ** 3 * iters / dt / 1e9:.0f} Giter/s)")
        vol.close()
    elif args.what == "frame":
        mode = args.mode or "exact"
        w, h = args.width, args.height
        c = clone(cam)
        lp.scene_cam_recalculate(c, w, h, 1)
        rgba, pts = ld.PeerBuffer(w * h * 4), ld.PeerBuffer(w * h * 36)     # rank 0 zero-fills them
        if world > 1:
            dist.barrier()
        ev = torch.zeros(1, dtype=torch.int64, device=dev)
        t0 = time.perf_counter()
        ld.render_frame_sharded_peer(rgba, pts, c, prm, seq, d_lights, n_lights, w, h, mode=mode, evals=ev)
        dt = time.perf_counter() - t0
        if world > 1:
            dist.all_reduce(ev)
        if rank == 0:
            os.makedirs(args.out, exist_ok=True)
            secs = int(dt)
            stem = api.format_filename("Render", stamp, w, h, seq_s, c, prm) + "_time=%dh%02dm%02ds" % (secs // 3600, secs // 60 % 60, secs % 60)
            img = rgba.view((h, w, 4), "|u1").tensor().cpu().numpy()
            api.write_png(os.path.join(args.out, stem + ".png"), img)
            if args.ppm:
                api.write_ppm(os.path.join(args.out, stem + ".ppm"), img)
            if args.points:
                api.write_raw(os.path.join(args.out, stem.replace("Render_", "Points_", 1) + ".raw"),
                              pts.view((h, w, 36), "|u1").tensor().cpu().numpy())
            print(f"{w}x{h} frame on {world} GPU(s): {dt * 1e3:.1f} ms, {int(ev.item())} evaluations, "
                  f"{int(ev.item()) * iters / dt / 1e9:.0f} Giter/s -> {stem}.png")
        rgba.close()
        pts.close()
    else:
        mode = args.mode or "exact"
        w, h = args.width, args.height
        if rank == 0:
            os.makedirs(args.out, exist_ok=True)
        t0 = time.perf_counter()

        def save(f, frame):
            api.write_png(os.path.join(args.out, "orbit_%04d.png" % f), frame.cpu().numpy())

        ld.render_animation_sharded(args.frames, w, h, prm, cam, seq, d_lights, n_lights, mode=mode, on_frame=save)
        if rank == 0:
            dt = time.perf_counter() - t0
            print(f"{args.frames} frames {w}x{h} on {world} GPU(s) in {dt:.2f} s ({args.frames / dt:.2f} frames/s incl. PNG encoding)")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
