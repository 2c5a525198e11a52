#!/usr/bin/env python
"""Quick GPU check: peaks, bake parity/timing in all modes, 1080p frame timing."""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lyapunov3d_b200 as lp
from lyapunov3d_b200 import api
from lyapunov3d_b200.structs import clone
from oracle import Oracle

def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return r, best

pk = api.probe_peaks(); print("peaks", {k: (v/1e12 if 'ops' in k else v) for k, v in pk.items()}, flush=True)
o = Oracle()
prm, cam, lights, nl, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, nl); seq = lp.scene_convert_sequence(s)
iters = prm.settle + prm.accum
for dims in ((64, 64, 64), (21, 13, 7)):
    ref = o.bake(prm, seq, *dims)
    for mode in ("fast", "exact", "host"):
        v = lp.bake(prm, seq, *dims, mode=mode).cpu().numpy()
        print("bake", dims, mode, "nan_same", bool((np.isnan(v) == np.isnan(ref)).all()), "maxerr", float(np.nanmax(np.abs(v - ref))), flush=True)
p2 = clone(prm); p2.settle, p2.accum = 72, 4032
for sq in ("A6B6C6", "A9B9C9D9", "AB"):
    sq_ = lp.scene_convert_sequence(sq); ref = o.bake(p2, sq_, 16)
    v = lp.bake(p2, sq_, 16, mode="fast").cpu().numpy()
    print("bake long", sq, "nan_same", bool((np.isnan(v) == np.isnan(ref)).all()), "maxerr", float(np.nanmax(np.abs(v - ref))), flush=True)
for mode, dt in (("fast", "f32"), ("fast", "f16"), ("exact", "f32")):
    buf = torch.empty((512, 512, 512), dtype=torch.float16 if dt == "f16" else torch.float32, device="cuda")
    _, ms = timed(lambda: lp.bake(prm, seq, 512, mode=mode, dtype=dt, out=buf))
    git = 512 ** 3 * iters / ms / 1e6
    peak = (pk["ffma_lane_ops_per_s"] / 3.965 if mode == "fast" else pk["mufu_lane_ops_per_s"] * 1026 / 1008) / 1e9
    print("bake512", mode, dt, "ms %.3f Giter/s %.1f frac %.3f" % (ms, git, git / peak), flush=True)
    del buf
c = clone(cam); lp.scene_cam_recalculate(c, 1920, 1080, 1)
for mode in sys.argv[1:] or ("exact", "fast"):
    (rgba, pts, ev), ms = timed(lambda: lp.render(c, prm, seq, lights, nl, 1920, 1080, mode=mode), reps=2)
    git = int(ev.item()) * iters / ms / 1e6
    peak = (pk["ffma_lane_ops_per_s"] / 3.965 if mode == "fast" else pk["mufu_lane_ops_per_s"] * 1026 / 1008) / 1e9
    print("frame1080", mode, "ms %.2f Giter/s %.1f frac %.3f" % (ms, git, git / peak), flush=True)
