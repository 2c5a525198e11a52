#!/usr/bin/env python
"""First-light GPU script: peaks, parity numbers and timings in one gpurun call.
Writes gpurun_out/first_light.json.  Development tool (uses the oracle as checker)."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import lyapunov3d_b200 as lp  # noqa: E402
from helpers import frac_within, same_floats  # noqa: E402
from lyapunov3d_b200 import api  # noqa: E402
from lyapunov3d_b200.structs import POINT_DTYPE, clone  # noqa: E402
from oracle import Oracle, RefCuda  # noqa: E402

out = {}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)


def save():
    with open(os.path.join(ROOT, "gpurun_out", "first_light.json"), "w") as f:
        json.dump(out, f, indent=1, default=float)


def timed(fn, reps=3):
    best = 1e30
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return r, best


print(torch.cuda.get_device_name(0), flush=True)
pk = api.probe_peaks()
pk["ffma_per_clk_sm"] = pk["ffma_lane_ops_per_s"] / pk["sm_clock_hz"] / pk["sm_count"]
pk["mufu_per_clk_sm"] = pk["mufu_lane_ops_per_s"] / pk["sm_clock_hz"] / pk["sm_count"]
out["peaks"] = pk
print("peaks", pk, flush=True)
save()

o = Oracle()
prm, cam, lights, nl, seq_s, _ = lp.params_init()
lp.scene_lights_recalculate(lights, nl)
seq = lp.scene_convert_sequence(seq_s)
iters = prm.settle + prm.accum

# ---- bake parity at 64^3
ref64 = o.bake(prm, seq, 64)
bk = {}
for mode in ("exact", "fast", "host"):
    v = lp.bake(prm, seq, 64, mode=mode).cpu().numpy()
    nan_same = bool((np.isnan(v) == np.isnan(ref64)).all())
    err = float(np.nanmax(np.abs(v - ref64)))
    bk[mode] = {"nan_same": nan_same, "max_abs_err": err, "bit_equal": same_floats(v, ref64)}
lp.api.set_option("force_generic", 1)
for mode in ("exact", "fast", "host"):
    v = lp.bake(prm, seq, 64, mode=mode).cpu().numpy()
    bk[mode + "_generic"] = {"max_abs_err": float(np.nanmax(np.abs(v - ref64))), "bit_equal": same_floats(v, ref64),
                             "nan_same": bool((np.isnan(v) == np.isnan(ref64)).all())}
lp.api.set_option("force_generic", 0)
out["bake_parity_64"] = bk
print("bake parity", bk, flush=True)
save()

# ---- bake throughput
bt = {}
for n in (256, 512):
    for mode in ("exact", "fast", "host"):
        if mode == "host" and n == 512:
            continue
        buf = torch.empty((n, n, n), dtype=torch.float32, device="cuda")
        _, ms = timed(lambda: lp.bake(prm, seq, n, mode=mode, out=buf))
        git = n ** 3 * iters / ms / 1e6
        bt[f"{mode}_{n}"] = {"ms": ms, "giter_s": git, "iter_per_clk_sm": git * 1e9 / pk["sm_clock_hz"] / pk["sm_count"]}
        print("bake", n, mode, bt[f"{mode}_{n}"], flush=True)
        del buf
buf = torch.empty((512, 512, 512), dtype=torch.float16, device="cuda")
_, ms = timed(lambda: lp.bake(prm, seq, 512, mode="fast", dtype="f16", out=buf))
bt["fast_512_f16"] = {"ms": ms, "giter_s": 512 ** 3 * iters / ms / 1e6}
del buf
out["bake_throughput"] = bt
save()

# ---- reference CUDA kernel
rc = RefCuda()
rt = {}
vref, ms = rc.bake(prm, seq, 256, reps=2)
rt["bake_256"] = {"ms": ms, "giter_s": 256 ** 3 * iters / ms / 1e6}
v = lp.bake(prm, seq, 256, mode="exact").cpu().numpy()
rt["bake_256_exact_vs_refcuda"] = {"bit_equal": same_floats(v, vref), "max_abs_err": float(np.nanmax(np.abs(v - vref))),
                                   "n_diff": int((~((v.view(np.uint32) == vref.view(np.uint32)) | (np.isnan(v) & np.isnan(vref)))).sum())}
print("ref cuda bake", rt, flush=True)
out["ref_cuda"] = rt
save()

# ---- render parity: exact vs reference CUDA kernel, host vs oracle
rp = {}
for (w, h) in ((256, 256), (640, 360)):
    c = clone(cam)
    lp.scene_cam_recalculate(c, w, h, 1)
    r_rgba, r_pts, r_ms = rc.render(c, prm, seq, lights, nl, w, h, reps=1)
    (g_rgba, g_pts, g_ev), g_ms = timed(lambda: lp.render(c, prm, seq, lights, nl, w, h, mode="exact"), reps=2)
    g_rgba = g_rgba.cpu().numpy()
    g_pts = g_pts.cpu().numpy().view(POINT_DTYPE)[..., 0]
    pts_equal = float((g_pts.view(np.uint8).reshape(h * w, 36) == r_pts.view(np.uint8).reshape(h * w, 36)).all(axis=1).mean())
    evals = int(g_ev.item())
    rp[f"exact_vs_refcuda_{w}x{h}"] = {
        "pixels_identical": float((g_rgba == r_rgba).all(-1).mean()), "within_2": frac_within(g_rgba, r_rgba),
        "points_bit_identical": pts_equal, "ref_ms": r_ms, "ours_ms": g_ms, "evals": evals,
        "ours_giter_s": evals * iters / g_ms / 1e6, "ref_giter_s": evals * iters / r_ms / 1e6}
    print("render", w, h, rp[f"exact_vs_refcuda_{w}x{h}"], flush=True)
    (f_rgba, _, f_ev), f_ms = timed(lambda: lp.render(c, prm, seq, lights, nl, w, h, mode="fast"), reps=2)
    f_rgba = f_rgba.cpu().numpy()
    rp[f"fast_vs_refcuda_{w}x{h}"] = {"within_2": frac_within(f_rgba, r_rgba), "ms": f_ms,
                                      "giter_s": int(f_ev.item()) * iters / f_ms / 1e6}
    print("render fast", rp[f"fast_vs_refcuda_{w}x{h}"], flush=True)
    out["render_parity"] = rp
    save()

w = h = 128
c = clone(cam)
lp.scene_cam_recalculate(c, w, h, 1)
t = time.time()
o_rgba, o_pts, o_calls = o.render(c, prm, seq, lights, nl, w, h)
cpu_s = time.time() - t
(g_rgba, g_pts, g_ev), g_ms = timed(lambda: lp.render(c, prm, seq, lights, nl, w, h, mode="host"), reps=2)
g_rgba = g_rgba.cpu().numpy()
g_pts = g_pts.cpu().numpy().view(POINT_DTYPE)[..., 0]
rp["host_vs_oracle_128"] = {
    "pixels_identical": float((g_rgba == o_rgba).all(-1).mean()), "within_2": frac_within(g_rgba, o_rgba),
    "points_bit_identical": float((g_pts.view(np.uint8).reshape(h * w, 36) == o_pts.view(np.uint8).reshape(h * w, 36)).all(axis=1).mean()),
    "oracle_calls": o_calls, "gpu_evals": int(g_ev.item()), "ms": g_ms, "cpu_s": cpu_s, "cpu_threads": o.threads(),
    "cpu_giter_s": o_calls * iters / cpu_s / 1e9, "gpu_giter_s": int(g_ev.item()) * iters / g_ms / 1e6}
print("host vs oracle", rp["host_vs_oracle_128"], flush=True)
e_rgba = lp.render(c, prm, seq, lights, nl, w, h, mode="exact")[0].cpu().numpy()
rp["exact_vs_oracle_128"] = {"within_2": frac_within(e_rgba, o_rgba), "identical": float((e_rgba == o_rgba).all(-1).mean())}
print("exact vs oracle(host)", rp["exact_vs_oracle_128"], flush=True)
out["render_parity"] = rp
save()

# ---- 1080p timing, warps/SM sweep
rt2 = {}
w, h = 1920, 1080
c = clone(cam)
lp.scene_cam_recalculate(c, w, h, 1)
for mode in ("exact", "fast"):
    for wps in (8, 12, 16, 20):
        api.set_option("render_warps_per_sm", wps)
        (rgba, pts, ev), ms = timed(lambda: lp.render(c, prm, seq, lights, nl, w, h, mode=mode), reps=2)
        git = int(ev.item()) * iters / ms / 1e6
        rt2[f"{mode}_w{wps}"] = {"ms": ms, "giter_s": git, "iter_per_clk_sm": git * 1e9 / pk["sm_clock_hz"] / pk["sm_count"], "evals": int(ev.item())}
        print("1080p", mode, wps, rt2[f"{mode}_w{wps}"], flush=True)
        out["render_1080p"] = rt2
        save()
api.set_option("render_warps_per_sm", 0)
r_rgba, r_pts, r_ms = rc.render(c, prm, seq, lights, nl, w, h, reps=1)
rgba = lp.render(c, prm, seq, lights, nl, w, h, mode="exact")[0].cpu().numpy()
rt2["refcuda"] = {"ms": r_ms, "within_2_vs_exact": frac_within(rgba, r_rgba), "identical": float((rgba == r_rgba).all(-1).mean())}
print("1080p ref", rt2["refcuda"], flush=True)
out["render_1080p"] = rt2
api.write_png(os.path.join(ROOT, "gpurun_out", "first_light_1080p.png"), rgba)
save()
print("DONE")
