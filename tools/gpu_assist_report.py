#!/usr/bin/env python
"""Volume-assisted march: evaluations saved and parity against the non-assisted hybrid
render, over grid size / margin / upper bound / dilation -> gpurun_out/r02_assist_report.json."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import lyapunov3d_b200 as lp
from lyapunov3d_b200 import api
from lyapunov3d_b200.structs import POINT_DTYPE, clone

prm, cam, lights, nl, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, nl); seq = lp.scene_convert_sequence(s)
prm.jitter = 0.0
w, h = 1920, 1080
lp.scene_cam_recalculate(cam, w, h, 1)
dl = api.upload_lights(lights)
def timed(fn, reps=2):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.perf_counter(); out = fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t)
    return best * 1e3, out
ms_exact, ex = timed(lambda: lp.render(cam, prm, seq, dl, nl, w, h, mode="exact"))
ms_hyb, hy = timed(lambda: lp.render(cam, prm, seq, dl, nl, w, h, mode="hybrid"))
base_pts = hy[1].cpu().numpy().view(POINT_DTYPE)[..., 0]; base_rgba = hy[0].cpu().numpy(); base_ev = int(hy[2].item())
rep = {"frame": "1920x1080 default scene, jitter 0", "exact_ms": ms_exact, "hybrid_ms": ms_hyb, "evaluations": base_ev,
       "evaluations_per_pixel": base_ev / (w * h), "cases": []}
print(json.dumps({k: v for k, v in rep.items() if k != "cases"}), flush=True)
for n in (128, 256, 512):
    ms_bake, vol = timed(lambda: lp.bake(prm, seq, n, mode="fast", dtype="f16"))
    for margin, upper, dil in ((0.25, 0.0, 1), (0.25, float("inf"), 1), (0.1, float("inf"), 1), (0.25, float("inf"), 0), (0.5, float("inf"), 2)):
        ms_build, bits = timed(lambda: api.assist_build(vol, prm, margin=margin, upper=upper, dilate=dil))
        ms, out = timed(lambda: api.render_assisted(cam, prm, seq, dl, nl, w, h, vol, bits))
        pts = out[1].cpu().numpy().view(POINT_DTYPE)[..., 0]
        same = np.ones(pts.shape, bool)
        for f in ("P", "N", "l"):
            eq = (pts[f].view(np.uint32) == base_pts[f].view(np.uint32)) | (np.isnan(pts[f]) & np.isnan(base_pts[f]))
            same &= eq.reshape(pts.shape + (-1,)).all(-1)
        rgba = out[0].cpu().numpy()
        d = np.abs(rgba.astype(np.int16) - base_rgba.astype(np.int16)).max(-1)
        case = {"grid": n, "margin": margin, "upper": upper if np.isfinite(upper) else "inf", "dilate": dil, "bake_ms": ms_bake, "build_ms": ms_build,
                "safe_cells_frac": float(sum(bin(int(x) & 0xffffffff).count("1") for x in bits[::997].cpu().numpy()) / (32.0 * len(bits[::997]))),
                "render_ms": ms, "speedup_vs_hybrid": ms_hyb / ms, "speedup_vs_exact": ms_exact / ms,
                "evaluations": int(out[2].item()), "skipped": int(out[3].item()), "evaluations_saved_per_pixel": (base_ev - int(out[2].item())) / (w * h),
                "records_P_N_l_identical": float(same.mean()), "pixels_identical": float((d == 0).mean()), "pixels_within_2_of_255": float((d <= 2).mean())}
        rep["cases"].append(case)
        print(json.dumps(case), flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "r02_assist_report.json"), "w"), indent=1)
