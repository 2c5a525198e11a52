#!/usr/bin/env python
"""Diagnostic: field-by-field comparison of LyapPoint between ours(exact), ref CUDA, oracle."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import lyapunov3d_b200 as lp
from helpers import frac_within
from lyapunov3d_b200.structs import POINT_DTYPE, clone
from oracle import Oracle, RefCuda
o = Oracle(); rc = RefCuda()
prm, cam, lights, nl, seq_s, _ = lp.params_init()
lp.scene_lights_recalculate(lights, nl)
seq = lp.scene_convert_sequence(seq_s)
w = h = int(sys.argv[1]) if len(sys.argv) > 1 else 96
c = clone(cam); lp.scene_cam_recalculate(c, w, h, 1)
r_rgba, r_pts, _ = rc.render(c, prm, seq, lights, nl, w, h)
g_rgba, g_pts, _ = lp.render(c, prm, seq, lights, nl, w, h, mode="exact")
g_rgba = g_rgba.cpu().numpy(); g_pts = g_pts.cpu().numpy().view(POINT_DTYPE)[..., 0]
o_rgba, o_pts, _ = o.render(c, prm, seq, lights, nl, w, h)
def cmp(name, A, B, Ai, Bi):
    print("==", name)
    for f in ("P", "N", "a", "c", "l"):
        a, b = A[f], B[f]
        eq = (a.view(np.uint32) == b.view(np.uint32))
        if eq.ndim == 3: eq = eq.all(-1)
        close = np.abs(a - b) < 1e-4
        if close.ndim == 3: close = close.all(-1)
        print(f"  {f}: bit-equal {eq.mean():.4f}  |d|<1e-4 {close.mean():.4f}")
    print("  rgba identical %.4f within2 %.4f" % ((Ai == Bi).all(-1).mean(), frac_within(Ai, Bi)))
cmp("ours-exact vs refcuda", g_pts, r_pts, g_rgba, r_rgba)
cmp("ours-exact vs oracle", g_pts, o_pts, g_rgba, o_rgba)
cmp("refcuda vs oracle", r_pts, o_pts, r_rgba, o_rgba)
for (y, x) in ((h // 2, w // 2), (h // 3, w // 3), (5, 7), (h - 3, w - 9)):
    print("pixel", x, y)
    for n, P, I in (("ours", g_pts, g_rgba), ("refc", r_pts, r_rgba), ("orac", o_pts, o_rgba)):
        p = P[y, x]
        print("  ", n, [float.hex(float(v)) for v in p["P"]], [round(float(v), 6) for v in p["N"]], float(p["a"]), float(p["c"]), float(p["l"]), I[y, x].tolist())
# shade-only check: shade the reference's points with our shade kernel
t = torch.from_numpy(r_pts.view(np.uint8).reshape(h, w, 36).copy()).cuda()
s_rgba = lp.shade_points(t, c, lights, nl, mode="exact").cpu().numpy()
print("our shade(exact) of ref points vs ref rgba: identical %.4f within2 %.4f" % ((s_rgba == r_rgba).all(-1).mean(), frac_within(s_rgba, r_rgba)))
t = torch.from_numpy(o_pts.view(np.uint8).reshape(h, w, 36).copy()).cuda()
s_rgba = lp.shade_points(t, c, lights, nl, mode="host").cpu().numpy()
print("our shade(host) of oracle points vs oracle rgba: identical %.4f within2 %.4f" % ((s_rgba == o_rgba).all(-1).mean(), frac_within(s_rgba, o_rgba)))
