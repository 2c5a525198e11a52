import os, sys
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np, torch, lyapunov3d_b200 as lp
from lyapunov3d_b200.structs import POINT_DTYPE, clone
from oracle import Oracle
o = Oracle(); o.set_threads(len(os.sched_getaffinity(0)))
prm, cam, lights, nl, s, _ = lp.params_init(); lp.scene_lights_recalculate(lights, nl); seq = lp.scene_convert_sequence(s)
w = h = 192
c = clone(cam); lp.scene_cam_recalculate(c, w, h, 1)
o_rgba, o_pts, calls = o.render(c, prm, seq, lights, nl, w, h)
rgba, pts, ev = lp.render(c, prm, seq, lights, nl, w, h, mode="host")
g = pts.cpu().numpy().view(POINT_DTYPE)[..., 0]
same = (g.view(np.uint8).reshape(-1, 36) == o_pts.view(np.uint8).reshape(-1, 36)).all(axis=1).reshape(h, w)
ys, xs = np.nonzero(~same)
print("mismatching records:", len(ys), "of", w * h)
for y, x in list(zip(ys, xs))[:12]:
    a, b = g[y, x], o_pts[y, x]
    fields = [f for f in ("P", "N", "a", "c", "l") if not np.array_equal(np.atleast_1d(a[f]).view(np.uint32), np.atleast_1d(b[f]).view(np.uint32))]
    print((x, y), "fields", fields, "gpu N", a["N"], "cpu N", b["N"], "dN", np.abs(a["N"] - b["N"]).max(), "l", float(a["l"]), float(b["l"]), "P eq", np.array_equal(a["P"], b["P"]))
    # evaluate the 6 normal samples' exponents on both sides at the GPU's hit point
    P = b["P"]
    # dt = Fdt or Fdt/2: try both magnitudes
# direct exponent comparison on many points near the surface
rng = np.random.default_rng(3)
hit = o_pts["P"].reshape(-1, 3)
hit = hit[(hit != 0).any(1)]
xyz = (hit[rng.integers(0, len(hit), 400000)] + rng.normal(0, 3e-6, (400000, 3))).astype(np.float32)
lo = o.lyap4d_many(xyz, prm.d, prm.settle, prm.accum, seq)
lg = lp.exponent_points(torch.from_numpy(xyz).cuda(), prm, seq, mode="host").cpu().numpy()
neq = (lo.view(np.uint32) != lg.view(np.uint32)) & ~(np.isnan(lo) & np.isnan(lg))
print("exponent mismatches near the surface:", int(neq.sum()), "of", len(xyz), "max abs diff", float(np.abs(lo - lg)[neq].max()) if neq.any() else 0.0)
xyz2 = rng.uniform(0, 4, (400000, 3)).astype(np.float32)
lo = o.lyap4d_many(xyz2, prm.d, prm.settle, prm.accum, seq)
lg = lp.exponent_points(torch.from_numpy(xyz2).cuda(), prm, seq, mode="host").cpu().numpy()
neq2 = (lo.view(np.uint32) != lg.view(np.uint32)) & ~(np.isnan(lo) & np.isnan(lg))
print("exponent mismatches, uniform points:", int(neq2.sum()), "of", len(xyz2))
if neq.any():
    i = np.nonzero(neq)[0][0]
    print("example", xyz[i], float.hex(float(lo[i])), float.hex(float(lg[i])))
    np.save("gpurun_out/host_mismatch_points.npy", xyz[neq][:64])
