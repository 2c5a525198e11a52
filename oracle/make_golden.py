#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref_host.so).

TEST INFRASTRUCTURE.  Run in the build container (needs /root/reference to have
been compiled by `make -C oracle`):

    PYTHONPATH=. python oracle/make_golden.py

The fixtures are the reference's own outputs on fixed inputs; the C restatement
(oracle/lyap_oracle.c) and, on the GPU, the CUDA path are checked against them.
Nothing here is random except the exponent sample points, which come from a
fixed numpy seed and are stored alongside the results.
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from lyapunov3d_b200.structs import Color, Quat, Vec3, clone, struct_bytes  # noqa: E402
from oracle import RefHost  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def raw(s):
    return np.frombuffer(struct_bytes(s), np.uint8).copy()


def two_light_scene(lights):
    """The reference's alternative light rig (params.cu:65-93, the `if (false)` branch)."""
    L = lights
    L[0].C = Vec3(5.0, 7.0, 3.0)
    L[0].Q = Quat(0.710595, 0.282082, -0.512168, 0.391368)
    L[0].M = 0.5
    L[0].lightRange = 1.0
    L[0].ambient = Color(0, 0, 0, 0)
    L[0].diffuseColor = Color(0.30, 0.40, 0.50, 1)
    L[0].diffusePower = 10.0
    L[0].specularColor = Color(0.90, 0.90, 0.90, 1)
    L[0].specularPower = 10.0
    L[0].specularHardness = 10.0
    L[0].chaosColor = Color(0, 0, 1, 0.1)
    L[1].C = Vec3(3, 7, 5)
    L[1].Q = Quat(0.039640, 0.840027, -0.538582, -0.052093)
    L[1].M = 1.6772
    L[1].lightRange = 0.5
    L[1].ambient = Color(0, 0, 0, 0)
    L[1].diffuseColor = Color(0.3, 0.374694, 0.2, 1)
    L[1].diffusePower = 10.0
    L[1].specularColor = Color(1, 1, 1, 1)
    L[1].specularPower = 10.0
    L[1].specularHardness = 10.0
    L[1].chaosColor = Color(0, 0, 1, 1)
    return 2


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = RefHost()
    prm, cam, lights, n_lights, seq_str, (dw, dh) = ref.params_init()

    # ---- host scene --------------------------------------------------------
    g = {"prm": raw(prm), "cam": raw(cam), "lights": raw(lights), "n_lights": n_lights,
         "sequence": np.array(seq_str), "default_size": np.array([dw, dh])}
    seq_cases = ["BCABA", "A6B6C6", "A6B6C6D6", "ab", "3A", "AbC9d1", "D", "AAB2"]
    g["seq_strings"] = np.array(seq_cases)
    for k, s in enumerate(seq_cases):
        g[f"seq_{k}"] = ref.convert_sequence(s)
    sizes = [(256, 256, 1), (1920, 1080, 1), (3840, 2160, 1), (3840, 2160, 2), (64, 36, 1), (100, 50, 3)]
    g["cam_sizes"] = np.array(sizes)
    for k, (w, h, d) in enumerate(sizes):
        c = clone(cam)
        ref.cam_recalculate(c, w, h, d)
        g[f"cam_recalc_{k}"] = raw(c)
    lr = clone(lights)
    ref.lights_recalculate(lr, n_lights)
    g["lights_recalc"] = raw(lr)
    path_i = np.concatenate([np.linspace(0.0, 1.0, 21), [1e-7, 0.9999999, 0.123456789012345]])
    g["campath_i"] = path_i
    g["campath_cams"] = np.stack([raw((lambda c: (ref.campath(i, c), c)[1])(clone(cam))) for i in path_i])
    np.savez_compressed(os.path.join(OUT, "scene.npz"), **g)

    # ---- exponent ------------------------------------------------------------
    rng = np.random.default_rng(20180518)
    e = {}
    xyz = rng.uniform(0.0, 4.0, (4096, 3)).astype(np.float32)
    xyz[:64] *= np.float32(0.25)             # some points in the r < 1 corner (orbit decays to 0)
    xyz[64:72, 0] = 0.0                      # on the x = 0 face: zero derivative -> NaN
    xyz[72:80] = np.float32(2.0)             # r = 2 fixed point v = 0.5: exponent 0 / superstable
    e["xyz_default"] = xyz
    e["l_default"] = ref.lyap4d_many(xyz, prm.d, prm.settle, prm.accum, ref.convert_sequence("BCABA"))
    xyz2 = rng.uniform(0.0, 4.0, (768, 3)).astype(np.float32)
    e["xyz_long"] = xyz2
    e["l_long"] = ref.lyap4d_many(xyz2, prm.d, 72, 4032, ref.convert_sequence("A6B6C6"))
    e["l_d_symbol"] = ref.lyap4d_many(xyz2, 3.7, 10, 500, ref.convert_sequence("A6B6C6D6"))
    e["l_odd_counts"] = ref.lyap4d_many(xyz2, prm.d, 7, 333, ref.convert_sequence("AAB2"))
    e["l_no_settle"] = ref.lyap4d_many(xyz2, prm.d, 0, 100, ref.convert_sequence("AB"))
    np.savez_compressed(os.path.join(OUT, "exponent.npz"), **e)

    # ---- bake ------------------------------------------------------------------
    b = {"default_32": ref.bake(prm, ref.convert_sequence("BCABA"), 32),
         "ragged_20x12x9": ref.bake(prm, ref.convert_sequence("BCABA"), 20, 12, 9)}
    p2 = clone(prm)
    p2.settle, p2.accum = 72, 4032
    b["long_16"] = ref.bake(p2, ref.convert_sequence("A6B6C6"), 16)
    np.savez_compressed(os.path.join(OUT, "bake.npz"), **b)

    # ---- frames ------------------------------------------------------------------
    f = {}

    def frame(name, w, h, prm_, seq_s, lights_, n_l, cam_=None):
        c = clone(cam if cam_ is None else cam_)
        ref.cam_recalculate(c, w, h, 1)
        rgba, pts, _ = ref.render(c, prm_, ref.convert_sequence(seq_s), lights_, n_l, w, h)
        f[name + "_rgba"] = rgba
        f[name + "_points"] = pts.view(np.uint8).reshape(h, w, 36)
        f[name + "_cam"] = raw(c)
        f[name + "_prm"] = raw(prm_)
        f[name + "_lights"] = raw(lights_)
        f[name + "_n_lights"] = n_l
        f[name + "_seq"] = np.array(seq_s)

    frame("default_48", 48, 48, prm, "BCABA", lr, n_lights)
    frame("default_40x24", 40, 24, prm, "BCABA", lr, n_lights)
    pj = clone(prm)
    pj.jitter = 0.0
    frame("nojitter_32", 32, 32, pj, "BCABA", lr, n_lights)
    pm = clone(prm)
    pm.stepMethod = 1
    pm.depth = 512
    frame("method1_24", 24, 24, pm, "BCABA", lr, n_lights)
    pl = clone(prm)
    pl.settle, pl.accum = 72, 4032
    frame("long_24x16", 24, 16, pl, "A6B6C6", lr, n_lights)
    l2 = clone(lights)
    n2 = two_light_scene(l2)
    ref.lights_recalculate(l2, n2)
    frame("twolights_32", 32, 32, prm, "BCABA", l2, n2)
    # a camera far enough out that many rays miss the cube entirely
    cfar = clone(cam)
    ref.campath(0.0, cfar)
    cfar.M = 1.2
    frame("wide_32", 32, 32, prm, "BCABA", lr, n_lights, cfar)
    np.savez_compressed(os.path.join(OUT, "frames.npz"), **f)

    # ---- shade / pixel conversion unit vectors ---------------------------------------
    s = {}
    pts = np.frombuffer(f["default_48_points"].tobytes(), dtype=np.uint8).reshape(-1, 36)[::37][:48].copy()
    c48 = clone(cam)
    ref.cam_recalculate(c48, 48, 48, 1)
    cols = np.stack([ref.shade(p, c48, l2, n2) for p in pts])
    s["points"] = pts
    s["colors_twolights"] = cols
    s["colors_default"] = np.stack([ref.shade(p, c48, lr, n_lights) for p in pts])
    vals = np.array([[-3.5, 0.5, 1.7, 300.0], [np.nan, np.inf, -np.inf, 1e10], [0.999, 1.0, 1.004, -0.0],
                     [0.1, 0.0, 0.0, 0.0], [2.5, 1.0039216, 0.99999994, 16777216.0]], np.float32)
    s["rgba_in"] = vals
    s["rgba_out"] = np.stack([ref.to_rgba(v) for v in vals])
    np.savez_compressed(os.path.join(OUT, "shade.npz"), **s)

    for fn in sorted(os.listdir(OUT)):
        print(fn, os.path.getsize(os.path.join(OUT, fn)))


if __name__ == "__main__":
    main()
