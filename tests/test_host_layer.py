"""Host scene layer of the product library (params / scene / camera path / file
formats) against the reference's golden outputs, and the C-ABI surface itself.
CPU only: nothing here launches a kernel."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import lyapunov3d_b200 as lp
from helpers import from_raw
from lyapunov3d_b200 import api
from lyapunov3d_b200.structs import Cam, LightArray, Params, clone, struct_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    """Every function include/lyap/*.h declares must be exported by the built .so."""
    hdr = "".join(open(os.path.join(ROOT, "include", "lyap", h)).read() for h in ("abi.h", "scene.h"))
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)          # prose in comments mentions calls too
    names = sorted(set(re.findall(r"\b(lyap_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 40 and "lyap_scene_load" in names and "lyap_ray_probe" in names
    L = api.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert b"sm_100a" in L.lyap_version()


def test_params_init_matches_reference(golden):
    g = golden["scene"]
    prm, cam, lights, n, seq, size = lp.params_init()
    assert struct_bytes(prm) == g["prm"].tobytes()
    assert struct_bytes(cam) == g["cam"].tobytes()
    assert struct_bytes(lights) == g["lights"].tobytes()
    assert (n, seq, list(size)) == (int(g["n_lights"]), str(g["sequence"]), list(g["default_size"]))


def test_sequence_parser(golden):
    g = golden["scene"]
    for k, s in enumerate(g["seq_strings"]):
        assert lp.scene_convert_sequence(str(s)).tolist() == g[f"seq_{k}"].tolist(), s
    with pytest.raises(lp.LyapError):
        lp.scene_convert_sequence("ABX")          # reference exits; the library reports
    with pytest.raises(lp.LyapError):
        lp.scene_convert_sequence("")


def test_camera_lights_and_path(golden):
    g = golden["scene"]
    cam = from_raw(Cam, g["cam"])
    for k, (w, h, d) in enumerate(g["cam_sizes"]):
        c = clone(cam)
        lp.scene_cam_recalculate(c, int(w), int(h), int(d))
        assert struct_bytes(c) == g[f"cam_recalc_{k}"].tobytes(), (w, h, d)
    lights = from_raw(LightArray, g["lights"])
    lp.scene_lights_recalculate(lights, int(g["n_lights"]))
    assert struct_bytes(lights) == g["lights_recalc"].tobytes()
    for i, want in zip(g["campath_i"], g["campath_cams"]):
        c = clone(cam)
        lp.campath_orbit(float(i), c)
        assert struct_bytes(c) == want.tobytes(), i


def test_campath_frames_agree_with_oracle(oracle):
    """120-frame orbit (BASELINE config 5): frame f of n uses i = ease(f/(n-1)) rounded
    to 15 significant digits, as scale.pl's string interpolation would."""
    _, cam, *_ = lp.params_init()
    for f in (0, 1, 37, 59, 60, 118, 119):
        a, b = clone(cam), clone(cam)
        lp.campath_frame(f, 120, a)
        i = float("%.15g" % oracle.ease(f / 119.0))
        oracle.campath(i, b)
        assert struct_bytes(a) == struct_bytes(b), f
    last = clone(cam)
    lp.campath_frame(119, 120, last)
    assert struct_bytes(last) == struct_bytes(cam)   # the shipped params.cu is the final frame


def test_plan_picks_period_instantiations():
    conv = lp.scene_convert_sequence
    assert api.plan_period(conv("BCABA"), 18, 1008) == 5
    assert api.plan_period(conv("A6B6C6"), 72, 4032) == 21
    assert api.plan_period(conv("ABAB"), 18, 1008) == 2          # reduced to its true period
    assert api.plan_period(conv("A8B8"), 18, 1008) == 18
    assert api.plan_period(conv("A9B9C9D9"), 18, 1008) == 40     # 36 and 40 symbols still have a register table
    assert api.plan_period(conv("A9B9C9D9B3"), 18, 1008) == 0    # 44 symbols: generic path
    assert api.plan_period(conv("A8B7"), 18, 1008) == 17         # every period up to 32 has an instantiation
    assert api.plan_period(conv("A9A8B9B9"), 18, 1008) == 0      # 39 symbols: generic path
    assert api.plan_period(np.array([0, 5, -1], np.int32), 1, 1) == -1
    assert api.plan_period(np.array([-1], np.int32), 1, 1) == -1
    assert api.tile_count(1920, 1080, 8, 0, 1) == 240 * 135 * 64
    assert sum(api.tile_count(100, 50, 8, r, 3) for r in range(3)) == 13 * 7 * 64


def test_file_formats(tmp_path, golden):
    rgba = golden["frames"]["default_40x24_rgba"]
    h, w = rgba.shape[:2]
    ppm = tmp_path / "f.ppm"
    api.write_ppm(str(ppm), rgba)
    # what the reference's save_ppm loop prints (lyap_interactive.cu:595-606)
    want = "P3\n%d %d\n%d\n" % (w, h, 255) + "".join("%3d %3d %3d " % (p[0], p[1], p[2]) for p in rgba.reshape(-1, 4)) + "\n"
    assert ppm.read_text() == want
    png = tmp_path / "f.png"
    api.write_png(str(png), rgba)
    from PIL import Image
    assert np.array_equal(np.asarray(Image.open(png).convert("RGB")), rgba[..., :3])
    raw = tmp_path / "p.raw"
    pts = golden["frames"]["default_40x24_points"]
    api.write_raw(str(raw), pts)
    assert raw.read_bytes() == pts.tobytes()
    prm, cam, *_ = lp.params_init()
    name = api.format_filename("Render", 1526604923, 8192, 8192, "BCABA", cam, prm)
    assert name.startswith("Render_1526604923_8192x8192_BCABA_cx=3.9217899_cy=3.5029025_cz=3.5029025_")
    assert name.endswith("_step=2_D=2.1_i=18,1008_d=4096_j=0.5_r=32_ot=-0.75")


def test_no_cpu_fallback():
    """Without a GPU the device entry points must refuse, not compute something else."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    prm, cam, lights, n, s, _ = lp.params_init()
    with pytest.raises(lp.LyapError):
        lp.bake(prm, lp.scene_convert_sequence(s), 8)
    with pytest.raises(lp.LyapError):
        lp.render(cam, prm, lp.scene_convert_sequence(s), lights, n, 16, 16)
    with pytest.raises(lp.LyapError):
        lp.bake_host(prm, lp.scene_convert_sequence(s), 8)


def test_scene_helpers_match_oracle_on_random_inputs(oracle):
    """Product host layer vs the (reference-pinned) oracle on random cameras, lights and path
    positions: byte-identical structs, including non-unit quaternions (the reference's
    normalize() quirk) and tiny magnifications (the 1e-6 clamp)."""
    rng = np.random.default_rng(2018)
    _, cam, lights, _, _, _ = lp.params_init()
    for _ in range(200):
        a, b = clone(cam), clone(cam)
        q = rng.normal(size=4) * rng.choice([1.0, 1.0, 1.0, 0.3, 2.5])
        if rng.random() < 0.7:
            q = q / np.linalg.norm(q)
        mag = float(np.float32(rng.choice([0.45, 1.2, 1e-7, rng.uniform(0.01, 3.0)])))
        pos = [float(np.float32(v)) for v in rng.uniform(-2, 8, 3)]
        for c in (a, b):
            c.Q.x, c.Q.y, c.Q.z, c.Q.w = (float(np.float32(v)) for v in q)
            c.M = mag
            c.C.x, c.C.y, c.C.z = pos
        w, h, d = int(rng.integers(1, 4097)), int(rng.integers(1, 2161)), int(rng.integers(1, 4))
        lp.scene_cam_recalculate(a, w, h, d)
        oracle.cam_recalculate(b, w, h, d)
        assert struct_bytes(a) == struct_bytes(b)
    for _ in range(50):
        la, lb = clone(lights), clone(lights)
        for k in range(16):
            q = rng.normal(size=4)
            q /= np.linalg.norm(q)
            for L in (la, lb):
                L[k].Q.x, L[k].Q.y, L[k].Q.z, L[k].Q.w = (float(np.float32(v)) for v in q)
                L[k].M = float(np.float32(0.1 + 0.1 * k))
        lp.scene_lights_recalculate(la, 16)
        oracle.lights_recalculate(lb, 16)
        assert struct_bytes(la) == struct_bytes(lb)
    for i in np.concatenate([rng.uniform(0, 1, 300), [0.0, 1.0, 1e-7, 1 - 1e-7, 0.5]]):
        a, b = clone(cam), clone(cam)
        lp.campath_orbit(float(i), a)
        oracle.campath(float(i), b)
        assert struct_bytes(a) == struct_bytes(b), i
    for t in rng.uniform(0, 1, 100):
        assert api.ease_in_out_quart(float(t)) == oracle.ease(float(t))
