"""The C restatement (oracle/lyap_oracle.c) against the reference's own outputs.

tests/golden/*.npz were produced by oracle/make_golden.py from the UNMODIFIED
reference sources compiled as host C++ (oracle/_ref/libref_host.so).  The
restatement must reproduce every one of them bit for bit; that is what pins the
oracle the GPU parity tests lean on.  CPU only.
"""
import numpy as np
import pytest

from helpers import FRAME_NAMES, frame_inputs, from_raw, same_floats
from lyapunov3d_b200.structs import Cam, LightArray, Params, clone, struct_bytes


def test_params_init_matches_reference(oracle, golden):
    g = golden["scene"]
    prm, cam, lights, n, seq, size = oracle.params_init()
    assert struct_bytes(prm) == g["prm"].tobytes()
    assert struct_bytes(cam) == g["cam"].tobytes()
    assert struct_bytes(lights) == g["lights"].tobytes()
    assert n == int(g["n_lights"]) and seq == str(g["sequence"]) and list(size) == list(g["default_size"])


def test_sequence_parser(oracle, golden):
    g = golden["scene"]
    for k, s in enumerate(g["seq_strings"]):
        assert oracle.convert_sequence(str(s)).tolist() == g[f"seq_{k}"].tolist(), s
    # SURVEY.md section 4 known answers: digits add n MORE copies (A6 -> 7 A's)
    assert oracle.convert_sequence("BCABA").tolist() == [1, 2, 0, 1, 0, -1]
    assert oracle.convert_sequence("A6B6C6").tolist() == [0] * 7 + [1] * 7 + [2] * 7 + [-1]
    assert oracle.convert_sequence("3A").tolist() == [1, 1, 1, 0, -1]  # leading digit repeats 'B'


def test_camera_and_lights_recalculate(oracle, golden):
    g = golden["scene"]
    cam = from_raw(Cam, g["cam"])
    for k, (w, h, d) in enumerate(g["cam_sizes"]):
        c = clone(cam)
        oracle.cam_recalculate(c, int(w), int(h), int(d))
        assert struct_bytes(c) == g[f"cam_recalc_{k}"].tobytes(), (w, h, d)
    lights = from_raw(LightArray, g["lights"])
    oracle.lights_recalculate(lights, int(g["n_lights"]))
    assert struct_bytes(lights) == g["lights_recalc"].tobytes()


def test_camera_path(oracle, golden):
    g = golden["scene"]
    cam = from_raw(Cam, g["cam"])
    for i, want in zip(g["campath_i"], g["campath_cams"]):
        c = clone(cam)
        oracle.campath(float(i), c)
        assert struct_bytes(c) == want.tobytes(), i
    # the shipped params.cu is the i == 1 frame of scale.pl
    c = clone(cam)
    oracle.campath(1.0, c)
    assert struct_bytes(c) == g["cam"].tobytes()
    assert oracle.ease(0.0) == 0.0 and oracle.ease(1.0) == 1.0 and abs(oracle.ease(0.5) - 0.5) < 1e-15


def test_known_answers_from_survey(oracle):
    """values obtained by host-compiling the reference verbatim."""
    seq = oracle.convert_sequence("BCABA")
    assert abs(oracle.lyap4d(3.3, 3.6, 3.1, 2.1, 18, 1008, seq) - 0.0696999356) < 1e-9
    prm, cam, lights, n, _, _ = oracle.params_init()
    oracle.cam_recalculate(cam, 256, 256, 1)
    np.testing.assert_allclose(cam.C.tuple(), (3.92179, 3.5029, 3.5029), rtol=2e-6)
    np.testing.assert_allclose(cam.Q.tuple(), (0.368691, -0.752009, 0, 0.546396), rtol=2e-6, atol=1e-7)
    ret, pt, calls = oracle.raymarch(128, 128, cam, prm, seq)
    assert ret == 0
    np.testing.assert_allclose(pt["P"], (3.58282, 3.33671, 3.33671), rtol=2e-6)
    np.testing.assert_allclose(pt["N"], (0.51584, -0.848829, -0.115757), rtol=2e-5)
    np.testing.assert_allclose([pt["a"], pt["c"], pt["l"]], (-2.93403, 35.6067, -0.747593), rtol=2e-6)


def test_exponent_vectors(oracle, golden):
    e = golden["exponent"]
    cases = [("xyz_default", "l_default", 2.1, 18, 1008, "BCABA"),
             ("xyz_long", "l_long", 2.1, 72, 4032, "A6B6C6"),
             ("xyz_long", "l_d_symbol", 3.7, 10, 500, "A6B6C6D6"),
             ("xyz_long", "l_odd_counts", 2.1, 7, 333, "AAB2"),
             ("xyz_long", "l_no_settle", 2.1, 0, 100, "AB")]
    for xk, lk, d, settle, accum, s in cases:
        got = oracle.lyap4d_many(e[xk], np.float32(d), settle, accum, oracle.convert_sequence(s))
        assert same_floats(got, e[lk]), lk
    # edge cases the reference defines: zero derivative -> NaN, v == 0.5 -> exactly 0
    assert np.isnan(e["l_default"][64:72]).all()
    assert (e["l_default"][72:80] == 0).all()


def test_bake_volumes(oracle, golden):
    b = golden["bake"]
    prm, *_ = oracle.params_init()
    seq = oracle.convert_sequence("BCABA")
    assert same_floats(oracle.bake(prm, seq, 32), b["default_32"])
    assert same_floats(oracle.bake(prm, seq, 20, 12, 9), b["ragged_20x12x9"])
    # a z-slab is the same numbers as the same planes of the whole volume
    slab = oracle.bake(prm, seq, 20, 12, 9, 3, 7)
    assert same_floats(slab[3:7], b["ragged_20x12x9"][3:7]) and (slab[:3] == 0).all() and (slab[7:] == 0).all()
    prm.settle, prm.accum = 72, 4032
    assert same_floats(oracle.bake(prm, oracle.convert_sequence("A6B6C6"), 16), b["long_16"])


@pytest.mark.parametrize("name", FRAME_NAMES)
def test_frames(oracle, golden, name):
    cam, prm, lights, n_lights, seq_s, rgba, pts = frame_inputs(golden["frames"], name)
    h, w = rgba.shape[:2]
    got_rgba, got_pts, calls = oracle.render(cam, prm, oracle.convert_sequence(seq_s), lights, n_lights, w, h)
    assert got_pts.tobytes() == pts.tobytes()
    assert np.array_equal(got_rgba, rgba)
    assert calls >= w * h  # every ray that enters the cube evaluates at least its entry point


def test_shade_and_pixel_conversion(oracle, golden):
    s, g = golden["shade"], golden["scene"]
    cam = from_raw(Cam, golden["frames"]["default_48_cam"])
    l1 = from_raw(LightArray, golden["frames"]["default_48_lights"])
    l2 = from_raw(LightArray, golden["frames"]["twolights_32_lights"])
    for p, want1, want2 in zip(s["points"], s["colors_default"], s["colors_twolights"]):
        assert same_floats(oracle.shade(p, cam, l1, 1), want1)
        assert same_floats(oracle.shade(p, cam, l2, 2), want2)
    for v, want in zip(s["rgba_in"], s["rgba_out"]):
        assert oracle.to_rgba(v).tolist() == want.tolist()
    # no clamp: 1.004 * 255 = 256.02 wraps to 0 (reference color.hpp:169-175)
    assert oracle.to_rgba([1.004, 0.1, 0, 0]).tolist() == [0, 25, 0, 0]
