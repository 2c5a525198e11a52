"""The C restatement against the live host-compiled reference (oracle/_ref).

Wider, seeded sweeps than the committed fixtures can hold.  Skipped where
oracle/_ref/libref_host.so is absent.  CPU only.
"""
import numpy as np

from helpers import same_floats
from lyapunov3d_b200.structs import clone, struct_bytes


def test_exponent_sweep(oracle, refhost):
    rng = np.random.default_rng(7)
    for s, d, settle, accum in [("BCABA", 2.1, 18, 1008), ("A6B6C6", 2.1, 72, 1000), ("ABCD", 3.3, 5, 257), ("B", 2.1, 1, 64)]:
        xyz = rng.uniform(-0.1, 4.1, (6000, 3)).astype(np.float32)
        seq = oracle.convert_sequence(s)
        assert (seq == refhost.convert_sequence(s)).all()
        assert same_floats(oracle.lyap4d_many(xyz, np.float32(d), settle, accum, seq),
                           refhost.lyap4d_many(xyz, np.float32(d), settle, accum, seq)), s


def test_frame_and_bake_live(oracle, refhost):
    prm, cam, lights, n, s, _ = oracle.params_init()
    seq = oracle.convert_sequence(s)
    oracle.lights_recalculate(lights, n)
    for i, (w, h) in [(0.3, (40, 30)), (0.77, (36, 36))]:
        c = clone(cam)
        oracle.campath(i, c)
        c2 = clone(cam)
        refhost.campath(i, c2)
        assert struct_bytes(c) == struct_bytes(c2)
        oracle.cam_recalculate(c, w, h, 1)
        a_rgba, a_pts, _ = oracle.render(c, prm, seq, lights, n, w, h)
        b_rgba, b_pts, _ = refhost.render(c, prm, seq, lights, n, w, h)
        assert a_pts.tobytes() == b_pts.tobytes() and np.array_equal(a_rgba, b_rgba)
    assert same_floats(oracle.bake(prm, seq, 24, 20, 10), refhost.bake(prm, seq, 24, 20, 10))
