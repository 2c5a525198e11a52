"""The run-time scene file (include/lyap/scene.h): the replacement for editing params.cu
(reference params.cu:21-114) or pressing keys in the viewer (lyap_interactive.cu:144-463).
CPU-only: parser, writer and the derived-field recomputation, checked against the golden
fixtures of the host-compiled reference."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import lyapunov3d_b200 as lp
from helpers import FRAME_NAMES, frame_inputs
from lyapunov3d_b200 import api
from lyapunov3d_b200.structs import Scene, clone, struct_bytes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_defaults_are_params_init_and_round_trip():
    sc = api.scene_defaults()
    prm, cam, lights, n, seq, (w, h) = lp.params_init()
    assert struct_bytes(sc.prm) == struct_bytes(prm) and struct_bytes(sc.cam) == struct_bytes(cam)
    assert struct_bytes(sc.lights) == struct_bytes(lights) and sc.num_lights == n
    assert sc.sequence.decode() == seq == "BCABA" and (sc.width, sc.height) == (w, h) == (3840, 2160)
    text = api.scene_format(sc)
    again = api.scene_parse(text)
    assert struct_bytes(again) == struct_bytes(sc)
    # the reference's literal defaults, to the digit (params.cu:23-36,63)
    for line in ("d = 2.0999999", "settle = 18", "accum = 1008", "depth = 4096", "jitter = 0.5", "opaqueThreshold = -0.75",
                 "cam.M = 0.449999988", "lights = 1", "light0.C = 6 5 3"):
        assert line in text, line


@pytest.mark.parametrize("name", FRAME_NAMES)
def test_scene_file_reproduces_the_golden_frames_inputs(golden, tmp_path, name):
    """Each golden frame's camera / params / lights (as the unmodified reference built them) written
    as a scene file and read back: every byte of the structs the kernels receive is reproduced,
    derived fields included."""
    cam, prm, lights, n_lights, seq_s, rgba, _ = frame_inputs(golden["frames"], name)
    h, w = rgba.shape[:2]
    sc = Scene()
    sc.prm, sc.cam, sc.num_lights, sc.width, sc.height, sc.sequence = clone(prm), clone(cam), n_lights, w, h, seq_s.encode()
    for k in range(16):
        sc.lights[k] = lights[k]
    path = tmp_path / (name + ".scene")
    api.scene_save(sc, path)
    back = api.scene_finalize(api.scene_load(path))
    assert struct_bytes(back.prm) == struct_bytes(prm)
    assert struct_bytes(back.cam) == struct_bytes(cam)
    for k in range(n_lights):
        assert struct_bytes(back.lights[k]) == struct_bytes(lights[k]), k
    assert back.sequence.decode() == seq_s and (back.width, back.height) == (w, h)


def test_partial_files_comments_and_orbit():
    sc = api.scene_parse("""
        # only what changes
        jitter = 0          # no stochastic sampling
        sequence = A6B6C6
        settle=72
        accum =4032
        cam.orbit_frame = 59 120
        lights = 2
        light1.C = 1, 2, 3
        light1.diffuseColor = 0.1 0.2 0.3 1
    """)
    assert sc.prm.jitter == 0.0 and sc.prm.settle == 72 and sc.prm.accum == 4032 and sc.sequence == b"A6B6C6"
    assert sc.num_lights == 2 and sc.lights[1].C.tuple() == (1.0, 2.0, 3.0)
    assert abs(sc.lights[1].diffuseColor.b - 0.3) < 1e-7
    assert sc.prm.depth == 4096.0                                   # untouched default
    want = clone(sc.cam)
    lp.campath_frame(59, 120, want)
    assert sc.cam.C.tuple() == want.C.tuple() and sc.cam.Q.tuple() == want.Q.tuple()
    a, b = api.scene_parse("cam.orbit = 1"), api.scene_defaults()
    assert a.cam.C.tuple() == b.cam.C.tuple()                       # the shipped params.cu is the i = 1 frame


@pytest.mark.parametrize("text,needle", [
    ("depth = 4096\nfoo = 1", "line 2: unknown key 'foo'"),
    ("jitter = abc", "bad value"),
    ("settle = -3", "bad value"),
    ("settle = 1.5", "bad value"),
    ("sequence = AXB", "bad value"),
    ("lights = 17", "bad value"),
    ("light16.C = 1 2 3", "unknown key"),
    ("cam.C = 1 2", "bad value"),
    ("cam.Q = 1 2 3 4 5", "bad value"),
    ("just words", "expected key = value"),
    ("width = 0", "bad value"),
])
def test_parse_errors_name_the_line(text, needle):
    with pytest.raises(lp.LyapError) as e:
        api.scene_parse(text)
    assert needle in str(e.value)


def test_missing_file_is_an_io_error(tmp_path):
    with pytest.raises(lp.LyapError):
        api.scene_load(tmp_path / "nope.scene")


def test_apps_reject_bad_scene_files(tmp_path):
    exe = os.path.join(ROOT, "lyapunov3d_b200", "bin", "lyap_render")
    bad = tmp_path / "bad.scene"
    bad.write_text("opaqueThreshold = -0.75\nbogus = 3\n")
    r = subprocess.run([exe, "-scene", str(bad)], capture_output=True, text=True)
    assert r.returncode == 2 and "line 2: unknown key 'bogus'" in r.stderr
    r = subprocess.run([exe, "-set", "depth=x"], capture_output=True, text=True)
    assert r.returncode == 2 and "bad value" in r.stderr
