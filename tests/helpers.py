"""Shared helpers for the parity tests."""
import ctypes as C

import numpy as np

from lyapunov3d_b200.structs import POINT_DTYPE, Cam, LightArray, Params


def from_raw(cls, arr):
    """Rebuild a ctypes struct from the uint8 image stored in a fixture."""
    obj = cls()
    data = np.ascontiguousarray(arr, np.uint8).tobytes()
    assert len(data) == C.sizeof(obj), (len(data), C.sizeof(obj))
    C.memmove(C.byref(obj), data, len(data))
    return obj


def frame_inputs(frames, name):
    cam = from_raw(Cam, frames[name + "_cam"])
    prm = from_raw(Params, frames[name + "_prm"])
    lights = from_raw(LightArray, frames[name + "_lights"])
    n_lights = int(frames[name + "_n_lights"])
    seq_str = str(frames[name + "_seq"])
    rgba = frames[name + "_rgba"]
    pts = np.ascontiguousarray(frames[name + "_points"]).view(POINT_DTYPE)[..., 0]
    return cam, prm, lights, n_lights, seq_str, rgba, pts


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def same_floats(a, b):
    """Bit equality, with every NaN equal to every other NaN."""
    a = np.ascontiguousarray(a, np.float32)
    b = np.ascontiguousarray(b, np.float32)
    return bool(np.all((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))))


def frac_within(a, b, tol=2):
    """Fraction of pixels whose every channel differs by at most tol (of 255)."""
    d = np.abs(a.astype(np.int16) - b.astype(np.int16)).max(axis=-1)
    return float((d <= tol).mean())


FRAME_NAMES = ["default_48", "default_40x24", "nojitter_32", "method1_24", "long_24x16", "twolights_32", "wide_32"]
