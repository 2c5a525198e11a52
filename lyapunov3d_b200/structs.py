"""ctypes / numpy mirrors of include/lyap/types.h.

Layout-identical to the reference's PODs (reference structs.hpp:26-80; offsets
measured in SURVEY.md appendix A), so a buffer filled by the reference can be
viewed with these types and vice versa.
"""
import ctypes as C

import numpy as np


class Vec3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    def tuple(self):
        return (self.x, self.y, self.z)


class Quat(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]

    def tuple(self):
        return (self.x, self.y, self.z, self.w)


class Color(C.Structure):
    _fields_ = [("r", C.c_float), ("g", C.c_float), ("b", C.c_float), ("a", C.c_float)]

    def tuple(self):
        return (self.r, self.g, self.b, self.a)


class CamLight(C.Structure):
    """LyapCam == LyapLight (reference structs.hpp:26-51); 224 bytes, 16-aligned."""

    _fields_ = [
        ("C", Vec3),
        ("_pad0", C.c_uint32),
        ("Q", Quat),
        ("M", C.c_float),
        ("V", Vec3),
        ("S0", Vec3),
        ("SDX", Vec3),
        ("SDY", Vec3),
        ("textureWidth", C.c_uint32),
        ("textureHeight", C.c_uint32),
        ("renderWidth", C.c_uint32),
        ("renderHeight", C.c_uint32),
        ("renderDenominator", C.c_uint32),
        ("lightInnerCone", C.c_float),
        ("lightOuterCone", C.c_float),
        ("lightRange", C.c_float),
        ("_pad1", C.c_uint32 * 3),
        ("ambient", Color),
        ("diffuseColor", Color),
        ("diffusePower", C.c_float),
        ("_pad2", C.c_uint32 * 3),
        ("specularColor", Color),
        ("specularPower", C.c_float),
        ("specularHardness", C.c_float),
        ("_pad3", C.c_uint32 * 2),
        ("chaosColor", Color),
    ]


Cam = CamLight
Light = CamLight
MAX_LIGHTS = 16
LightArray = CamLight * MAX_LIGHTS


class Params(C.Structure):
    """LyapParams (reference structs.hpp:53-68); 56 bytes."""

    _fields_ = [
        ("d", C.c_float),
        ("settle", C.c_uint32),
        ("accum", C.c_uint32),
        ("stepMethod", C.c_uint32),
        ("nearThreshold", C.c_float),
        ("nearMultiplier", C.c_float),
        ("opaqueThreshold", C.c_float),
        ("chaosThreshold", C.c_float),
        ("depth", C.c_float),
        ("jitter", C.c_float),
        ("refine", C.c_float),
        ("gradient", C.c_float),
        ("lMin", C.c_float),
        ("lMax", C.c_float),
    ]


class Scene(C.Structure):
    """lyap_scene (include/lyap/scene.h): the run-time parameter surface."""

    _fields_ = [("prm", Params), ("_pad0", C.c_uint32 * 2), ("cam", CamLight), ("lights", LightArray), ("num_lights", C.c_uint32),
                ("width", C.c_uint32), ("height", C.c_uint32), ("sequence", C.c_char * 1024)]


class Point(C.Structure):
    """LyapPoint (reference structs.hpp:70-76); 36 bytes."""

    _fields_ = [("P", Vec3), ("N", Vec3), ("a", C.c_float), ("c", C.c_float), ("l", C.c_float)]


class RGBA(C.Structure):
    _fields_ = [("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8), ("a", C.c_uint8)]


assert C.sizeof(Vec3) == 12 and C.sizeof(Quat) == 16 and C.sizeof(Color) == 16
assert C.sizeof(CamLight) == 224, C.sizeof(CamLight)
assert CamLight.Q.offset == 16 and CamLight.M.offset == 32 and CamLight.V.offset == 36
assert CamLight.S0.offset == 48 and CamLight.SDX.offset == 60 and CamLight.SDY.offset == 72
assert CamLight.textureWidth.offset == 84 and CamLight.renderDenominator.offset == 100
assert CamLight.lightInnerCone.offset == 104 and CamLight.lightRange.offset == 112
assert CamLight.ambient.offset == 128 and CamLight.diffuseColor.offset == 144
assert CamLight.diffusePower.offset == 160 and CamLight.specularColor.offset == 176
assert CamLight.specularPower.offset == 192 and CamLight.specularHardness.offset == 196
assert CamLight.chaosColor.offset == 208
assert C.sizeof(Params) == 56 and C.sizeof(Point) == 36 and C.sizeof(RGBA) == 4
assert Scene.cam.offset == 64 and Scene.lights.offset == 288 and Scene.num_lights.offset == 288 + 16 * 224

# numpy views of the two bulk outputs
POINT_DTYPE = np.dtype(
    [("P", np.float32, 3), ("N", np.float32, 3), ("a", np.float32), ("c", np.float32), ("l", np.float32)]
)
assert POINT_DTYPE.itemsize == 36


def struct_bytes(s) -> bytes:
    return bytes(memoryview(s).cast("B"))


def clone(s):
    out = type(s)()
    C.memmove(C.byref(out), C.byref(s), C.sizeof(s))
    return out
