"""oracle -- TEST INFRASTRUCTURE ONLY.

ctypes front-ends for the CPU checkers:

* ``Oracle()``   -- oracle/liblyap_oracle.so, the C restatement (lyap_oracle.c)
* ``RefHost()``  -- oracle/_ref/libref_host.so, the unmodified reference sources
                    compiled as host C++ (ref_host_shim.cpp); optional
* ``RefCuda()``  -- oracle/_ref/libref_cuda.so, the unmodified reference kernel.cu
                    compiled with the reference's nvcc flags; needs a GPU; optional

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  Nothing under lyapunov3d_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from lyapunov3d_b200.structs import POINT_DTYPE, Cam, LightArray, Params

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liblyap_oracle.so")
REF_HOST_SO = os.path.join(HERE, "_ref", "libref_host.so")
REF_CUDA_SO = os.path.join(HERE, "_ref", "libref_cuda.so")


def build(verbose=False):
    """Compile the C restatement and, if /root/reference is present, oracle/_ref."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("oracle build failed")


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _seq_array(seq):
    if isinstance(seq, (bytes, str)):
        raise TypeError("pass a parsed sequence (int32 array ending in -1)")
    return np.ascontiguousarray(seq, dtype=np.int32)


class _HostChecker:
    """Shared surface of the C restatement and the host-compiled reference."""

    prefix = None

    def __init__(self, path):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        L, p = self.lib, self.prefix
        self._f = lambda name: getattr(L, p + name)
        self._f("convert_sequence").restype = C.c_size_t
        self._f("convert_sequence").argtypes = [C.c_char_p, C.c_void_p, C.c_size_t]
        self._f("lyap4d").restype = C.c_float
        self._f("lyap4d").argtypes = [C.c_float] * 4 + [C.c_uint32] * 2 + [C.c_void_p]
        self._f("lyap4d_many").argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        self._f("campath").argtypes = [C.c_double, C.c_void_p]
        self._f("cam_recalculate").argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]
        self._f("lights_recalculate").argtypes = [C.c_void_p, C.c_size_t]
        self._f("raymarch").restype = C.c_int

    # -- host scene ---------------------------------------------------------
    def convert_sequence(self, s):
        if isinstance(s, str):
            s = s.encode()
        buf = np.zeros(10 * len(s) + 2, np.int32)
        n = self._f("convert_sequence")(s, _p(buf), buf.size)
        return buf[:n].copy()

    def params_init(self):
        prm, cam, lights = Params(), Cam(), LightArray()
        n, w, h = C.c_uint32(), C.c_uint32(), C.c_uint32()
        sb = C.create_string_buffer(256)
        self._f("params_init")(C.byref(prm), C.byref(cam), C.byref(lights), C.byref(n), sb, C.c_size_t(256), C.byref(w), C.byref(h))
        return prm, cam, lights, n.value, sb.value.decode(), (w.value, h.value)

    def cam_recalculate(self, cam, tw, th, td):
        self._f("cam_recalculate")(C.byref(cam), tw, th, td)

    def lights_recalculate(self, lights, n):
        self._f("lights_recalculate")(C.byref(lights), n)

    def campath(self, i, cam):
        self._f("campath")(float(i), C.byref(cam))

    # -- exponent -------------------------------------------------------------
    def lyap4d(self, x, y, z, d, settle, accum, seq):
        seq = _seq_array(seq)
        return self._f("lyap4d")(x, y, z, d, settle, accum, _p(seq))

    def lyap4d_many(self, xyz, d, settle, accum, seq):
        seq = _seq_array(seq)
        xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
        out = np.empty(len(xyz), np.float32)
        self._f("lyap4d_many")(_p(xyz), len(xyz), d, settle, accum, _p(seq), _p(out))
        return out

    def shade(self, point, cam, lights, n):
        out = np.zeros(4, np.float32)
        pt = np.ascontiguousarray(point)
        self._f("shade")(_p(pt), C.byref(cam), C.byref(lights), C.c_uint32(n), _p(out))
        return out

    def to_rgba(self, rgba4):
        c = np.ascontiguousarray(rgba4, np.float32)
        out = np.zeros(4, np.uint8)
        self._f("to_rgba")(_p(c), _p(out))
        return out


class Oracle(_HostChecker):
    """The C restatement (kind == "port" in bench.py's cpu_baseline)."""

    prefix = "oracle_"

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build()
        super().__init__(ORACLE_SO)
        self.lib.oracle_render_rows.restype = C.c_uint64
        self.lib.oracle_render_pixels.restype = C.c_uint64
        self.lib.oracle_ease_in_out_quart.restype = C.c_double
        self.lib.oracle_ease_in_out_quart.argtypes = [C.c_double] * 4

    def threads(self):
        return self.lib.oracle_num_threads()

    def set_threads(self, n):
        self.lib.oracle_set_num_threads(int(n))

    def ease(self, t):
        return self.lib.oracle_ease_in_out_quart(t, 0.0, 1.0, 1.0)

    def raymarch(self, sx, sy, cam, prm, seq):
        seq = _seq_array(seq)
        pt = np.zeros(1, POINT_DTYPE)
        calls = C.c_uint64(0)
        ret = self.lib.oracle_raymarch(_p(pt), C.c_uint32(sx), C.c_uint32(sy), C.byref(cam), C.byref(prm), _p(seq), C.byref(calls))
        return ret, pt[0], calls.value

    def render(self, cam, prm, seq, lights, n_lights, w, h, y0=0, y1=None, points=None):
        """Returns (rgba[h,w,4] u8, points[h,w] POINT_DTYPE, n_calls). points zero-filled unless given."""
        seq = _seq_array(seq)
        y1 = h if y1 is None else y1
        rgba = np.zeros((h, w, 4), np.uint8)
        pts = np.zeros((h, w), POINT_DTYPE) if points is None else points
        calls = self.lib.oracle_render_rows(_p(rgba), _p(pts), C.byref(cam), C.byref(prm), _p(seq), C.byref(lights),
                                            C.c_uint32(n_lights), C.c_uint32(w), C.c_uint32(h), C.c_uint32(y0), C.c_uint32(y1))
        return rgba, pts, int(calls)

    def render_pixels(self, cam, prm, seq, lights, n_lights, w, h, pix):
        seq = _seq_array(seq)
        pix = np.ascontiguousarray(pix, np.uint32)
        rgba = np.zeros((h, w, 4), np.uint8)
        pts = np.zeros((h, w), POINT_DTYPE)
        calls = self.lib.oracle_render_pixels(_p(rgba), _p(pts), C.byref(cam), C.byref(prm), _p(seq), C.byref(lights),
                                              C.c_uint32(n_lights), C.c_uint32(w), C.c_uint32(h), _p(pix), C.c_size_t(pix.size))
        return rgba, pts, int(calls)

    def bake(self, prm, seq, nx, ny=None, nz=None, z0=0, z1=None):
        seq = _seq_array(seq)
        ny = nx if ny is None else ny
        nz = nx if nz is None else nz
        z1 = nz if z1 is None else z1
        vol = np.zeros((nz, ny, nx), np.float32)
        self.lib.oracle_bake_slab(_p(vol), C.byref(prm), _p(seq), C.c_uint32(nx), C.c_uint32(ny), C.c_uint32(nz), C.c_uint32(z0), C.c_uint32(z1))
        return vol


class RefHost(_HostChecker):
    """The unmodified reference, host-compiled (kind == "reference")."""

    prefix = "ref_"

    def __init__(self):
        super().__init__(REF_HOST_SO)

    @staticmethod
    def available():
        return os.path.exists(REF_HOST_SO)

    def set_threads(self, n):
        self.lib.ref_set_num_threads(int(n))

    def raymarch(self, sx, sy, cam, prm, seq):
        seq = _seq_array(seq)
        pt = np.zeros(1, POINT_DTYPE)
        ret = self.lib.ref_raymarch(_p(pt), C.c_uint32(sx), C.c_uint32(sy), C.byref(cam), C.byref(prm), _p(seq))
        return ret, pt[0], None

    def render(self, cam, prm, seq, lights, n_lights, w, h, y0=0, y1=None, points=None):
        seq = _seq_array(seq)
        y1 = h if y1 is None else y1
        rgba = np.zeros((h, w, 4), np.uint8)
        pts = np.zeros((h, w), POINT_DTYPE) if points is None else points
        self.lib.ref_render_rows(_p(rgba), _p(pts), C.byref(cam), C.byref(prm), _p(seq), C.byref(lights),
                                 C.c_uint32(n_lights), C.c_uint32(w), C.c_uint32(h), C.c_uint32(y0), C.c_uint32(y1))
        return rgba, pts, None

    def bake(self, prm, seq, nx, ny=None, nz=None, z0=0, z1=None):
        seq = _seq_array(seq)
        ny = nx if ny is None else ny
        nz = nx if nz is None else nz
        z1 = nz if z1 is None else z1
        vol = np.zeros((nz, ny, nx), np.float32)
        self.lib.ref_bake_slab(_p(vol), C.byref(prm), _p(seq), C.c_uint32(nx), C.c_uint32(ny), C.c_uint32(nz), C.c_uint32(z0), C.c_uint32(z1))
        return vol


class RefCuda:
    """The unmodified reference kernel.cu, nvcc --use_fast_math -arch=sm_100 (needs a GPU)."""

    def __init__(self):
        if not os.path.exists(REF_CUDA_SO):
            raise FileNotFoundError(REF_CUDA_SO)
        self.lib = C.CDLL(REF_CUDA_SO)

    @staticmethod
    def available():
        return os.path.exists(REF_CUDA_SO)

    def render(self, cam, prm, seq, lights, n_lights, w, h, reps=1):
        """Returns (rgba[h,w,4], points[h,w], best kernel ms)."""
        seq = _seq_array(seq)
        rgba = np.zeros((h, w, 4), np.uint8)
        pts = np.zeros((h, w), POINT_DTYPE)
        ms = C.c_float(0)
        rc = self.lib.ref_cuda_render(_p(rgba), _p(pts), C.byref(cam), C.byref(prm), _p(seq), C.c_size_t(seq.size),
                                      C.byref(lights), C.c_uint32(n_lights), C.c_uint32(w), C.c_uint32(h),
                                      C.c_int(reps), C.byref(ms))
        if rc != 0:
            raise RuntimeError(f"ref_cuda_render failed: cudaError {rc}")
        return rgba, pts, ms.value

    def bake(self, prm, seq, n, reps=1, want_output=True):
        seq = _seq_array(seq)
        vol = np.zeros((n, n, n), np.float32) if want_output else None
        ms = C.c_float(0)
        rc = self.lib.ref_cuda_bake(_p(vol) if want_output else None, C.byref(prm), _p(seq), C.c_size_t(seq.size),
                                    C.c_uint32(n), C.c_int(reps), C.byref(ms))
        if rc != 0:
            raise RuntimeError(f"ref_cuda_bake failed: {rc}")
        return vol, ms.value
