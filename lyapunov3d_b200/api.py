"""Python face of liblyap_b200.so (ctypes over the C ABI in include/lyap/abi.h).

The function names mirror the reference's host interface (params_init,
scene_convert_sequence, scene_cam_recalculate, scene_lights_recalculate,
kernel_calc_render -> render, kernel_calc_volume -> bake).  torch is used only for
device memory and streams.  There is no CPU fallback: if the shared library is
missing or has no CUDA device to run on, calls raise.
"""
import ctypes as C
import os

import numpy as np

from .structs import MAX_LIGHTS, POINT_DTYPE, Cam, LightArray, Params, Scene

MODE_EXACT, MODE_FAST, MODE_HOST, MODE_HYBRID, MODE_HYBRID_HOST = 0, 1, 2, 3, 4
MODES = {"exact": MODE_EXACT, "fast": MODE_FAST, "host": MODE_HOST, "hybrid": MODE_HYBRID, "hybrid_host": MODE_HYBRID_HOST}
F32, F16 = 0, 1

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblyap_b200.so")
_lib = None


class LyapError(RuntimeError):
    pass


def lib():
    """Load the CUDA extension; never falls back to anything else."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise LyapError(f"{_LIB_PATH} is missing: build it with `python -m lyapunov3d_b200._build` "
                        "(there is no CPU fallback for the hot path)")
    L = C.CDLL(_LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.lyap_version.restype = C.c_char_p
    L.lyap_error_string.restype = C.c_char_p
    L.lyap_error_string.argtypes = [i32]
    L.lyap_set_option.argtypes = [C.c_char_p, C.c_long]
    L.lyap_plan_period.argtypes = [vp, u32, u32]
    L.lyap_plan_describe.argtypes = [vp, u32, u32, vp, vp, vp, vp]
    L.lyap_params_init.argtypes = [vp, vp, vp, vp, C.c_char_p, C.c_size_t, vp, vp]
    L.lyap_params_init.restype = None
    L.lyap_scene_convert_sequence.restype = C.c_size_t
    L.lyap_scene_convert_sequence.argtypes = [C.POINTER(C.POINTER(C.c_int32)), C.c_char_p]
    L.lyap_scene_cam_recalculate.argtypes = [vp, u32, u32, u32]
    L.lyap_scene_cam_recalculate.restype = None
    L.lyap_scene_lights_recalculate.argtypes = [vp, C.c_size_t]
    L.lyap_scene_lights_recalculate.restype = None
    L.lyap_ease_in_out_quart.restype = C.c_double
    L.lyap_ease_in_out_quart.argtypes = [C.c_double] * 4
    L.lyap_campath_orbit.argtypes = [C.c_double, vp]
    L.lyap_campath_orbit.restype = None
    L.lyap_campath_frame.argtypes = [u32, u32, vp]
    L.lyap_campath_frame.restype = None
    L.lyap_render.argtypes = [vp, vp, vp, vp, vp, vp, u32, u32, u32, i32, vp, vp]
    L.lyap_render_tiles.argtypes = [vp, vp, vp, vp, vp, vp, u32, u32, u32, u32, u32, u32, i32, i32, vp, vp]
    L.lyap_tile_count.restype = u64
    L.lyap_tile_count.argtypes = [u32, u32, u32, u32, u32]
    L.lyap_scatter_tiles.argtypes = [vp, vp, u32, u32, u32, u32, u32, u32, vp]
    L.lyap_assist_bits_bytes.restype = u64
    L.lyap_assist_bits_bytes.argtypes = [u32]
    L.lyap_assist_build.argtypes = [vp, vp, i32, u32, vp, C.c_float, C.c_float, u32, vp]
    L.lyap_render_assisted.argtypes = [vp, vp, vp, vp, vp, vp, u32, u32, u32, u32, u32, u32, i32, i32, vp, i32, u32, vp, vp, vp, vp]
    L.lyap_shade_points.argtypes = [vp, vp, vp, vp, u32, u64, i32, vp]
    L.lyap_bake.argtypes = [vp, i32, vp, vp, u32, u32, u32, u32, u32, i32, vp]
    L.lyap_exponent_points.argtypes = [vp, vp, u64, vp, vp, i32, vp]
    L.lyap_ray_probe.argtypes = [vp, vp, u64, vp, vp, i32, vp]
    L.lyap_normalize_vectors.argtypes = [vp, u64, i32, vp]
    L.lyap_render_host.argtypes = [vp, vp, vp, vp, vp, vp, u32, u32, u32, i32, i32, vp]
    L.lyap_bake_host.argtypes = [vp, i32, vp, vp, u32, u32, u32, u32, u32, i32, i32]
    L.lyap_write_ppm.argtypes = [C.c_char_p, vp, u32, u32]
    L.lyap_write_png.argtypes = [C.c_char_p, vp, u32, u32]
    L.lyap_write_raw.argtypes = [C.c_char_p, vp, u64]
    L.lyap_format_filename.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_ulong, u32, u32, C.c_char_p, vp, vp]
    L.lyap_probe_peaks.argtypes = [vp, vp, vp, vp]
    L.lyap_scene_defaults.argtypes = [vp]
    L.lyap_scene_defaults.restype = None
    L.lyap_scene_parse.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_size_t]
    L.lyap_scene_load.argtypes = [vp, C.c_char_p, C.c_char_p, C.c_size_t]
    L.lyap_scene_finalize.argtypes = [vp, u32, u32]
    L.lyap_scene_finalize.restype = None
    L.lyap_scene_format.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.lyap_scene_format.restype = C.c_size_t
    L.lyap_scene_save.argtypes = [vp, C.c_char_p]
    L.lyap_host_workspace_release.restype = None
    L.lyap_peer_alloc.argtypes = [C.POINTER(vp), u64]
    L.lyap_peer_free.argtypes = [vp]
    L.lyap_peer_export.argtypes = [vp, vp]
    L.lyap_peer_open.argtypes = [vp, C.POINTER(vp)]
    L.lyap_peer_close.argtypes = [vp]
    L.lyap_probe_ffma2.argtypes = [vp]
    _lib = L
    return L


def _check(code, what):
    if code != 0:
        raise LyapError(f"{what} failed: [{code}] {lib().lyap_error_string(code).decode()}")


def _mode(mode):
    return MODES[mode] if isinstance(mode, str) else int(mode)


def set_option(key, value):
    _check(lib().lyap_set_option(key.encode(), int(value)), f"set_option({key})")


# ------------------------------------------------------------------ host scene
def params_init():
    """reference params.cu:21-114 -> (prm, cam, lights[16], num_lights, sequence, (w, h))."""
    prm, cam, lights = Params(), Cam(), LightArray()
    n, w, h = C.c_uint32(), C.c_uint32(), C.c_uint32()
    sb = C.create_string_buffer(64)
    lib().lyap_params_init(C.byref(prm), C.byref(cam), C.byref(lights), C.byref(n), sb, 64, C.byref(w), C.byref(h))
    return prm, cam, lights, n.value, sb.value.decode(), (w.value, h.value)


def scene_convert_sequence(s):
    """reference scene.cu:69-108 -> int32 array ending in -1."""
    if isinstance(s, str):
        s = s.encode()
    ptr = C.POINTER(C.c_int32)()
    n = lib().lyap_scene_convert_sequence(C.byref(ptr), s)
    if n == 0:
        raise LyapError(f"bad sequence string {s!r}")
    out = np.ctypeslib.as_array(ptr, shape=(n,)).copy()
    C.CDLL(None).free(ptr)
    return out


def scene_cam_recalculate(cam, tw, th, td=1):
    lib().lyap_scene_cam_recalculate(C.byref(cam), tw, th, td)


def scene_lights_recalculate(lights, n):
    lib().lyap_scene_lights_recalculate(C.byref(lights), n)


def ease_in_out_quart(t, b=0.0, c=1.0, d=1.0):
    return lib().lyap_ease_in_out_quart(t, b, c, d)


def campath_orbit(i, cam):
    lib().lyap_campath_orbit(float(i), C.byref(cam))


def campath_frame(f, n_frames, cam):
    lib().lyap_campath_frame(f, n_frames, C.byref(cam))


# ------------------------------------------------------------------ scene files
def scene_defaults():
    """params_init() as one lyap_scene (include/lyap/scene.h)."""
    sc = Scene()
    lib().lyap_scene_defaults(C.byref(sc))
    return sc


def scene_parse(text, scene=None):
    """Apply `key = value` lines on top of `scene` (default: the params_init() scene)."""
    sc = scene if scene is not None else scene_defaults()
    err = C.create_string_buffer(256)
    rc = lib().lyap_scene_parse(C.byref(sc), text.encode(), err, 256)
    if rc != 0:
        raise LyapError(f"scene: {err.value.decode()}")
    return sc


def scene_load(path):
    sc = Scene()
    err = C.create_string_buffer(256)
    rc = lib().lyap_scene_load(C.byref(sc), str(path).encode(), err, 256)
    if rc != 0:
        raise LyapError(f"scene {path}: {err.value.decode() or lib().lyap_error_string(rc).decode()}")
    return sc


def scene_finalize(scene, width=0, height=0):
    """Recompute the derived camera / light fields (the reference's update_scene())."""
    lib().lyap_scene_finalize(C.byref(scene), width, height)
    return scene


def scene_format(scene):
    n = lib().lyap_scene_format(C.byref(scene), None, 0)
    buf = C.create_string_buffer(n + 1)
    lib().lyap_scene_format(C.byref(scene), buf, n + 1)
    return buf.value.decode()


def scene_save(scene, path):
    _check(lib().lyap_scene_save(C.byref(scene), str(path).encode()), "lyap_scene_save")


def plan_period(seq, settle, accum):
    seq = np.ascontiguousarray(seq, np.int32)
    return lib().lyap_plan_period(seq.ctypes.data, settle, accum)


def plan_describe(seq, settle, accum):
    """The iteration schedule (SeqPlan) the library builds for a sequence; for tests."""
    seq = np.ascontiguousarray(seq, np.int32)
    hdr = np.zeros(11, np.uint32)
    sym, rot, runs = np.zeros(1024, np.uint8), np.zeros(64, np.uint8), np.zeros(2048, np.uint8)
    _check(lib().lyap_plan_describe(seq.ctypes.data, settle, accum, hdr.ctypes.data, sym.ctypes.data, rot.ctypes.data,
                                    runs.ctypes.data), "lyap_plan_describe")
    P, ln, sh, sp_, ap, at = (int(v) for v in hdr[:6])
    return {"P": P, "len": ln, "settle_head": sh, "settle_periods": sp_, "accum_periods": ap, "accum_tail": at,
            "cnt": [int(v) for v in hdr[6:10]], "sym": sym[:ln].tolist(), "rot": rot[:min(ln, 40)].tolist(),
            "runs": runs[:2 * int(hdr[10])].reshape(-1, 2).tolist()}


def tile_count(width, height, tile, rank, world):
    return int(lib().lyap_tile_count(width, height, tile, rank, world))


# ------------------------------------------------------------------ device path
def _torch():
    import torch

    if not torch.cuda.is_available():
        raise LyapError("no CUDA device: the Lyapunov hot path has no CPU fallback")
    return torch


def _stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def upload_lights(lights, device=None):
    torch = _torch()
    host = torch.frombuffer(bytearray(bytes(memoryview(lights).cast("B"))), dtype=torch.uint8)
    return host.to(device if device is not None else torch.device("cuda", torch.cuda.current_device()))


def render(cam, prm, seq, lights, num_lights, width, height, mode="exact", points=None, rgba=None,
           tile=None, rank=0, world=1, compact=False, count_evals=True):
    """kernel_calc_render replacement.  Returns (rgba u8[h,w,4] | [n,4], points u8[..,36], evals).

    `lights` is either a LightArray (uploaded here) or a uint8 CUDA tensor holding it.
    Tensors live on the current CUDA device; work is queued on the current stream.
    """
    torch = _torch()
    dev = torch.device("cuda", torch.cuda.current_device())
    seq = np.ascontiguousarray(seq, np.int32)
    d_lights = lights if hasattr(lights, "data_ptr") else upload_lights(lights, dev)
    if tile is None:
        n = width * height
        shape = (height, width)
    else:
        n = tile_count(width, height, tile, rank, world) if compact else width * height
        shape = (n,) if compact else (height, width)
    if points is None:
        points = torch.zeros(shape + (36,), dtype=torch.uint8, device=dev)
    if rgba is None:
        rgba = torch.zeros(shape + (4,), dtype=torch.uint8, device=dev)
    evals = torch.zeros(1, dtype=torch.int64, device=dev) if count_evals else None
    ev_ptr = C.c_void_p(evals.data_ptr()) if evals is not None else None
    L = lib()
    if tile is None:
        rc = L.lyap_render(rgba.data_ptr(), points.data_ptr(), C.byref(cam), C.byref(prm), seq.ctypes.data,
                           d_lights.data_ptr(), num_lights, width, height, _mode(mode), ev_ptr, _stream_ptr(torch))
    else:
        rc = L.lyap_render_tiles(rgba.data_ptr(), points.data_ptr(), C.byref(cam), C.byref(prm), seq.ctypes.data,
                                 d_lights.data_ptr(), num_lights, width, height, tile, rank, world, int(compact),
                                 _mode(mode), ev_ptr, _stream_ptr(torch))
    _check(rc, "lyap_render")
    return rgba, points, evals


def assist_build(volume, prm, margin=0.25, upper=float("inf"), dilate=1):
    """Classify the cells of a baked cubic volume (float32 or float16 CUDA tensor [n,n,n]) for the
    volume-assisted march.  Returns the bit mask (int32 CUDA tensor)."""
    torch = _torch()
    n = volume.shape[0]
    assert volume.is_contiguous() and tuple(volume.shape) == (n, n, n)
    bits = torch.zeros(lib().lyap_assist_bits_bytes(n) // 4, dtype=torch.int32, device=volume.device)
    _check(lib().lyap_assist_build(bits.data_ptr(), volume.data_ptr(), F16 if volume.dtype == torch.float16 else F32, n, C.byref(prm),
                                   float(margin), float(upper), int(dilate), _stream_ptr(torch)), "lyap_assist_build")
    return bits


def render_assisted(cam, prm, seq, lights, num_lights, width, height, volume, safe_bits, mode="hybrid", points=None, rgba=None,
                    tile=8, rank=0, world=1, compact=False):
    """lyap_render_assisted: the hybrid render with march samples in safe cells stepped over.
    Returns (rgba, points, evals, skipped)."""
    torch = _torch()
    dev = torch.device("cuda", torch.cuda.current_device())
    seq = np.ascontiguousarray(seq, np.int32)
    d_lights = lights if hasattr(lights, "data_ptr") else upload_lights(lights, dev)
    n_out = tile_count(width, height, tile, rank, world) if compact else width * height
    shape = (n_out,) if compact else (height, width)
    if points is None:
        points = torch.zeros(shape + (36,), dtype=torch.uint8, device=dev)
    if rgba is None:
        rgba = torch.zeros(shape + (4,), dtype=torch.uint8, device=dev)
    evals = torch.zeros(1, dtype=torch.int64, device=dev)
    skipped = torch.zeros(1, dtype=torch.int64, device=dev)
    _check(lib().lyap_render_assisted(rgba.data_ptr(), points.data_ptr(), C.byref(cam), C.byref(prm), seq.ctypes.data, d_lights.data_ptr(),
                                      num_lights, width, height, tile, rank, world, int(compact), _mode(mode), volume.data_ptr(),
                                      F16 if volume.dtype == torch.float16 else F32, volume.shape[0], safe_bits.data_ptr(),
                                      evals.data_ptr(), skipped.data_ptr(), _stream_ptr(torch)), "lyap_render_assisted")
    return rgba, points, evals, skipped


def scatter_tiles(image, compact, width, height, tile, rank, world):
    torch = _torch()
    elem = compact.shape[-1] * compact.element_size()
    _check(lib().lyap_scatter_tiles(image.data_ptr(), compact.data_ptr(), elem, width, height, tile, rank, world,
                                    _stream_ptr(torch)), "lyap_scatter_tiles")


def shade_points(points, cam, lights, num_lights, mode="exact"):
    torch = _torch()
    d_lights = lights if hasattr(lights, "data_ptr") else upload_lights(lights, points.device)
    count = points.numel() // 36
    rgba = torch.zeros(points.shape[:-1] + (4,), dtype=torch.uint8, device=points.device)
    _check(lib().lyap_shade_points(rgba.data_ptr(), points.data_ptr(), C.byref(cam), d_lights.data_ptr(), num_lights,
                                   count, _mode(mode), _stream_ptr(torch)), "lyap_shade_points")
    return rgba


def bake(prm, seq, nx, ny=None, nz=None, z0=0, z1=None, mode="fast", dtype="f32", out=None):
    """kernel_calc_volume replacement: fills planes z0..z1 of the full [nz,ny,nx] volume."""
    torch = _torch()
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    z1 = nz if z1 is None else z1
    seq = np.ascontiguousarray(seq, np.int32)
    tdt = torch.float16 if dtype in ("f16", F16) else torch.float32
    if out is None:
        out = torch.zeros((nz, ny, nx), dtype=tdt, device=torch.device("cuda", torch.cuda.current_device()))
    _check(lib().lyap_bake(out.data_ptr(), F16 if tdt == torch.float16 else F32, C.byref(prm), seq.ctypes.data,
                           nx, ny, nz, z0, z1, _mode(mode), _stream_ptr(torch)), "lyap_bake")
    return out


def bake_into(slab, slab_z0, prm, seq, nx, ny, nz, z0, z1, mode="fast", dtype="f32"):
    """Bake planes z0..z1 of an [nz,ny,nx] volume into `slab`, a tensor whose plane 0 is plane
    `slab_z0` of the volume (a rank that owns one z-slab need not allocate the whole volume)."""
    torch = _torch()
    seq = np.ascontiguousarray(seq, np.int32)
    f16 = slab.dtype == torch.float16
    assert slab.is_contiguous() and slab.shape[1:] == (ny, nx) and slab_z0 <= z0 and z1 - slab_z0 <= slab.shape[0]
    base = slab.data_ptr() - slab_z0 * ny * nx * slab.element_size()
    _check(lib().lyap_bake(base, F16 if f16 else F32, C.byref(prm), seq.ctypes.data, nx, ny, nz, z0, z1, _mode(mode),
                           _stream_ptr(torch)), "lyap_bake")
    return slab


def exponent_points(xyz, prm, seq, mode="exact"):
    torch = _torch()
    seq = np.ascontiguousarray(seq, np.int32)
    xyz = xyz.contiguous().float()
    out = torch.empty(xyz.shape[0], dtype=torch.float32, device=xyz.device)
    _check(lib().lyap_exponent_points(out.data_ptr(), xyz.data_ptr(), xyz.shape[0], C.byref(prm), seq.ctypes.data,
                                      _mode(mode), _stream_ptr(torch)), "lyap_exponent_points")
    return out


def ray_probe(pixels, cam, prm, mode="exact"):
    """Ray set-up in the mode's own arithmetic: pixels int[n,2] (x, y) -> float32[n,12]
    {hit, V(3), t0, t1, Fdt, Ndt, P0(3), Fdt*gradient}."""
    torch = _torch()
    px = torch.as_tensor(np.ascontiguousarray(pixels, np.uint32).view(np.int32)).cuda()
    out = torch.empty((px.shape[0], 12), dtype=torch.float32, device=px.device)
    _check(lib().lyap_ray_probe(out.data_ptr(), px.data_ptr(), px.shape[0], C.byref(cam), C.byref(prm), _mode(mode),
                                _stream_ptr(torch)), "lyap_ray_probe")
    return out


def normalize_vectors(xyz, mode="exact"):
    """Vec::normalize in the mode's arithmetic, on a float32 CUDA tensor [n,3] (returns a new tensor)."""
    torch = _torch()
    out = xyz.contiguous().float().clone()
    _check(lib().lyap_normalize_vectors(out.data_ptr(), out.shape[0], _mode(mode), _stream_ptr(torch)), "lyap_normalize_vectors")
    return out


# ---------------------------------------------------------------- raw pointers
class DevicePointer:
    """A device pointer that is not owned by torch (a peer-mapped buffer, say), exposed through
    __cuda_array_interface__ so that torch.as_tensor() can view it and, duck-typed with
    data_ptr(), so that render()/bake() accept it as an output buffer."""

    def __init__(self, ptr, shape, typestr):
        self.ptr, self.shape, self.typestr = int(ptr), tuple(shape), typestr
        self.__cuda_array_interface__ = {"shape": self.shape, "typestr": typestr, "data": (self.ptr, False), "version": 2}

    def data_ptr(self):
        return self.ptr

    def tensor(self):
        return _torch().as_tensor(self, device="cuda")


def render_into(d_rgba, d_points, cam, prm, seq, d_lights, num_lights, width, height, mode="exact", tile=8, rank=0, world=1,
                evals=None):
    """lyap_render_tiles on raw device pointers (ints): in-place (non-compact) tile render."""
    torch = _torch()
    seq = np.ascontiguousarray(seq, np.int32)
    ev_ptr = C.c_void_p(evals.data_ptr()) if evals is not None else None
    _check(lib().lyap_render_tiles(int(d_rgba), int(d_points), C.byref(cam), C.byref(prm), seq.ctypes.data, d_lights.data_ptr(),
                                   num_lights, width, height, tile, rank, world, 0, _mode(mode), ev_ptr, _stream_ptr(torch)),
           "lyap_render_tiles")


def bake_ptr(d_exps, f16, prm, seq, nx, ny, nz, z0, z1, mode="fast"):
    """lyap_bake on a raw device pointer to the FULL volume."""
    torch = _torch()
    seq = np.ascontiguousarray(seq, np.int32)
    _check(lib().lyap_bake(int(d_exps), F16 if f16 else F32, C.byref(prm), seq.ctypes.data, nx, ny, nz, z0, z1, _mode(mode),
                           _stream_ptr(torch)), "lyap_bake")


# ------------------------------------------------------------- host-buffer path
def render_host(cam, prm, seq, lights, num_lights, width, height, mode="exact", device=0, want_points=True,
                rgba=None, points=None):
    """Whole frame with HOST buffers: allocs, H2D, kernel, D2H and sync inside the call."""
    seq = np.ascontiguousarray(seq, np.int32)
    if rgba is None:
        rgba = np.zeros((height, width, 4), np.uint8)
    if points is None and want_points:
        points = np.zeros((height, width), POINT_DTYPE)
    ev = C.c_ulonglong(0)
    rc = lib().lyap_render_host(rgba.ctypes.data, points.ctypes.data if points is not None else None, C.byref(cam),
                                C.byref(prm), seq.ctypes.data, C.byref(lights), num_lights, width, height,
                                _mode(mode), device, C.byref(ev))
    _check(rc, "lyap_render_host")
    return rgba, points, ev.value


def bake_host(prm, seq, nx, ny=None, nz=None, z0=0, z1=None, mode="fast", dtype="f32", device=0, out=None):
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    z1 = nz if z1 is None else z1
    seq = np.ascontiguousarray(seq, np.int32)
    f16 = dtype in ("f16", F16)
    if out is None:
        out = np.zeros((nz, ny, nx), np.float16 if f16 else np.float32)
    _check(lib().lyap_bake_host(out.ctypes.data, F16 if f16 else F32, C.byref(prm), seq.ctypes.data, nx, ny, nz, z0, z1,
                                _mode(mode), device), "lyap_bake_host")
    return out


# ------------------------------------------------------------------------ output
def write_ppm(path, rgba):
    rgba = np.ascontiguousarray(rgba, np.uint8)
    _check(lib().lyap_write_ppm(path.encode(), rgba.ctypes.data, rgba.shape[1], rgba.shape[0]), "lyap_write_ppm")


def write_png(path, rgba):
    rgba = np.ascontiguousarray(rgba, np.uint8)
    _check(lib().lyap_write_png(path.encode(), rgba.ctypes.data, rgba.shape[1], rgba.shape[0]), "lyap_write_png")


def write_raw(path, arr):
    arr = np.ascontiguousarray(arr)
    _check(lib().lyap_write_raw(path.encode(), arr.ctypes.data, arr.nbytes), "lyap_write_raw")


def format_filename(prefix, timestamp, width, height, sequence, cam, prm):
    buf = C.create_string_buffer(512)
    _check(lib().lyap_format_filename(buf, 512, prefix.encode(), timestamp, width, height, sequence.encode(),
                                      C.byref(cam), C.byref(prm)), "lyap_format_filename")
    return buf.value.decode()


def probe_peaks():
    """Register-only FFMA and MUFU.LG2 loops -> dict of measured lane-ops/s and SM clock."""
    _torch()
    f, m, c, n = C.c_double(), C.c_double(), C.c_double(), C.c_int()
    _check(lib().lyap_probe_peaks(C.byref(f), C.byref(m), C.byref(c), C.byref(n)), "lyap_probe_peaks")
    f2 = C.c_double()
    _check(lib().lyap_probe_ffma2(C.byref(f2)), "lyap_probe_ffma2")
    return {"ffma2_lane_ops_per_s": f2.value, "ffma_lane_ops_per_s": f.value, "mufu_lane_ops_per_s": m.value, "sm_clock_hz": c.value, "sm_count": n.value}
