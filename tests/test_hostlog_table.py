"""The glibc logf restatement used by LYAP_MODE_HOST (csrc/kernels/hostlog.cuh), pinned on the CPU.

The 16-entry (invc, logc) table and the polynomial are parsed out of the CUDA header, the algorithm
(glibc 2.39 sysdeps/ieee754/flt-32/e_logf.c) is restated in numpy float64, and the result is compared
bit for bit with this machine's libm logf -- the function the reference's host build calls per step.
(The fused and unfused double forms round to the same float except within ~1e-16 of a tie.)"""
import ctypes
import ctypes.util
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse_header():
    text = open(os.path.join(ROOT, "lyapunov3d_b200", "csrc", "kernels", "hostlog.cuh")).read()
    tab_src = text[text.index("kLogfTab[32]"):]
    tab_src = tab_src[tab_src.index("{") + 1:tab_src.index("}")]
    tab = [float.fromhex(t) for t in re.findall(r"-?0x[0-9a-fA-F.]+p[+-]?\d+", tab_src)]
    assert len(tab) == 32
    poly_src = text[text.index("kLogfPoly[5]"):]
    poly_src = poly_src[poly_src.index("{") + 1:poly_src.index("}")]
    poly = [float.fromhex(t) for t in re.findall(r"-?0x[0-9a-fA-F.]+p[+-]?\d+", poly_src)]
    assert len(poly) == 4
    return np.array(tab).reshape(16, 2), poly


def logf_model(x, tab, poly):
    a0, a1, a2, ln2 = poly
    ix = x.view(np.uint32).astype(np.int64)
    tmp = (ix - 0x3F330000) & 0xFFFFFFFF
    i = (tmp >> 19) & 15
    k = ((tmp.astype(np.uint32)).view(np.int32) >> 23).astype(np.float64)
    iz = ((ix - (tmp & 0xFF800000)) & 0xFFFFFFFF).astype(np.uint32)
    z = iz.view(np.float32).astype(np.float64)
    invc, logc = tab[i, 0], tab[i, 1]
    r = z * invc - 1.0
    y0 = logc + k * ln2
    r2 = r * r
    y = a1 * r + a2
    y = a0 * r2 + y
    y = y * r2 + (y0 + r)
    return y.astype(np.float32)


def test_logf_restatement_equals_this_machines_libm():
    tab, poly = parse_header()
    libm = ctypes.CDLL(ctypes.util.find_library("m"))
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(11)
    # normal positive floats, with emphasis on (0, 4]: the range of |r(1-2v)|
    bits = np.concatenate([rng.integers(0x00800000, 0x7F800000, 60000), rng.integers(0x30000000, 0x40800000, 140000)]).astype(np.uint32)
    x = bits.view(np.float32)
    want = np.array([libm.logf(float(v)) for v in x], np.float32)
    got = logf_model(x, tab, poly)
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    assert logf_model(np.array([1.0], np.float32), tab, poly)[0] == 0.0      # exact zero at x == 1, as glibc returns
    assert tab[9, 0] == 1.0 and tab[9, 1] == 0.0


def test_merged_table_hot_loop_form_equals_libm_and_poisons_everything_else():
    """The hot-loop form of hostlog.cuh: one table indexed by (k - kLogfKmin, i) holding invc*2^-k and
    logc + k*ln2, x used unreduced, the entry index clamped by an unsigned minimum to a poison entry
    (0, NaN) for every input the table does not cover.  Restated in numpy float64: inside the table's
    range it must give libm's logf bit for bit; outside (zero, subnormal, inf, nan, tiny, huge) it
    must give NaN, which is what sends the sample to the careful form."""
    tab, poly = parse_header()
    a0, a1, a2, ln2 = poly
    text = open(os.path.join(ROOT, "lyapunov3d_b200", "csrc", "kernels", "hostlog.cuh")).read()
    kmin = int(re.search(r"kLogfKmin\s*=\s*(-?\d+)", text).group(1))
    knum = int(re.search(r"kLogfKnum\s*=\s*(\d+)", text).group(1))
    n_ent = knum * 16
    e = np.arange(n_ent)
    i, k = e & 15, (e >> 4) + kmin
    invc_s = np.append(tab[i, 0] * np.ldexp(1.0, -k), 0.0)  # exact: a power-of-two scale; last = poison
    y0 = np.append(k.astype(np.float64) * ln2 + tab[i, 1], np.nan)
    off_k = 0x3F330000 - (-kmin) * 0x00800000
    assert re.search(r"kLogfOffK\s*=\s*0x3f330000u\s*-\s*\(uint32_t\)\(-kLogfKmin\)\s*\*\s*0x00800000u", text)

    def hot(x):
        bits = x.view(np.uint32).astype(np.int64)
        t = ((bits << 1) - 2 * off_k) & 0xFFFFFFFF              # shl, sub (mod 2^32)
        idx = np.minimum(t >> 20, n_ent)                        # shr, unsigned min
        xd = np.abs(x).astype(np.float64)
        with np.errstate(all="ignore"):
            r = xd * invc_s[idx] - 1.0
            r2 = r * r
            y = a1 * r + a2
            y = a0 * r2 + y
            y = y * r2 + (y0[idx] + r)
            return y.astype(np.float32), idx == n_ent

    libm = ctypes.CDLL(ctypes.util.find_library("m"))
    libm.logf.restype = ctypes.c_float
    libm.logf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(12)
    x_lo = np.array([off_k], np.uint32).view(np.float32)[0]          # smallest covered |x|: 0x1.66p-1 * 2^kmin
    x_hi = np.array([off_k + knum * 0x00800000 - 1], np.uint32).view(np.float32)[0]
    assert x_hi >= 4.0 and x_lo <= 2.0 ** kmin                       # |r (1 - 2v)| <= 4 on the cube
    lo, hi = int(off_k), int(off_k + knum * 0x00800000)
    bits = np.concatenate([rng.integers(lo, hi, 60000), rng.integers(0x3C000000, 0x40800000, 140000),
                           [lo, hi - 1]]).astype(np.uint32)
    x = bits.view(np.float32).copy()
    x[::2] *= -1                                            # the hot loop sees r(1-2v) with its sign
    got, poisoned = hot(x)
    want = np.array([libm.logf(abs(float(v))) for v in x], np.float32)
    assert not poisoned.any()
    assert (got.view(np.uint32) == want.view(np.uint32)).all()
    # every input outside the range must come back NaN (the careful form then redoes the sample)
    bad = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1e-45, 1e-39, -1e-40, 2.0 ** (kmin - 2), 2.0 ** (kmin - 40), 2.0 ** -126, 16.0, 100.0,
                    3e38, -12.0, 2.0 ** 60, 2.0 ** 67, 5.6, -5.7], np.float32)
    bad = np.concatenate([bad, np.array([lo - 1, hi], np.uint32).view(np.float32)])
    y, poisoned = hot(bad)
    assert poisoned.all() and np.isnan(y).all()
    # in-range boundaries are served by the table
    edge = np.array([x_lo, x_hi, 1.0, 0.7, 1.4, 4.0, -4.0, 5.59], np.float32)
    y, poisoned = hot(edge)
    assert not poisoned.any() and not np.isnan(y).any()
