// hostlog.cuh -- glibc 2.39 logf(), bit for bit, on the device.
//
// LYAP_MODE_HOST has to reproduce the exponent of the reference's HOST build,
// whose `l += logf(r)` calls glibc.  The algorithm lives in a third-party
// dependency that is not vendored by the reference: GNU libc 2.39
// (Ubuntu 2.39-0ubuntu8.5 in this image), sysdeps/ieee754/flt-32/e_logf.c with the
// table of e_logf_data.c (Szabolcs Nagy's "optimized routines" logf):
//
//   x = 2^k * z, z in [0x1.66p-1, 0x1.66p0) picked by subtracting OFF = 0x3f330000
//   from the bits; i = top 4 mantissa bits of the shifted value;
//   r = z*invc[i] - 1;  y0 = logc[i] + k*ln2;
//   y = (A0*r^2 + (A1*r + A2)) * r^2 + (y0 + r);  return (float)y   (all in double)
//
// The 16-entry (invc, logc) table, ln2 and A[] below were read out of this image's
// libm.so.6 (.rodata, located by the ln2 bit pattern) and agree with the upstream
// source.  With or without FMA contraction of the double expressions the float
// result is the same except when y lies within ~1e-16 of a rounding boundary:
// 0 differences against the installed libm over 3e8 inputs for both forms, so the
// fused form is used here.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lyap {

static __device__ const double kLogfTab[32] = {
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2,
    0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,
    0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4,
    0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1p+0,               0x0p+0,
    0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5,  0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4,
    0x1.b2036576afce6p-1, 0x1.526e57720db08p-3,  0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3,
    0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,  0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2};

// Hot-loop form.  glibc splits x = 2^k * z and looks (invc, logc) up by the top four mantissa bits of
// z; here ONE shared-memory table, indexed by (k, i), holds the pair already combined with k:
//     invc' = invc[i] * 2^-k      (a power-of-two scale: exact)     so  z*invc - 1 == x*invc' - 1
//     y0    = logc[i] + k*ln2     (the same fused double expression the scalar form evaluates)
// which removes the reduction of x to z, the int->double conversion of k and two of the eight FP64
// operations from every step: r = fma((double)x, invc', -1); y = (A0 r^2 + (A1 r + A2)) r^2 + (y0 + r)
// -- bit for bit the value of the scalar form for every x the table covers.
//
// Layout.  The index differs per lane, and a 16-byte entry read by eight lanes of a quarter-warp costs
// one shared-memory wavefront per lane that shares a bank group with another one: measured on B200
// (tools/probe_hostlog.cu, profiles/r02_probe_hostlog.md) a plain [k][i] table takes 9.5 wavefronts per
// LDS.128 instead of 4 and the kernel is bound by the shared-memory pipe (98 % busy, 38.6 SMSP-cycles per
// step against 24.8 without conflicts).  So every entry is stored EIGHT times, replica (lane & 7) in bank
// group (lane & 7): conflict-free whatever the lanes look up.  That costs 128 bytes per entry, so the
// table covers only the k that occur in practice -- kLogfKnum values from kLogfKmin up, i.e.
// 2^kLogfKmin * 0.7 <= |x| < 5.6 (|r (1 - 2v)| <= 4 on the cube) -- plus ONE poison entry (0, NaN) to
// which every other input (zero, subnormal, tiny, huge, inf, nan) is clamped by an unsigned minimum:
// its logarithm comes out NaN and turns the caller's float sum NaN; the caller looks at the sum once
// per unrolled group of ~20 steps and replays a spoiled group with the careful form (glibc's special
// cases included; Accum<kHost>::run_group in exponent.cuh).  On a chaotic orbit ~3e-6 of the steps do
// that, i.e. one group in ~10^4.
constexpr int kLogfKmin = -17;
constexpr int kLogfKnum = 20;
constexpr uint32_t kLogfEntries = (uint32_t)kLogfKnum * 16u;                       // + 1 poison entry
constexpr uint32_t kLogfOffK = 0x3f330000u - (uint32_t)(-kLogfKmin) * 0x00800000u;  // glibc's OFF moved down to k = kLogfKmin
constexpr uint32_t kLogfSmemBytes = (kLogfEntries + 1u) * 128u;                     // dynamic shared memory of every HOST-mode kernel
extern __shared__ __align__(128) unsigned char lyap_dyn_smem[];

__device__ __forceinline__ void hostlog_init()
{
    double2 *tab = reinterpret_cast<double2 *>(lyap_dyn_smem);
    for (unsigned j = threadIdx.x; j < (kLogfEntries + 1u) * 8u; j += blockDim.x) {
        const unsigned e = j >> 3;                       // replica j & 7 of entry e
        double2 v = make_double2(0.0, __longlong_as_double(0x7ff8000000000000ll));
        if (e < kLogfEntries) {
            const int i = (int)(e & 15u), k = (int)(e >> 4) + kLogfKmin;
            const double scale = __hiloint2double((1023 - k) << 20, 0);                   // 2^-k
            v = make_double2(__dmul_rn(kLogfTab[2 * i], scale), __fma_rn((double)k, 0x1.62e42fefa39efp-1, kLogfTab[2 * i + 1]));
        }
        tab[j] = v;
    }
    __syncthreads();
}

// Polynomial, ln2 and the int->double magic as GLOBAL data: the values are loaded once per
// sample and then live in registers.  (As literals the compiler re-materialises every 64-bit
// constant with two uniform moves per use; as __constant__ it re-loads them per use.)
static __device__ const double kLogfPoly[5] = {-0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2,
                                               0x1.62e42fefa39efp-1, 4503601774854144.0 /* 2^52 + 2^31 */};

struct LogfCtx {
    double a0, a1, a2;
    uint32_t base;     // shared-window address of this lane's replica of entry 0
    __device__ __forceinline__ void init()
    {
        const volatile double *p = kLogfPoly;
        a0 = p[0]; a1 = p[1]; a2 = p[2];
        base = (uint32_t)__cvta_generic_to_shared(lyap_dyn_smem) + (threadIdx.x & 7u) * 16u;
    }
};

// Rare inputs: zero, subnormal, inf, nan, negative.  bit 32 of the result set: the low word is
// the final answer; clear: the low word is the normalised bit pattern to continue with.
static __device__ __noinline__ unsigned long long glibc_logf_special(uint32_t ix)
{
    const unsigned long long fin = 1ull << 32;
    if (ix * 2 == 0) return fin | 0xff800000u;                                   // log(0) = -inf
    if (ix == 0x7f800000u) return fin | ix;                                       // log(inf) = inf
    if ((ix & 0x80000000u) || ix * 2 >= 0xff000000u) return fin | 0x7fc00000u;    // negative or nan
    return __float_as_uint(__fmul_rn(__uint_as_float(ix), 8388608.0f)) - (23u << 23);   // subnormal * 2^23 (exact)
}

// The scalar form for a normal positive x given as its bit pattern: glibc's own sequence of operations
// (rare path: the careful redo of a sample, and shading).
__device__ __forceinline__ float glibc_logf_bits(uint32_t ix)
{
    const uint32_t tmp = ix - 0x3f330000u;
    const uint32_t i = (tmp >> 19) & 15u;
    const int k = (int)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double invc = kLogfTab[2 * i], logc = kLogfTab[2 * i + 1];
    const double z = (double)__uint_as_float(iz);
    const double r = __fma_rn(z, invc, -1.0);
    const double y0 = __fma_rn((double)k, kLogfPoly[3], logc);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(kLogfPoly[1], r, kLogfPoly[2]);
    y = __fma_rn(kLogfPoly[0], r2, y);
    y = __fma_rn(y, r2, __dadd_rn(y0, r));
    return __double2float_rn(y);
}

// |x| with full glibc semantics (any input).
static __device__ __noinline__ float glibc_logf_careful(float x)
{
    uint32_t ix = __float_as_uint(x);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u) {
        const unsigned long long s = glibc_logf_special(ix);
        if (s >> 32) return __uint_as_float((uint32_t)s);
        ix = (uint32_t)s;
    }
    return glibc_logf_bits(ix);
}

// Speculative form for the hot loop: the merged-table path on |x| whatever x is.  An input outside the
// table comes back as NaN (poison entry); the caller tests its float SUM once per group of steps and
// replays the group with glibc_logf_careful when it is not finite.
// Per step: 4 integer ops, 1 LDS.128 (conflict-free), 6 FP64 ops, 2 conversions.
__device__ __forceinline__ uint32_t glibc_logf_entry_addr(float x, uint32_t base)
{
    // (bits << 1) drops the sign; t = 2 (|x| bits - OffK) is below kLogfKnum << 24 exactly for the x the
    // table covers (any other x wraps to a larger unsigned value) and t >> 20 is the entry index.
    // LEA, SHF, VIMNMX, LEA.  (Inline PTX: written in C the compiler rewrites the shift pair into add,
    // shift and an extra mask.)
    uint32_t e;
    asm("{ .reg .u32 t;\n\t"
        "shl.b32 t, %1, 1;\n\t"
        "sub.u32 t, t, %2;\n\t"
        "shr.u32 %0, t, 20; }"
        : "=r"(e)
        : "r"(__float_as_uint(x)), "n"(2u * kLogfOffK));
    return base + (min(e, kLogfEntries) << 7);
}

__device__ __forceinline__ float glibc_logf_speculative(float x, const LogfCtx &c)
{
    double2 e;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(e.x), "=d"(e.y) : "r"(glibc_logf_entry_addr(x, c.base)));
    const double xd = (double)fabsf(x);
    const double r = __fma_rn(xd, e.x, -1.0);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(c.a1, r, c.a2);
    y = __fma_rn(c.a0, r2, y);
    y = __fma_rn(y, r2, __dadd_rn(e.y, r));
    return __double2float_rn(y);
}

// Single-lane variant for divergent callers (shading).
__device__ __forceinline__ float glibc_logf_lane(float x) { return glibc_logf_careful(x); }

} // namespace lyap
