// exponent.cuh -- the per-sample Lyapunov exponent (reference kernel.cu:108-154, lyap4d).
//
// One call = settle + accum steps of the sequence-forced logistic map for ONE
// sample point per lane.  The trip count is uniform across a warp, so every lane
// of a warp stays busy for the whole call; all divergence lives in the callers.
//
// Trajectory.  The reference evaluates v <- r*v*(1.0 - v) with a float product
// r*v and a double tail.  p = r*v; v' = fma(-p, v, p) gives the same float, bit
// for bit, as does fma(-2r, v', r) for the derivative
// r - 2.0*r*v'.  Two FP32 instructions per step, no FP64, no conversions.
//
// Three accumulators share that trajectory:
//   kExact  l = fma(lg2.approx(|d|), ln2, l) every step -- the exact PTX ops of the
//           reference's fast-math build, so l is bit-identical to the reference
//           kernel on the same GPU.  4 FP32 + 1 MUFU per step: SFU-bound.
//   kFast   the derivative's magnitude is multiplied up instead:
//           prod *= |1 - 2v'|, exponent folded out with integer ops every <= 20
//           steps (see Accum<kFast>), one lg2 at the end; sum(log r) is added analytically from the
//           per-symbol counts.  4 FP32 per step, no MUFU in the loop: FP32-bound.
//   kHost   IEEE adds of a glibc-exact logf per step (reference host build).
//
// Sequence.  The symbol sequence is warp-uniform.  For a period P <= 32 (and 36, 40) the
// per-position multipliers live in P registers and the period is fully
// unrolled (template parameter P); anything else takes the generic path (P == 0):
// a per-lane multiplier table in shared memory, or a run-length loop.
#pragma once
#include "arith.cuh"
#include "hostlog.cuh"

namespace lyap {

constexpr int kMaxSeq = 1024;       // symbols per period accepted by the ABI
constexpr int kMaxPeriodRegs = 40;  // longest period served by the register-table path
constexpr int kSymTabThreads = 256; // largest block that calls the packed evaluator (its per-thread multiplier columns)

enum Mode { kExact = 0, kFast = 1, kHost = 2 };

// Uniform description of the iteration schedule; lives in kernel parameter space.
struct SeqPlan {
    double ln2_over_accum;    // fast evaluator: ln 2 / accum (a warp-uniform double division otherwise paid per sample)
    double cnt_d[4];          // cnt[] as doubles (likewise)
    uint32_t len;             // period in symbols
    uint32_t settle, accum;   // reference prm.settle / prm.accum
    uint32_t settle_head;     // settle % len: steps taken before the rotated periods start
    uint32_t settle_periods;  // settle / len
    uint32_t accum_periods;   // accum / len
    uint32_t accum_tail;      // accum % len
    uint32_t cnt[4];          // how often each symbol occurs among the accum steps
    uint32_t fold_bias;       // float bits of 2^kBias; a run-time value so that each fold's
                              // (bits & mantissa) | bias stays ONE three-input LOP3
    uint32_t redo;            // fast mode: re-evaluate samples whose fold saw a zero exponent field (default 1)
    uint32_t n_runs;          // run-length form of the period (generic path): number of (symbol, length) pairs
    uint32_t table_stride;    // generic path: entries per lane of the shared-memory multiplier table (len | 1), 0 = run-length loop
    uint8_t rot[kMaxPeriodRegs];  // sym rotated left by settle_head: the order both period loops see
    uint8_t sym[kMaxSeq];         // the period as written (register-table path)
    uint8_t runs[2 * kMaxSeq];    // (symbol, length <= 255) pairs covering one period (generic path)
};

// Cursor over the run-length encoded period: which run, how far into it.
struct RunCursor {
    uint32_t run, off;
};

// Take `count` steps of the sequence from the cursor, calling step(r) once per step with the
// multiplier of the current run, fold() after every 8 steps and at the end of every run segment.
// Everything here is warp-uniform; the inner 8-step loop is unrolled with r in a register.
template <class Sel, class Step, class Fold>
__device__ __forceinline__ void run_steps(const SeqPlan &sp, RunCursor &cur, uint32_t count, Sel sel, Step step, Fold fold)
{
    while (count) {
        const uint32_t sym = sp.runs[2 * cur.run], len = sp.runs[2 * cur.run + 1];
        const uint32_t n = min(len - cur.off, count);
        const auto r = sel(sym);
        uint32_t i = 0;
#pragma unroll 1
        for (; i + 8 <= n; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) step(r);
            fold();
        }
#pragma unroll 1
        for (; i < n; ++i) step(r);
        fold();
        count -= n;
        cur.off += n;
        if (cur.off == len) {
            cur.off = 0;
            cur.run = (cur.run + 1 == sp.n_runs) ? 0 : cur.run + 1;
        }
    }
}

__device__ __forceinline__ float sel4(uint32_t s, float x, float y, float z, float d)
{
    return s == 0 ? x : (s == 1 ? y : (s == 2 ? z : d));
}

// ------------------------------------------------- generic path: per-lane multiplier table
// A period longer than the 32 registers of the unrolled path: the multiplier of every position of the
// (rotated) period is written once per evaluation into a per-lane strip of dynamic shared memory and
// read back with one LDS per step -- immediate offsets inside an 8-step unrolled block, the strip
// stride (len | 1 entries) odd so that the 32 lanes of a warp hit 32 different banks.  An SFU-bound
// exact step has 8 issue slots for its 6 instructions, so the load rides for free; the packed fast step
// (two rays per lane, 8-byte entries) has 8 FMA-pipe cycles for 5 issue slots.  Periods whose strips
// do not fit (launch.hpp) fall back to the run-length loop below.
__device__ __forceinline__ uint32_t seq_table_base(const SeqPlan &sp, uint32_t entry_bytes)
{
    return (uint32_t)__cvta_generic_to_shared(lyap_dyn_smem) + threadIdx.x * sp.table_stride * entry_bytes;
}
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned long long lds_b64(uint32_t addr)
{
    unsigned long long v;
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts_b64(uint32_t addr, unsigned long long v) { asm volatile("st.shared.b64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }

// `count` consecutive table entries starting at `addr`: step(address) per entry, fold() after every 8 and at the end.
template <int ESZ, class Step, class Fold>
__device__ __forceinline__ void table_span(uint32_t addr, uint32_t count, Step step, Fold fold)
{
#pragma unroll 1
    for (uint32_t n = count >> 3; n; --n, addr += 8 * ESZ) {
#pragma unroll
        for (int j = 0; j < 8; ++j) step(addr + j * ESZ);
        fold();
    }
#pragma unroll 1
    for (uint32_t n = count & 7u; n; --n, addr += ESZ) step(addr);
    fold();
}

// ---------------------------------------------------------------- trajectory
template <int MODE>
__device__ __forceinline__ void logistic_step(float r, float &v)
{
    if constexpr (MODE == kExact) {
        float p = ArithDev::mul(r, v);
        v = ArithDev::fma(-p, v, p);
    } else {
        float p = __fmul_rn(r, v);
        v = __fmaf_rn(-p, v, p);
    }
}

// --------------------------------------------------------------- accumulators
template <int MODE>
struct Accum;

template <>
struct Accum<kExact> {
    float l;
    __device__ __forceinline__ void init() { l = 0.0f; }
    __device__ __forceinline__ void step(float r, float &v)
    {
        logistic_step<kExact>(r, v);
        float d = ArithDev::fma(-(r + r), v, r);
        l = ArithDev::fma(ArithDev::lg2(fabsf(d)), 0.693147182f, l);
    }
    __device__ __forceinline__ void renorm() {}
    template <class Body>
    __device__ __forceinline__ void run_group(float &v, Body body)
    {
        body([&](float r, float &vv) { step(r, vv); });
    }
    // +-inf and NaN are absorbing under the adds above, so the reference's per-step
    // isfinite() early-out (kernel.cu:147) equals one test here.
    __device__ __forceinline__ float finish(const SeqPlan &sp, float, float, float, float, float)
    {
        return is_finite(l) ? ArithDev::div(l, __uint2float_rn(sp.accum)) : quiet_nan();
    }
};

template <>
struct Accum<kHost> {
    float l;
    bool dead;      // the sum went non-finite for good (a true log(0) = -inf, inf or nan): the sample is NAN whatever follows
    LogfCtx ctx;
    __device__ __forceinline__ void init() { l = 0.0f; dead = false; ctx.init(); }
    __device__ __forceinline__ void step(float r, float &v)
    {
        logistic_step<kHost>(r, v);
        const float d = __fmaf_rn(-(r + r), v, r);
        l = __fadd_rn(l, glibc_logf_speculative(d, ctx));
    }
    __device__ __forceinline__ void step_careful(float r, float &v)
    {
        logistic_step<kHost>(r, v);
        const float d = __fmaf_rn(-(r + r), v, r);
        l = __fadd_rn(l, glibc_logf_careful(fabsf(d)));
    }
    // generic (run-length) loop: the speculative logf with an immediate careful retry of a poisoned step
    __device__ __forceinline__ void step_checked(float r, float &v)
    {
        logistic_step<kHost>(r, v);
        const float d = __fmaf_rn(-(r + r), v, r);
        float y = glibc_logf_speculative(d, ctx);
        if (y != y) y = glibc_logf_careful(fabsf(d));
        l = __fadd_rn(l, y);
    }
    __device__ __forceinline__ void renorm() {}
    // One unrolled group of steps (a few periods, ~20 steps).  The hot loop's logf answers NaN for an
    // input its table does not cover (zero, |d| < ~5e-6, inf, nan: ~3e-6 of the steps of a chaotic orbit),
    // which turns the float sum NaN: the sum is looked at once per group, and a group that spoiled it
    // is replayed from the saved (v, l) with the careful logf -- a few hundred instructions for the
    // lanes concerned, once per ~10^3 groups, instead of a test per step or a replay of the sample.
    // A sum that is still not finite after the careful replay is final (the reference returns NAN for
    // it, kernel.cu:147-149); such a lane is not looked at again.
    template <class Body>
    __device__ __forceinline__ void run_group(float &v, Body body)
    {
        const float v0 = v, l0 = l;
        body([&](float r, float &vv) { step(r, vv); });
        if (!dead && !is_finite(l)) {
            v = v0;
            l = l0;
            body([&](float r, float &vv) { step_careful(r, vv); });
            dead = !is_finite(l);
        }
    }
    __device__ __forceinline__ float finish(const SeqPlan &sp, float, float, float, float, float)
    {
        return is_finite(l) ? __fdiv_rn(l, __uint2float_rn(sp.accum)) : quiet_nan();
    }
};

template <>
struct Accum<kFast> {
    // prod is kept in [2^kBias, 2^(kBias+1)) after each fold.  |1-2v| is 0 or at
    // least 2^-24 for any float v in [0,1], and never above 1, so TEN steps can
    // take prod no lower than 2^(kBias-240) > 2^-126: folding every <= 10 steps
    // (kSafeFold; the run-length loop folds every <= 8) cannot underflow, and a zero
    // exponent field then can only mean a true zero derivative.
    // The unrolled-period loops fold every kFoldEvery = 20 steps instead, which halves
    // the integer work: that is safe unless |1-2v| averages below 2^-12 over twenty
    // consecutive steps (orbits within ~1e-4 of the superstable point for the whole
    // block).  Such a sample -- like a true zero -- shows up as emin == 0 and is simply
    // evaluated again by the guaranteed-safe loop (fast_redo below).
    static constexpr int kBias = 120;
    static constexpr int kSafeFold = 10;
    static constexpr int kFoldEvery = 20;
    float prod;
    int esum, emin;
    __device__ __forceinline__ void init()
    {
        prod = __int_as_float((127 + kBias) << 23);
        esum = 0;
        emin = 255;
    }
    __device__ __forceinline__ void step(float r, float &v)
    {
        logistic_step<kFast>(r, v);
        float q = __fmaf_rn(-2.0f, v, 1.0f);
        prod = __fmul_rn(prod, fabsf(q));
    }
    int bias_bits;   // = (127 + kBias) << 23, from SeqPlan::fold_bias
    template <class Body>
    __device__ __forceinline__ void run_group(float &v, Body body)
    {
        body([&](float r, float &vv) { step(r, vv); });
    }
    __device__ __forceinline__ void renorm()
    {
        int bits = __float_as_int(prod);
        int e = bits >> 23;  // prod >= 0: no sign bit
        esum += e - (127 + kBias);
        emin = min(emin, e);
        prod = __int_as_float((bits & 0x007fffff) | bias_bits);
    }
    __device__ __forceinline__ float finish(const SeqPlan &sp, float x, float y, float z, float d, float v)
    {
        renorm();
        float mant = __int_as_float((__float_as_int(prod) & 0x007fffff) | 0x3f800000);
        double l2 = (double)esum + (double)__log2f(mant);
        if (sp.cnt[0]) l2 += (double)sp.cnt[0] * (double)__log2f(fabsf(x));
        if (sp.cnt[1]) l2 += (double)sp.cnt[1] * (double)__log2f(fabsf(y));
        if (sp.cnt[2]) l2 += (double)sp.cnt[2] * (double)__log2f(fabsf(z));
        if (sp.cnt[3]) l2 += (double)sp.cnt[3] * (double)__log2f(fabsf(d));
        float l = (float)(l2 * sp.ln2_over_accum);
        // zero derivative somewhere (log -> -inf) or an orbit that left [0,1] for good
        bool bad = (emin == 0) || !is_finite(v) || !is_finite(l);
        return bad ? quiet_nan() : l;
    }
};

// -------------------------------------------------------------------- driver
// Out-of-line, rarely taken: the guaranteed-safe evaluation of one sample (defined below).
static __device__ __noinline__ float fast_redo(const SeqPlan &sp, float x, float y, float z, float d, uint32_t strip_entry = 4);

template <int P>
struct PeriodUnroll {
    static constexpr int U = (P >= 11) ? 1 : (20 / (P > 0 ? P : 1));  // periods per loop body
    // packed fast evaluator: ~40-step bodies (two folds), which halves the loop control and the register
    // moves the compiler needs to close a body (3 of the 16 non-FMA instructions per 20 steps)
    static constexpr int U2 = (P >= 21) ? 1 : (40 / (P > 0 ? P : 1));
};

// `strip_entry`: bytes per entry of the calling kernel's per-lane strips in the generic path's multiplier
// table (4, or 8 in the two-rays-per-lane march kernel, whose warps alternate between this evaluator and
// the packed one independently of each other: every thread must stay inside its own strip).
template <int MODE, int P>
__device__ __forceinline__ float exponent(const SeqPlan &sp, float x, float y, float z, float d, uint32_t strip_entry = 4)
{
    float v = 0.5f;
    Accum<MODE> acc;
    acc.init();
    if constexpr (MODE == kFast) acc.bias_bits = (int)sp.fold_bias;
    float v_settled;

    if constexpr (P > 0) {
        float r[P];
#pragma unroll
        for (int k = 0; k < P; k++) r[k] = sel4(sp.rot[k], x, y, z, d);

        // settle: a partial period in written order, then whole periods in rotated order
        for (uint32_t n = 0; n < sp.settle_head; n++) logistic_step<MODE>(sel4(sp.sym[n], x, y, z, d), v);
#pragma unroll 1
        for (uint32_t i = 0; i < sp.settle_periods; i++) {
#pragma unroll
            for (int k = 0; k < P; k++) logistic_step<MODE>(r[k], v);
        }
        v_settled = v;

        // accumulate: whole periods (rotated order continues seamlessly), then a partial one
        constexpr int U = PeriodUnroll<P>::U;
        const uint32_t groups = sp.accum_periods / U;
#pragma unroll 1
        for (uint32_t g = 0; g < groups; g++) {
            acc.run_group(v, [&](auto step) {
#pragma unroll
                for (int s = 0; s < U * P; s++) {
                    step(r[s % P], v);
                    if ((s + 1) % Accum<kFast>::kFoldEvery == 0 || s + 1 == U * P) acc.renorm();
                }
            });
        }
        if constexpr (U > 1) {
#pragma unroll 1
            for (uint32_t i = groups * U; i < sp.accum_periods; i++) {
                acc.run_group(v, [&](auto step) {
#pragma unroll
                    for (int k = 0; k < P; k++) step(r[k], v);
                    acc.renorm();
                });
            }
        }
        acc.run_group(v, [&](auto step) {
#pragma unroll 1
            for (uint32_t n = 0; n < sp.accum_tail; n++) {
                step(sel4(sp.rot[n], x, y, z, d), v);
                if ((n & 7) == 7) acc.renorm();
            }
        });
    } else {
        bool tabled = false;
        if constexpr (MODE != kHost) tabled = sp.table_stride != 0;   // HOST mode's dynamic shared memory holds its logf table
        if (tabled) {
            const uint32_t T = seq_table_base(sp, strip_entry);
            {
                uint32_t pos = sp.settle_head, a = T;   // table in rotated order: entry k = position settle_head + k
                for (uint32_t k = 0; k < sp.len; ++k, a += 4) {
                    sts_f32(a, sel4(sp.sym[pos], x, y, z, d));
                    pos = (pos + 1 == sp.len) ? 0 : pos + 1;
                }
            }
            for (uint32_t n = 0; n < sp.settle_head; n++) logistic_step<MODE>(sel4(sp.sym[n], x, y, z, d), v);
#pragma unroll 1
            for (uint32_t i = 0; i < sp.settle_periods; i++)
                table_span<4>(T, sp.len, [&](uint32_t a) { logistic_step<MODE>(lds_f32(a), v); }, [] {});
            v_settled = v;
            // (a hand-pipelined variant of this loop -- next block's multipliers prefetched, previous block's
            // logs taken while the trajectory runs -- was slower: 0.71 against 0.79 of the SFU roofline)
#pragma unroll 1
            for (uint32_t i = 0; i < sp.accum_periods; i++)
                table_span<4>(T, sp.len, [&](uint32_t a) { acc.step(lds_f32(a), v); }, [&] { acc.renorm(); });
            table_span<4>(T, sp.accum_tail, [&](uint32_t a) { acc.step(lds_f32(a), v); }, [&] { acc.renorm(); });
        } else {
            RunCursor cur{0, 0};
            auto sel = [&](uint32_t s) { return sel4(s, x, y, z, d); };
            run_steps(sp, cur, sp.settle, sel, [&](float r) { logistic_step<MODE>(r, v); }, [] {});
            v_settled = v;
            if constexpr (MODE == kHost) run_steps(sp, cur, sp.accum, sel, [&](float r) { acc.step_checked(r, v); }, [] {});
            else run_steps(sp, cur, sp.accum, sel, [&](float r) { acc.step(r, v); }, [&] { acc.renorm(); });
        }
    }

    float l = acc.finish(sp, x, y, z, d, v);
    if constexpr (MODE == kFast && P > 0) {
        if (acc.emin == 0 && sp.redo) return fast_redo(sp, x, y, z, d);   // zero or underflow: ask the safe loop
    }
    // reference kernel.cu:138: an orbit sitting on v == 0.5 after settling skips the
    // accumulation and reports l = 0 / accum
    if (v_settled == 0.5f) {
        if constexpr (MODE == kExact) l = ArithDev::div(0.0f, __uint2float_rn(sp.accum));
        else l = __fdiv_rn(0.0f, __uint2float_rn(sp.accum));
    }
    return l;
}

// `strip_entry`: as for exponent() -- the safe loop of a generic-period launch reads the multipliers from the
// calling thread's own strip of the shared-memory table, whose width the calling kernel fixed.
static __device__ __noinline__ float fast_redo(const SeqPlan &sp, float x, float y, float z, float d, uint32_t strip_entry)
{
    return exponent<kFast, 0>(sp, x, y, z, d, strip_entry);
}

// ------------------------------------------------- two samples per lane (packed f32x2)
// sm_100 has packed single-precision FMA/MUL on 64-bit register pairs (FFMA2/FMUL2).  They
// occupy one ISSUE slot for two lanes' worth of work, which matters here because the fast
// exponent is issue-bound (four FP32 instructions per step plus folds and loop control on a
// one-instruction-per-clock scheduler).  The packed forms take no negate/abs modifiers, so
// the trajectory is carried as w = -v:  p' = r*w = -(r*v);  w' = fma(p', w, p') = -(p - p*v),
// bit-identical to the scalar form by the sign symmetry of round-to-nearest; the derivative
// factor is q = fma(2, w', 1) = 1 - 2v' and its sign is simply dropped at fold time.
typedef unsigned long long f32x2;

__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

struct AccumFast2 {
    static constexpr int kBias = Accum<kFast>::kBias;
    // Folds of the unrolled-period loops (renorm_pos) cost three integer instructions per sample:
    // the step before the fold multiplies with |prod| * |q| (two scalar FMULs with source modifiers
    // in place of one FMUL2: the same two issue cycles), so prod has no sign bit at the fold and
    //   esum += bits >> 23          one LEA.HI; the bias is taken off once at the end (nfold of them)
    //   emin  = min(emin, bits)     raw bits: an exponent field of zero <=> emin < 2^23
    //   prod  = (bits & mantissa) | bias
    // The sign-agnostic renorm() (five instructions) serves the remainder loops and the generic paths.
    f32x2 prod;
    int esum0, esum1, emin0, emin1, bias_bits, nfold;
    __device__ __forceinline__ void init(uint32_t fold_bias)
    {
        bias_bits = (int)fold_bias;
        const float b = __int_as_float((127 + kBias) << 23);
        prod = pack2(b, b);
        esum0 = esum1 = 0;
        emin0 = emin1 = 0x7fffffff;
        nfold = 0;
    }
    __device__ __forceinline__ void step(f32x2 r, f32x2 &w, f32x2 two, f32x2 one)
    {
        const f32x2 p = mul2(r, w);
        w = fma2(p, w, p);
        prod = mul2(prod, fma2(two, w, one));
    }
    // the step before a renorm_pos(): leaves prod >= 0
    __device__ __forceinline__ void step_abs(f32x2 r, f32x2 &w, f32x2 two, f32x2 one)
    {
        const f32x2 p = mul2(r, w);
        w = fma2(p, w, p);
        float pa, pb, qa, qb;
        unpack2(prod, pa, pb);
        unpack2(fma2(two, w, one), qa, qb);
        prod = pack2(__fmul_rn(fabsf(pa), fabsf(qa)), __fmul_rn(fabsf(pb), fabsf(qb)));
    }
    __device__ __forceinline__ void renorm_pos()
    {
        float a, b;
        unpack2(prod, a, b);
        const int ba = __float_as_int(a), bb = __float_as_int(b);
        esum0 += (int)((unsigned)ba >> 23);
        esum1 += (int)((unsigned)bb >> 23);
        emin0 = min(emin0, ba);
        emin1 = min(emin1, bb);
        nfold += 1;
        prod = pack2(__int_as_float((ba & 0x007fffff) | bias_bits), __int_as_float((bb & 0x007fffff) | bias_bits));
    }
    __device__ __forceinline__ void renorm()
    {
        float a, b;
        unpack2(prod, a, b);
        const int ba = __float_as_int(a) & 0x7fffffff, bb = __float_as_int(b) & 0x7fffffff;   // the sign of prod is not tracked: drop it
        esum0 += (ba >> 23) - (127 + kBias);
        esum1 += (bb >> 23) - (127 + kBias);
        emin0 = min(emin0, ba);
        emin1 = min(emin1, bb);
        prod = pack2(__int_as_float((ba & 0x007fffff) | bias_bits), __int_as_float((bb & 0x007fffff) | bias_bits));
    }
    // after the last fold: the biases of the renorm_pos() folds come off, emin becomes an exponent field
    __device__ __forceinline__ void close()
    {
        esum0 -= nfold * (127 + kBias);
        esum1 -= nfold * (127 + kBias);
        emin0 >>= 23;
        emin1 >>= 23;
    }
};

// `guard` (optional): half-width of the band around a threshold inside which a PARITY evaluator's value
// of this sample may lie on the other side (hybrid modes).  The fast value is accurate to ~1e-7; the
// parity evaluators add `accum` float logs one by one, each add rounding by at most half an ulp of the
// partial sum S_i, so they are off by at most ulp(max |S_i|) / 2 in units of l.  S_i = A_i + B_i with
// A_i = sum log|1 - 2v| (every term <= 0: monotone, so A_i lies in [A, 0] for the final A the fast
// evaluator holds) and B_i = sum log r over the symbols seen so far (between -Bneg and Bpos): hence
// max |S_i| <= max(|A| + Bneg, Bpos).  The caller adds the parity evaluator's per-term error.
__device__ __forceinline__ float lg2_ftz(float x)   // for normal x (a mantissa in [1,2)): no denormal pre-scaling
{
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float fast_finish(const SeqPlan &sp, int esum, int emin, float prod, float x, float y, float z,
                                             float d, float v, float *guard = nullptr)
{
    const float mant = __int_as_float((__float_as_int(prod) & 0x007fffff) | 0x3f800000);
    const float lgm = lg2_ftz(mant);
    double l2 = (double)esum + (double)lgm;
    float bpos = 0.0f, bneg = 0.0f;   // float is plenty for the guard: only the binade of the bound matters
    const float r4[4] = {x, y, z, d};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (sp.cnt[k]) {
            const float lg = __log2f(fabsf(r4[k]));
            l2 = __fma_rn(sp.cnt_d[k], (double)lg, l2);
            if (guard) {
                const float t = __uint2float_rn(sp.cnt[k]) * lg;
                bpos += fmaxf(t, 0.0f);
                bneg += fmaxf(-t, 0.0f);
            }
        }
    }
    const float l = (float)(l2 * sp.ln2_over_accum);
    const bool bad = (emin == 0) || !is_finite(v) || !is_finite(l);
    if (guard) {
        // |A| = -(esum + log2 mant) >= 0; 1.001: the float arithmetic above must not land just under a binade
        const float smax = fmaxf(bneg - ((float)esum + lgm), bpos) * (0.693147182f * 1.001f);
        // half an ulp of smax: its exponent field, scaled by 2^-24 (0 for smax == 0, inf/nan stay so)
        *guard = bad ? __int_as_float(0x7f800000) : __int_as_float(__float_as_int(smax) & 0x7f800000) * 5.9604644775390625e-8f;
    }
    return bad ? quiet_nan() : l;
}

// Exponents of two sample points A and B in one pass (fast mode only).
template <int P>
__device__ __forceinline__ void exponent_fast2(const SeqPlan &sp, float xa, float ya, float za, float xb, float yb, float zb,
                                               float d, float &la, float &lb, float *ga = nullptr, float *gb = nullptr)
{
    const f32x2 two = pack2(2.0f, 2.0f), one = pack2(1.0f, 1.0f);
    f32x2 w = pack2(-0.5f, -0.5f);
    AccumFast2 acc;
    acc.init(sp.fold_bias);
    float vsa, vsb;
    auto rpair = [&](uint32_t s) { return pack2(sel4(s, xa, ya, za, d), sel4(s, xb, yb, zb, d)); };
    auto settle_step = [&](f32x2 r) {
        const f32x2 p = mul2(r, w);
        w = fma2(p, w, p);
    };

    if constexpr (P > 0) {
        // Multiplier pair of a symbol: the four candidate pairs go to this thread's column of a small
        // shared array once per evaluation and come back with ONE LDS.64 per use at a warp-uniform row
        // offset -- the plan is decoded per evaluation, and a compare-and-branch select cost ~10
        // instructions per position of the period (profiles/r02_render_fast_1080p.md).  Only the
        // owning thread touches its column: no synchronisation.
        __shared__ f32x2 symtab[4][kSymTabThreads];
        const uint32_t tb = (uint32_t)__cvta_generic_to_shared(&symtab[0][threadIdx.x]);
        constexpr uint32_t kRow = kSymTabThreads * (uint32_t)sizeof(f32x2);
        sts_b64(tb, pack2(xa, xb));
        sts_b64(tb + kRow, pack2(ya, yb));
        sts_b64(tb + 2 * kRow, pack2(za, zb));
        sts_b64(tb + 3 * kRow, pack2(d, d));
        auto rsel = [&](uint32_t s) { return lds_b64(tb + s * kRow); };
        f32x2 r[P];
#pragma unroll
        for (int k = 0; k < P; k++) r[k] = rsel(sp.rot[k]);
        for (uint32_t n = 0; n < sp.settle_head; n++) settle_step(rsel(sp.sym[n]));
#pragma unroll 1
        for (uint32_t i = 0; i < sp.settle_periods; i++) {
#pragma unroll
            for (int k = 0; k < P; k++) settle_step(r[k]);
        }
        unpack2(w, vsa, vsb);
        constexpr int U = PeriodUnroll<P>::U2;
        const uint32_t groups = sp.accum_periods / U;
#pragma unroll 1
        for (uint32_t g = 0; g < groups; g++) {
#pragma unroll
            for (int s = 0; s < U * P; s++) {
                if ((s + 1) % Accum<kFast>::kFoldEvery == 0 || s + 1 == U * P) {
                    acc.step_abs(r[s % P], w, two, one);
                    acc.renorm_pos();
                } else {
                    acc.step(r[s % P], w, two, one);
                }
            }
        }
        if constexpr (U > 1) {
#pragma unroll 1
            for (uint32_t i = groups * U; i < sp.accum_periods; i++) {
#pragma unroll
                for (int k = 0; k < P; k++) acc.step(r[k], w, two, one);
                acc.renorm();
            }
        }
#pragma unroll 1
        for (uint32_t n = 0; n < sp.accum_tail; n++) {
            acc.step(rsel(sp.rot[n]), w, two, one);
            if ((n & 7) == 7) acc.renorm();
        }
    } else if (sp.table_stride != 0) {
        const uint32_t T = seq_table_base(sp, 8);
        {
            uint32_t pos = sp.settle_head, a = T;
            for (uint32_t k = 0; k < sp.len; ++k, a += 8) {
                sts_b64(a, rpair(sp.sym[pos]));
                pos = (pos + 1 == sp.len) ? 0 : pos + 1;
            }
        }
        for (uint32_t n = 0; n < sp.settle_head; n++) settle_step(rpair(sp.sym[n]));
#pragma unroll 1
        for (uint32_t i = 0; i < sp.settle_periods; i++) table_span<8>(T, sp.len, [&](uint32_t a) { settle_step(lds_b64(a)); }, [] {});
        unpack2(w, vsa, vsb);
        // 16-step blocks closed by the three-instruction fold (same cadence class as the unrolled periods:
        // an underflow shows as a zero exponent field and sends the sample to the safe loop below)
        auto span16 = [&](uint32_t addr, uint32_t count) {
#pragma unroll 1
            for (uint32_t n = count >> 4; n; --n, addr += 16 * 8) {
#pragma unroll
                for (int j = 0; j < 15; ++j) acc.step(lds_b64(addr + j * 8), w, two, one);
                acc.step_abs(lds_b64(addr + 15 * 8), w, two, one);
                acc.renorm_pos();
            }
#pragma unroll 1
            for (uint32_t n = count & 15u; n; --n, addr += 8) acc.step(lds_b64(addr), w, two, one);
            acc.renorm();
        };
#pragma unroll 1
        for (uint32_t i = 0; i < sp.accum_periods; i++) span16(T, sp.len);
        span16(T, sp.accum_tail);
    } else {
        RunCursor cur{0, 0};
        run_steps(sp, cur, sp.settle, rpair, settle_step, [] {});
        unpack2(w, vsa, vsb);
        run_steps(sp, cur, sp.accum, rpair, [&](f32x2 r) { acc.step(r, w, two, one); }, [&] { acc.renorm(); });
    }
    acc.renorm();
    acc.close();
    float pa, pb, wa, wb;
    unpack2(acc.prod, pa, pb);
    unpack2(w, wa, wb);
    la = fast_finish(sp, acc.esum0, acc.emin0, pa, xa, ya, za, d, wa, ga);
    lb = fast_finish(sp, acc.esum1, acc.emin1, pb, xb, yb, zb, d, wb, gb);
    if (P > 0 || sp.table_stride != 0) {
        // zero derivative or (with the 16/20-step folds) a possible underflow: ask the safe loop
        // (its guard stays infinite: a hybrid caller leaves such a sample to the parity evaluator)
        if (acc.emin0 == 0 && sp.redo) la = fast_redo(sp, xa, ya, za, d, 8);
        if (acc.emin1 == 0 && sp.redo) lb = fast_redo(sp, xb, yb, zb, d, 8);
    }
    const float zero = sp.accum ? 0.0f : quiet_nan();   // 0 / accum
    if (vsa == -0.5f) la = zero;   // w == -0.5 <=> v == 0.5 after settling (kernel.cu:138)
    if (vsb == -0.5f) lb = zero;
}

} // namespace lyap
