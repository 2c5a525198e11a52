// probe_hostlog.cu -- which form of the bit-exact glibc logf step is fastest on a B200?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -lineinfo -o build/probe_hostlog tools/probe_hostlog.cu
//
// Every variant runs the HOST-mode accumulate step of the exponent (FMUL, FFMA, FFMA, logf, FADD) on
// chaotic BCABA orbits, 128 threads x 4 blocks per SM like render_kernel, and is checked bit for bit
// against the scalar glibc form.  Printed: T lane-steps/s and SMSP cycles per warp-step.
//
// The table constants are those of lyapunov3d_b200/csrc/kernels/hostlog.cuh (glibc 2.39 e_logf_data.c).
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

static __device__ const double kTab[32] = {
    0x1.661ec79f8f3bep+0, -0x1.57bf7808caadep-2, 0x1.571ed4aaf883dp+0, -0x1.2bef0a7c06ddbp-2,
    0x1.49539f0f010b0p+0, -0x1.01eae7f513a67p-2, 0x1.3c995b0b80385p+0, -0x1.b31d8a68224e9p-3,
    0x1.30d190c8864a5p+0, -0x1.6574f0ac07758p-3, 0x1.25e227b0b8ea0p+0, -0x1.1aa2bc79c8100p-3,
    0x1.1bb4a4a1a343fp+0, -0x1.a4e76ce8c0e5ep-4, 0x1.12358f08ae5bap+0, -0x1.1973c5a611cccp-4,
    0x1.0953f419900a7p+0, -0x1.252f438e10c1ep-5, 0x1p+0,               0x0p+0,
    0x1.e608cfd9a47acp-1, 0x1.aa5aa5df25984p-5,  0x1.ca4b31f026aa0p-1, 0x1.c5e53aa362eb4p-4,
    0x1.b2036576afce6p-1, 0x1.526e57720db08p-3,  0x1.9c2d163a1aa2dp-1, 0x1.bc2860d224770p-3,
    0x1.886e6037841edp-1, 0x1.1058bc8a07ee1p-2,  0x1.767dcf5534862p-1, 0x1.4043057b6ee09p-2};
static __device__ const double kPoly[5] = {-0x1.00ea348b88334p-2, 0x1.5575b0be00b6ap-2, -0x1.ffffef20a4123p-2,
                                           0x1.62e42fefa39efp-1, 4503601774854144.0 /* 2^52 + 2^31 */};

extern __shared__ __align__(16) unsigned char smem[];

struct Ctx {
    double a0, a1, a2, ln2, magic;
    uint32_t tab, tab2, lane16, lane8;
};

// scalar glibc form (FMA-contracted): the reference every variant must equal
__device__ __forceinline__ float logf_scalar(float x, const Ctx &)
{
    const uint32_t ix = __float_as_uint(fabsf(x));
    const uint32_t tmp = ix - 0x3f330000u;
    const uint32_t i = (tmp >> 19) & 15u;
    const int k = (int)tmp >> 23;
    const uint32_t iz = ix - (tmp & 0xff800000u);
    const double invc = kTab[2 * i], logc = kTab[2 * i + 1];
    const double z = (double)__uint_as_float(iz);
    const double r = __fma_rn(z, invc, -1.0);
    const double y0 = __fma_rn((double)k, kPoly[3], logc);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(kPoly[1], r, kPoly[2]);
    y = __fma_rn(kPoly[0], r2, y);
    y = __fma_rn(y, r2, __dadd_rn(y0, r));
    return __double2float_rn(y);
}

enum { V_SCALAR = 0, V_MERGED = 1, V_MERGED_UNIFORM = 2, V_ITAB_ARITHK = 3, V_ITAB_KTAB = 4, V_ITAB_ARITHK_INTCVT = 5, V_MERGED16_REPL = 6,
       V_ITAB_NOREPL = 7, V_MERGED_INTCVT = 8, V_FINAL = 9, V_COUNT = 10 };
static const char *kNames[V_COUNT] = {"scalar glibc form, global table", "merged (k,i) table 16 KB [current]", "merged table, warp-uniform orbit (no conflicts)",
                                      "i-table x8 replicas + arithmetic k", "i-table x8 + k-table x16 (LDS.64), unfused y0", "i-table x8 + arithmetic k, int f32->f64",
                                      "merged table, 16 k-values, x8 replicas (32 KB)", "i-table 256 B unreplicated + arithmetic k", "merged (k,i) table, int f32->f64",
                                      "merged, 20 k-values, x8 replicas, clamp to a NaN entry [adopted]"};

template <int V>
__device__ __forceinline__ void table_init(Ctx &c)
{
    c.a0 = kPoly[0]; c.a1 = kPoly[1]; c.a2 = kPoly[2]; c.ln2 = kPoly[3]; c.magic = kPoly[4];
    const unsigned lane = threadIdx.x & 31;
    c.lane16 = (lane & 7u) * 16u;
    c.lane8 = (lane & 15u) * 8u;
    c.tab = (uint32_t)__cvta_generic_to_shared(smem);
    c.tab2 = c.tab;
    double2 *t = reinterpret_cast<double2 *>(smem);
    if constexpr (V == V_MERGED || V == V_MERGED_UNIFORM || V == V_MERGED_INTCVT) {
        for (unsigned e = threadIdx.x; e < 1024; e += blockDim.x) {
            const int i = (int)(e & 15u);
            const int k = (int)(((e >> 4) + 60u) & 63u) - 60;
            const double scale = __hiloint2double((1023 - k) << 20, 0);
            t[e] = make_double2(__dmul_rn(kTab[2 * i], scale), __fma_rn((double)k, kPoly[3], kTab[2 * i + 1]));
        }
    } else if constexpr (V == V_ITAB_ARITHK || V == V_ITAB_ARITHK_INTCVT || V == V_ITAB_KTAB) {
        for (unsigned e = threadIdx.x; e < 128; e += blockDim.x) t[e] = make_double2(kTab[2 * (e >> 3)], kTab[2 * (e >> 3) + 1]);   // [i][replica]
        if constexpr (V == V_ITAB_KTAB) {
            double *kt = reinterpret_cast<double *>(smem + 2048);   // [k & 63][16 replicas]
            c.tab2 = c.tab + 2048;
            for (unsigned e = threadIdx.x; e < 1024; e += blockDim.x) {
                const int k = (int)(((e >> 4) + 60u) & 63u) - 60;
                kt[e] = __dmul_rn((double)k, kPoly[3]);
            }
        }
    } else if constexpr (V == V_MERGED16_REPL) {
        for (unsigned e = threadIdx.x; e < 2048; e += blockDim.x) {   // [(k & 15) * 16 + i][replica]
            const unsigned ent = e >> 3;
            const int i = (int)(ent & 15u);
            const int k = (int)(((ent >> 4) + 13u) & 15u) - 13;   // k in [-13, 2]
            const double scale = __hiloint2double((1023 - k) << 20, 0);
            t[e] = make_double2(__dmul_rn(kTab[2 * i], scale), __fma_rn((double)k, kPoly[3], kTab[2 * i + 1]));
        }
    } else if constexpr (V == V_FINAL) {
        for (unsigned j = threadIdx.x; j < 321u * 8u; j += blockDim.x) {   // hostlog.cuh: [(k + 17) * 16 + i][replica], entry 320 = poison
            const unsigned ent = j >> 3;
            double2 v = make_double2(0.0, __longlong_as_double(0x7ff8000000000000ll));
            if (ent < 320u) {
                const int i = (int)(ent & 15u), k = (int)(ent >> 4) - 17;
                v = make_double2(__dmul_rn(kTab[2 * i], __hiloint2double((1023 - k) << 20, 0)), __fma_rn((double)k, kPoly[3], kTab[2 * i + 1]));
            }
            t[j] = v;
        }
        c.tab += c.lane16;
    } else if constexpr (V == V_ITAB_NOREPL) {
        for (unsigned e = threadIdx.x; e < 16; e += blockDim.x) t[e] = make_double2(kTab[2 * e], kTab[2 * e + 1]);
    }
    __syncthreads();
}

__device__ __forceinline__ double2 lds128(uint32_t addr)
{
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds64(uint32_t addr)
{
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ double f32_to_f64_int(uint32_t ix)   // positive normal float bits -> double, integer ops only
{
    return __hiloint2double((int)((ix >> 3) + 0x38000000u), (int)(ix << 29));
}

template <int V>
__device__ __forceinline__ float logf_v(float x, const Ctx &c, bool &odd)
{
    if constexpr (V == V_SCALAR) {
        return logf_scalar(x, c);
    } else if constexpr (V == V_MERGED || V == V_MERGED_UNIFORM || V == V_MERGED_INTCVT) {
        const uint32_t idx = ((__float_as_uint(x) - 0x3f330000u) >> 15) & 0x3ff0u;
        const double2 e = lds128(c.tab + idx);
        const double xd = (V == V_MERGED_INTCVT) ? f32_to_f64_int(__float_as_uint(x) & 0x7fffffffu) : (double)fabsf(x);
        const double r = __fma_rn(xd, e.x, -1.0);
        const double r2 = __dmul_rn(r, r);
        odd = odd || ((uint32_t)__double2hiint(r2) >= 0x3fb00000u);
        double y = __fma_rn(c.a1, r, c.a2);
        y = __fma_rn(c.a0, r2, y);
        y = __fma_rn(y, r2, __dadd_rn(e.y, r));
        return __double2float_rn(y);
    } else if constexpr (V == V_FINAL) {
        uint32_t ent;
        asm("{ .reg .u32 t;\n\tshl.b32 t, %1, 1;\n\tsub.u32 t, t, %2;\n\tshr.u32 %0, t, 20; }" : "=r"(ent) : "r"(__float_as_uint(x)), "n"(2u * (0x3f330000u - 17u * 0x00800000u)));
        const double2 e = lds128(c.tab + (min(ent, 320u) << 7));
        const double xd = (double)fabsf(x);
        const double r = __fma_rn(xd, e.x, -1.0);
        const double r2 = __dmul_rn(r, r);
        double y = __fma_rn(c.a1, r, c.a2);
        y = __fma_rn(c.a0, r2, y);
        y = __fma_rn(y, r2, __dadd_rn(e.y, r));
        return __double2float_rn(y);
    } else if constexpr (V == V_MERGED16_REPL) {
        const uint32_t tmp = __float_as_uint(x) - 0x3f330000u;
        const uint32_t idx = ((tmp >> 12) & 0x7f80u) | c.lane16;
        const double2 e = lds128(c.tab + idx);
        const double xd = (double)fabsf(x);
        const double r = __fma_rn(xd, e.x, -1.0);
        const double r2 = __dmul_rn(r, r);
        odd = odd || ((uint32_t)__double2hiint(r2) >= 0x3fb00000u);
        double y = __fma_rn(c.a1, r, c.a2);
        y = __fma_rn(c.a0, r2, y);
        y = __fma_rn(y, r2, __dadd_rn(e.y, r));
        return __double2float_rn(y);
    } else {
        const uint32_t ix = __float_as_uint(x) & 0x7fffffffu;
        const uint32_t tmp = ix - 0x3f330000u;
        odd = odd || (ix - 0x00800000u >= 0x7f000000u);
        const uint32_t iz = ix - (tmp & 0xff800000u);
        const uint32_t ioff = (V == V_ITAB_NOREPL) ? ((tmp >> 15) & 0xf0u) : (((tmp >> 12) & 0x780u) | c.lane16);
        const double2 e = lds128(c.tab + ioff);
        const double z = (V == V_ITAB_ARITHK_INTCVT) ? f32_to_f64_int(iz) : (double)__uint_as_float(iz);
        const double r = __fma_rn(z, e.x, -1.0);
        double y0;
        if constexpr (V == V_ITAB_KTAB) {
            const uint32_t koff = ((tmp >> 16) & 0x1f80u) | c.lane8;
            y0 = __dadd_rn(e.y, lds64(c.tab2 + koff));
        } else {
            const int k = (int)tmp >> 23;
            const double kd = __dadd_rn(__hiloint2double(0x43300000, k ^ (int)0x80000000), -c.magic);
            y0 = __fma_rn(kd, c.ln2, e.y);
        }
        const double r2 = __dmul_rn(r, r);
        double y = __fma_rn(c.a1, r, c.a2);
        y = __fma_rn(c.a0, r2, y);
        y = __fma_rn(y, r2, __dadd_rn(y0, r));
        return __double2float_rn(y);
    }
}

__device__ __forceinline__ uint32_t hash(uint32_t a)
{
    a ^= a >> 16; a *= 0x7feb352dU; a ^= a >> 15; a *= 0x846ca68bU; a ^= a >> 16;
    return a;
}

template <int V>
__global__ void __launch_bounds__(128, 4) probe(float *out, unsigned *odd_out, int groups)
{
    Ctx c;
    table_init<V>(c);
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (V == V_MERGED_UNIFORM) id &= ~31u;
    const float x = 2.9f + 1.1f * (hash(3 * id) >> 8) * (1.0f / 16777216.0f);
    const float y = 2.9f + 1.1f * (hash(3 * id + 1) >> 8) * (1.0f / 16777216.0f);
    const float z = 2.9f + 1.1f * (hash(3 * id + 2) >> 8) * (1.0f / 16777216.0f);
    const float r[5] = {y, z, x, y, x};   // BCABA
    float v = 0.5f, l = 0.0f;
    bool odd = false;
    for (int s = 0; s < 18; ++s) { const float p = __fmul_rn(r[s % 5], v); v = __fmaf_rn(-p, v, p); }
#pragma unroll 1
    for (int g = 0; g < groups; ++g) {
#pragma unroll
        for (int s = 0; s < 20; ++s) {
            const float rr = r[s % 5];
            const float p = __fmul_rn(rr, v);
            v = __fmaf_rn(-p, v, p);
            const float d = __fmaf_rn(-(rr + rr), v, rr);
            l = __fadd_rn(l, logf_v<V>(d, c, odd));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = l;
    odd_out[blockIdx.x * blockDim.x + threadIdx.x] = odd || (V == V_FINAL && !(fabsf(l) <= 3.0e38f));   // the adopted form flags through a NaN sum
}

template <int V>
static void run(int groups, float *d_out, unsigned *d_odd, float *h_out, unsigned *h_odd, const float *h_ref, const unsigned *h_ref_odd, double clock_hz, int sms)
{
    const int blocks = sms * 4, threads = 128;
    const size_t sm = 321 * 128;
    cudaFuncSetAttribute(probe<V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    const size_t need = (V == V_MERGED || V == V_MERGED_UNIFORM || V == V_MERGED_INTCVT) ? 16384 : (V == V_MERGED16_REPL) ? 32768 : (V == V_FINAL) ? 321 * 128 : (V == V_ITAB_KTAB) ? 2048 + 8192 : 2048;
    probe<V><<<blocks, threads, need>>>(d_out, d_odd, groups);   // warm-up
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        probe<V><<<blocks, threads, need>>>(d_out, d_odd, groups);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    const size_t n = (size_t)blocks * threads;
    cudaMemcpy(h_out, d_out, n * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(h_odd, d_odd, n * 4, cudaMemcpyDeviceToHost);
    size_t diff = 0, nodd = 0, compared = 0;
    if (h_ref && V != V_MERGED_UNIFORM) {
        for (size_t i = 0; i < n; ++i) {
            nodd += h_odd[i] != 0;
            if (h_odd[i] || h_ref_odd[i]) continue;
            ++compared;
            uint32_t a, b;
            memcpy(&a, &h_out[i], 4); memcpy(&b, &h_ref[i], 4);
            diff += a != b;
        }
    }
    const double steps = (double)groups * 20.0;
    const double lane_steps = steps * (double)n;
    const double cyc = best * 1e-3 * clock_hz / (steps * 4.0);   // 4 warps per SMSP
    printf("%-52s %8.3f ms  %6.3f T lane-steps/s  %6.2f SMSP-cycles per warp-step   mismatches %zu of %zu (flagged %zu)  %s\n", kNames[V], best,
           lane_steps / (best * 1e-3) / 1e12, cyc, diff, compared, nodd, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main(int argc, char **argv)
{
    const int groups = argc > 1 ? atoi(argv[1]) : 1000;
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clock_hz = khz * 1e3;
    const int sms = p.multiProcessorCount;
    const size_t n = (size_t)sms * 4 * 128;
    float *d_out; unsigned *d_odd;
    cudaMalloc(&d_out, n * 4); cudaMalloc(&d_odd, n * 4);
    float *h_ref = (float *)malloc(n * 4), *h_out = (float *)malloc(n * 4);
    unsigned *h_ref_odd = (unsigned *)calloc(n, 4), *h_odd = (unsigned *)malloc(n * 4);
    printf("%s, %d SMs, %.0f MHz, %d groups of 20 steps\n", p.name, sms, clock_hz / 1e6, groups);
    run<V_SCALAR>(groups, d_out, d_odd, h_ref, h_odd, nullptr, nullptr, clock_hz, sms);
    run<V_MERGED>(groups, d_out, d_odd, h_out, h_odd, h_ref, h_ref_odd, clock_hz, sms);
    run<V_MERGED_UNIFORM>(groups, d_out, d_odd, h_out, h_odd, h_ref, h_ref_odd, clock_hz, sms);
    run<V_MERGED_INTCVT>(groups, d_out, d_odd, h_out, h_odd, h_ref, h_ref_odd, clock_hz, sms);
    run<V_MERGED16_REPL>(groups, d_out, d_odd, h_out, h_odd, h_ref, h_ref_odd, clock_hz, sms);
    run<V_ITAB_NOREPL>(groups, d_out, d_odd, h_out, h_odd, h_ref, h_ref_odd, clock_hz, sms);
    run<V_ITAB_ARITHK>(groups, d_out, d_odd, h_out, h_odd, h_ref, h_ref_odd, clock_hz, sms);
    run<V_ITAB_ARITHK_INTCVT>(groups, d_out, d_odd, h_out, h_odd, h_ref, h_ref_odd, clock_hz, sms);
    run<V_ITAB_KTAB>(groups, d_out, d_odd, h_out, h_odd, h_ref, h_ref_odd, clock_hz, sms);
    run<V_FINAL>(groups, d_out, d_odd, h_out, h_odd, h_ref, h_ref_odd, clock_hz, sms);
    return 0;
}
