// misc.cu -- shade-only pass, tile scatter, and the roofline probe kernels.
#include <cstdio>
#include <cstdlib>

#include <cub/device/device_radix_sort.cuh>

#include "launch.hpp"

namespace lyap {

// ---- tile order of a frame launch -----------------------------------------------------------
// Key of a tile: the chord of its centre ray through the cube in march steps' own unit (t1 - t0 of
// ray_begin), an upper bound on how long that ray can march with a fixed step (stepMethod 2); with
// stepMethod 1 every ray takes `depth` steps, so all hitting tiles get the same key and keep image
// order.  Tiles whose centre ray misses the cube get 0 and go last.  A stable descending radix sort:
// same keys on every rank (same arithmetic on the same inputs), same permutation.
struct TileKeyArgs {
    lyap_cam cam;
    lyap_params prm;
    uint32_t width, height, tile, tiles_x, n_tiles;
    float *keys;
    uint32_t *ids;
};

__global__ void __launch_bounds__(256) tile_key_kernel(const __grid_constant__ TileKeyArgs a)
{
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= a.n_tiles) return;
    const uint32_t px = min((j % a.tiles_x) * a.tile + a.tile / 2, a.width - 1);
    const uint32_t py = min((j / a.tiles_x) * a.tile + a.tile / 2, a.height - 1);
    MarchState st;
    float key = 0.0f;
    if (ray_begin<ArithDev>(st, px, py, a.cam, a.prm)) {
        key = a.prm.stepMethod == 1 ? 1.0f : st.t1 - st.t;
        if (!(key > 0.0f)) key = 0.0f;   // NaN or a degenerate chord
    }
    a.keys[j] = key;
    a.ids[j] = j;
}

static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }

static size_t tile_sort_temp_bytes(uint32_t n_tiles)
{
    size_t temp = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, temp, (const float *)nullptr, (float *)nullptr, (const uint32_t *)nullptr,
                                              (uint32_t *)nullptr, (int)n_tiles);
    return temp;
}

size_t tile_order_scratch_bytes(uint32_t n_tiles)
{
    return 4 * align256((size_t)n_tiles * 4) + align256(tile_sort_temp_bytes(n_tiles));
}

cudaError_t launch_tile_order(const RenderArgs &r, void *scratch, const uint32_t **order, cudaStream_t s)
{
    const size_t arr = align256((size_t)r.n_tiles * 4);
    unsigned char *base = static_cast<unsigned char *>(scratch);
    float *keys_in = reinterpret_cast<float *>(base), *keys_out = reinterpret_cast<float *>(base + arr);
    uint32_t *ids_in = reinterpret_cast<uint32_t *>(base + 2 * arr), *ids_out = reinterpret_cast<uint32_t *>(base + 3 * arr);
    TileKeyArgs a;
    a.cam = r.cam;
    a.prm = r.prm;
    a.width = r.width;
    a.height = r.height;
    a.tile = r.tile;
    a.tiles_x = r.tiles_x;
    a.n_tiles = r.n_tiles;
    a.keys = keys_in;
    a.ids = ids_in;
    tile_key_kernel<<<(r.n_tiles + 255) / 256, 256, 0, s>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    size_t temp = tile_sort_temp_bytes(r.n_tiles);
    e = cub::DeviceRadixSort::SortPairsDescending(base + 4 * arr, temp, keys_in, keys_out, ids_in, ids_out, (int)r.n_tiles, 0, 32, s);
    *order = ids_out;
    return e;
}

cudaError_t launch_shade(int mode, const ShadeArgs &a, unsigned grid, cudaStream_t s)
{
    if (mode == kHost) shade_kernel<kHost><<<grid, 256, 0, s>>>(a);
    else shade_kernel<kExact><<<grid, 256, 0, s>>>(a);
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256) scatter_kernel(const __grid_constant__ ScatterArgs a)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t tt = a.tile * a.tile;
    for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < a.n_items; k += stride) {
        const uint32_t m = (uint32_t)(k / tt), p = (uint32_t)(k % tt);
        const uint32_t j = m * a.world + a.rank;
        const uint32_t x = (j % a.tiles_x) * a.tile + p % a.tile;
        const uint32_t y = (j / a.tiles_x) * a.tile + p / a.tile;
        if (j >= a.n_tiles || x >= a.width || y >= a.height) continue;
        const uint8_t *src = a.compact + k * a.elem;
        uint8_t *dst = a.image + ((uint64_t)y * a.width + x) * a.elem;
        if (a.elem % 4 == 0) {
            for (uint32_t b = 0; b < a.elem; b += 4) *reinterpret_cast<uint32_t *>(dst + b) = *reinterpret_cast<const uint32_t *>(src + b);
        } else {
            for (uint32_t b = 0; b < a.elem; ++b) dst[b] = src[b];
        }
    }
}

__global__ void __launch_bounds__(256) assist_build_kernel(const __grid_constant__ AssistBuildArgs a)
{
    const uint64_t total = (uint64_t)a.n * a.n * a.n;
    const uint64_t words = (total + 31) / 32;
    const uint64_t stride = (uint64_t)gridDim.x * (blockDim.x / 32);
    const unsigned lane = threadIdx.x & 31;
    const int n = (int)a.n, d = (int)a.dilate;
    for (uint64_t w = (uint64_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32; w < words; w += stride) {
        const uint64_t cell = w * 32 + lane;
        bool safe = cell < total;
        if (safe) {
            const int ix = (int)(cell % a.n), iy = (int)((cell / a.n) % a.n), iz = (int)(cell / ((uint64_t)a.n * a.n));
            if (ix - d < 0 || iy - d < 0 || iz - d < 0 || ix + 1 + d >= n || iy + 1 + d >= n || iz + 1 + d >= n) {
                safe = false;
            } else {
                for (int z = iz - d; z <= iz + 1 + d && safe; ++z)
                    for (int y = iy - d; y <= iy + 1 + d && safe; ++y)
                        for (int x = ix - d; x <= ix + 1 + d; ++x) {
                            const uint64_t k = (uint64_t)x + (uint64_t)a.n * ((uint64_t)y + (uint64_t)a.n * (uint64_t)z);
                            const float v = a.f16 ? __half2float(reinterpret_cast<const __half *>(a.vol)[k]) : reinterpret_cast<const float *>(a.vol)[k];
                            if (!(v > a.lo && v < a.hi)) { safe = false; break; }      // NaN fails both compares
                        }
            }
        }
        const unsigned word = __ballot_sync(0xffffffffu, safe);
        if (lane == 0) a.bits[w] = word;
    }
}

cudaError_t launch_assist_build(const AssistBuildArgs &a, unsigned grid, cudaStream_t s)
{
    assist_build_kernel<<<grid, 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_scatter(const ScatterArgs &a, unsigned grid, cudaStream_t s)
{
    scatter_kernel<<<grid, 256, 0, s>>>(a);
    return cudaGetLastError();
}

// ---- diagnostics: the ray set-up and the normalisation in a mode's own arithmetic ----------
// What a parity test cannot recompute on the CPU in EXACT mode (div.approx / sqrt.approx results).
template <class A>
__global__ void __launch_bounds__(128) ray_probe_kernel(float *out, const uint32_t *pixels, uint64_t n, const __grid_constant__ lyap_cam cam,
                                                        const __grid_constant__ lyap_params prm)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    RayState st;
    const bool hit = ray_begin<A>(st, pixels[2 * i], pixels[2 * i + 1], cam, prm);
    float *o = out + 12 * i;
    o[0] = hit ? 1.0f : 0.0f;
    if (!hit) { for (int k = 1; k < 12; ++k) o[k] = 0.0f; return; }
    o[1] = st.Vx; o[2] = st.Vy; o[3] = st.Vz; o[4] = st.t; o[5] = st.t1; o[6] = st.Fdt; o[7] = st.Ndt;
    o[8] = st.Px; o[9] = st.Py; o[10] = st.Pz; o[11] = A::mul(st.Fdt, prm.gradient);
}

template <class A>
__global__ void __launch_bounds__(128) normalize_kernel(float *xyz, uint64_t n)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    normalize3<A>(x, y, z);
    xyz[3 * i] = x; xyz[3 * i + 1] = y; xyz[3 * i + 2] = z;
}

cudaError_t launch_ray_probe(int mode, float *out, const uint32_t *pixels, uint64_t n, const lyap_cam &cam, const lyap_params &prm, cudaStream_t s)
{
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (mode == kHost) ray_probe_kernel<ArithHost><<<grid, 128, 0, s>>>(out, pixels, n, cam, prm);
    else ray_probe_kernel<ArithDev><<<grid, 128, 0, s>>>(out, pixels, n, cam, prm);
    return cudaGetLastError();
}

cudaError_t launch_normalize(int mode, float *xyz, uint64_t n, cudaStream_t s)
{
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (mode == kHost) normalize_kernel<ArithHost><<<grid, 128, 0, s>>>(xyz, n);
    else normalize_kernel<ArithDev><<<grid, 128, 0, s>>>(xyz, n);
    return cudaGetLastError();
}

// ---- roofline probes ------------------------------------------------------------
// Register-only loops: 8 independent chains per thread so the pipes, not latency,
// set the pace.  The kernels also report elapsed SM cycles so the caller can turn
// the event time into the clock the SMs actually ran at.
__global__ void __launch_bounds__(256) probe_ffma_kernel(float *sink, float b, float c, int iters, long long *cycles)
{
    float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
    float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    const long long t0 = clock64();
    if (t0 == 0x7fffffffffffffffLL) x0 = 0.f;   // order the loop after the first clock read
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = __fmaf_rn(x0, b, c); x1 = __fmaf_rn(x1, b, c); x2 = __fmaf_rn(x2, b, c); x3 = __fmaf_rn(x3, b, c);
            x4 = __fmaf_rn(x4, b, c); x5 = __fmaf_rn(x5, b, c); x6 = __fmaf_rn(x6, b, c); x7 = __fmaf_rn(x7, b, c);
        }
    }
    const float sum = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    sink[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    long long t1 = clock64();
    if (sum == 123.456f) t1 = 0;                // and the second read after the results exist
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

__global__ void __launch_bounds__(256) probe_mufu_kernel(float *sink, float b, float c, int iters, long long *cycles)
{
    float x0 = 1.5f + threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
    float x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    const long long t0 = clock64();
    if (t0 == 0x7fffffffffffffffLL) x0 = 0.f;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            // one MUFU.LG2 + one FFMA per link, the mix of the exact-mode exponent step
            x0 = __fmaf_rn(ArithDev::lg2(x0), b, c); x1 = __fmaf_rn(ArithDev::lg2(x1), b, c);
            x2 = __fmaf_rn(ArithDev::lg2(x2), b, c); x3 = __fmaf_rn(ArithDev::lg2(x3), b, c);
            x4 = __fmaf_rn(ArithDev::lg2(x4), b, c); x5 = __fmaf_rn(ArithDev::lg2(x5), b, c);
            x6 = __fmaf_rn(ArithDev::lg2(x6), b, c); x7 = __fmaf_rn(ArithDev::lg2(x7), b, c);
        }
    }
    const float sum = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    sink[blockIdx.x * blockDim.x + threadIdx.x] = sum;
    long long t1 = clock64();
    if (sum == 123.456f) t1 = 0;                // and the second read after the results exist
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int VARIANT>
__global__ void __launch_bounds__(256) probe_packed_kernel(float *sink, float b, float c, int iters)
{
    f32x2 x0 = pack2(threadIdx.x * 1e-3f, 1.f), x1 = pack2(2.f, 3.f), x2 = pack2(4.f, 5.f), x3 = pack2(6.f, 7.f);
    f32x2 x4 = pack2(8.f, 9.f), x5 = pack2(10.f, 11.f), x6 = pack2(12.f, 13.f), x7 = pack2(14.f, 15.f);
    const f32x2 bb = pack2(b, b), cc = pack2(c, c), two = pack2(2.f, 2.f), one = pack2(1.f, 1.f);
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if constexpr (VARIANT == 1) {
                x0 = mul2(x0, bb); x1 = mul2(x1, bb); x2 = mul2(x2, bb); x3 = mul2(x3, bb);
                x4 = mul2(x4, bb); x5 = mul2(x5, bb); x6 = mul2(x6, bb); x7 = mul2(x7, bb);
            } else {
                // four independent (w, prod) pairs doing the fast exponent step: 2 mul2 + 2 fma2 each
                f32x2 p;
                p = mul2(bb, x0); x0 = fma2(p, x0, p); x1 = mul2(x1, fma2(two, x0, one));
                p = mul2(cc, x2); x2 = fma2(p, x2, p); x3 = mul2(x3, fma2(two, x2, one));
                p = mul2(bb, x4); x4 = fma2(p, x4, p); x5 = mul2(x5, fma2(two, x4, one));
                p = mul2(cc, x6); x6 = fma2(p, x6, p); x7 = mul2(x7, fma2(two, x6, one));
            }
        }
    }
    float lo, hi, s = 0.f;
    unpack2(x0, lo, hi); s += lo + hi; unpack2(x1, lo, hi); s += lo + hi; unpack2(x2, lo, hi); s += lo + hi;
    unpack2(x3, lo, hi); s += lo + hi; unpack2(x4, lo, hi); s += lo + hi; unpack2(x5, lo, hi); s += lo + hi;
    unpack2(x6, lo, hi); s += lo + hi; unpack2(x7, lo, hi); s += lo + hi;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) probe_ffma2_kernel(float *sink, float b, float c, int iters)
{
    f32x2 x0 = pack2(threadIdx.x * 1e-3f, 1.f), x1 = pack2(2.f, 3.f), x2 = pack2(4.f, 5.f), x3 = pack2(6.f, 7.f);
    f32x2 x4 = pack2(8.f, 9.f), x5 = pack2(10.f, 11.f), x6 = pack2(12.f, 13.f), x7 = pack2(14.f, 15.f);
    const f32x2 bb = pack2(b, b), cc = pack2(c, c);
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma2(x0, bb, cc); x1 = fma2(x1, bb, cc); x2 = fma2(x2, bb, cc); x3 = fma2(x3, bb, cc);
            x4 = fma2(x4, bb, cc); x5 = fma2(x5, bb, cc); x6 = fma2(x6, bb, cc); x7 = fma2(x7, bb, cc);
        }
    }
    float lo, hi, s = 0.f;
    unpack2(x0, lo, hi); s += lo + hi; unpack2(x1, lo, hi); s += lo + hi; unpack2(x2, lo, hi); s += lo + hi;
    unpack2(x3, lo, hi); s += lo + hi; unpack2(x4, lo, hi); s += lo + hi; unpack2(x5, lo, hi); s += lo + hi;
    unpack2(x6, lo, hi); s += lo + hi; unpack2(x7, lo, hi); s += lo + hi;
    sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Packed-FMA rate in scalar-lane-operations per second (two per packed lane-op).
cudaError_t probe_ffma2(double *lane_ops)
{
    int dev = 0, n_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = n_sm * 8, threads = 256, iters = 2048;
    float *sink = nullptr;
    if ((e = cudaMalloc(&sink, sizeof(float) * blocks * threads)) != cudaSuccess) return e;
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        float ms = 0;
        cudaEventRecord(ev0);
        probe_ffma2_kernel<<<blocks, threads>>>(sink, 0.999f, 0.001f, iters);
        cudaEventRecord(ev1);
        if ((e = cudaEventSynchronize(ev1)) != cudaSuccess) break;
        cudaEventElapsedTime(&ms, ev0, ev1);
        const double ops = (double)blocks * threads * iters * 64.0 * 2.0 / (ms * 1e-3);
        if (rep && ops > best) best = ops;
        if (rep == 3 && getenv("LYAP_PROBE_VERBOSE")) {
            cudaEventRecord(ev0);
            probe_packed_kernel<1><<<blocks, threads>>>(sink, 0.999f, 0.5f, iters);
            cudaEventRecord(ev1);
            cudaEventSynchronize(ev1);
            cudaEventElapsedTime(&ms, ev0, ev1);
            printf("probe mul2 chains: %.2f T lane-ops/s\n", (double)blocks * threads * iters * 64.0 * 2.0 / (ms * 1e-3) / 1e12);
            cudaEventRecord(ev0);
            probe_packed_kernel<2><<<blocks, threads>>>(sink, 3.7f, 3.2f, iters);
            cudaEventRecord(ev1);
            cudaEventSynchronize(ev1);
            cudaEventElapsedTime(&ms, ev0, ev1);
            printf("probe fast-step mix (2 mul2 + 2 fma2, 4 chains): %.2f T lane-ops/s\n", (double)blocks * threads * iters * 8 * 16.0 * 2.0 / (ms * 1e-3) / 1e12);
        }
    }
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    cudaFree(sink);
    *lane_ops = best;
    return e;
}

cudaError_t probe_peaks(double *ffma_ops, double *mufu_ops, double *clock_hz, int *sms)
{
    int dev = 0, n_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = n_sm * 8, threads = 256;
    float *sink = nullptr;
    long long *cyc = nullptr;
    if ((e = cudaMalloc(&sink, sizeof(float) * blocks * threads)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&cyc, sizeof(long long) * blocks)) != cudaSuccess) { cudaFree(sink); return e; }
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    double best_f = 0, best_m = 0, clk = 0;
    const int it_f = 4096, it_m = 1024;
    for (int rep = 0; rep < 4; ++rep) {
        float ms = 0;
        cudaEventRecord(ev0);
        probe_ffma_kernel<<<blocks, threads>>>(sink, 0.999f, 0.001f, it_f, cyc);
        cudaEventRecord(ev1);
        if ((e = cudaEventSynchronize(ev1)) != cudaSuccess) break;
        cudaEventElapsedTime(&ms, ev0, ev1);
        const double ops = (double)blocks * threads * it_f * 64.0 / (ms * 1e-3);
        if (rep && ops > best_f) {
            best_f = ops;
            long long c0 = 0;
            cudaMemcpy(&c0, cyc, sizeof c0, cudaMemcpyDeviceToHost);
            // a block's share of the run: blocks are all resident at once (8 per SM)
            clk = (double)c0 / (ms * 1e-3);
        }
        cudaEventRecord(ev0);
        probe_mufu_kernel<<<blocks, threads>>>(sink, 0.37f, 2.5f, it_m, cyc);
        cudaEventRecord(ev1);
        if ((e = cudaEventSynchronize(ev1)) != cudaSuccess) break;
        cudaEventElapsedTime(&ms, ev0, ev1);
        const double mops = (double)blocks * threads * it_m * 32.0 / (ms * 1e-3);
        if (rep && mops > best_m) best_m = mops;
    }
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    cudaFree(sink);
    cudaFree(cyc);
    if (e != cudaSuccess) return e;
    *ffma_ops = best_f;
    *mufu_ops = best_m;
    *clock_hz = clk;
    *sms = n_sm;
    return cudaSuccess;
}

} // namespace lyap
