// oracle/ref_cuda_launcher.cu -- TEST INFRASTRUCTURE, not product code.
//
// Headless launcher for the UNMODIFIED reference kernels (kernel.cu, compiled in
// place from /root/reference with the reference's own flags, GNUmakefile:8).  It does
// what lyap_interactive.cu:111-136,711 and lyap_calculate.cu:58-91 do around the two
// launches, minus GLUT/GL: upload sequence and lights, launch
// <<<(W/16,H/16),(16,16)>>> resp. <<<(N/8)^3,(8,8,8)>>>, copy the result back.
//
// The reference kernels have no bounds checks and derive the row stride from the
// launch grid, so frames are rendered into buffers padded up to multiples of 16 and
// the visible part is copied out.  Its point buffer is zero-filled first (the
// reference leaves it uninitialised; miss pixels shade whatever is there).
//
// Output: oracle/_ref/libref_cuda.so -- the image-parity oracle for LYAP_MODE_EXACT
// and the "reference CUDA kernel on 1 B200" timing arm.  Needs a GPU to run.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>

#include "kernel.hpp"

#define CK(x)                                   \
    do {                                        \
        cudaError_t e_ = (x);                   \
        if (e_ != cudaSuccess) { rc = (int)e_; goto done; } \
    } while (0)

extern "C" {

int ref_cuda_render(void *h_rgba, void *h_points, const void *camP, const void *prmP, const int *h_seq, size_t n_seq,
                    const void *h_lights, unsigned n_lights, unsigned w, unsigned h, int reps, float *best_ms)
{
    int rc = 0;
    const unsigned wp = (w + 15) / 16 * 16, hp = (h + 15) / 16 * 16;
    RGBA *d_rgba = 0;
    LyapPoint *d_points = 0;
    LyapLight *d_lights = 0;
    Int *d_seq = 0;
    cudaEvent_t e0 = 0, e1 = 0;
    float best = 1e30f;
    LyapCam cam = *(const LyapCam *)camP;
    LyapParams prm = *(const LyapParams *)prmP;
    CK(cudaMalloc(&d_rgba, sizeof(RGBA) * wp * hp));
    CK(cudaMalloc(&d_points, sizeof(LyapPoint) * wp * hp));
    CK(cudaMalloc(&d_lights, sizeof(LyapLight) * 16));
    CK(cudaMalloc(&d_seq, sizeof(Int) * n_seq));
    CK(cudaMemcpy(d_seq, h_seq, sizeof(Int) * n_seq, cudaMemcpyHostToDevice));
    if (n_lights) CK(cudaMemcpy(d_lights, h_lights, sizeof(LyapLight) * n_lights, cudaMemcpyHostToDevice));
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
        float ms = 0;
        CK(cudaMemset(d_points, 0, sizeof(LyapPoint) * wp * hp));
        CK(cudaMemset(d_rgba, 0, sizeof(RGBA) * wp * hp));
        CK(cudaEventRecord(e0));
        kernel_calc_render<<<dim3(wp / 16, hp / 16), dim3(16, 16)>>>(d_rgba, d_points, cam, prm, d_seq, d_lights, n_lights);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    CK(cudaMemcpy2D(h_rgba, sizeof(RGBA) * w, d_rgba, sizeof(RGBA) * wp, sizeof(RGBA) * w, h, cudaMemcpyDeviceToHost));
    if (h_points)
        CK(cudaMemcpy2D(h_points, sizeof(LyapPoint) * w, d_points, sizeof(LyapPoint) * wp, sizeof(LyapPoint) * w, h, cudaMemcpyDeviceToHost));
    if (best_ms) *best_ms = best;
done:
    cudaFree(d_rgba);
    cudaFree(d_points);
    cudaFree(d_lights);
    cudaFree(d_seq);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

// n must be a multiple of 8 (the reference's block edge, lyap_calculate.cu:22-24).
int ref_cuda_bake(float *h_exps, const void *prmP, const int *h_seq, size_t n_seq, unsigned n, int reps, float *best_ms)
{
    int rc = 0;
    float *d_exps = 0;
    Int *d_seq = 0;
    cudaEvent_t e0 = 0, e1 = 0;
    float best = 1e30f;
    LyapParams prm = *(const LyapParams *)prmP;
    const size_t bytes = sizeof(float) * n * n * n;
    if (n % 8) return -1;
    CK(cudaMalloc(&d_exps, bytes));
    CK(cudaMalloc(&d_seq, sizeof(Int) * n_seq));
    CK(cudaMemcpy(d_seq, h_seq, sizeof(Int) * n_seq, cudaMemcpyHostToDevice));
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) {
        float ms = 0;
        CK(cudaEventRecord(e0));
        kernel_calc_volume<<<dim3(n / 8, n / 8, n / 8), dim3(8, 8, 8)>>>(d_exps, prm, d_seq);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    if (h_exps) CK(cudaMemcpy(h_exps, d_exps, bytes, cudaMemcpyDeviceToHost));
    if (best_ms) *best_ms = best;
done:
    cudaFree(d_exps);
    cudaFree(d_seq);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return rc;
}

} // extern "C"
