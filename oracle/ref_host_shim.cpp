// oracle/ref_host_shim.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Builds the UNMODIFIED reference sources (kernel.cu, scene.cu, params.cu and
// the headers they pull in) as ordinary host C++ by including them, in place,
// from /root/reference (-I on the command line; nothing is copied into this
// repository).  Outside nvcc, cuda_runtime.h leaves __device__/__host__/
// __global__ empty, so the reference's device functions and even its two
// __global__ kernels become plain functions; the launch geometry built-ins
// (blockIdx & co.) are provided here as thread-local variables so the kernels
// can be "launched" one thread at a time from an OpenMP loop.
//
// Output: oracle/_ref/libref_host.so (git-ignored).  It is the ground truth the
// C restatement in oracle/lyap_oracle.c is pinned against, the generator of
// tests/golden/*, and the `cpu_baseline.kind == "reference"` arm of bench.py.
//
// Build recipe: oracle/Makefile (g++ -O2 -fopenmp -ffp-contract=off).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>

using std::isfinite;
using std::isnan;

// Launch geometry the reference kernels read (kernel.cu:502-504,520-529).
static thread_local uint3 blockIdx;
static thread_local uint3 threadIdx;
static thread_local dim3 blockDim;
static thread_local dim3 gridDim;

// 24-bit multiply intrinsic used for indexing (kernel.cu:502).
static inline unsigned int __umul24(unsigned int a, unsigned int b)
{
    return (a & 0xffffffu) * (b & 0xffffffu);
}

#include "kernel.cu"
#include "scene.cu"
#include "params.cu"

#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {

void ref_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

size_t ref_sizeof_camlight(void) { return sizeof(LyapCam); }
size_t ref_sizeof_params(void) { return sizeof(LyapParams); }
size_t ref_sizeof_point(void) { return sizeof(LyapPoint); }

// params.cu:21-114.  Copies the reference globals out.
void ref_params_init(void *prm_out, void *cam_out, void *lights_out, unsigned *n_lights_out,
                     char *seq_out, size_t seq_cap, unsigned *w_out, unsigned *h_out)
{
    memset(&prm, 0, sizeof(prm));
    memset(&cam, 0, sizeof(cam));
    memset(lights, 0, sizeof(LyapLight) * MAX_LIGHTS);
    params_init();
    memcpy(prm_out, &prm, sizeof(prm));
    memcpy(cam_out, &cam, sizeof(cam));
    memcpy(lights_out, lights, sizeof(LyapLight) * MAX_LIGHTS);
    *n_lights_out = num_lights;
    snprintf(seq_out, seq_cap, "%s", (const char *)sequence);
    *w_out = default_imageWidth;
    *h_out = default_imageHeight;
}

// scene.cu:69-108.  Returns the element count including the -1 terminator.
size_t ref_convert_sequence(const char *str, int *out, size_t cap)
{
    Int *seq = 0;
    size_t n = scene_convert_sequence(&seq, (unsigned char *)str);
    for (size_t i = 0; i < n && i < cap; i++) out[i] = seq[i];
    free(seq);
    return n;
}

void ref_cam_recalculate(void *camP, unsigned tw, unsigned th, unsigned td)
{
    scene_cam_recalculate((LyapCam *)camP, tw, th, td);
}

void ref_lights_recalculate(void *lightsP, size_t n)
{
    scene_lights_recalculate((LyapLight *)lightsP, n);
}

// The scale.pl:33-48 camera block, evaluated with the reference's own Vec/Quat.
// scale.pl pastes `$i` into the source as a decimal literal, i.e. a double: nlerp
// narrows it to float at the call, the cam.C expression keeps it in double.
void ref_campath(double i, void *camP)
{
    LyapCam *c = (LyapCam *)camP;
    Vec dir = Vec(4, 4, 4);
    Vec side = Vec(-4, 4, 4);
    side.normalize();
    Vec up = side * dir.normalized();
    up.normalize();
    Quat rot0 = Quat(up, -20, 1);
    Quat rot1 = Quat(up, 20, 1);
    Quat nrot = rot0.nlerp(rot1, i);
    Vec nd = nrot.transform(dir).normalized() * -1.0;
    c->C = Vec(4.0 - 0.9 * i, 4.0 - 0.9 * i, 4.0 - 0.9 * i) - nd;
    c->Q = Quat(Vec(0, 0, 1), nd, 1.0);
}

float ref_lyap4d(float x, float y, float z, float d, unsigned settle, unsigned accum, const int *seq)
{
    return lyap4d(Vec(x, y, z), d, settle, accum, seq);
}

void ref_lyap4d_many(const float *xyz, size_t n, float d, unsigned settle, unsigned accum,
                     const int *seq, float *out)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < (long long)n; i++)
        out[i] = lyap4d(Vec(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), d, settle, accum, seq);
}

int ref_raymarch(void *point, unsigned sx, unsigned sy, const void *camP, const void *prmP, const int *seq)
{
    return raymarch((LyapPoint *)point, sx, sy, *(const LyapCam *)camP, *(const LyapParams *)prmP, (Int *)seq);
}

void ref_shade(const void *point, const void *camP, const void *lightsP, unsigned n, float *rgba4)
{
    Color c = shade(*(const LyapPoint *)point, *(const LyapCam *)camP, (LyapLight *)lightsP, n);
    rgba4[0] = c.x; rgba4[1] = c.y; rgba4[2] = c.z; rgba4[3] = c.w;
}

void ref_to_rgba(const float *rgba4, unsigned char *out)
{
    Color c(rgba4[0], rgba4[1], rgba4[2], rgba4[3]);
    c.to_rgba(out);
}

// kernel_calc_render (kernel.cu:500-516), one "thread" per pixel, rows y0..y1.
// The caller owns `points` (the reference never clears it; zero-fill it to get
// defined miss pixels) and passes an already recalculated camera and lights.
void ref_render_rows(void *rgba, void *points, const void *camP, const void *prmP, const int *seq,
                     const void *lightsP, unsigned n_lights, unsigned w, unsigned h,
                     unsigned y0, unsigned y1)
{
    const LyapCam c = *(const LyapCam *)camP;
    const LyapParams p = *(const LyapParams *)prmP;
    const long long first = (long long)y0 * w, last = (long long)y1 * w;
#pragma omp parallel for schedule(dynamic, 16)
    for (long long i = first; i < last; i++) {
        blockDim = dim3(1, 1, 1);
        gridDim = dim3(w, h, 1);
        threadIdx = make_uint3(0, 0, 0);
        blockIdx = make_uint3((unsigned)(i % w), (unsigned)(i / w), 0);
        kernel_calc_render((RGBA *)rgba, (LyapPoint *)points, c, p, (Int *)seq, (LyapLight *)lightsP, n_lights);
    }
}

// kernel_calc_volume (kernel.cu:518-532), one "thread" per voxel, planes z0..z1
// of an nx*ny*nz grid; exps is the FULL volume (index x + (y + z*ny)*nx).
void ref_bake_slab(float *exps, const void *prmP, const int *seq, unsigned nx, unsigned ny, unsigned nz,
                   unsigned z0, unsigned z1)
{
    const LyapParams p = *(const LyapParams *)prmP;
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (long long z = z0; z < (long long)z1; z++) {
        for (long long y = 0; y < (long long)ny; y++) {
            blockDim = dim3(1, 1, 1);
            gridDim = dim3(nx, ny, nz);
            threadIdx = make_uint3(0, 0, 0);
            for (unsigned x = 0; x < nx; x++) {
                blockIdx = make_uint3(x, (unsigned)y, (unsigned)z);
                kernel_calc_volume(exps, p, (Int *)seq);
            }
        }
    }
}

} // extern "C"
