// ray.cuh -- per-pixel ray logic around the exponent: cube clip, march, refine,
// normal, shade (reference kernel.cu:25-106 `shade`, :156-486 `raymarch`).
//
// The reference runs one thread per pixel, start to finish.  Here a ray is a small
// state machine that is advanced by ONE exponent evaluation at a time, so that a
// warp can keep 32 different rays -- each at its own stage -- in flight and spend
// all its time in the (uniform, fixed trip count) exponent loop.  `ray_begin`
// produces the first sample point, `ray_advance` consumes one exponent and either
// produces the next sample point or finishes the pixel.
//
// Every floating-point operation goes through the arithmetic profile A (arith.cuh):
//   A = ArithDev  reproduces, op for op, the PTX nvcc emits for the reference with
//                 its own flags (checked against `nvcc --use_fast_math -ptx kernel.cu`;
//                 the contraction pattern of each expression is noted inline);
//   A = ArithHost reproduces the reference's host build (checked against the oracle).
#pragma once
#include "arith.cuh"
#include "exponent.cuh"
#include "lyap/types.h"

namespace lyap {

enum RayPhase : int {
    kNeedRay = 0,   // lane is idle: wants a new pixel
    kMarchFirst,    // waiting for the exponent at the entry point t0        (kernel.cu:318)
    kMarch,         // waiting for the exponent of a march sample             (:366)
    kRefine,        // waiting for the exponent of a refinement sample        (:427)
    kNormal0,       // waiting for the k-th finite-difference sample, k=0..5  (:467)
    kNormal1, kNormal2, kNormal3, kNormal4, kNormal5,
};

struct RayState {
    int phase;
    uint32_t out;            // output slot of the pixel in flight
    float sx, sy, sz;        // sample point the next exponent is wanted at
    float Px, Py, Pz;        // current point on the ray
    float Vx, Vy, Vz;        // ray direction, length 1/M
    float t, t1;             // ray parameter and exit
    float dt, Fdt, Ndt;      // current / far / near step
    float a, c, l;           // accumulated alpha, chaos, last on-ray exponent
    bool near;
    // refinement
    float Qdt, QVx, QVy, QVz, Qt0, Qt1, minQ;
    int sign;
    // normal
    float mag, lprev, Nx, Ny, Nz;
};

// The march-only part of a ray (hybrid mode's first kernel): the sample point of a march step is
// the point on the ray itself, and nothing of the refinement or the normal is needed yet.
struct MarchState {
    int phase;               // kNeedRay, kMarchFirst, kMarch; bit kGuardBit: parked, wants the parity evaluator
    uint32_t item;           // work item (-> pixel and output slot) of the ray in flight
    float Px, Py, Pz;
    float Vx, Vy, Vz;
    float t, t1;
    float dt, Fdt, Ndt;
    float a, c, l;
    bool near;
};
constexpr int kGuardBit = 0x100;

__device__ __forceinline__ void set_sample(RayState &st) { st.sx = st.Px; st.sy = st.Py; st.sz = st.Pz; }
__device__ __forceinline__ void set_sample(MarchState &) {}

template <class A>
__device__ __forceinline__ void normalize3(float &x, float &y, float &z)
{
    // vec3.hpp:141-154; the epsilon compares are done in double in both builds
    const float m2 = A::dot3(x, y, z, x, y, z);
    const double m2d = A::f2d(m2);
    if (m2d < 1e-12) {
        x = y = z = 0.0f;
    } else if (m2 == 1.0f || (m2d > (double)1.0f - 1e-12 && m2d < (double)1.0f + 1e-12)) {
        // unit already
    } else {
        const float s = A::sqrt(m2);
        x = A::div(x, s);
        y = A::div(y, s);
        z = A::div(z, s);
    }
}

// jitter factor of kernel.cu:338-342.  Device build: the double expression
// 1.0 +- jit*jitter is demoted to one fma(frac, +-jitter, 1); host build: float
// product, double add, narrowed.
template <class A>
__device__ __forceinline__ float jitter_factor(float l, float jitter);

template <>
__device__ __forceinline__ float jitter_factor<ArithDev>(float l, float jitter)
{
    const float frac = ArithDev::sub(l, truncf(l));
    const float s = frac < 0.0f ? -jitter : jitter;
    return ArithDev::fma(frac, s, 1.0f);
}

template <>
__device__ __forceinline__ float jitter_factor<ArithHost>(float l, float jitter)
{
    const float frac = __fsub_rn(l, truncf(l));  // exact
    const float prod = __fmul_rn(frac, jitter);
    return frac < 0.0f ? __double2float_rn(1.0 - (double)prod) : __double2float_rn(1.0 + (double)prod);
}

// ---------------------------------------------------------------------- shade
template <class A>
__device__ __forceinline__ float clamp01(float v)
{
    // Vec::clamp (vec3.hpp:190-193): that<0 ? 0 : that>1 ? 1 : that  (NaN passes through)
    return v < 0.0f ? 0.0f : (v >= 1.0f ? 1.0f : v);
}

template <class A>
__device__ __forceinline__ float shade_pow(float x, float y);
template <>
__device__ __forceinline__ float shade_pow<ArithDev>(float x, float y) { return ArithDev::pow(x, y); }
template <>
__device__ __forceinline__ float shade_pow<ArithHost>(float x, float y) { return powf(x, y); }

template <class A>
__device__ __forceinline__ float shade_log(float x);
template <>
__device__ __forceinline__ float shade_log<ArithDev>(float x) { return ArithDev::log(x); }
template <>
__device__ __forceinline__ float shade_log<ArithHost>(float x)
{
    return x < 0.0f ? quiet_nan() : glibc_logf_lane(x);
}

// kernel.cu:25-106 followed by Color::to_rgba (color.hpp:169-175)
template <class A>
__device__ __forceinline__ uint32_t shade_pixel(const lyap_point &pt, const lyap_cam &cam,
                                                const lyap_light *__restrict__ lights, uint32_t n_lights)
{
    if (isnan(pt.a) || n_lights == 0) return 0u;  // Color() with x = w = 0  (:29-33)

    float cr = 0.0f, cg = 0.0f, cb = 0.0f, ca = 0.0f;
    const float camVx = A::sub(cam.C.x, pt.P.x), camVy = A::sub(cam.C.y, pt.P.y), camVz = A::sub(cam.C.z, pt.P.z);

    for (uint32_t k = 0; k < n_lights; ++k) {
        const lyap_light &L = lights[k];
        float Lx = A::sub(L.C.x, pt.P.x), Ly = A::sub(L.C.y, pt.P.y), Lz = A::sub(L.C.z, pt.P.z);
        const float d2 = A::dot3(Lx, Ly, Lz, Lx, Ly, Lz);
        normalize3<A>(Lx, Ly, Lz);
        float i = A::dot3(pt.N.x, pt.N.y, pt.N.z, Lx, Ly, Lz);     // dev: N.y*L.y first, then fma x, fma z
        const float j = -A::dot3(Lx, Ly, Lz, L.V.x, L.V.y, L.V.z);

        float pr, pg, pb, pa;
        if (j > L.lightOuterCone) {
            i = clamp01<A>(i);
            const float kd = A::mul(i, L.diffusePower);
            float Hx = A::add(camVx, Lx), Hy = A::add(camVy, Ly), Hz = A::add(camVz, Lz);  // camV is NOT normalised (:49,74)
            normalize3<A>(Hx, Hy, Hz);
            float s = clamp01<A>(A::dot3(pt.N.x, pt.N.y, pt.N.z, Hx, Hy, Hz));
            s = shade_pow<A>(s, L.specularHardness);
            const float ks = A::mul(s, L.specularPower);
            // (specular + diffuse): dev contracts to fma(specularColor, ks, diffuseColor*kd)
            pr = A::madd(L.specularColor.r, ks, A::mul(L.diffuseColor.r, kd));
            pg = A::madd(L.specularColor.g, ks, A::mul(L.diffuseColor.g, kd));
            pb = A::madd(L.specularColor.b, ks, A::mul(L.diffuseColor.b, kd));
            pa = A::madd(L.specularColor.a, ks, A::mul(L.diffuseColor.a, kd));
            const float fall = A::div(L.lightRange, d2);
            pr = A::mul(pr, fall); pg = A::mul(pg, fall); pb = A::mul(pb, fall); pa = A::mul(pa, fall);
            if (j < L.lightInnerCone) {
                const float cone = A::div(A::sub(j, L.lightOuterCone), A::sub(L.lightInnerCone, L.lightOuterCone));
                pr = A::mul(pr, cone); pg = A::mul(pg, cone); pb = A::mul(pb, cone); pa = A::mul(pa, cone);
            }
            pr = A::add(pr, L.ambient.r); pg = A::add(pg, L.ambient.g); pb = A::add(pb, L.ambient.b); pa = A::add(pa, L.ambient.a);
        } else {
            pr = L.ambient.r; pg = L.ambient.g; pb = L.ambient.b; pa = L.ambient.a;
        }

        if (pt.c > 0.0f) {
            // chaosColor * (0.1125 / log(c)): the division is in double in both builds (:98)
            const float tint = A::d2f(0.1125 / A::f2d(shade_log<A>(pt.c)));
            pr = A::madd(L.chaosColor.r, tint, pr);
            pg = A::madd(L.chaosColor.g, tint, pg);
            pb = A::madd(L.chaosColor.b, tint, pb);
            pa = A::madd(L.chaosColor.a, tint, pa);
        }
        cr = A::add(cr, pr); cg = A::add(cg, pg); cb = A::add(cb, pb); ca = A::add(ca, pa);
    }
    return (uint32_t)A::to_byte(cr) | ((uint32_t)A::to_byte(cg) << 8) | ((uint32_t)A::to_byte(cb) << 16) |
           ((uint32_t)A::to_byte(ca) << 24);
}

// ------------------------------------------------------------------ ray setup
// kernel.cu:160-318 up to (not including) the first exponent.  Returns false when
// the ray misses the cube (the reference's `return 1`, :262-264).
template <class A>
__device__ __forceinline__ void ray_direction(float &Vx, float &Vy, float &Vz, uint32_t px, uint32_t py, const lyap_cam &cam)
{
    const float fx = __uint2float_rn(px), fy = __uint2float_rn(py);
    // V = S0 + SDX*sx + SDY*sy  (dev: two fma per component)
    Vx = A::madd(cam.SDY.x, fy, A::madd(cam.SDX.x, fx, cam.S0.x));
    Vy = A::madd(cam.SDY.y, fy, A::madd(cam.SDX.y, fx, cam.S0.y));
    Vz = A::madd(cam.SDY.z, fy, A::madd(cam.SDX.z, fx, cam.S0.z));
    normalize3<A>(Vx, Vy, Vz);
    Vx = A::div(Vx, cam.M);
    Vy = A::div(Vy, cam.M);
    Vz = A::div(Vz, cam.M);
}

template <class A, class S>
__device__ __forceinline__ bool ray_begin(S &st, uint32_t px, uint32_t py, const lyap_cam &cam, const lyap_params &prm)
{
    float Vx, Vy, Vz;
    ray_direction<A>(Vx, Vy, Vz, px, py, cam);

    const float C[3] = {cam.C.x, cam.C.y, cam.C.z};
    const float V[3] = {Vx, Vy, Vz};
    const float inf = __int_as_float(0x7f800000);

    // :187-193 plane hits against the double literals 0.0 / 4.0
    float ts[6];
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        if (V[ax] != 0.0f) {
            const double cd = A::f2d(C[ax]), vd = A::f2d(V[ax]);
            ts[2 * ax] = A::d2f(__ddiv_rn(0.0 - cd, vd));
            ts[2 * ax + 1] = A::d2f(__ddiv_rn(4.0 - cd, vd));
        } else {
            ts[2 * ax] = inf;
            ts[2 * ax + 1] = inf;
        }
    }
    // :204-244 drop hits whose other two coordinates leave [0,4]
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (ts[k] != inf) {
            const int u = (k / 2 == 0) ? 1 : 0;
            const int w = (k / 2 == 2) ? 1 : 2;
            const float hu = A::madd(V[u], ts[k], C[u]);
            const float hw = A::madd(V[w], ts[k], C[w]);
            if (hu < 0.0f || hu > 4.0f || hw < 0.0f || hw > 4.0f) ts[k] = quiet_nan();
        }
    }
    // :249-258 nearest and farthest surviving hit
    float t0 = 3.40282347e+38f, t1 = 0.0f;
    int i0 = -1, i1 = -1;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        if (is_finite(ts[k])) {
            if (i0 == -1 || ts[k] < t0) { t0 = ts[k]; i0 = k; }
            if (i1 == -1 || ts[k] > t1) { t1 = ts[k]; i1 = k; }
        }
    }
    if (i0 == -1 && i1 == -1) return false;
    if (i1 == -1 || i0 == i1) { t1 = t0; t0 = 0.0f; }   // :268-273
    if (t0 < 0.0f) t0 = 0.0f;                           // :277-279

    st.Vx = Vx; st.Vy = Vy; st.Vz = Vz;
    st.t = t0;
    st.t1 = t1;
    st.Px = A::madd(Vx, t0, C[0]);
    st.Py = A::madd(Vy, t0, C[1]);
    st.Pz = A::madd(Vz, t0, C[2]);
    st.a = 0.0f;
    st.c = 0.0f;

    // :299-315
    float Fdt;
    if (prm.stepMethod == 1) Fdt = A::div(A::sub(t1, t0), prm.depth);
    else Fdt = A::div(A::sqrt(A::dot3(Vx, Vy, Vz, Vx, Vy, Vz)), prm.depth);
    st.Fdt = Fdt;
    st.dt = Fdt;
    st.Ndt = A::div(Fdt, prm.nearMultiplier);
    st.near = false;

    set_sample(st);
    st.phase = kMarchFirst;
    return true;
}

enum RayEvent { kContinue = 0, kHit = 1, kMiss = 2 };

// One march step (kernel.cu:334-363).  Returns kMiss when the ray leaves the cube.
template <class A, class S>
__device__ __forceinline__ RayEvent march_step(S &st, const lyap_params &prm)
{
    float step = st.dt;
    if (prm.jitter != 0.0f) {
        const float jf = jitter_factor<A>(st.l, prm.jitter);
        if (is_finite(jf)) step = A::mul(st.dt, jf);
    }
    st.Pz = A::madd(st.Vz, step, st.Pz);
    st.Py = A::madd(st.Vy, step, st.Py);
    st.Px = A::madd(st.Vx, step, st.Px);
    st.t = A::add(st.t, step);
    if (st.t > st.t1) return kMiss;
    set_sample(st);
    st.phase = kMarch;
    return kContinue;
}

// :370-384 the bookkeeping every march sample gets before the loop test: cloud accumulation and
// the near/far step switch.
template <class A, class S>
__device__ __forceinline__ void march_account(S &st, float l, const lyap_params &prm)
{
    if (l > prm.chaosThreshold) st.c = A::add(st.c, l);
    else if (l > prm.opaqueThreshold) st.a = A::add(st.a, l);
    if (l <= prm.nearThreshold && !st.near) { st.near = true; st.dt = st.Ndt; }
    else if (l > prm.nearThreshold && st.near) { st.near = false; st.dt = st.Fdt; }
}

// Hybrid mode, first kernel: consume the exponent of a march sample.  kContinue: the ray moved on
// to its next march sample; kMiss: it left the cube; kHit: the march loop ended inside the cube
// (kernel.cu:393) -- refinement and normal are left to the second kernel.
template <class A>
__device__ __forceinline__ RayEvent march_consume(MarchState &st, float l, const lyap_params &prm)
{
    if (st.phase == kMarch) march_account<A>(st, l, prm);
    st.l = l;
    if (l > prm.opaqueThreshold) return march_step<A>(st, prm);
    return st.t > st.t1 ? kMiss : kHit;
}

template <class A>
__device__ __forceinline__ void normals_begin(RayState &st, const lyap_params &prm)
{
    st.mag = A::mul(st.dt, prm.gradient);  // :456
    st.sx = A::sub(st.Px, st.mag); st.sy = st.Py; st.sz = st.Pz;
    st.phase = kNormal0;
}

// Loop test of the refinement (kernel.cu:420) and, if it holds, the move of :423-424.
template <class A>
__device__ __forceinline__ void refine_next(RayState &st, const lyap_params &prm)
{
    const bool in_window = st.t <= st.Qt1 && st.t >= st.Qt0;
    const bool big_enough = st.Qdt <= -st.minQ || st.Qdt >= st.minQ;
    if (in_window && big_enough) {
        st.t = A::add(st.t, st.Qdt);
        st.Px = A::add(st.Px, st.QVx);
        st.Py = A::add(st.Py, st.QVy);
        st.Pz = A::add(st.Pz, st.QVz);
        st.sx = st.Px; st.sy = st.Py; st.sz = st.Pz;
        st.phase = kRefine;
    } else {
        normals_begin<A>(st, prm);
    }
}

// After the march loop ends without leaving the cube (kernel.cu:393-416).
template <class A>
__device__ __forceinline__ RayEvent refine_begin(RayState &st, const lyap_params &prm)
{
    if (st.t > st.t1) return kMiss;
    st.sign = 0;
    st.Qdt = A::mul(st.dt, -0.5f);
    st.QVx = A::mul(st.Vx, st.Qdt);
    st.QVy = A::mul(st.Vy, st.Qdt);
    st.QVz = A::mul(st.Vz, st.Qdt);
    st.Qt1 = st.t;
    st.Qt0 = A::sub(st.t, st.dt);
    st.minQ = A::div(st.dt, prm.refine);
    refine_next<A>(st, prm);
    return kContinue;
}

// Hybrid mode, second kernel: pick a ray up where the march kernel left it.  The march kernel
// parks its state in the pixel's own LyapPoint slot (P = point on the ray, N = (t, dt, -), a, c, l);
// the direction is recomputed from the pixel (same operations, same bits).
template <class A>
__device__ __forceinline__ void ray_resume(RayState &st, uint32_t px, uint32_t py, const lyap_cam &cam, const lyap_params &prm,
                                           const lyap_point &rec)
{
    ray_direction<A>(st.Vx, st.Vy, st.Vz, px, py, cam);
    st.Px = rec.P.x; st.Py = rec.P.y; st.Pz = rec.P.z;
    st.t = rec.N.x;
    st.dt = rec.N.y;
    st.t1 = st.t;            // t <= t1 was established before the hand-over
    st.Fdt = st.Ndt = st.dt; // the step switch belongs to the march
    st.near = false;
    st.a = rec.a; st.c = rec.c; st.l = rec.l;
    refine_begin<A>(st, prm);
}

// Consume the exponent `l` of the pending sample.  On kHit st.{P,a,c,l} hold the
// LyapPoint fields (kernel.cu:479-483) and st.N the un-normalised central differences.
template <class A>
__device__ __forceinline__ RayEvent ray_advance(RayState &st, float l, const lyap_params &prm)
{
    switch (st.phase) {
    case kMarch:
        march_account<A>(st, l, prm);
        // fall through to the loop test
    case kMarchFirst:
        st.l = l;
        if (l > prm.opaqueThreshold) return march_step<A>(st, prm);   // :326 (NaN ends the march)
        return refine_begin<A>(st, prm);
    case kRefine:
        st.l = l;
        if (l == prm.opaqueThreshold) { normals_begin<A>(st, prm); return kContinue; }   // :430
        {
            const int s = (l < prm.opaqueThreshold) ? 0 : 1;   // :434 (NaN -> 1)
            if (s != st.sign) {
                st.Qdt = A::mul(st.Qdt, -0.5f);
                st.QVx = A::mul(st.QVx, -0.5f);
                st.QVy = A::mul(st.QVy, -0.5f);
                st.QVz = A::mul(st.QVz, -0.5f);
            }
            st.sign = s;
        }
        refine_next<A>(st, prm);
        return kContinue;
    case kNormal0:
        st.lprev = l;
        st.sx = A::add(st.Px, st.mag);
        st.phase = kNormal1;
        return kContinue;
    case kNormal1:
        st.Nx = A::sub(l, st.lprev);
        st.sx = st.Px; st.sy = A::sub(st.Py, st.mag);
        st.phase = kNormal2;
        return kContinue;
    case kNormal2:
        st.lprev = l;
        st.sy = A::add(st.Py, st.mag);
        st.phase = kNormal3;
        return kContinue;
    case kNormal3:
        st.Ny = A::sub(l, st.lprev);
        st.sy = st.Py; st.sz = A::sub(st.Pz, st.mag);
        st.phase = kNormal4;
        return kContinue;
    case kNormal4:
        st.lprev = l;
        st.sz = A::add(st.Pz, st.mag);
        st.phase = kNormal5;
        return kContinue;
    case kNormal5:
        st.Nz = A::sub(l, st.lprev);
        return kHit;   // caller normalises N (:473)
    default:
        return kContinue;
    }
}

} // namespace lyap
