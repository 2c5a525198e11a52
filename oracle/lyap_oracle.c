/*
 * oracle/lyap_oracle.c -- TEST INFRASTRUCTURE ONLY (see lyap_oracle.h).
 *
 * CPU restatement of the reference hot path in plain C, host-build arithmetic.
 * Each function cites the reference lines it follows (paths under
 * /root/reference).  Must be compiled with -ffp-contract=off so that no
 * multiply-add is fused: the reference host build the fixtures came from
 * rounds every float operation separately.
 *
 * Parity: PINNED against oracle/_ref/libref_host.so and tests/golden/.
 */
#include "lyap_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef lyap_vec3 v3;
typedef lyap_quat q4;

/* ------------------------------------------------------------------ vec3 */
/* vec3.hpp:30-124: component-wise float ops, sums evaluated left to right. */
static inline v3 v3_of(float x, float y, float z) { v3 r; r.x = x; r.y = y; r.z = z; return r; }
static inline v3 v3_add(v3 a, v3 b) { return v3_of(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 v3_sub(v3 a, v3 b) { return v3_of(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 v3_mul(v3 a, float s) { return v3_of(a.x * s, a.y * s, a.z * s); }
static inline v3 v3_div(v3 a, float s) { return v3_of(a.x / s, a.y / s, a.z / s); }
static inline float v3_dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float v3_mag2(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
static inline v3 v3_cross(v3 a, v3 b)
{
    return v3_of(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

/* vec3.hpp:128-154: zero below 1e-12, untouched at (double-compared) unit length. */
static v3 v3_unit(v3 a)
{
    float m2 = v3_mag2(a);
    if ((double)m2 < 1e-12) return v3_of(0, 0, 0);
    if (m2 == 1.0f || ((double)m2 > ((double)1.0f - 1e-12) && (double)m2 < ((double)1.0f + 1e-12))) return a;
    return v3_div(a, sqrtf(m2));
}

/* ------------------------------------------------------------------ quat */
static inline q4 q4_of(float x, float y, float z, float w) { q4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

/* quat.hpp:148-165: "normalize" MULTIPLIES by the magnitude (reference quirk B4). */
static q4 q4_refnormalize(q4 q)
{
    float m2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    if ((double)m2 < 1e-6) return q4_of(0, 0, 0, 1.0f);
    if (m2 == 1.0f || ((double)m2 > ((double)1.0f - 1e-6) && (double)m2 < ((double)1.0f + 1e-6))) return q;
    float m = sqrtf(m2);
    return q4_of(q.x * m, q.y * m, q.z * m, q.w * m);
}

/* quat.hpp:227-238: axis/angle constructor. */
static q4 q4_axis_angle(v3 axis, float ang, int degrees)
{
    ang = degrees ? (float)((double)ang * 3.14159265358979323846264338327950288 / (double)360.0f) : (ang * 0.5f);
    float s = sinf(ang);
    return q4_refnormalize(q4_of(axis.x * s, axis.y * s, axis.z * s, cosf(ang)));
}

/* quat.hpp:194-220: rotation taking p to q, scaled. */
static q4 q4_between(v3 p, v3 q, float scale)
{
    float cosa = v3_dot(p, q);
    if ((double)cosa < -1.0) cosa = -1.0f;
    else if ((double)cosa > 1.0) cosa = 1.0f;
    if (cosa == 0 || ((double)cosa >= -1e-6 && (double)cosa <= 1e-6)) return q4_of(0, 0, 0, 1.0f);
    float ang = acosf(cosa);
    v3 axis = v3_cross(p, q);
    float half = (float)((double)ang * 0.5 * (double)scale);
    float k = sinf(half) / sinf(ang);
    return q4_of(axis.x * k, axis.y * k, axis.z * k, cosf(half));
}

/* quat.hpp:172-192 */
static q4 q4_nlerp(q4 a, q4 b, float t)
{
    if ((double)t == 0.0 || (double)t < 1e-6) return a;
    if (t == 1.0f || (double)t > (double)1.0f - 1e-6) return b;
    float dot = a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
    float tA = dot >= 0 ? t : -t;
    float tI = 1.0f - t;
    return q4_refnormalize(q4_of(a.x * tI + b.x * tA, a.y * tI + b.y * tA, a.z * tI + b.z * tA, a.w * tI + b.w * tA));
}

/* quat.hpp:279-284: the eight-term expansion, summed left to right in float. */
static v3 q4_rotate(q4 q, v3 v)
{
    float x = q.x, y = q.y, z = q.z, w = q.w;
    float rx = w * w * v.x + 2 * y * w * v.z - 2 * z * w * v.y + x * x * v.x + 2 * y * x * v.y + 2 * z * x * v.z - z * z * v.x - y * y * v.x;
    float ry = 2 * x * y * v.x + y * y * v.y + 2 * z * y * v.z + 2 * w * z * v.x - z * z * v.y + w * w * v.y - 2 * x * w * v.z - x * x * v.y;
    float rz = 2 * x * z * v.x + 2 * y * z * v.y + z * z * v.z - 2 * w * y * v.x - y * y * v.z + 2 * w * x * v.y - x * x * v.z + w * w * v.z;
    return v3_of(rx, ry, rz);
}

/* ----------------------------------------------------------- host scene */
/* scene.cu:69-108.  Digits append that many MORE copies of the last symbol. */
size_t oracle_convert_sequence(const char *str, int32_t *out, size_t cap)
{
    size_t n = 0;
    int last = 1;
    const unsigned char *p = (const unsigned char *)str;
    do {
        int c = *p;
        if (c >= '1' && c <= '9') {
            for (int k = 0; k < c - '0'; k++) { if (n < cap) out[n] = last; n++; }
        } else {
            switch (c) {
            case 'a': case 'A': last = 0; break;
            case 'b': case 'B': last = 1; break;
            case 'c': case 'C': last = 2; break;
            case 'd': case 'D': last = 3; break;
            default:
                fprintf(stderr, "Bad sequence letter '%c'\n", c);
                exit(1);
            }
            if (n < cap) out[n] = last;
            n++;
        }
    } while (*(++p));
    if (n < cap) out[n] = -1;
    return n + 1;
}

/* scale.pl:5-11 */
double oracle_ease_in_out_quart(double t, double b, double c, double d)
{
    t /= d / 2;
    if (t < 1) return c / 2 * t * t + b;
    t -= 1;
    return -c / 2 * (t * (t - 2) - 1) + b;
}

/* params.cu:42-57 / scale.pl:33-48.  `i` is the double the source literal would hold. */
void oracle_campath(double i, lyap_cam *cam)
{
    v3 dir = v3_of(4, 4, 4);
    v3 side = v3_unit(v3_of(-4, 4, 4));
    v3 up = v3_unit(v3_cross(side, v3_unit(dir)));
    q4 rot0 = q4_axis_angle(up, -20, 1);
    q4 rot1 = q4_axis_angle(up, 20, 1);
    q4 nrot = q4_nlerp(rot0, rot1, (float)i);
    v3 nd = v3_mul(v3_unit(q4_rotate(nrot, dir)), (float)-1.0);
    float c = (float)(4.0 - 0.9 * i);
    cam->C = v3_sub(v3_of(c, c, c), nd);
    cam->Q = q4_between(v3_of(0, 0, 1), nd, (float)1.0);
}

/* params.cu:21-114 (the `else` light branch is the live one). */
void oracle_params_init(lyap_params *prm, lyap_cam *cam, lyap_light *lights, uint32_t *n_lights,
                        char *seq_out, size_t seq_cap, uint32_t *w, uint32_t *h)
{
    memset(prm, 0, sizeof(*prm));
    memset(cam, 0, sizeof(*cam));
    memset(lights, 0, sizeof(lyap_light) * LYAP_MAX_LIGHTS);
    prm->d = (float)2.1;
    prm->settle = 18;
    prm->accum = 1008;
    prm->stepMethod = 2;
    prm->nearThreshold = (float)-1.0;
    prm->nearMultiplier = (float)2.0;
    prm->opaqueThreshold = (float)-0.75;
    prm->chaosThreshold = (float)-0.5;
    prm->depth = 4096;
    prm->jitter = (float)0.5;
    prm->refine = 32;
    prm->gradient = (float)0.01;
    prm->lMin = (float)0.0;
    prm->lMax = (float)4.0;
    snprintf(seq_out, seq_cap, "%s", "BCABA");

    oracle_campath(1.0, cam);
    cam->M = (float)0.45;

    lyap_light *L = &lights[0];
    L->C = v3_of(6.0f, 5.0f, 3.0f);
    L->Q = q4_of(0.710595f, 0.282082f, -0.512168f, 0.391368f);
    L->M = (float)0.500000;
    L->lightInnerCone = 0.904535f;
    L->lightOuterCone = 0.816497f;
    L->lightRange = (float)1.0;
    L->ambient = (lyap_color){(float)0.1, 0, 0, 0};
    L->diffuseColor = (lyap_color){(float)1.0, (float)0.25, (float)0.125, 1};
    L->diffusePower = (float)10.0;
    L->specularColor = (lyap_color){(float)1.0, (float)1.0, (float)1.0, 1};
    L->specularPower = (float)10.0;
    L->specularHardness = (float)10.0;
    L->chaosColor = (lyap_color){0, 0, 0, 0};
    *n_lights = 1;
    *w = 3840;
    *h = 2160;
}

/* scene.cu:20-29 */
void oracle_lights_recalculate(lyap_light *lights, size_t n)
{
    for (size_t k = 0; k < n; k++) {
        lyap_light *L = &lights[k];
        L->V = v3_unit(q4_rotate(L->Q, v3_of(0, 0, 1)));
        L->lightInnerCone = v3_dot(L->V, v3_unit(q4_rotate(L->Q, v3_of(-L->M, -L->M, (float)1.5))));
        L->lightOuterCone = v3_dot(L->V, v3_unit(q4_rotate(L->Q, v3_of(-L->M, -L->M, 1))));
    }
}

/* scene.cu:31-63 */
void oracle_cam_recalculate(lyap_cam *c, uint32_t tw, uint32_t th, uint32_t td)
{
    if ((double)c->M < 1e-6) c->M = (float)1e-6;
    c->Q = q4_refnormalize(c->Q);
    if (td > 0) c->renderDenominator = td;
    if (tw > 0) { c->textureWidth = tw; c->renderWidth = c->textureWidth / c->renderDenominator; }
    if (th > 0) { c->textureHeight = th; c->renderHeight = c->textureHeight / c->renderDenominator; }
    c->V = v3_unit(q4_rotate(c->Q, v3_of(0, 0, 1)));
    c->S0 = q4_rotate(c->Q, v3_of(-c->M, -c->M, 1));
    c->lightInnerCone = v3_dot(c->V, v3_unit(q4_rotate(c->Q, v3_of(-c->M, -c->M, (float)1.5))));
    c->lightOuterCone = v3_dot(c->V, v3_unit(q4_rotate(c->Q, v3_of(-c->M, -c->M, 1))));
    c->SDX = q4_rotate(c->Q, v3_of(2 * c->M / (float)c->renderWidth, 0, 0));
    c->SDY = q4_rotate(c->Q, v3_of(0, 2 * c->M / (float)c->renderHeight, 0));
}

/* ------------------------------------------------------------- exponent */
/* kernel.cu:108-154.  The double literals 1.0 and 2.0 promote the marked
 * sub-expressions to double; everything else is float. */
static float lyap4d(v3 P, float d, uint32_t settle, uint32_t accum, const int32_t *seq)
{
    const float abcd[4] = {P.x, P.y, P.z, d};
    uint32_t pos = 0;
    float v = 0.5f, l = 0.0f;

    for (uint32_t n = 0; n < settle; n++) {
        float r = abcd[seq[pos++]];
        if (seq[pos] == -1) pos = 0;
        v = (float)((double)(r * v) * (1.0 - (double)v));
    }
    double off = (double)v - 0.5;
    if (off <= -1e-8 || off >= 1e-8) {
        for (uint32_t n = 0; n < accum; n++) {
            float r = abcd[seq[pos++]];
            if (seq[pos] == -1) pos = 0;
            v = (float)((double)(r * v) * (1.0 - (double)v));
            float dv = (float)((double)r - 2.0 * (double)r * (double)v);
            if (dv < 0) dv = -dv;
            l += logf(dv);
            if (!isfinite(l)) return NAN;
        }
    }
    return l / (float)accum;
}

float oracle_lyap4d(float x, float y, float z, float d, uint32_t settle, uint32_t accum, const int32_t *seq)
{
    return lyap4d(v3_of(x, y, z), d, settle, accum, seq);
}

void oracle_lyap4d_many(const float *xyz, size_t n, float d, uint32_t settle, uint32_t accum,
                        const int32_t *seq, float *out)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < (long long)n; i++)
        out[i] = lyap4d(v3_of(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]), d, settle, accum, seq);
}

/* ------------------------------------------------------------- raymarch */
/* kernel.cu:156-486.  Returns 1 (and leaves *pt untouched) on a miss. */
int oracle_raymarch(lyap_point *pt, uint32_t sx, uint32_t sy, const lyap_cam *cam, const lyap_params *prm,
                    const int32_t *seq, uint64_t *n_calls)
{
    uint64_t calls = 0;
#define EXPONENT(p) (calls++, lyap4d((p), prm->d, prm->settle, prm->accum, seq))
#define LEAVE(code) do { if (n_calls) *n_calls += calls; return (code); } while (0)

    /* :160-163 ray through the pixel's corner, length 1/M */
    v3 V = v3_add(v3_add(cam->S0, v3_mul(cam->SDX, (float)sx)), v3_mul(cam->SDY, (float)sy));
    V = v3_unit(V);
    V = v3_div(V, cam->M);

    /* :187-244 six plane hits, computed in double against the literals 0.0 / 4.0 */
    const float Cc[3] = {cam->C.x, cam->C.y, cam->C.z};
    const float Vc[3] = {V.x, V.y, V.z};
    float ts[6];
    for (int k = 0; k < 6; k++) {
        int ax = k >> 1;
        double plane = (k & 1) ? 4.0 : 0.0;
        ts[k] = (Vc[ax] != 0.0f) ? (float)((plane - (double)Cc[ax]) / (double)Vc[ax]) : INFINITY;
    }
    for (int k = 0; k < 6; k++) {
        if (ts[k] != INFINITY) {
            v3 H = v3_add(cam->C, v3_mul(V, ts[k]));
            const float Hc[3] = {H.x, H.y, H.z};
            int u = ((k >> 1) + 1) % 3, w = ((k >> 1) + 2) % 3;
            if (u > w) { int s = u; u = w; w = s; }
            if ((double)Hc[u] < 0.0 || (double)Hc[u] > 4.0 || (double)Hc[w] < 0.0 || (double)Hc[w] > 4.0) ts[k] = NAN;
        }
    }

    /* :249-279 nearest / farthest finite hit */
    float t0 = 3.40282347e+38F, t1 = 0;
    int i0 = -1, i1 = -1;
    for (int k = 0; k < 6; k++) {
        if (isfinite(ts[k])) {
            if (i0 == -1 || ts[k] < t0) { i0 = k; t0 = ts[k]; }
            if (i1 == -1 || ts[k] > t1) { i1 = k; t1 = ts[k]; }
        }
    }
    if (i0 == -1 && i1 == -1) LEAVE(1);
    if (i1 == -1 || i0 == i1) { i1 = i0; t1 = t0; i0 = 0; t0 = 0; }
    if (t0 < 0) t0 = 0;

    float t = t0;
    v3 P = v3_add(cam->C, v3_mul(V, t));
    float a = 0, c = 0;

    /* :299-315 step sizes */
    float Fdt;
    if (prm->stepMethod == 1) Fdt = (t1 - t0) / prm->depth;
    else Fdt = sqrtf(v3_mag2(V)) / prm->depth;
    float dt = Fdt;
    float Ndt = dt / prm->nearMultiplier;
    int near = 0;

    float l = EXPONENT(P);

    /* :326-385 march through transparent space */
    while (l > prm->opaqueThreshold) {
        if ((double)prm->jitter != 0.0) {
            float jit = (float)((double)l - trunc((double)l));
            if (jit < 0) jit = (float)(1.0 - (double)(jit * prm->jitter));
            else jit = (float)(1.0 + (double)(jit * prm->jitter));
            if (isfinite(jit)) {
                t += dt * jit;
                P = v3_add(P, v3_mul(V, dt * jit));
            } else {
                t += dt;
                P = v3_add(P, v3_mul(V, dt));
            }
        } else {
            t += dt;
            P = v3_add(P, v3_mul(V, dt));
        }
        if (t > t1) LEAVE(1);

        l = EXPONENT(P);

        if (l > prm->chaosThreshold) c += l;
        else if (l > prm->opaqueThreshold) a += l;

        if (l <= prm->nearThreshold && !near) { near = 1; dt = Ndt; }
        else if (l > prm->nearThreshold && near) { near = 0; dt = Fdt; }
    }
    if (t > t1) LEAVE(1);

    /* :402-441 bisect back and forth across the threshold */
    int sign = 0, osign;
    float Qdt = dt * -0.5f;
    v3 QdV = v3_mul(V, Qdt);
    float Qt1 = t, Qt0 = t - dt;
    float min_Qdt = dt / prm->refine;
    while (t <= Qt1 && t >= Qt0 && (Qdt <= -min_Qdt || Qdt >= min_Qdt)) {
        t += Qdt;
        P = v3_add(P, QdV);
        l = EXPONENT(P);
        if (l == prm->opaqueThreshold) break;
        osign = sign;
        sign = (l < prm->opaqueThreshold) ? 0 : 1;
        if (sign != osign) { Qdt *= -0.5f; QdV = v3_mul(QdV, -0.5f); }
    }

    /* :456-473 central differences of the exponent */
    float mag = dt * prm->gradient;
    float ls[6];
    for (int k = 0; k < 6; k++) {
        v3 S = P;
        float *comp = (k >> 1) == 0 ? &S.x : ((k >> 1) == 1 ? &S.y : &S.z);
        if (k & 1) *comp += mag;
        else *comp -= mag;
        ls[k] = EXPONENT(S);
    }
    v3 N = v3_unit(v3_of(ls[1] - ls[0], ls[3] - ls[2], ls[5] - ls[4]));

    pt->P = P;
    pt->N = N;
    pt->a = a;
    pt->c = c;
    pt->l = l;
    LEAVE(0);
#undef EXPONENT
#undef LEAVE
}

/* ---------------------------------------------------------------- shade */
/* kernel.cu:25-106 */
void oracle_shade(const lyap_point *pt, const lyap_cam *cam, const lyap_light *lights, uint32_t n, float *out)
{
    float col[4] = {0, 0, 0, 0};
    if (isnan(pt->a)) { out[0] = out[1] = out[2] = out[3] = 0; return; }

    for (uint32_t k = 0; k < n; k++) {
        const lyap_light *L = &lights[k];
        float ph[4];
        v3 camV = v3_sub(cam->C, pt->P);
        v3 lightV = v3_sub(L->C, pt->P);
        float d2 = v3_mag2(lightV);
        lightV = v3_unit(lightV);
        float i = v3_dot(lightV, pt->N);
        float j = -v3_dot(lightV, L->V);

        if (j > L->lightOuterCone) {
            i = (double)i < 0.0 ? 0.0f : ((double)i > 1.0 ? 1.0f : i);
            float kd = i * L->diffusePower;
            v3 halfV = v3_unit(v3_add(camV, lightV));
            float s = v3_dot(pt->N, halfV);
            s = (double)s < 0.0 ? 0.0f : ((double)s > 1.0 ? 1.0f : s);
            s = powf(s, L->specularHardness);
            float ks = s * L->specularPower;
            float fall = L->lightRange / d2;
            const float *dc = &L->diffuseColor.r, *sc = &L->specularColor.r, *am = &L->ambient.r;
            for (int q = 0; q < 4; q++) ph[q] = (sc[q] * ks + dc[q] * kd) * fall;
            if (j < L->lightInnerCone) {
                float cone = (j - L->lightOuterCone) / (L->lightInnerCone - L->lightOuterCone);
                for (int q = 0; q < 4; q++) ph[q] *= cone;
            }
            for (int q = 0; q < 4; q++) ph[q] += am[q];
        } else {
            const float *am = &L->ambient.r;
            for (int q = 0; q < 4; q++) ph[q] = am[q];
        }

        if ((double)pt->c > 0.0) {
            float tint = (float)(0.1125 / (double)logf(pt->c));
            const float *cc = &L->chaosColor.r;
            for (int q = 0; q < 4; q++) ph[q] += cc[q] * tint;
        }
        for (int q = 0; q < 4; q++) col[q] += ph[q];
    }
    for (int q = 0; q < 4; q++) out[q] = col[q];
}

/* color.hpp:169-175: no clamp; the x86 conversion goes through a 32-bit
 * truncation, so out-of-range channels wrap modulo 256. */
void oracle_to_rgba(const float *c, uint8_t *out)
{
    for (int q = 0; q < 4; q++) out[q] = (unsigned char)(255.0 * (double)c[q]);
}

/* kernel.cu:500-516 for one pixel. */
static uint64_t render_pixel(lyap_rgba *rgba, lyap_point *points, const lyap_cam *cam, const lyap_params *prm,
                             const int32_t *seq, const lyap_light *lights, uint32_t n_lights, uint32_t w, size_t idx)
{
    uint64_t calls = 0;
    float col[4];
    oracle_raymarch(&points[idx], (uint32_t)(idx % w), (uint32_t)(idx / w), cam, prm, seq, &calls);
    oracle_shade(&points[idx], cam, lights, n_lights, col);
    oracle_to_rgba(col, &rgba[idx].r);
    return calls;
}

uint64_t oracle_render_rows(lyap_rgba *rgba, lyap_point *points, const lyap_cam *cam, const lyap_params *prm,
                            const int32_t *seq, const lyap_light *lights, uint32_t n_lights,
                            uint32_t w, uint32_t h, uint32_t y0, uint32_t y1)
{
    (void)h;
    uint64_t calls = 0;
    const long long first = (long long)y0 * w, last = (long long)y1 * w;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : calls)
    for (long long i = first; i < last; i++)
        calls += render_pixel(rgba, points, cam, prm, seq, lights, n_lights, w, (size_t)i);
    return calls;
}

uint64_t oracle_render_pixels(lyap_rgba *rgba, lyap_point *points, const lyap_cam *cam, const lyap_params *prm,
                              const int32_t *seq, const lyap_light *lights, uint32_t n_lights,
                              uint32_t w, uint32_t h, const uint32_t *pix, size_t n_pix)
{
    (void)h;
    uint64_t calls = 0;
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : calls)
    for (long long i = 0; i < (long long)n_pix; i++)
        calls += render_pixel(rgba, points, cam, prm, seq, lights, n_lights, w, pix[i]);
    return calls;
}

/* kernel.cu:518-532 with lyap_calculate.cu:19-24's geometry generalised to nx*ny*nz. */
void oracle_bake_slab(float *exps, const lyap_params *prm, const int32_t *seq,
                      uint32_t nx, uint32_t ny, uint32_t nz, uint32_t z0, uint32_t z1)
{
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (long long z = z0; z < (long long)z1; z++) {
        for (long long y = 0; y < (long long)ny; y++) {
            float b = 4.0f * (float)y / (float)ny;
            float c = 4.0f * (float)z / (float)nz;
            float *row = exps + ((size_t)z * ny + (size_t)y) * nx;
            for (uint32_t x = 0; x < nx; x++) {
                float a = 4.0f * (float)x / (float)nx;
                row[x] = lyap4d(v3_of(a, b, c), prm->d, prm->settle, prm->accum, seq);
            }
        }
    }
}

/* torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants every host core. */
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
