// scene.cpp -- host scene layer of the C ABI: defaults, sequence parser, camera and
// light derived quantities, and the scale.pl camera path as a runtime function.
// Mirrors reference params.cu / scene.cu / scale.pl behaviour (cited per function).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "linalg.hpp"
#include "lyap/abi.h"

using namespace lyap_host;

extern "C" {

// scale.pl:5-11
double lyap_ease_in_out_quart(double t, double b, double c, double d)
{
    t /= d / 2;
    if (t < 1) return c / 2 * t * t + b;
    t -= 1;
    return -c / 2 * (t * (t - 2) - 1) + b;
}

// params.cu:42-57 with the literal `1` replaced by i, as scale.pl:33-48 does.  The
// script pastes i into the source as a decimal (double) literal: nlerp narrows it to
// float, the position expression stays in double.
void lyap_campath_orbit(double i, lyap_cam *cam)
{
    const V3 dir(4, 4, 4);
    const V3 side = unit(V3(-4, 4, 4));
    const V3 up = unit(cross(side, unit(dir)));
    const Q4 rot0 = from_axis_angle(up, -20, true);
    const Q4 rot1 = from_axis_angle(up, 20, true);
    const Q4 nrot = nlerp(rot0, rot1, (float)i);
    const V3 nd = unit(rotate(nrot, dir)) * (float)-1.0;
    const float c = (float)(4.0 - 0.9 * i);
    cam->C = V3(c, c, c) - nd;
    rotation_between(V3(0, 0, 1), nd, (float)1.0).store(cam->Q);
}

void lyap_campath_frame(uint32_t f, uint32_t n_frames, lyap_cam *cam)
{
    const double t = n_frames > 1 ? (double)f / (double)(n_frames - 1) : 1.0;
    double i = lyap_ease_in_out_quart(t, 0.0, 1.0, 1.0);
    // perl interpolates numbers into strings with %.15g; the compiler then reads that text
    char txt[64];
    snprintf(txt, sizeof txt, "%.15g", i);
    i = strtod(txt, nullptr);
    lyap_campath_orbit(i, cam);
}

// params.cu:21-114 (the live `else` light rig)
void lyap_params_init(lyap_params *prm, lyap_cam *cam, lyap_light *lights, uint32_t *num_lights,
                      char *sequence, size_t sequence_cap, uint32_t *image_width, uint32_t *image_height)
{
    memset(prm, 0, sizeof *prm);
    memset(cam, 0, sizeof *cam);
    memset(lights, 0, sizeof(lyap_light) * LYAP_MAX_LIGHTS);

    prm->d = (float)2.1;
    prm->settle = 18;
    prm->accum = 1008;
    prm->stepMethod = 2;
    prm->nearThreshold = -1.0f;
    prm->nearMultiplier = 2.0f;
    prm->opaqueThreshold = -0.75f;
    prm->chaosThreshold = -0.5f;
    prm->depth = 4096;
    prm->jitter = 0.5f;
    prm->refine = 32;
    prm->gradient = (float)0.01;
    prm->lMin = 0.0f;
    prm->lMax = 4.0f;
    if (sequence && sequence_cap) snprintf(sequence, sequence_cap, "%s", "BCABA");

    lyap_campath_orbit(1.0, cam);
    cam->M = (float)0.45;

    lyap_light &L = lights[0];
    L.C = lyap_vec3{6.0f, 5.0f, 3.0f};
    L.Q = lyap_quat{0.710595f, 0.282082f, -0.512168f, 0.391368f};
    L.M = 0.5f;
    L.lightInnerCone = 0.904535f;
    L.lightOuterCone = 0.816497f;
    L.lightRange = 1.0f;
    L.ambient = lyap_color{(float)0.1, 0, 0, 0};
    L.diffuseColor = lyap_color{1.0f, 0.25f, 0.125f, 1};
    L.diffusePower = 10.0f;
    L.specularColor = lyap_color{1.0f, 1.0f, 1.0f, 1};
    L.specularPower = 10.0f;
    L.specularHardness = 10.0f;
    L.chaosColor = lyap_color{0, 0, 0, 0};
    if (num_lights) *num_lights = 1;
    if (image_width) *image_width = 3840;
    if (image_height) *image_height = 2160;
}

// scene.cu:69-108.  'A'..'D' (either case) -> 0..3; a digit n appends n MORE copies of
// the last symbol (initially 'B'); terminator -1.
size_t lyap_scene_convert_sequence(int32_t **seqP, const unsigned char *seqStr)
{
    *seqP = nullptr;
    if (!seqStr || !*seqStr) return 0;
    const size_t cap = 10 * strlen((const char *)seqStr) + 1;
    int32_t *seq = (int32_t *)malloc(cap * sizeof(int32_t));
    if (!seq) return 0;
    size_t n = 0;
    int32_t last = 1;
    for (const unsigned char *p = seqStr; *p; ++p) {
        const unsigned char ch = *p;
        if (ch >= '1' && ch <= '9') {
            for (int k = ch - '0'; k > 0; --k) seq[n++] = last;
            continue;
        }
        switch (ch | 0x20) {
        case 'a': last = 0; break;
        case 'b': last = 1; break;
        case 'c': last = 2; break;
        case 'd': last = 3; break;
        default:
            fprintf(stderr, "Bad sequence letter '%c'\n", ch);
            free(seq);
            return 0;
        }
        seq[n++] = last;
    }
    seq[n++] = -1;
    *seqP = seq;
    return n;
}

// scene.cu:20-29
void lyap_scene_lights_recalculate(lyap_light *lights, size_t num_lights)
{
    for (size_t k = 0; k < num_lights; ++k) {
        lyap_light &L = lights[k];
        const Q4 q(L.Q);
        const V3 fwd = unit(rotate(q, V3(0, 0, 1)));
        L.V = fwd;
        L.lightInnerCone = dot(fwd, unit(rotate(q, V3(-L.M, -L.M, (float)1.5))));
        L.lightOuterCone = dot(fwd, unit(rotate(q, V3(-L.M, -L.M, 1))));
    }
}

// scene.cu:31-63
void lyap_scene_cam_recalculate(lyap_cam *c, uint32_t tw, uint32_t th, uint32_t td)
{
    if ((double)c->M < 1e-6) c->M = (float)1e-6;
    const Q4 q = ref_normalize(Q4(c->Q));
    q.store(c->Q);
    if (td > 0) c->renderDenominator = td;
    if (tw > 0) {
        c->textureWidth = tw;
        c->renderWidth = c->textureWidth / c->renderDenominator;
    }
    if (th > 0) {
        c->textureHeight = th;
        c->renderHeight = c->textureHeight / c->renderDenominator;
    }
    const V3 fwd = unit(rotate(q, V3(0, 0, 1)));
    c->V = fwd;
    c->S0 = rotate(q, V3(-c->M, -c->M, 1));
    c->lightInnerCone = dot(fwd, unit(rotate(q, V3(-c->M, -c->M, (float)1.5))));
    c->lightOuterCone = dot(fwd, unit(rotate(q, V3(-c->M, -c->M, 1))));
    c->SDX = rotate(q, V3(2 * c->M / (float)c->renderWidth, 0, 0));
    c->SDY = rotate(q, V3(0, 2 * c->M / (float)c->renderHeight, 0));
}

} // extern "C"
