/*
 * lyap/types.h -- plain-C mirrors of the reference's POD surface.
 *
 * Every struct here is layout-identical (size, alignment, field offsets) to
 * the type the reference passes across its host->kernel boundary, so a
 * reference-side caller can hand its own objects to this library by pointer:
 *
 *   lyap_vec3    <-> Vec   (VEC3<float,float3>)          reference vec3.hpp:13, structs.hpp:21
 *   lyap_quat    <-> Quat  (QUAT<float,float3,float4>)   reference quat.hpp:16, structs.hpp:20
 *   lyap_color   <-> Color (COLOR<float,float4>)         reference color.hpp:12, structs.hpp:22
 *   lyap_camlight<-> LyapCam == LyapLight                reference structs.hpp:26-51
 *   lyap_params  <-> LyapParams                          reference structs.hpp:53-68
 *   lyap_point   <-> LyapPoint                           reference structs.hpp:70-76
 *   lyap_rgba    <-> RGBA                                reference structs.hpp:78-80
 *
 * float4-derived reference types are 16-byte aligned (CUDA's float4), float3
 * ones are 4-byte aligned; the static asserts at the bottom pin the offsets
 * measured from the reference build.
 */
#ifndef LYAP_TYPES_H
#define LYAP_TYPES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__) || defined(__clang__) || defined(__CUDACC__)
#define LYAP_ALIGN16 __attribute__((aligned(16)))
#else
#define LYAP_ALIGN16
#endif

typedef struct lyap_vec3 {
    float x, y, z;
} lyap_vec3;

typedef struct LYAP_ALIGN16 lyap_quat {
    float x, y, z, w;
} lyap_quat;

typedef struct LYAP_ALIGN16 lyap_color {
    float r, g, b, a;
} lyap_color;

/* One struct serves as both camera and light, exactly as in the reference. */
typedef struct lyap_camlight {
    lyap_vec3 C;               /* position                                  */
    lyap_quat Q;               /* orientation                               */
    float     M;               /* screen half-extent ("magnification")      */
    lyap_vec3 V;               /* forward vector (derived)                  */
    lyap_vec3 S0;              /* screen origin (derived)                   */
    lyap_vec3 SDX;             /* screen step per pixel in x (derived)      */
    lyap_vec3 SDY;             /* screen step per pixel in y (derived)      */
    uint32_t  textureWidth;
    uint32_t  textureHeight;
    uint32_t  renderWidth;
    uint32_t  renderHeight;
    uint32_t  renderDenominator;
    float     lightInnerCone;  /* cosines, derived from Q and M             */
    float     lightOuterCone;
    float     lightRange;
    lyap_color ambient;
    lyap_color diffuseColor;
    float      diffusePower;
    lyap_color specularColor;
    float      specularPower;
    float      specularHardness;
    lyap_color chaosColor;
} lyap_camlight;

typedef lyap_camlight lyap_cam;
typedef lyap_camlight lyap_light;

typedef struct lyap_params {
    float    d;                /* fourth ("D") coordinate                   */
    uint32_t settle;           /* iterations run before accumulating        */
    uint32_t accum;            /* iterations accumulated into the exponent  */
    uint32_t stepMethod;       /* 1: split [t0,t1]; 2 (default): |V|/depth  */
    float    nearThreshold;
    float    nearMultiplier;
    float    opaqueThreshold;
    float    chaosThreshold;
    float    depth;
    float    jitter;
    float    refine;
    float    gradient;
    float    lMin;
    float    lMax;
} lyap_params;

typedef struct lyap_point {
    lyap_vec3 P;               /* hit point                                 */
    lyap_vec3 N;               /* surface normal                            */
    float     a;               /* accumulated "alpha" below chaos threshold */
    float     c;               /* accumulated chaos                         */
    float     l;               /* exponent at the hit point                 */
} lyap_point;

typedef struct lyap_rgba {
    uint8_t r, g, b, a;
} lyap_rgba;

enum { LYAP_MAX_LIGHTS = 16 };  /* reference params.hpp:21 */

#if defined(__cplusplus)
#define LYAP_SASSERT(c, m) static_assert(c, m)
#else
#define LYAP_SASSERT(c, m) _Static_assert(c, m)
#endif

LYAP_SASSERT(sizeof(lyap_vec3) == 12, "Vec is 12 bytes");
LYAP_SASSERT(sizeof(lyap_quat) == 16, "Quat is 16 bytes");
LYAP_SASSERT(sizeof(lyap_color) == 16, "Color is 16 bytes");
LYAP_SASSERT(sizeof(lyap_camlight) == 224, "LyapCam/LyapLight is 224 bytes");
LYAP_SASSERT(offsetof(lyap_camlight, Q) == 16, "Q");
LYAP_SASSERT(offsetof(lyap_camlight, M) == 32, "M");
LYAP_SASSERT(offsetof(lyap_camlight, V) == 36, "V");
LYAP_SASSERT(offsetof(lyap_camlight, S0) == 48, "S0");
LYAP_SASSERT(offsetof(lyap_camlight, SDX) == 60, "SDX");
LYAP_SASSERT(offsetof(lyap_camlight, SDY) == 72, "SDY");
LYAP_SASSERT(offsetof(lyap_camlight, textureWidth) == 84, "textureWidth");
LYAP_SASSERT(offsetof(lyap_camlight, renderDenominator) == 100, "renderDenominator");
LYAP_SASSERT(offsetof(lyap_camlight, lightInnerCone) == 104, "lightInnerCone");
LYAP_SASSERT(offsetof(lyap_camlight, lightRange) == 112, "lightRange");
LYAP_SASSERT(offsetof(lyap_camlight, ambient) == 128, "ambient");
LYAP_SASSERT(offsetof(lyap_camlight, diffuseColor) == 144, "diffuseColor");
LYAP_SASSERT(offsetof(lyap_camlight, diffusePower) == 160, "diffusePower");
LYAP_SASSERT(offsetof(lyap_camlight, specularColor) == 176, "specularColor");
LYAP_SASSERT(offsetof(lyap_camlight, specularPower) == 192, "specularPower");
LYAP_SASSERT(offsetof(lyap_camlight, specularHardness) == 196, "specularHardness");
LYAP_SASSERT(offsetof(lyap_camlight, chaosColor) == 208, "chaosColor");
LYAP_SASSERT(sizeof(lyap_params) == 56, "LyapParams is 56 bytes");
LYAP_SASSERT(offsetof(lyap_params, lMax) == 52, "lMax");
LYAP_SASSERT(sizeof(lyap_point) == 36, "LyapPoint is 36 bytes");
LYAP_SASSERT(offsetof(lyap_point, a) == 24, "a");
LYAP_SASSERT(sizeof(lyap_rgba) == 4, "RGBA is 4 bytes");

#ifdef __cplusplus
}
#endif

#endif /* LYAP_TYPES_H */
