/*
 * oracle/lyap_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * C restatement, on the CPU, of the reference's hot path with the arithmetic
 * of the reference's HOST build (g++ -O2 -ffp-contract=off: IEEE float ops, no
 * FMA contraction, double precision exactly where the reference's literals
 * force it, glibc logf/powf/sqrtf).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it; the product
 * library never links or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py checks every entry
 * point bit-for-bit against oracle/_ref/libref_host.so (the unmodified reference
 * sources compiled in place) when that library is present, and
 * tests/test_oracle_golden.py checks it against the fixtures in tests/golden/
 * that oracle/make_golden.py generated from the same reference build.
 */
#ifndef LYAP_ORACLE_H
#define LYAP_ORACLE_H

#include "lyap/types.h"

#ifdef __cplusplus
extern "C" {
#endif

size_t oracle_convert_sequence(const char *str, int32_t *out, size_t cap);
void oracle_params_init(lyap_params *prm, lyap_cam *cam, lyap_light *lights16, uint32_t *n_lights,
                        char *seq_out, size_t seq_cap, uint32_t *w, uint32_t *h);
void oracle_cam_recalculate(lyap_cam *cam, uint32_t tw, uint32_t th, uint32_t td);
void oracle_lights_recalculate(lyap_light *lights, size_t n);
double oracle_ease_in_out_quart(double t, double b, double c, double d);
void oracle_campath(double i, lyap_cam *cam);

float oracle_lyap4d(float x, float y, float z, float d, uint32_t settle, uint32_t accum, const int32_t *seq);
void oracle_lyap4d_many(const float *xyz, size_t n, float d, uint32_t settle, uint32_t accum,
                        const int32_t *seq, float *out);
int oracle_raymarch(lyap_point *pt, uint32_t sx, uint32_t sy, const lyap_cam *cam, const lyap_params *prm,
                    const int32_t *seq, uint64_t *n_calls);
void oracle_shade(const lyap_point *pt, const lyap_cam *cam, const lyap_light *lights, uint32_t n, float *rgba4);
void oracle_to_rgba(const float *rgba4, uint8_t *out);

/* Rows y0..y1 of a w*h frame; returns the number of exponent evaluations made. */
uint64_t oracle_render_rows(lyap_rgba *rgba, lyap_point *points, const lyap_cam *cam, const lyap_params *prm,
                            const int32_t *seq, const lyap_light *lights, uint32_t n_lights,
                            uint32_t w, uint32_t h, uint32_t y0, uint32_t y1);
/* An arbitrary list of pixel indices (x + y*w) of the same frame. */
uint64_t oracle_render_pixels(lyap_rgba *rgba, lyap_point *points, const lyap_cam *cam, const lyap_params *prm,
                              const int32_t *seq, const lyap_light *lights, uint32_t n_lights,
                              uint32_t w, uint32_t h, const uint32_t *pix, size_t n_pix);
/* Planes z0..z1 of an nx*ny*nz volume; exps is the full volume. */
void oracle_bake_slab(float *exps, const lyap_params *prm, const int32_t *seq,
                      uint32_t nx, uint32_t ny, uint32_t nz, uint32_t z0, uint32_t z1);
int oracle_num_threads(void);
void oracle_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
