/*
 * lyap/abi.h -- C ABI of liblyap_b200.so: the drop-in boundary for the reference's
 * hot path (the two __global__ kernels of kernel.hpp:17,19 and the host helpers of
 * scene.hpp:11-15 / params.hpp:26 that feed them).
 *
 * Plain pointers and sizes only; no torch, no C++ types.  Every entry point that
 * touches the GPU returns 0 on success or a cudaError_t / LYAP_ERR_* code and never
 * calls exit() (the reference's checkCudaErrors does, helper_cuda.h).  Device entry
 * points are stream-ordered and do not synchronise; `stream` is a cudaStream_t
 * passed as void* (NULL = the legacy default stream, as in the reference).
 *
 * INTEGRATION.md shows the two-line change in lyap_interactive.cu / lyap_calculate.cu.
 */
#ifndef LYAP_ABI_H
#define LYAP_ABI_H

#include "lyap/types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* How the per-sample exponent is evaluated. */
enum lyap_mode {
    /* Parity mode against the reference's CUDA build: the same PTX-level operations
     * in the same order as `nvcc --use_fast_math kernel.cu` (GNUmakefile:8), one
     * lg2.approx per step.  Bit-identical LyapPoint/RGBA to the reference kernel on
     * the same GPU.  SFU-bound. */
    LYAP_MODE_EXACT = 0,
    /* Throughput mode: derivative magnitudes are multiplied, the exponent folded out
     * with integer ops, log taken once per sample.  Same chaotic trajectory; volumes
     * within 1e-3 of the reference.  NOT an image-parity mode: with jitter on, frac(l)
     * is the reference's PRNG and normals are differences at float-noise level, so FAST
     * frames agree with the reference on only ~75 % of pixels.
     * For frames use EXACT / HOST, or the HYBRID modes below when jitter == 0. */
    LYAP_MODE_FAST = 1,
    /* Parity mode against the reference's HOST build (and the CPU oracle): IEEE
     * arithmetic without contraction and a bit-exact glibc logf per step. */
    LYAP_MODE_HOST = 2,
    /* EXACT's (resp. HOST's) hit points, normals, exponents and pixels at close to FAST's cost for
     * scenes with prm->jitter == 0: the march samples -- whose exponents are then only
     * compared with the three thresholds -- run on the fast evaluator, samples inside a guard band
     * around a threshold are re-evaluated, and refinement + normals run on the parity evaluator.
     * LyapPoint.P/N/l and RGBA equal the parity mode's; the cloud sums a and c are sums of the fast
     * march exponents and agree to ~1e-6 relative.  With jitter != 0 the call IS the parity mode.
     * lyap_bake / lyap_exponent_points / lyap_shade_points treat these as EXACT resp. HOST. */
    LYAP_MODE_HYBRID = 3,
    LYAP_MODE_HYBRID_HOST = 4,
};

enum lyap_dtype { LYAP_F32 = 0, LYAP_F16 = 1 };

enum lyap_error {
    LYAP_OK = 0,
    LYAP_ERR_BAD_SEQUENCE = 10001,   /* empty, symbol outside 0..3, or longer than LYAP_MAX_SEQUENCE */
    LYAP_ERR_BAD_ARGUMENT = 10002,
    LYAP_ERR_NO_DEVICE = 10003,
    LYAP_ERR_IO = 10004,
};

enum { LYAP_MAX_SEQUENCE = 1024 };
enum { LYAP_MAX_TILE = 4096 };   /* largest tile edge lyap_render_tiles / lyap_scatter_tiles / lyap_tile_count accept */

const char *lyap_version(void);
const char *lyap_error_string(int code);

/* Tuning / test knobs: "render_warps_per_sm" (persistent warps per SM, default 16),
 * "bake_blocks_per_sm" (cap; 0 = occupancy maximum), "force_generic" (1 = always use
 * the generic-period path instead of a period instantiation), "seq_table" (generic path:
 * 0 = run-length loop, 2 = per-lane multiplier table in shared memory wherever it fits,
 * 1 = default: the table for sequences whose runs average below 14 steps),
 * "emulate_ref_nvcc_normals" (test knob: reproduce the normals the reference's CUDA
 * build produces under nvcc 12.9, where ls[] aliases lyap4d's abcd[]),
 * "hybrid_guard_batch" (parked lanes per warp that trigger a parity pass; 0 = default = 1),
 * "hybrid_guard_percent" (test knob: guard band width in % of the derived bound),
 * "fast_redo" (1 = default: a FAST-mode sample whose exponent fold saw a zero exponent field
 * -- a zero derivative or an underflow between two folds -- is evaluated again by the
 * guaranteed-safe loop; 0 reports NaN for it at once, for measurements),
 * "tail_compaction" (1 = default: repack the last rays of a frame launch into fewer warps),
 * "tile_order" (1 = default: a frame launch queues its tiles by descending chord length of
 * the centre ray, an upper bound of the march length, so that the rays started last are
 * short; 0 = image order.  The output does not depend on either).
 * The knobs are process-global and read at launch time: set them before launching from
 * several threads, not concurrently with launches. */
int lyap_set_option(const char *key, long value);
/* Which unrolled-period instantiation a sequence runs on: its period's smallest
 * compiled multiple (1..32, 36, 40), 0 for the generic path, -1 for an invalid sequence. */
int lyap_plan_period(const int32_t *seq, uint32_t settle, uint32_t accum);
/* Test hook: dumps the iteration schedule built for a sequence (see csrc/abi.cu). */
int lyap_plan_describe(const int32_t *seq, uint32_t settle, uint32_t accum, uint32_t *header, uint8_t *sym, uint8_t *rot, uint8_t *runs);

/* ----------------------------------------------------------------------------
 * Host scene helpers -- same meaning as the reference functions they replace.
 * ------------------------------------------------------------------------- */

/* params_init() (params.cu:21-114): the hard-coded defaults.  The reference fills
 * globals; here the caller passes the objects.  `lights` must hold LYAP_MAX_LIGHTS
 * entries; `sequence` receives the NUL-terminated default string ("BCABA"). */
void lyap_params_init(lyap_params *prm, lyap_cam *cam, lyap_light *lights, uint32_t *num_lights,
                      char *sequence, size_t sequence_cap, uint32_t *image_width, uint32_t *image_height);

/* scene_convert_sequence() (scene.cu:69-108): "A6B6C6" -> malloc'd int array ending
 * in -1, returns the element count including the terminator; caller free()s.  A bad
 * letter returns 0 with *seqP = NULL instead of exit(1). */
size_t lyap_scene_convert_sequence(int32_t **seqP, const unsigned char *seqStr);

/* scene_cam_recalculate() (scene.cu:31-63) and scene_lights_recalculate() (:20-29). */
void lyap_scene_cam_recalculate(lyap_cam *camP, uint32_t tw, uint32_t th, uint32_t td);
void lyap_scene_lights_recalculate(lyap_light *lights, size_t num_lights);

/* scale.pl:5-11 and the camera block scale.pl:33-48 / params.cu:42-57 as a runtime
 * function: sets cam->C and cam->Q for eased path position i in [0,1]. */
double lyap_ease_in_out_quart(double t, double b, double c, double d);
void lyap_campath_orbit(double i, lyap_cam *cam);
/* Frame f of an n-frame orbit: t = f/(n-1), i = ease(t); i is rounded to 15
 * significant digits as scale.pl's string interpolation does. */
void lyap_campath_frame(uint32_t f, uint32_t n_frames, lyap_cam *cam);

/* ----------------------------------------------------------------------------
 * Device entry points (caller-owned device buffers, as in the reference).
 * ------------------------------------------------------------------------- */

/* Replaces  kernel_calc_render<<<blocks,threads>>>(cudaRGBA, cudaPoints, cam, prm,
 *           cudaSeq, cudaLights, num_lights)            (lyap_interactive.cu:711).
 *
 * d_rgba / d_points: width*height elements, row-major, index x + y*width.
 * seq: the -1-terminated array scene_convert_sequence produced -- either the HOST array
 *      (preferred: it is read on the host and baked into the launch) or the reference's
 *      device copy `cudaSeq` (fetched with a small blocking copy per call).
 * d_lights: device array of num_lights lights, already recalculated.
 * Miss pixels leave d_points[ind] untouched and shade whatever it holds, exactly as
 * the reference does (kernel.cu:508-512): zero-fill it for defined output.
 * d_evals: optional device counter; the number of exponent evaluations is added. */
int lyap_render(lyap_rgba *d_rgba, lyap_point *d_points, const lyap_cam *cam, const lyap_params *prm,
                const int32_t *seq, const lyap_light *d_lights, uint32_t num_lights,
                uint32_t width, uint32_t height, int mode, unsigned long long *d_evals, void *stream);

/* The same frame cut into tile x tile pixel tiles, dealt round-robin to `world`
 * ranks; this call renders rank `rank`'s tiles only.  With compact != 0 the outputs
 * are written densely in the rank's work order (lyap_tile_count() elements) instead
 * of at their image positions. */
int lyap_render_tiles(lyap_rgba *d_rgba, lyap_point *d_points, const lyap_cam *cam, const lyap_params *prm,
                      const int32_t *seq, const lyap_light *d_lights, uint32_t num_lights,
                      uint32_t width, uint32_t height, uint32_t tile, uint32_t rank, uint32_t world,
                      int compact, int mode, unsigned long long *d_evals, void *stream);
/* Number of work items (tile pixels, including padding of ragged edge tiles) of a rank. */
uint64_t lyap_tile_count(uint32_t width, uint32_t height, uint32_t tile, uint32_t rank, uint32_t world);
/* Rank 0: place a rank's compact buffer (elem_size bytes per item) into the full image. */
int lyap_scatter_tiles(void *d_image, const void *d_compact, uint32_t elem_size, uint32_t width, uint32_t height,
                       uint32_t tile, uint32_t rank, uint32_t world, void *stream);

/* ----------------------------------------------------------------------------
 * Volume-assisted rendering (beyond the reference).  A volume baked by
 * lyap_bake with the SAME sequence / d / settle / accum classifies the cells of its n^3 grid
 * (n a cube edge, samples at 4i/n); march samples that fall into a SAFE cell are stepped over
 * without evaluating the exponent.  Hybrid modes only, and only when prm->jitter == 0 (otherwise the
 * call renders exactly like lyap_render_tiles).  The ray positions are unchanged, so hit point,
 * normal, exponent and pixel equal the non-assisted frame's wherever no surface crossing hides inside
 * a "safe" neighbourhood; the cloud sums a / c take the baked value of each skipped sample's cell.
 *
 * lyap_assist_build: cell (i,j,k) is SAFE when every volume sample of the neighbourhood
 * i-dilate .. i+1+dilate (all three axes) exists, is finite, and lies strictly between
 * max(opaqueThreshold, nearThreshold) + margin and `upper` (pass +inf for no upper bound; e.g. 0
 * keeps the march exact inside chaotic regions, whose thin stable windows no grid resolves).
 * d_safe_bits holds lyap_assist_bits_bytes(n) bytes: one bit per cell, x fastest.
 * d_skipped: optional device counter, += number of samples stepped over.
 * ------------------------------------------------------------------------- */
uint64_t lyap_assist_bits_bytes(uint32_t n);
int lyap_assist_build(uint32_t *d_safe_bits, const void *d_volume, int dtype, uint32_t n, const lyap_params *prm,
                      float margin, float upper, uint32_t dilate, void *stream);
int lyap_render_assisted(lyap_rgba *d_rgba, lyap_point *d_points, const lyap_cam *cam, const lyap_params *prm,
                         const int32_t *seq, const lyap_light *d_lights, uint32_t num_lights,
                         uint32_t width, uint32_t height, uint32_t tile, uint32_t rank, uint32_t world, int compact, int mode,
                         const void *d_volume, int dtype, uint32_t n, const uint32_t *d_safe_bits,
                         unsigned long long *d_evals, unsigned long long *d_skipped, void *stream);

/* Re-shade a stored LyapPoint buffer without marching (lights/camera edits). */
int lyap_shade_points(lyap_rgba *d_rgba, const lyap_point *d_points, const lyap_cam *cam,
                      const lyap_light *d_lights, uint32_t num_lights, uint64_t count, int mode, void *stream);

/* Replaces  kernel_calc_volume<<<blocks,threads>>>(cudaExps, prm, cudaSeq)
 *                                                       (lyap_calculate.cu:72).
 * Voxel (x,y,z) samples (4x/nx, 4y/ny, 4z/nz) (kernel.cu:527-529) and is stored at
 * x + (y + z*ny)*nx of the FULL volume d_exps; this call fills planes z0 <= z < z1. */
int lyap_bake(void *d_exps, int dtype, const lyap_params *prm, const int32_t *seq,
              uint32_t nx, uint32_t ny, uint32_t nz, uint32_t z0, uint32_t z1, int mode, void *stream);

/* Exponent at arbitrary points (xyz: n*3 floats on the device) -> d_out[n]. */
int lyap_exponent_points(float *d_out, const float *d_xyz, uint64_t n, const lyap_params *prm,
                         const int32_t *seq, int mode, void *stream);

/* Diagnostics in a mode's own arithmetic (EXACT: the reference CUDA build's approximate divide and
 * square root, which a CPU cannot recompute; HOST: IEEE).  lyap_ray_probe runs the ray set-up of
 * kernel.cu:160-315 for n pixels (d_pixels: n x {x, y}) and writes 12 floats per pixel:
 * {hit, V.x, V.y, V.z, t0, t1, Fdt, Ndt, P0.x, P0.y, P0.z, Fdt*gradient}.  lyap_normalize_vectors
 * applies Vec::normalize (vec3.hpp:141-154) in place to n xyz triples. */
int lyap_ray_probe(float *d_out, const uint32_t *d_pixels, uint64_t n, const lyap_cam *cam, const lyap_params *prm, int mode, void *stream);
int lyap_normalize_vectors(float *d_xyz, uint64_t n, int mode, void *stream);

/* ----------------------------------------------------------------------------
 * Peer memory for one-process-per-GPU sharding: rank 0 allocates the result buffer and
 * exports a 64-byte CUDA IPC handle; the other ranks open it and pass the mapped pointer
 * as d_exps / d_rgba / d_points above.  Their kernels then store their shard directly
 * into rank 0's memory over NVLink -- there is no separate gather.
 * ------------------------------------------------------------------------- */
int lyap_peer_alloc(void **dptr, uint64_t bytes);              /* cudaMalloc + zero fill */
int lyap_peer_free(void *dptr);
int lyap_peer_export(void *dptr, unsigned char *handle64);     /* owner side */
int lyap_peer_open(const unsigned char *handle64, void **dptr);/* other ranks */
int lyap_peer_close(void *dptr);

/* ----------------------------------------------------------------------------
 * Whole-call convenience with HOST buffers: H2D of the lights, the kernel, D2H of the
 * results and the final synchronise happen inside (this is the end-to-end path bench.py
 * times).  Device buffers and a stream are kept per device between calls; one call at a
 * time per device.  LYAP_TRACE=1 in the environment prints stage timings to stderr.
 * ------------------------------------------------------------------------- */
int lyap_render_host(lyap_rgba *h_rgba, lyap_point *h_points /* may be NULL */, const lyap_cam *cam,
                     const lyap_params *prm, const int32_t *seq, const lyap_light *h_lights, uint32_t num_lights,
                     uint32_t width, uint32_t height, int mode, int device, unsigned long long *evals_out);
/* Frees the device buffers and stream the host-buffer calls keep between calls. */
void lyap_host_workspace_release(void);
int lyap_bake_host(void *h_exps, int dtype, const lyap_params *prm, const int32_t *seq,
                   uint32_t nx, uint32_t ny, uint32_t nz, uint32_t z0, uint32_t z1, int mode, int device);

/* ----------------------------------------------------------------------------
 * Headless output in the reference's formats.
 * ------------------------------------------------------------------------- */
/* ASCII "P3" exactly as save_ppm writes it (lyap_interactive.cu:595-606). */
int lyap_write_ppm(const char *path, const lyap_rgba *h_rgba, uint32_t width, uint32_t height);
/* 8-bit RGB PNG (zlib deflate). */
int lyap_write_png(const char *path, const lyap_rgba *h_rgba, uint32_t width, uint32_t height);
/* Raw dumps: LyapPoint[] (save_points, :642-647) and exps.raw (lyap_calculate.cu:84-86). */
int lyap_write_raw(const char *path, const void *data, uint64_t bytes);
/* The reference's self-describing file stem (lyap_interactive.cu:577-590), without extension. */
int lyap_format_filename(char *out, size_t cap, const char *prefix, unsigned long timestamp, uint32_t width, uint32_t height,
                         const char *sequence, const lyap_cam *cam, const lyap_params *prm);

/* ----------------------------------------------------------------------------
 * Roofline probes on the current device: register-only loops of eight independent chains
 * per thread.  lyap_probe_peaks reports the achieved FFMA and MUFU.LG2 lane-operations per
 * second (the denominators of bench.py's roofline); lyap_probe_ffma2 the same for packed
 * fma.rn.f32x2 in scalar-equivalent lane-operations.  sm_clock_hz_est is informational.
 * ------------------------------------------------------------------------- */
int lyap_probe_ffma2(double *packed_ffma_lane_ops_per_s);
int lyap_probe_peaks(double *ffma_lane_ops_per_s, double *mufu_lane_ops_per_s, double *sm_clock_hz_est, int *sm_count);

#ifdef __cplusplus
}
#endif

#endif /* LYAP_ABI_H */
