// launch.hpp -- host-visible launchers; one set per exponent mode, each compiled in
// its own translation unit (tu_*.cu) so the period instantiations build in parallel.
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

// Sequence periods with a register-table instantiation: every period up to 32 symbols, plus 36 and 40
// (A8B8C8D8, A9B9C9D9: the register table still fits; 48 spills).  Other periods take the generic path (0):
// a per-lane multiplier table in shared memory, or the run-length loop where that does not fit.
#define LYAP_PERIODS(X) \
    X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(14) X(15) X(16) \
    X(17) X(18) X(19) X(20) X(21) X(22) X(23) X(24) X(25) X(26) X(27) X(28) X(29) X(30) X(31) X(32) \
    X(36) X(40)

namespace lyap {

// Dynamic shared memory of a launch.  HOST-mode kernels keep their replicated logf table (hostlog.cuh)
// there; the generic-period path of the other modes keeps its per-lane multiplier table there
// (exponent.cuh: plan.table_stride entries of `entry_bytes` per thread).  Either can exceed the 48 KiB a
// launch gets without opting in.
constexpr size_t kSeqTableMaxBytes = 56 * 1024;   // per block: four render blocks per SM stay resident
constexpr size_t dyn_smem_of(int mode) { return mode == kHost ? (size_t)kLogfSmemBytes : 0; }
inline size_t seq_table_bytes(int mode, const SeqPlan &plan, size_t entry_bytes, size_t threads)
{
    return mode == kHost ? 0 : (size_t)plan.table_stride * entry_bytes * threads;
}
// Generic-period launches take their table when it fits and the run-length loop otherwise.
template <class Args>
inline size_t settle_seq_table(int mode, Args &a, size_t entry_bytes, size_t threads)
{
    const size_t bytes = seq_table_bytes(mode, a.plan, entry_bytes, threads);
    if (bytes > kSeqTableMaxBytes) {
        a.plan.table_stride = 0;
        return 0;
    }
    return bytes;
}
// The opt-in limit of a kernel is a constant per (mode, period) -- the most any launch of it may ask
// for -- so that concurrent launches from several host threads never lower it under each other.
constexpr size_t dyn_smem_cap(int mode, int P) { return dyn_smem_of(mode) + ((P == 0 && mode != kHost) ? kSeqTableMaxBytes : 0); }
template <class K>
inline void opt_in_dyn_smem(K kernel, size_t cap)
{
    if (cap) cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cap);
}

#define LYAP_DECLARE_MODE(NAME)                                                                          \
    cudaError_t launch_bake_##NAME(int P, const BakeArgs &a, unsigned grid, cudaStream_t s);              \
    cudaError_t launch_points_##NAME(int P, const PointsArgs &a, unsigned grid, cudaStream_t s);          \
    cudaError_t launch_render_##NAME(int P, const RenderArgs &a, unsigned grid, cudaStream_t s);          \
    int render_blocks_per_sm_##NAME(int P, const SeqPlan &plan);                                          \
    int bake_blocks_per_sm_##NAME(int P, const SeqPlan &plan);                                            \
    int bake_threads_##NAME(const SeqPlan &plan);

LYAP_DECLARE_MODE(exact)
LYAP_DECLARE_MODE(fast)
LYAP_DECLARE_MODE(host)

// hybrid mode's march kernel (packed fast evaluator + parity evaluator NAME for the guard bands)
cudaError_t launch_march_exact(int P, const RenderArgs &a, unsigned grid, cudaStream_t s);
cudaError_t launch_march_host(int P, const RenderArgs &a, unsigned grid, cudaStream_t s);
int march_blocks_per_sm_exact(int P, const SeqPlan &plan);
int march_blocks_per_sm_host(int P, const SeqPlan &plan);

// Tile permutation of a frame launch (RenderArgs::tile_order): `scratch` of tile_order_scratch_bytes(n_tiles)
// device bytes, used on stream s; *order points into it and is valid for work queued on s afterwards.
size_t tile_order_scratch_bytes(uint32_t n_tiles);
cudaError_t launch_tile_order(const RenderArgs &a, void *scratch, const uint32_t **order, cudaStream_t s);
cudaError_t launch_shade(int mode, const ShadeArgs &a, unsigned grid, cudaStream_t s);
cudaError_t launch_scatter(const ScatterArgs &a, unsigned grid, cudaStream_t s);
cudaError_t launch_assist_build(const AssistBuildArgs &a, unsigned grid, cudaStream_t s);
cudaError_t launch_ray_probe(int mode, float *out, const uint32_t *pixels, uint64_t n, const lyap_cam &cam, const lyap_params &prm, cudaStream_t s);
cudaError_t launch_normalize(int mode, float *xyz, uint64_t n, cudaStream_t s);
cudaError_t probe_peaks(double *ffma_ops, double *mufu_ops, double *clock_hz, int *sms);
cudaError_t probe_ffma2(double *lane_ops);

} // namespace lyap
