// tu_render_impl.cuh -- included by tu_render_<mode>.cu with LYAP_TU_MODE / LYAP_TU_NAME set.
#include "launch.hpp"

namespace lyap {

// fast mode renders two rays per lane through the packed evaluator
template <int MODE, int P>
struct RenderKernelOf {
    static constexpr void (*fn)(RenderArgs) = render_kernel<MODE, P>;
};
template <int P>
struct RenderKernelOf<kFast, P> {
    static constexpr void (*fn)(RenderArgs) = render_fast2_kernel<P>;
};
#define LYAP_RENDER_KERNEL(p) RenderKernelOf<LYAP_TU_MODE, p>::fn
#define LYAP_CAT2(a, b) a##b
#define LYAP_CAT(a, b) LYAP_CAT2(a, b)

// bytes per table entry: the two-rays-per-lane fast kernel stores multiplier pairs
constexpr size_t kRenderEntry = (LYAP_TU_MODE == kFast) ? 8 : 4;

cudaError_t LYAP_CAT(launch_render_, LYAP_TU_NAME)(int P, const RenderArgs &args, unsigned grid, cudaStream_t s)
{
    RenderArgs a = args;
    const size_t dyn = dyn_smem_of(LYAP_TU_MODE) + (P == 0 ? settle_seq_table(LYAP_TU_MODE, a, kRenderEntry, kRenderThreads) : 0);
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(LYAP_RENDER_KERNEL(p), dyn_smem_cap(LYAP_TU_MODE, p)); LYAP_RENDER_KERNEL(p)<<<grid, kRenderThreads, dyn, s>>>(a); break;
        LYAP_PERIODS(X)
#undef X
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

int LYAP_CAT(render_blocks_per_sm_, LYAP_TU_NAME)(int P, const SeqPlan &plan)
{
    int n = 0;
    size_t dyn = dyn_smem_of(LYAP_TU_MODE);
    if (P == 0 && seq_table_bytes(LYAP_TU_MODE, plan, kRenderEntry, kRenderThreads) <= kSeqTableMaxBytes)
        dyn += seq_table_bytes(LYAP_TU_MODE, plan, kRenderEntry, kRenderThreads);
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(LYAP_RENDER_KERNEL(p), dyn_smem_cap(LYAP_TU_MODE, p)); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, LYAP_RENDER_KERNEL(p), kRenderThreads, dyn); break;
        LYAP_PERIODS(X)
#undef X
    default: break;
    }
    return n;
}

} // namespace lyap
