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

cudaError_t LYAP_CAT(launch_render_, LYAP_TU_NAME)(int P, const RenderArgs &a, unsigned grid, cudaStream_t s)
{
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(LYAP_RENDER_KERNEL(p), dyn_smem_of(LYAP_TU_MODE)); LYAP_RENDER_KERNEL(p)<<<grid, kRenderThreads, dyn_smem_of(LYAP_TU_MODE), s>>>(a); break;
        LYAP_PERIODS(X)
#undef X
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

int LYAP_CAT(render_blocks_per_sm_, LYAP_TU_NAME)(int P)
{
    int n = 0;
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(LYAP_RENDER_KERNEL(p), dyn_smem_of(LYAP_TU_MODE)); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, LYAP_RENDER_KERNEL(p), kRenderThreads, dyn_smem_of(LYAP_TU_MODE)); break;
        LYAP_PERIODS(X)
#undef X
    default: break;
    }
    return n;
}

} // namespace lyap
