// tu_march_impl.cuh -- included by tu_march_<mode>.cu with LYAP_TU_MODE / LYAP_TU_NAME set.
#include "launch.hpp"

namespace lyap {

#define LYAP_CAT2(a, b) a##b
#define LYAP_CAT(a, b) LYAP_CAT2(a, b)

cudaError_t LYAP_CAT(launch_march_, LYAP_TU_NAME)(int P, const RenderArgs &a, unsigned grid, cudaStream_t s)
{
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(march_fast2_kernel<LYAP_TU_MODE, p>, dyn_smem_of(LYAP_TU_MODE)); march_fast2_kernel<LYAP_TU_MODE, p><<<grid, kRenderThreads, dyn_smem_of(LYAP_TU_MODE), s>>>(a); break;
        LYAP_PERIODS(X)
#undef X
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

int LYAP_CAT(march_blocks_per_sm_, LYAP_TU_NAME)(int P)
{
    int n = 0;
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(march_fast2_kernel<LYAP_TU_MODE, p>, dyn_smem_of(LYAP_TU_MODE)); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, march_fast2_kernel<LYAP_TU_MODE, p>, kRenderThreads, dyn_smem_of(LYAP_TU_MODE)); break;
        LYAP_PERIODS(X)
#undef X
    default: break;
    }
    return n;
}

} // namespace lyap
