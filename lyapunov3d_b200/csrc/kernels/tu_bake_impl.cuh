// tu_bake_impl.cuh -- included by tu_bake_<mode>.cu with LYAP_TU_MODE / LYAP_TU_NAME set.
#include "launch.hpp"

namespace lyap {

#define LYAP_CAT2(a, b) a##b
#define LYAP_CAT(a, b) LYAP_CAT2(a, b)

cudaError_t LYAP_CAT(launch_bake_, LYAP_TU_NAME)(int P, const BakeArgs &a, unsigned grid, cudaStream_t s)
{
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(bake_kernel<LYAP_TU_MODE, p>, dyn_smem_of(LYAP_TU_MODE)); bake_kernel<LYAP_TU_MODE, p><<<grid, 256, dyn_smem_of(LYAP_TU_MODE), s>>>(a); break;
        LYAP_PERIODS(X)
#undef X
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t LYAP_CAT(launch_points_, LYAP_TU_NAME)(int P, const PointsArgs &a, unsigned grid, cudaStream_t s)
{
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(points_kernel<LYAP_TU_MODE, p>, dyn_smem_of(LYAP_TU_MODE)); points_kernel<LYAP_TU_MODE, p><<<grid, 256, dyn_smem_of(LYAP_TU_MODE), s>>>(a); break;
        LYAP_PERIODS(X)
#undef X
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

int LYAP_CAT(bake_blocks_per_sm_, LYAP_TU_NAME)(int P)
{
    int n = 0;
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(bake_kernel<LYAP_TU_MODE, p>, dyn_smem_of(LYAP_TU_MODE)); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bake_kernel<LYAP_TU_MODE, p>, 256, dyn_smem_of(LYAP_TU_MODE)); break;
        LYAP_PERIODS(X)
#undef X
    default: break;
    }
    return n;
}

} // namespace lyap
