// tu_bake_impl.cuh -- included by tu_bake_<mode>.cu with LYAP_TU_MODE / LYAP_TU_NAME set.
#include "launch.hpp"

namespace lyap {

#define LYAP_CAT2(a, b) a##b
#define LYAP_CAT(a, b) LYAP_CAT2(a, b)

// The fast bake evaluates voxel pairs (8-byte table entries).  A generic-period launch that uses the
// shared-memory table runs 128-thread blocks, which halves the table of a block.
constexpr size_t kBakeEntry = (LYAP_TU_MODE == kFast) ? 8 : 4;

int LYAP_CAT(bake_threads_, LYAP_TU_NAME)(const SeqPlan &plan)
{
    const size_t b = seq_table_bytes(LYAP_TU_MODE, plan, kBakeEntry, 128);
    return (b != 0 && b <= kSeqTableMaxBytes) ? 128 : 256;
}

cudaError_t LYAP_CAT(launch_bake_, LYAP_TU_NAME)(int P, const BakeArgs &args, unsigned grid, cudaStream_t s)
{
    BakeArgs a = args;
    const unsigned threads = P == 0 ? (unsigned)LYAP_CAT(bake_threads_, LYAP_TU_NAME)(a.plan) : 256u;
    const size_t dyn = dyn_smem_of(LYAP_TU_MODE) + (P == 0 ? settle_seq_table(LYAP_TU_MODE, a, kBakeEntry, threads) : 0);
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(bake_kernel<LYAP_TU_MODE, p>, dyn_smem_cap(LYAP_TU_MODE, p)); bake_kernel<LYAP_TU_MODE, p><<<grid, threads, dyn, s>>>(a); break;
        LYAP_PERIODS(X)
#undef X
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t LYAP_CAT(launch_points_, LYAP_TU_NAME)(int P, const PointsArgs &args, unsigned grid, cudaStream_t s)
{
    PointsArgs a = args;
    const size_t tb = P == 0 ? seq_table_bytes(LYAP_TU_MODE, a.plan, 4, 128) : 0;
    const unsigned threads = (tb != 0 && tb <= kSeqTableMaxBytes) ? 128u : 256u;
    const size_t dyn = dyn_smem_of(LYAP_TU_MODE) + (P == 0 ? settle_seq_table(LYAP_TU_MODE, a, 4, threads) : 0);
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(points_kernel<LYAP_TU_MODE, p>, dyn_smem_cap(LYAP_TU_MODE, p)); points_kernel<LYAP_TU_MODE, p><<<grid, threads, dyn, s>>>(a); break;
        LYAP_PERIODS(X)
#undef X
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

int LYAP_CAT(bake_blocks_per_sm_, LYAP_TU_NAME)(int P, const SeqPlan &plan)
{
    int n = 0;
    const int threads = P == 0 ? LYAP_CAT(bake_threads_, LYAP_TU_NAME)(plan) : 256;
    size_t dyn = dyn_smem_of(LYAP_TU_MODE);
    if (P == 0 && threads == 128) dyn += seq_table_bytes(LYAP_TU_MODE, plan, kBakeEntry, 128);
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(bake_kernel<LYAP_TU_MODE, p>, dyn_smem_cap(LYAP_TU_MODE, p)); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, bake_kernel<LYAP_TU_MODE, p>, threads, dyn); break;
        LYAP_PERIODS(X)
#undef X
    default: break;
    }
    return n;
}

} // namespace lyap
