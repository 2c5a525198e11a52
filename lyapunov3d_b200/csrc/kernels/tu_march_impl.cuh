// tu_march_impl.cuh -- included by tu_march_<mode>.cu with LYAP_TU_MODE / LYAP_TU_NAME set.
#include "launch.hpp"

namespace lyap {

#define LYAP_CAT2(a, b) a##b
#define LYAP_CAT(a, b) LYAP_CAT2(a, b)

// the march kernel alternates between the packed fast evaluator (8-byte entries) and the parity one
// (4-byte entries) on the same strip; with a HOST parity evaluator the strip is not available
// (that mode's logf table lives there) and both take the run-length loop
cudaError_t LYAP_CAT(launch_march_, LYAP_TU_NAME)(int P, const RenderArgs &args, unsigned grid, cudaStream_t s)
{
    RenderArgs a = args;
    if (LYAP_TU_MODE == kHost) a.plan.table_stride = 0;
    const size_t dyn = dyn_smem_of(LYAP_TU_MODE) + (P == 0 ? settle_seq_table(LYAP_TU_MODE, a, 8, kRenderThreads) : 0);
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(march_fast2_kernel<LYAP_TU_MODE, p>, dyn_smem_cap(LYAP_TU_MODE, p)); march_fast2_kernel<LYAP_TU_MODE, p><<<grid, kRenderThreads, dyn, s>>>(a); break;
        LYAP_PERIODS(X)
#undef X
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

int LYAP_CAT(march_blocks_per_sm_, LYAP_TU_NAME)(int P, const SeqPlan &plan)
{
    int n = 0;
    size_t dyn = dyn_smem_of(LYAP_TU_MODE);
    if (P == 0 && seq_table_bytes(LYAP_TU_MODE, plan, 8, kRenderThreads) <= kSeqTableMaxBytes) dyn += seq_table_bytes(LYAP_TU_MODE, plan, 8, kRenderThreads);
    switch (P) {
#define X(p) case p: opt_in_dyn_smem(march_fast2_kernel<LYAP_TU_MODE, p>, dyn_smem_cap(LYAP_TU_MODE, p)); cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, march_fast2_kernel<LYAP_TU_MODE, p>, kRenderThreads, dyn); break;
        LYAP_PERIODS(X)
#undef X
    default: break;
    }
    return n;
}

} // namespace lyap
