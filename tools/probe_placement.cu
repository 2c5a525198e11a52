// Where do the blocks and warps of a persistent 592 x 128 launch land?  (tail compaction wants each
// co-resident block to keep a different scheduler: see kernels.cuh)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(128, 4) probe(unsigned *out, long long spin)
{
    unsigned smid, warpid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    const long long t0 = clock64();
    while (clock64() - t0 < spin) {}
    if ((threadIdx.x & 31) == 0) {
        out[(blockIdx.x * 4 + threadIdx.x / 32) * 2] = smid;
        out[(blockIdx.x * 4 + threadIdx.x / 32) * 2 + 1] = warpid;
    }
}
int main()
{
    const int blocks = 592;
    unsigned *d, h[blocks * 8];
    cudaMalloc(&d, sizeof h);
    probe<<<blocks, 128, 12708>>>(d, 2000000);
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("block: smid warpids\n");
    for (int b = 0; b < blocks; b += 1)
        if (b < 8 || (b % 148) < 2 || b > 586) printf("%3d: sm %3u  warps %2u %2u %2u %2u\n", b, h[b * 8], h[b * 8 + 1], h[b * 8 + 3], h[b * 8 + 5], h[b * 8 + 7]);
    // per SM: the blocks it hosts and their first warp ids
    int shown = 0;
    for (unsigned sm = 0; sm < 148 && shown < 6; ++sm) {
        printf("sm %u:", sm);
        for (int b = 0; b < blocks; ++b) if (h[b * 8] == sm) printf("  b%d(w%u,%u,%u,%u)", b, h[b * 8 + 1], h[b * 8 + 3], h[b * 8 + 5], h[b * 8 + 7]);
        printf("\n");
        ++shown;
    }
    // summary: is blockIdx/148 distinct among co-resident blocks?  is (warpid>>2)&3 ?
    int ok_div = 0, ok_hw = 0, ok_aligned = 0;
    for (unsigned sm = 0; sm < 148; ++sm) {
        int seen_div = 0, seen_hw = 0, n = 0, aligned = 1;
        for (int b = 0; b < blocks; ++b) if (h[b * 8] == sm) {
            seen_div |= 1 << (b / 148);
            seen_hw |= 1 << ((h[b * 8 + 1] >> 2) & 3);
            for (int w = 0; w < 4; ++w) if ((h[b * 8 + 1 + 2 * w] & 3) != (unsigned)w) aligned = 0;
            ++n;
        }
        ok_div += seen_div == 15; ok_hw += seen_hw == 15; ok_aligned += aligned;
    }
    printf("SMs where blockIdx/148 is distinct among the 4 co-resident blocks: %d of 148\n", ok_div);
    printf("SMs where (warpid>>2)&3 is distinct among them:                      %d of 148\n", ok_hw);
    printf("SMs where every block's warp w has warpid&3 == w:                    %d of 148\n", ok_aligned);
    return 0;
}
