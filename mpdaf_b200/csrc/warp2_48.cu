// Two-run warp-tiled sigma clip for stacks of up to 96 exposures (runs of 48).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "launch.h"
#include "sigclip_warp2.cuh"

int mb_launch_warp2_48(const mb::Job& j, bool unit, mb::LaunchCtx& ctx)
{
    using namespace mb;
    std::snprintf(ctx.name, ctx.name_len, "sigclip_warp2<2x48,%s>", unit ? "unit" : "scaled");
    static const int flags = std::getenv("MPDAF_B200_WARP_FLAGS") ? std::atoi(std::getenv("MPDAF_B200_WARP_FLAGS")) : 1;
    auto launch = [&](auto kernel) -> int {
        const int smem = (int)sizeof(Block2Smem<48>);
        MB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int per_sm = 0;
        MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWT, smem));
        if (per_sm < 1) per_sm = 1;
        const long long ntiles = (j.a.nvox + 31) / 32;
        const long long nblocks = (ntiles + (kWT / 32) - 1) / (kWT / 32);
        const int grid = (int)std::max(1LL, std::min(nblocks, (long long)ctx.sm_count * per_sm));
        kernel<<<grid, kWT, smem, ctx.stream>>>(j.a, (const StackTable*)j.table_dev, flags, ctx.slow);
        MB_CUDA(cudaGetLastError());
        ctx.launches++;
        return MPDAF_B200_OK;
    };
    return unit ? launch(sigclip_warp2_kernel<48, true>) : launch(sigclip_warp2_kernel<48, false>);
}
