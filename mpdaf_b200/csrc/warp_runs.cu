// Multi-run warp-tiled sigma clip for stacks of 129..512 exposures (runs of 64, sigclip_runs.cuh).
#include <algorithm>
#include <cstdio>

#include "launch.h"
#include "sigclip_runs.cuh"

int mb_launch_runs(const mb::Job& j, bool unit, mb::LaunchCtx& ctx)
{
    using namespace mb;
    const int runs = j.nfiles <= 256 ? 4 : 8;
    std::snprintf(ctx.name, ctx.name_len, "sigclip_runs<%dx64,%s>", runs, unit ? "unit" : "scaled");
    auto launch = [&](auto kernel, int smem) -> int {
        MB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int per_sm = 0;
        MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 32, smem));
        if (per_sm < 1) per_sm = 1;
        const long long ntiles = (j.a.nvox + 31) / 32;
        const int grid = (int)std::max(1LL, std::min(ntiles, (long long)ctx.sm_count * per_sm));
        kernel<<<grid, 32, smem, ctx.stream>>>(j.a, (const StackTable*)j.table_dev, ctx.slow);
        MB_CUDA(cudaGetLastError());
        ctx.launches++;
        return MPDAF_B200_OK;
    };
    constexpr int kRunBytes = 64 * 32 * (int)sizeof(float);
    if (runs == 4) return unit ? launch(sigclip_runs_kernel<4, true>, 4 * kRunBytes) : launch(sigclip_runs_kernel<4, false>, 4 * kRunBytes);
    return unit ? launch(sigclip_runs_kernel<8, true>, 8 * kRunBytes) : launch(sigclip_runs_kernel<8, false>, 8 * kRunBytes);
}
