// Multi-run warp-tiled sigma clip for stacks of 129..512 exposures (runs of 64, sigclip_runs.cuh).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "launch.h"
#include "sigclip_runs.cuh"

int mb_launch_runs(const mb::Job& j, bool unit, mb::LaunchCtx& ctx)
{
    using namespace mb;
    const int runs = j.nfiles <= 256 ? 4 : 8;
    std::snprintf(ctx.name, ctx.name_len, "sigclip_runs<%dx64,%s>", runs, unit ? "unit" : "scaled");
    // equally spaced rows (one [N][V] allocation per HDU, the library's own staging slots): no per-sample table lookups
    auto stride_of = [&](const void* const* rows) -> unsigned long long {
        if (rows == nullptr || j.nfiles < 2) return 0ull;
        const long long step = static_cast<const char*>(rows[1]) - static_cast<const char*>(rows[0]);
        if (step <= 0 || (step & 3) != 0) return 0ull;
        for (int i = 2; i < j.nfiles; i++)
            if (static_cast<const char*>(rows[i]) - static_cast<const char*>(rows[i - 1]) != step) return 0ull;
        return (unsigned long long)(step >> 2);
    };
    RunRows rr{};
    rr.data0 = static_cast<const float*>(j.data[0]);
    rr.var0 = (j.var != nullptr && j.a.typ_var == 0) ? static_cast<const float*>(j.var[0]) : nullptr;
    rr.dstride = stride_of(j.data);
    rr.vstride = rr.var0 != nullptr ? stride_of(j.var) : 0ull;
    rr.next_tile = ctx.slow_n;     // the pair kernels' queue counter doubles as the tile counter here
    static long long* dbg_clk = nullptr;
    static const bool want_clk = std::getenv("MPDAF_B200_RUNS_CLOCKS") != nullptr;
    if (want_clk && dbg_clk == nullptr) { cudaMalloc(&dbg_clk, 5 * sizeof(long long)); }
    if (want_clk) cudaMemsetAsync(dbg_clk, 0, 5 * sizeof(long long), ctx.stream);
    rr.phase_clk = want_clk ? dbg_clk : nullptr;
    const bool strided = rr.dstride != 0 && (rr.var0 == nullptr || rr.vstride != 0) && std::getenv("MPDAF_B200_NO_STRIDED") == nullptr;
    if (strided) std::snprintf(ctx.name, ctx.name_len, "sigclip_runs<%dx64,%s,strided>", runs, unit ? "unit" : "scaled");
    if (j.median) {
        std::snprintf(ctx.name, ctx.name_len, "median_runs<%dx64%s>", runs, strided ? ",strided" : "");
        auto launch_median = [&](auto kernel, int smem, int wpb) -> int {
            MB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            const long long ntiles = (j.a.nvox + 31) / 32;
            const int grid = (int)std::max(1LL, std::min((ntiles + wpb - 1) / wpb, (long long)ctx.sm_count));
            if (ctx.slow_n == nullptr) return fail(MPDAF_B200_ERR_ARG, "runs kernel needs the device tile counter");
            MB_CUDA(cudaMemsetAsync(ctx.slow_n, 0, sizeof(unsigned), ctx.stream));
            kernel<<<grid, 32 * wpb, smem, ctx.stream>>>(j.a, (const StackTable*)j.table_dev, rr);
            MB_CUDA(cudaGetLastError());
            ctx.launches++;
            return MPDAF_B200_OK;
        };
        if (runs == 4) return strided ? launch_median(median_runs_kernel<4, true>, RunsCfg<4>::smem, RunsCfg<4>::WPB)
                                      : launch_median(median_runs_kernel<4, false>, RunsCfg<4>::smem, RunsCfg<4>::WPB);
        return strided ? launch_median(median_runs_kernel<8, true>, RunsCfg<8>::smem, RunsCfg<8>::WPB)
                       : launch_median(median_runs_kernel<8, false>, RunsCfg<8>::smem, RunsCfg<8>::WPB);
    }
    auto launch = [&](auto kernel, int smem, const auto& consts, int wpb) -> int {
        MB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const long long ntiles = (j.a.nvox + 31) / 32;
        const int grid = (int)std::max(1LL, std::min((ntiles + wpb - 1) / wpb, (long long)ctx.sm_count));   // one block per SM
        if (ctx.slow_n == nullptr) return fail(MPDAF_B200_ERR_ARG, "runs kernel needs the device tile counter");
        MB_CUDA(cudaMemsetAsync(ctx.slow_n, 0, sizeof(unsigned), ctx.stream));
        kernel<<<grid, 32 * wpb, smem, ctx.stream>>>(j.a, (const StackTable*)j.table_dev, ctx.slow, rr, consts);
        MB_CUDA(cudaGetLastError());
        if (want_clk) {
            long long h[5];
            cudaMemcpyAsync(h, dbg_clk, sizeof h, cudaMemcpyDeviceToHost, ctx.stream);
            cudaStreamSynchronize(ctx.stream);
            std::fprintf(stderr, "runs clocks per tile: gather+pass1+sort %lld, clip %lld, exact %lld, stat+rest %lld (tiles %lld)\n",
                         h[0] / std::max(1LL, h[4]), h[1] / std::max(1LL, h[4]), h[2] / std::max(1LL, h[4]), h[3] / std::max(1LL, h[4]), h[4]);
        }
        ctx.launches++;
        return MPDAF_B200_OK;
    };
    auto fill = [&](auto& k, int nb) {
        for (int i = 0; i < nb; i++) {
            const bool in = i < j.nfiles;
            k.scale[i] = (in && j.scale) ? j.scale[i] : 1.0;
            k.offset[i] = (in && j.offset) ? j.offset[i] : 0.0;
            k.scale2[i] = k.scale[i] * k.scale[i];
        }
    };
    static RunConsts<256> k4;      // launches are serialized by the library's call mutex; the launch copies the parameters
    static RunConsts<512> k8;
    if (runs == 4) fill(k4, 256); else fill(k8, 512);
    if (strided) {
        if (runs == 4) return unit ? launch(sigclip_runs_kernel<4, true, true>, RunsCfg<4>::smem, k4, RunsCfg<4>::WPB) : launch(sigclip_runs_kernel<4, false, true>, RunsCfg<4>::smem, k4, RunsCfg<4>::WPB);
        return unit ? launch(sigclip_runs_kernel<8, true, true>, RunsCfg<8>::smem, k8, RunsCfg<8>::WPB) : launch(sigclip_runs_kernel<8, false, true>, RunsCfg<8>::smem, k8, RunsCfg<8>::WPB);
    }
    if (runs == 4) return unit ? launch(sigclip_runs_kernel<4, true>, RunsCfg<4>::smem, k4, RunsCfg<4>::WPB) : launch(sigclip_runs_kernel<4, false>, RunsCfg<4>::smem, k4, RunsCfg<4>::WPB);
    return unit ? launch(sigclip_runs_kernel<8, true>, RunsCfg<8>::smem, k8, RunsCfg<8>::WPB) : launch(sigclip_runs_kernel<8, false>, RunsCfg<8>::smem, k8, RunsCfg<8>::WPB);
}
