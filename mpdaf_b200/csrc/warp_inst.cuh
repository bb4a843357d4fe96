// Launchers of the warp-tiled kernels for one stack-depth bucket; included by warp_<NB>.cu.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "launch.h"
#include "sigclip_warp.cuh"

namespace mb {

template <int NB>
void fill_tile_args(TileArgs<NB>& s, const Job& j)
{
    double maxos = 0.0;
    for (int i = 0; i < NB; i++) {
        const bool in = i < j.nfiles;
        s.data[i] = in ? static_cast<const float*>(j.data[i]) : nullptr;
        s.var[i] = (in && j.var) ? static_cast<const float*>(j.var[i]) : nullptr;
        s.scale[i] = (in && j.scale) ? j.scale[i] : 1.0;
        s.offset[i] = (in && j.offset) ? j.offset[i] : 0.0;
        s.scale2[i] = s.scale[i] * s.scale[i];
        s.os[i] = make_double2(s.offset[i], s.scale[i]);
        s.kscale[i] = (float)s.scale[i];
        s.koffset[i] = (float)(s.offset[i] * s.scale[i]);
        s.scale2f[i] = (float)s.scale2[i];
        maxos = std::max(maxos, std::fabs(s.offset[i] * s.scale[i]));
    }
    // key = fma(d, kscale, koffset) in float32 vs float32((offset + d) * scale): the three float32 roundings
    // and the two rounded constants are each <= 2^-24 of (|w| + 2 |offset * scale|); the |w| part is inside
    // the relative key margin (KEPS), the offset part is this absolute term (x2 for slack)
    s.koff = 0x1p-20 * maxos + 0x1p-120;
    // equally spaced rows (one [N][V] allocation): the STRIDED kernels step from row to row
    auto stride_of = [&](const void* const* rows) -> unsigned {
        if (rows == nullptr || j.nfiles != NB || NB < 2) return 0u;
        const long long step = static_cast<const char*>(rows[1]) - static_cast<const char*>(rows[0]);
        if (step <= 0 || (step & 3) != 0 || (step >> 2) > 0xFFFFFFFFLL) return 0u;
        for (int i = 2; i < NB; i++)
            if (static_cast<const char*>(rows[i]) - static_cast<const char*>(rows[i - 1]) != step) return 0u;
        return (unsigned)(step >> 2);
    };
    s.dstride = stride_of(j.data);
    s.vstride = stride_of(j.var);
}

// Warp-tiled sigma clip (sigclip_warp.cuh): the main kernel for float32 stacks up to 128 deep.
template <int NB>
int launch_warp_impl(const Job& j, bool unit, LaunchCtx& ctx)
{
    TileArgs<NB> s;
    fill_tile_args(s, j);
    const bool full = (j.nfiles == NB);   // every slot in use: the per-row predicates compile away
    const bool mad = j.a.mad != 0;
    std::snprintf(ctx.name, ctx.name_len, "sigclip_warp<%d,%s%s%s>", NB, unit ? "unit" : "scaled", full ? "" : ",partial",
                  mad ? ",mad" : "");
    static const int flags = std::getenv("MPDAF_B200_WARP_FLAGS") ? std::atoi(std::getenv("MPDAF_B200_WARP_FLAGS")) : (NB >= 24 ? 1 : 0);
    auto launch = [&](auto kernel, int smem) -> int {
        MB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int per_sm = 0;
        MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWT, smem));
        if (per_sm < 1) per_sm = 1;
        const long long ntiles = (j.a.nvox + 31) / 32;
        const long long nblocks = (ntiles + (kWT / 32) - 1) / (kWT / 32);
        const int grid = (int)std::max(1LL, std::min(nblocks, (long long)ctx.sm_count * per_sm));
        kernel<<<grid, kWT, smem, ctx.stream>>>(s, j.a, 0LL, ntiles, flags, ctx.slow);
        MB_CUDA(cudaGetLastError());
        ctx.launches++;
        return MPDAF_B200_OK;
    };
    if (mad) {   // MAD: one variant per scale mode (row predicates kept: FULL = false)
        constexpr int smem_mad = (int)sizeof(BlockSmem<NB, true>);
        if constexpr (smem_mad <= 227 * 1024) {
            if (unit) return launch(sigclip_warp_kernel<NB, true, false, true>, smem_mad);
            return launch(sigclip_warp_kernel<NB, false, false, true>, smem_mad);
        } else {
            return fail(MPDAF_B200_ERR_ARG, "stack too deep for the MAD warp kernel");
        }
    }
    constexpr int smem = (int)sizeof(BlockSmem<NB>);
    static const bool no_strided = std::getenv("MPDAF_B200_NO_STRIDED") != nullptr;
    const bool strided = full && !no_strided && s.dstride != 0 && (j.a.typ_var != 0 || s.vstride != 0);
    if (strided) {
        std::snprintf(ctx.name, ctx.name_len, "sigclip_warp<%d,%s,strided>", NB, unit ? "unit" : "scaled");
        if (unit) return launch(sigclip_warp_kernel<NB, true, true, false, true>, smem);
        return launch(sigclip_warp_kernel<NB, false, true, false, true>, smem);
    }
    if (unit) return full ? launch(sigclip_warp_kernel<NB, true, true>, smem) : launch(sigclip_warp_kernel<NB, true, false>, smem);
    return full ? launch(sigclip_warp_kernel<NB, false, true>, smem) : launch(sigclip_warp_kernel<NB, false, false>, smem);
}

template <int NB>
int launch_median_warp_impl(const Job& j, LaunchCtx& ctx)
{
    TileArgs<NB> s;
    fill_tile_args(s, j);
    const bool full = (j.nfiles == NB);
    std::snprintf(ctx.name, ctx.name_len, "median_warp<%d%s>", NB, full ? "" : ",partial");
    auto launch = [&](auto kernel) -> int {
        int per_sm = 0;
        MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kThreads, 0));
        if (per_sm < 1) per_sm = 1;
        const long long nblocks = ((j.a.nvox + 31) / 32 + (kThreads / 32) - 1) / (kThreads / 32);
        const int grid = (int)std::max(1LL, std::min(nblocks, (long long)ctx.sm_count * per_sm));
        kernel<<<grid, kThreads, 0, ctx.stream>>>(s, j.a);
        MB_CUDA(cudaGetLastError());
        ctx.launches++;
        return MPDAF_B200_OK;
    };
    return full ? launch(median_warp_kernel<NB, true>) : launch(median_warp_kernel<NB, false>);
}

}  // namespace mb

#define MB_DEFINE_DEPTH(NB)                                                                                   \
    int mb_launch_warp_##NB(const mb::Job& j, bool unit, mb::LaunchCtx& ctx) { return mb::launch_warp_impl<NB>(j, unit, ctx); } \
    int mb_launch_median_warp_##NB(const mb::Job& j, mb::LaunchCtx& ctx) { return mb::launch_median_warp_impl<NB>(j, ctx); }
