// Launchers of the warp-tiled kernels for one stack-depth bucket; included by warp_<NB>.cu.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#include <cuda.h>   // CUtensorMap types only: the encoder is fetched from the driver at run time

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

// 2-D tensor map over a stack of equally pitched float32 rows: dim 0 = voxels, dim 1 = rows, box 32 x nrows,
// NaN beyond the last voxel (so the lanes of a ragged last tile see invalid samples).
inline bool encode_stack_map(TensorMap2D& out, const void* base, unsigned long long nvox, int nrows, unsigned long long pitch_bytes)
{
    using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return reinterpret_cast<EncodeFn>(fn);
    }();
    if (encode == nullptr) return false;
    static_assert(sizeof(CUtensorMap) == sizeof(TensorMap2D), "tensor map image size");
    CUtensorMap map;
    const cuuint64_t dims[2] = {nvox, (cuuint64_t)nrows};
    const cuuint64_t strides[1] = {pitch_bytes};
    const cuuint32_t box[2] = {32u, (cuuint32_t)nrows};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult rc = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA);
    if (rc != CUDA_SUCCESS) return false;
    std::memcpy(&out, &map, sizeof(out));
    return true;
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
    static const bool have_flags = std::getenv("MPDAF_B200_WARP_FLAGS") != nullptr;
    static const int env_flags = have_flags ? std::atoi(std::getenv("MPDAF_B200_WARP_FLAGS")) : 0;
    int flags = have_flags ? env_flags : (NB >= 24 ? 1 : 0);
    TmaMaps maps;
    std::memset(&maps, 0, sizeof(maps));
    auto launch = [&](auto kernel, int smem, int kWT) -> int {
        MB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int per_sm = 0;
        MB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWT, smem));
        if (per_sm < 1) per_sm = 1;
        const long long ntiles = (j.a.nvox + 31) / 32;
        const long long nblocks = (ntiles + (kWT / 32) - 1) / (kWT / 32);
        const int grid = (int)std::max(1LL, std::min(nblocks, (long long)ctx.sm_count * per_sm));
        kernel<<<grid, kWT, smem, ctx.stream>>>(s, j.a, 0LL, ntiles, flags, ctx.slow, maps);
        MB_CUDA(cudaGetLastError());
        ctx.launches++;
        return MPDAF_B200_OK;
    };
    if (mad) {   // MAD: one variant per scale mode (row predicates kept: FULL = false)
        constexpr int smem_mad = (int)sizeof(BlockSmem<NB, true>);
        if constexpr (smem_mad <= 227 * 1024) {
            if (unit) return launch(sigclip_warp_kernel<NB, true, false, true>, smem_mad, WarpThreads<NB, true, false>::value);
            return launch(sigclip_warp_kernel<NB, false, false, true>, smem_mad, WarpThreads<NB, true, false>::value);
        } else {
            return fail(MPDAF_B200_ERR_ARG, "stack too deep for the MAD warp kernel");
        }
    }
    constexpr int smem = (int)sizeof(BlockSmem<NB>);
    static const bool no_strided = std::getenv("MPDAF_B200_NO_STRIDED") != nullptr;
    const bool strided = full && !no_strided && s.dstride != 0 && (j.a.typ_var != 0 || s.vstride != 0);
    // TMA variant: equally pitched rows whose base and pitch are multiples of 16 bytes
    // opt-in (MPDAF_B200_TMA=1): correct and tested, but measured slower than the register gather on B200
    // (1.37 vs 1.02 ms for 64 planes x 48 exposures: its warps wait on the tile barriers)
    static const bool no_tma = std::getenv("MPDAF_B200_TMA") == nullptr;
    if constexpr (NB <= 96) {
        auto tma_ok = [&](const void* const* rows, unsigned stride) {
            return stride != 0 && (stride & 3u) == 0 && (reinterpret_cast<uintptr_t>(rows[0]) & 15u) == 0;
        };
        const bool need_var = j.a.typ_var == 0;
        if (strided && !no_tma && j.a.nvox < 0x7FFFFFE0LL && tma_ok(j.data, s.dstride) && (!need_var || tma_ok(j.var, s.vstride)) &&
            encode_stack_map(maps.data, j.data[0], (unsigned long long)j.a.nvox, NB, 4ull * s.dstride) &&
            (!need_var || encode_stack_map(maps.stat, j.var[0], (unsigned long long)j.a.nvox, NB, 4ull * s.vstride))) {
            constexpr int smem_tma = (int)sizeof(BlockSmem<NB, false, true>);
            if (!have_flags) flags = 0;   // no phase lock: the loop body fits the instruction cache
            std::snprintf(ctx.name, ctx.name_len, "sigclip_warp<%d,%s,tma>", NB, unit ? "unit" : "scaled");
            if (unit) return launch(sigclip_warp_kernel<NB, true, true, false, true, true>, smem_tma, WarpThreads<NB, false, true>::value);
            return launch(sigclip_warp_kernel<NB, false, true, false, true, true>, smem_tma, WarpThreads<NB, false, true>::value);
        }
    }
    if (strided) {
        std::snprintf(ctx.name, ctx.name_len, "sigclip_warp<%d,%s,strided>", NB, unit ? "unit" : "scaled");
        if (unit) return launch(sigclip_warp_kernel<NB, true, true, false, true>, smem, WarpThreads<NB, false, false>::value);
        return launch(sigclip_warp_kernel<NB, false, true, false, true>, smem, WarpThreads<NB, false, false>::value);
    }
    if (unit) return full ? launch(sigclip_warp_kernel<NB, true, true>, smem, WarpThreads<NB, false, false>::value) : launch(sigclip_warp_kernel<NB, true, false>, smem, WarpThreads<NB, false, false>::value);
    return full ? launch(sigclip_warp_kernel<NB, false, true>, smem, WarpThreads<NB, false, false>::value) : launch(sigclip_warp_kernel<NB, false, false>, smem, WarpThreads<NB, false, false>::value);
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
