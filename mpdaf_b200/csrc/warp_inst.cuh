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
#include "sigclip_pair.cuh"
#include "sigclip_duo.cuh"
#include "median_pipe.cuh"

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
inline bool encode_stack_map(TensorMap2D& out, const void* base, unsigned long long nvox, int nrows, unsigned long long pitch_bytes,
                             unsigned box_voxels = 32u, int box_rows = 0, bool nan_fill = true, bool f64 = false)
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
    const cuuint32_t box[2] = {box_voxels, (cuuint32_t)(box_rows > 0 ? box_rows : nrows)};
    const cuuint32_t estr[2] = {1u, 1u};
    const CUresult rc = encode(&map, f64 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                               nan_fill ? CU_TENSOR_MAP_FLOAT_OOB_FILL_NAN_REQUEST_ZERO_FMA : CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) return false;
    std::memcpy(&out, &map, sizeof(out));
    return true;
}

// Pair-packed TMA-pipelined sigma clip (sigclip_pair.cuh): float32 / float64 stacks up to 64 deep whose rows
// are equally pitched and 16-byte aligned.  Returns MPDAF_B200_OK and sets `taken` when the stack qualifies.
template <int NB, typename ET = float>
int launch_pair_impl(const Job& j, bool unit, LaunchCtx& ctx, bool& taken)
{
    taken = false;
    constexpr long long ES = (long long)sizeof(ET);
    constexpr bool F64 = sizeof(ET) == 8;
    if constexpr (NB > 64) {
        return MPDAF_B200_OK;
    } else {
        const bool off = std::getenv("MPDAF_B200_NO_PAIR") != nullptr;   // read per call: tests switch kernel families
        const bool need_var = j.a.typ_var == 0;
        if (off || j.median || ctx.slow_q == nullptr || j.a.nvox < 1 || j.a.nvox >= 0x7FFFFF00LL) return MPDAF_B200_OK;
        // MAD: with var = 'propagate' only (sigma then feeds nothing but decisions; the 'stat' variances would need the
        // exact MAD of the last set and stay on the warp kernel, which is bit-identical to the reference there)
        if (j.a.mad && !need_var) return MPDAF_B200_OK;
        auto pitch_of = [&](const void* const* rows) -> unsigned long long {   // elements between rows, 0 = not equally pitched / unaligned
            if (rows == nullptr || (reinterpret_cast<uintptr_t>(rows[0]) & 15u)) return 0ull;
            if (j.nfiles == 1) return (unsigned long long)((j.a.nvox + 3) / 4 * 4);
            const long long step = static_cast<const char*>(rows[1]) - static_cast<const char*>(rows[0]);
            if (step < j.a.nvox * ES || (step & 15) != 0) return 0ull;
            for (int i = 2; i < j.nfiles; i++)
                if (static_cast<const char*>(rows[i]) - static_cast<const char*>(rows[i - 1]) != step) return 0ull;
            return (unsigned long long)(step / ES);
        };
        const unsigned long long dp = pitch_of(j.data), vp = need_var ? pitch_of(j.var) : 0ull;
        if (dp == 0 || (need_var && vp == 0)) return MPDAF_B200_OK;
        if ((reinterpret_cast<uintptr_t>(j.a.out) & 15u) || (reinterpret_cast<uintptr_t>(j.a.outvar) & 15u) || (reinterpret_cast<uintptr_t>(j.a.expmap) & 7u))
            return MPDAF_B200_OK;
        TmaMaps maps;
        std::memset(&maps, 0, sizeof(maps));
        if (!encode_stack_map(maps.data, j.data[0], (unsigned long long)j.a.nvox, j.nfiles, (unsigned long long)ES * dp, 64u, NB, true, F64)) return MPDAF_B200_OK;
        if (need_var && !encode_stack_map(maps.stat, j.var[0], (unsigned long long)j.a.nvox, j.nfiles, (unsigned long long)ES * vp, 64u, NB, false, F64)) return MPDAF_B200_OK;
        PairArgs<NB> s;
        std::memset(&s, 0, sizeof(s));
        s.data0 = j.data[0];
        s.var0 = need_var ? j.var[0] : nullptr;
        s.dstride = dp;
        s.vstride = vp;
        double osmax = 0.0;
        for (int i = 0; i < NB; i++) {
            const bool in = i < j.nfiles;
            const double sc = (in && j.scale) ? j.scale[i] : 1.0, of = (in && j.offset) ? j.offset[i] : 0.0;
            s.os[i] = make_double2(of, sc);
            s.osp[i] = of * sc;
            s.kscale[i] = (float)sc;
            s.koffset[i] = (float)(of * sc);
            s.scale2[i] = sc * sc;
            s.scale2f[i] = (float)(sc * sc);
            osmax = std::max(osmax, std::fabs(of * sc));
        }
        s.osmax = osmax;
        s.hist = ctx.hist;
        ctx.hist_fused = ctx.hist != nullptr;
        using Cfg = PairCfg<NB, sizeof(ET)>;
        std::snprintf(ctx.name, ctx.name_len, "sigclip_pair<%d,%s%s>", NB, unit ? "unit" : "scaled", F64 ? ",f64" : "");
        const long long ntiles = (j.a.nvox + 63) / 64;
        auto launch = [&](auto kernel, auto slow_kernel) -> int {
            const int smem = (int)sizeof(PairSmem<NB, ET>);
            MB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            MB_CUDA(cudaMemsetAsync(ctx.slow_n, 0, sizeof(unsigned), ctx.stream));
            const int grid = (int)std::max(1LL, std::min(ntiles, (long long)ctx.sm_count));
            static const int pair_flags = std::getenv("MPDAF_B200_PAIR_FLAGS") ? std::atoi(std::getenv("MPDAF_B200_PAIR_FLAGS")) : (2 | 256);   // two phase-locked groups, re-aligned every second round
            kernel<<<grid, Cfg::threads, smem, ctx.stream>>>(s, j.a, ntiles, ctx.slow_q, ctx.slow_n, ctx.slow_cap, pair_flags, maps);
            MB_CUDA(cudaGetLastError());
            // queued voxels (uncertifiable from their keys): the reference algorithm restated literally, one thread each
            slow_kernel<<<ctx.sm_count * 8, 128, 0, ctx.stream>>>(s, j.a, ctx.slow_q, ctx.slow_n, ctx.slow_cap, ctx.slow);
            MB_CUDA(cudaGetLastError());
            ctx.launches += 2;
            return MPDAF_B200_OK;
        };
        taken = true;
        if (j.a.mad) {
            std::snprintf(ctx.name, ctx.name_len, "sigclip_pair<%d,%s,mad%s>", NB, unit ? "unit" : "scaled", F64 ? ",f64" : "");
            if (unit) return launch(sigclip_pair_kernel<NB, true, true, true, ET>, sigclip_pair_slow_kernel<NB, true, ET>);
            return launch(sigclip_pair_kernel<NB, false, true, true, ET>, sigclip_pair_slow_kernel<NB, false, ET>);
        }
        if (unit) return need_var ? launch(sigclip_pair_kernel<NB, true, true, false, ET>, sigclip_pair_slow_kernel<NB, true, ET>)
                                  : launch(sigclip_pair_kernel<NB, true, false, false, ET>, sigclip_pair_slow_kernel<NB, true, ET>);
        return need_var ? launch(sigclip_pair_kernel<NB, false, true, false, ET>, sigclip_pair_slow_kernel<NB, false, ET>)
                        : launch(sigclip_pair_kernel<NB, false, false, false, ET>, sigclip_pair_slow_kernel<NB, false, ET>);
    }
}

// Pair kernel with the stack split between two lanes (sigclip_duo.cuh): float32 stacks of 65 .. 128 exposures on equally
// pitched, 16-byte aligned rows (MAD with var = 'propagate' only).
template <int NB>
int launch_duo_impl(const Job& j, bool unit, LaunchCtx& ctx, bool& taken)
{
    taken = false;
    if constexpr (NB != 96 && NB != 128) {
        return MPDAF_B200_OK;
    } else {
        constexpr int HB = NB / 2;
        const bool off = std::getenv("MPDAF_B200_NO_PAIR") != nullptr || std::getenv("MPDAF_B200_NO_DUO") != nullptr;
        const bool need_var = j.a.typ_var == 0;
        if (off || j.median || ctx.slow_q == nullptr || j.a.nvox < 1 || j.a.nvox >= 0x7FFFFF00LL || j.nfiles <= HB) return MPDAF_B200_OK;
        if (j.a.mad && !need_var) return MPDAF_B200_OK;      // MAD: with var = 'propagate' only (as in the pair kernel)
        auto pitch_of = [&](const void* const* rows) -> unsigned long long {
            if (rows == nullptr || (reinterpret_cast<uintptr_t>(rows[0]) & 15u)) return 0ull;
            const long long step = static_cast<const char*>(rows[1]) - static_cast<const char*>(rows[0]);
            if (step < j.a.nvox * 4 || (step & 15) != 0) return 0ull;
            for (int i = 2; i < j.nfiles; i++)
                if (static_cast<const char*>(rows[i]) - static_cast<const char*>(rows[i - 1]) != step) return 0ull;
            return (unsigned long long)(step >> 2);
        };
        const unsigned long long dp = pitch_of(j.data), vp = need_var ? pitch_of(j.var) : 0ull;
        if (dp == 0 || (need_var && vp == 0)) return MPDAF_B200_OK;
        if ((reinterpret_cast<uintptr_t>(j.a.out) & 15u) || (reinterpret_cast<uintptr_t>(j.a.outvar) & 15u) || (reinterpret_cast<uintptr_t>(j.a.expmap) & 7u))
            return MPDAF_B200_OK;
        TmaMaps maps;
        std::memset(&maps, 0, sizeof(maps));
        if (!encode_stack_map(maps.data, j.data[0], (unsigned long long)j.a.nvox, j.nfiles, 4ull * dp, 32u, NB, true)) return MPDAF_B200_OK;
        if (need_var && !encode_stack_map(maps.stat, j.var[0], (unsigned long long)j.a.nvox, j.nfiles, 4ull * vp, 32u, NB, false)) return MPDAF_B200_OK;
        static PairArgs<NB> s;     // (12 KB: kept off the stack; launches are serialized by the library's call mutex)
        std::memset(&s, 0, sizeof(s));
        s.data0 = j.data[0];
        s.var0 = need_var ? j.var[0] : nullptr;
        s.dstride = dp;
        s.vstride = vp;
        double osmax = 0.0;
        for (int i = 0; i < NB; i++) {
            const bool in = i < j.nfiles;
            const double sc = (in && j.scale) ? j.scale[i] : 1.0, of = (in && j.offset) ? j.offset[i] : 0.0;
            s.os[i] = make_double2(of, sc);
            s.osp[i] = of * sc;
            s.kscale[i] = (float)sc;
            s.koffset[i] = (float)(of * sc);
            s.scale2[i] = sc * sc;
            s.scale2f[i] = (float)(sc * sc);
            osmax = std::max(osmax, std::fabs(of * sc));
        }
        s.osmax = osmax;
        s.hist = ctx.hist;
        ctx.hist_fused = ctx.hist != nullptr;
        using Cfg = DuoCfg<HB>;
        std::snprintf(ctx.name, ctx.name_len, "sigclip_duo<%d,%s>", NB, unit ? "unit" : "scaled");
        const long long ntiles = (j.a.nvox + 31) / 32;
        auto launch = [&](auto kernel, auto slow_kernel) -> int {
            const int smem = (int)sizeof(DuoSmem<HB>);
            MB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            MB_CUDA(cudaMemsetAsync(ctx.slow_n, 0, sizeof(unsigned), ctx.stream));
            const int grid = (int)std::max(1LL, std::min(ntiles, (long long)ctx.sm_count));
            static const int pair_flags = std::getenv("MPDAF_B200_PAIR_FLAGS") ? std::atoi(std::getenv("MPDAF_B200_PAIR_FLAGS")) : (2 | 256);
            kernel<<<grid, Cfg::threads, smem, ctx.stream>>>(s, j.a, ntiles, ctx.slow_q, ctx.slow_n, ctx.slow_cap, pair_flags, maps);
            MB_CUDA(cudaGetLastError());
            slow_kernel<<<ctx.sm_count * 8, 128, 0, ctx.stream>>>(s, j.a, ctx.slow_q, ctx.slow_n, ctx.slow_cap, ctx.slow);
            MB_CUDA(cudaGetLastError());
            ctx.launches += 2;
            return MPDAF_B200_OK;
        };
        taken = true;
        if (j.a.mad) {
            std::snprintf(ctx.name, ctx.name_len, "sigclip_duo<%d,%s,mad>", NB, unit ? "unit" : "scaled");
            if (unit) return launch(sigclip_duo_kernel<HB, true, true, true>, sigclip_pair_slow_kernel<NB, true>);
            return launch(sigclip_duo_kernel<HB, false, true, true>, sigclip_pair_slow_kernel<NB, false>);
        }
        if (unit) return need_var ? launch(sigclip_duo_kernel<HB, true, true>, sigclip_pair_slow_kernel<NB, true>)
                                  : launch(sigclip_duo_kernel<HB, true, false>, sigclip_pair_slow_kernel<NB, true>);
        return need_var ? launch(sigclip_duo_kernel<HB, false, true>, sigclip_pair_slow_kernel<NB, false>)
                        : launch(sigclip_duo_kernel<HB, false, false>, sigclip_pair_slow_kernel<NB, false>);
    }
}

// Warp-tiled sigma clip (sigclip_warp.cuh): the main kernel for float32 stacks up to 128 deep.
template <int NB>
int launch_warp_impl(const Job& j, bool unit, LaunchCtx& ctx)
{
    {
        bool taken = false;
        int rc = launch_pair_impl<NB>(j, unit, ctx, taken);
        if (rc != MPDAF_B200_OK || taken) return rc;
        rc = launch_duo_impl<NB>(j, unit, ctx, taken);
        if (rc != MPDAF_B200_OK || taken) return rc;
    }
    TileArgs<NB> s;
    fill_tile_args(s, j);
    const bool full = (j.nfiles == NB);   // every slot in use: the per-row predicates compile away
    const bool mad = j.a.mad != 0;
    std::snprintf(ctx.name, ctx.name_len, "sigclip_warp<%d,%s%s%s>", NB, unit ? "unit" : "scaled", full ? "" : ",partial",
                  mad ? ",mad" : "");
    static const bool have_flags = std::getenv("MPDAF_B200_WARP_FLAGS") != nullptr;
    static const int env_flags = have_flags ? std::atoi(std::getenv("MPDAF_B200_WARP_FLAGS")) : 0;
    int flags = have_flags ? env_flags : (NB >= 24 ? 1 : 0);
    auto launch = [&](auto kernel, int smem, int kWT) -> int {
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
            if (unit) return launch(sigclip_warp_kernel<NB, true, false, true>, smem_mad, WarpThreads<NB, true>::value);
            return launch(sigclip_warp_kernel<NB, false, false, true>, smem_mad, WarpThreads<NB, true>::value);
        } else {
            return fail(MPDAF_B200_ERR_ARG, "stack too deep for the MAD warp kernel");
        }
    }
    constexpr int smem = (int)sizeof(BlockSmem<NB>);
    static const bool no_strided = std::getenv("MPDAF_B200_NO_STRIDED") != nullptr;
    const bool strided = full && !no_strided && s.dstride != 0 && (j.a.typ_var != 0 || s.vstride != 0);
    if (strided) {
        std::snprintf(ctx.name, ctx.name_len, "sigclip_warp<%d,%s,strided>", NB, unit ? "unit" : "scaled");
        if (unit) return launch(sigclip_warp_kernel<NB, true, true, false, true>, smem, WarpThreads<NB, false>::value);
        return launch(sigclip_warp_kernel<NB, false, true, false, true>, smem, WarpThreads<NB, false>::value);
    }
    if (unit) return full ? launch(sigclip_warp_kernel<NB, true, true>, smem, WarpThreads<NB, false>::value) : launch(sigclip_warp_kernel<NB, true, false>, smem, WarpThreads<NB, false>::value);
    return full ? launch(sigclip_warp_kernel<NB, false, true>, smem, WarpThreads<NB, false>::value) : launch(sigclip_warp_kernel<NB, false, false>, smem, WarpThreads<NB, false>::value);
}

template <int NB>
int launch_median_warp_impl(const Job& j, LaunchCtx& ctx)
{
    // TMA-pipelined kernel (median_pipe.cuh) when the rows are equally pitched and 16-byte aligned: 8 .. 48 exposures
    // (profiles/r02_depth_sweep.txt: 4650 / 5530 / 5430 / 5270 / 5530 / 4070 GB/s at N = 8 / 12 / 16 / 24 / 32 / 48 against
    // 4210 / 4500 / 4440 / 4300 / 3990 / 3340 of the register-gather kernel, which stays ahead at N = 4 and from N = 64 on)
    if constexpr (NB >= 8 && NB <= 48) {
        const bool off = std::getenv("MPDAF_B200_NO_PAIR") != nullptr;
        auto pitch_of = [&](const void* const* rows) -> unsigned long long {
            if (rows == nullptr || (reinterpret_cast<uintptr_t>(rows[0]) & 15u)) return 0ull;
            if (j.nfiles == 1) return (unsigned long long)((j.a.nvox + 3) / 4 * 4);
            const long long step = static_cast<const char*>(rows[1]) - static_cast<const char*>(rows[0]);
            if (step < j.a.nvox * 4 || (step & 15) != 0) return 0ull;
            for (int i = 2; i < j.nfiles; i++)
                if (static_cast<const char*>(rows[i]) - static_cast<const char*>(rows[i - 1]) != step) return 0ull;
            return (unsigned long long)(step >> 2);
        };
        const unsigned long long dp = off ? 0ull : pitch_of(j.data);
        TmaMaps maps;
        std::memset(&maps, 0, sizeof(maps));
        if (dp != 0 && j.a.nvox >= 1 && j.a.nvox < 0x7FFFFF00LL && !(reinterpret_cast<uintptr_t>(j.a.out) & 15u) &&
            !(reinterpret_cast<uintptr_t>(j.a.expmap) & 7u) &&
            encode_stack_map(maps.data, j.data[0], (unsigned long long)j.a.nvox, j.nfiles, 4ull * dp, (unsigned)MedianPipeCfg<NB>::TV, NB, true)) {
            // two forms of the pipeline (median_pipe.cuh; both tile widths are the same): a producer lane + 15 consumer
            // warps up to 12 exposures, warps that own and refill their stages from 16 on; MPDAF_B200_MEDIAN_PIPE=1|2 forces one
            const char* force = std::getenv("MPDAF_B200_MEDIAN_PIPE");
            const bool own = force ? std::atoi(force) == 2 : NB > 12;
            static_assert(MedianPipeCfg<NB>::TV == MedianOwnCfg<NB>::TV, "tile width");
            const long long ntiles = (j.a.nvox + MedianPipeCfg<NB>::TV - 1) / MedianPipeCfg<NB>::TV;
            const int grid = (int)std::max(1LL, std::min(ntiles, (long long)ctx.sm_count));
            if (own) {
                using Cfg = MedianOwnCfg<NB>;
                std::snprintf(ctx.name, ctx.name_len, "median_pipe<%d,own>", NB);
                const int smem = (int)sizeof(MedianOwnSmem<NB>);
                MB_CUDA(cudaFuncSetAttribute(median_own_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                median_own_kernel<NB><<<grid, Cfg::threads, smem, ctx.stream>>>(j.a, ntiles, maps);
            } else {
                using Cfg = MedianPipeCfg<NB>;
                std::snprintf(ctx.name, ctx.name_len, "median_pipe<%d>", NB);
                const int smem = (int)sizeof(MedianPipeSmem<NB>);
                MB_CUDA(cudaFuncSetAttribute(median_pipe_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
                median_pipe_kernel<NB><<<grid, Cfg::threads, smem, ctx.stream>>>(j.a, ntiles, maps);
            }
            MB_CUDA(cudaGetLastError());
            ctx.launches++;
            return MPDAF_B200_OK;
        }
    }
    // 49 .. 96 exposures on pitched rows: the two-lane form (median_pipe.cuh, median_duo_kernel): 4275 / 2843 GB/s at N = 64 / 96
    // against 3264 / 2406 of the register-gather kernel, which stays ahead at 128 (2122 against 1931: six warps per SM there)
    if constexpr (NB == 64 || NB == 96) {
        constexpr int HB = NB / 2;
        const bool off = std::getenv("MPDAF_B200_NO_PAIR") != nullptr || std::getenv("MPDAF_B200_NO_DUO") != nullptr;
        unsigned long long dp = 0ull;
        if (!off && j.nfiles > HB && j.data != nullptr && !(reinterpret_cast<uintptr_t>(j.data[0]) & 15u)) {
            const long long step = static_cast<const char*>(j.data[1]) - static_cast<const char*>(j.data[0]);
            bool ok = step >= j.a.nvox * 4 && (step & 15) == 0;
            for (int i = 2; ok && i < j.nfiles; i++) ok = static_cast<const char*>(j.data[i]) - static_cast<const char*>(j.data[i - 1]) == step;
            if (ok) dp = (unsigned long long)(step >> 2);
        }
        TmaMaps maps;
        std::memset(&maps, 0, sizeof(maps));
        if (dp != 0 && j.a.nvox >= 1 && j.a.nvox < 0x7FFFFF00LL && !(reinterpret_cast<uintptr_t>(j.a.out) & 15u) &&
            !(reinterpret_cast<uintptr_t>(j.a.expmap) & 7u) &&
            encode_stack_map(maps.data, j.data[0], (unsigned long long)j.a.nvox, j.nfiles, 4ull * dp, 32u, NB, true)) {
            using Cfg = MedianDuoCfg<HB>;
            std::snprintf(ctx.name, ctx.name_len, "median_duo<%d>", NB);
            const int smem = (int)sizeof(MedianDuoSmem<HB>);
            MB_CUDA(cudaFuncSetAttribute(median_duo_kernel<HB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
            const long long ntiles = (j.a.nvox + 31) / 32;
            const int grid = (int)std::max(1LL, std::min(ntiles, (long long)ctx.sm_count));
            median_duo_kernel<HB><<<grid, Cfg::threads, smem, ctx.stream>>>(j.a, ntiles, maps);
            MB_CUDA(cudaGetLastError());
            ctx.launches++;
            return MPDAF_B200_OK;
        }
    }
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

// float64 samples (BITPIX -64): the pair kernel with float64 stages; `taken` false = the layout does not qualify
#define MB_DEFINE_PAIR64(NB)                                                                                  \
    int mb_launch_pair64_##NB(const mb::Job& j, bool unit, mb::LaunchCtx& ctx, bool& taken) { return mb::launch_pair_impl<NB, double>(j, unit, ctx, taken); }

#define MB_DEFINE_DEPTH(NB)                                                                                   \
    int mb_launch_warp_##NB(const mb::Job& j, bool unit, mb::LaunchCtx& ctx) { return mb::launch_warp_impl<NB>(j, unit, ctx); } \
    int mb_launch_median_warp_##NB(const mb::Job& j, mb::LaunchCtx& ctx) { return mb::launch_median_warp_impl<NB>(j, ctx); }
