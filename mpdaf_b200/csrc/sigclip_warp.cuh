// Warp-tiled sigma-clip kernel for sm_100a: the main combine kernel (float32 stacks <= 128 deep).
//
// Reference behaviour restated (never copied): src/merging.c:423-477 (gather, n == 0 / 1,
// outputs, counters) and src/tools.c:82-118 (median-centred iterative clip).
//
// Work decomposition.  A warp owns tiles of 32 consecutive voxels; one lane owns one voxel.  Warps share no
// data; all resident warps of an SM form one block and (for stacks of 24 and more) start every tile together,
// because the unrolled loop body is larger than the instruction cache (see WarpThreads below).  Every
// exposure load is one coalesced 128-byte line per warp; the rows of the next tile are prefetched into L2 a
// whole tile ahead (two prefetches per line), the STAT rows of the current tile after its sort.
//
// Per voxel:
//  1. exposure order, float64, no FMA: w = (offset + d) * scale and sum(w); the first mean is bit-identical
//     to the reference (tools.c:13-17).  Alongside, shifted moments T1 = sum(w - p), S2 = sum((w - p)^2)
//     about a pivot sample p.  Invalid (NaN) samples cost selects, not branches.
//  2. 32-bit keys: float32(w) with the low 6 (7) mantissa bits replaced by the exposure index, ordered by a
//     register sorting network.  A compare-exchange is FMNMX (min) plus integer a + b - min (the keys are
//     distinct bit patterns, so that is exactly the max).  Invalid samples are sentinel keys below every
//     real key, so the largest samples sit in fixed registers after the sort.
//  3. the exact values of the two largest samples are looked up speculatively; the sorted keys go to the
//     lane's shared-memory column; the median and four probe keys at each end of the clip window are read
//     from it (a binary search only when four or more samples fall on one side).
//  4. later means / sigmas: the rejected samples (exact w from the speculative values or re-read through
//     L1 / L2) are subtracted from the shifted moments.
//  5. propagated variance: STAT rows loaded after the clip loop, float32 FMAs in short chains.
//  6. per-exposure counters: the lanes' "invalid" and "dropped" bit masks are transposed across the warp
//     with 5 shuffle stages per 32 exposures, then one POPC per lane.
// Decisions are CERTIFIED: key-derived median and moment-derived sigma carry rigorous error
// bounds; a voxel with a sample closer to a clip bound than that uncertainty, or with an ill-conditioned
// variance, is recomputed by `exact_voxel_w` (the reference algorithm restated literally), so expmap and the
// counters are always bit-exact.
// Tuning builds: -DMB_TIMING=1 (per-phase clock totals), -DMB_ABLATE=1 (memory ablation flags), see tools/variant.py.
#pragma once

#include "combine_kernels.cuh"
#include "stack_common.cuh"

#ifndef MB_WARP_THREADS
#define MB_WARP_THREADS 0     // 0: per-depth default (WarpThreads below); otherwise threads per block of the warp kernel
#endif
#ifndef MB_SORT_FMA_EVERY
#define MB_SORT_FMA_EVERY 0   // every k-th compare-exchange computes its max with 2 IMAD (FMA pipe) instead of IADD3
#endif
#ifndef MB_EXACT_SUM
#define MB_EXACT_SUM 1   // 1: also carry sum(w) in exposure order, so an unclipped mean is the reference's to the bit
#endif
#ifndef MB_KEY_FFMA
#define MB_KEY_FFMA 0    // 1: key = float32 FMA of the raw sample (FMA pipe + 2 constant loads) instead of float32(w) (XU pipe)
#endif
#ifndef MB_PREFETCH_FULL_LINE
#define MB_PREFETCH_FULL_LINE 1   // prefetch both 64-byte halves of every line
#endif
#ifndef MB_TIMING
#define MB_TIMING 0      // 1 (tuning builds only): per-phase clock64 totals of blocks 0 and 400, printed at the end
#endif
#if MB_TIMING
#define MB_STAMP(k) do { const long long c_ = clock64(); if (lane == 0) sm.tph[wib][k] += (unsigned long long)(c_ - tprev); tprev = c_; } while (0)
#define MB_TOUCH(x) do { unsigned u_; asm volatile("mov.b32 %0, %1;" : "=r"(u_) : "f"(x)); } while (0)
#else
#define MB_STAMP(k) do { } while (0)
#define MB_TOUCH(x) do { } while (0)
#endif
#ifndef MB_ABLATE
#define MB_ABLATE 0      // 1 (tuning builds only): flags 32 / 64 / 128 replace memory traffic by L1-resident accesses
#endif
#ifndef MB_P1_CHAINS
#define MB_P1_CHAINS 1
#endif
#ifndef MB_VAR_F32
#define MB_VAR_F32 1     // 1: propagated variance accumulated by float32 FMAs in short chains (<= 4e-7 relative for N <= 64)
#endif
#ifndef MB_LATE_STAT
#define MB_LATE_STAT 1   // 1: load the STAT rows after the clip loop (fewer registers, more blocks per SM; +7 %)
#endif
#ifndef MB_WARP_BLOCKS
#define MB_WARP_BLOCKS 6   // resident 128-thread blocks per SM for stacks up to 48 deep (80 registers)
#endif

namespace mb {

// Key layout: float32(w) with the low kIdxBits mantissa bits replaced by the exposure index.
template <int NB> struct KeyBits { static constexpr int idx_bits = NB <= 64 ? 6 : 7; };
template <int NB> __device__ __forceinline__ constexpr unsigned idx_mask() { return (1u << KeyBits<NB>::idx_bits) - 1u; }

// Per-voxel exposure set (bit i = exposure i), 32-bit words so that constant-index updates stay
// single LOP3s for any depth up to 128.
template <int NB>
struct ExpSet {
    static constexpr int W = (NB + 31) / 32;
    unsigned w[W];
    __device__ __forceinline__ ExpSet()
    {
#pragma unroll
        for (int k = 0; k < W; k++) w[k] = 0u;
    }
    __device__ __forceinline__ void set(int i) { w[i >> 5] |= 1u << (i & 31); }              // i constant after unrolling
    __device__ __forceinline__ bool test(int i) const { return (w[i >> 5] >> (i & 31)) & 1u; }
    __device__ __forceinline__ bool test_dyn(int i) const
    {
        unsigned x = 0u;
#pragma unroll
        for (int k = 0; k < W; k++) x = ((i >> 5) == k) ? w[k] : x;
        return (x >> (i & 31)) & 1u;
    }
    __device__ __forceinline__ void set_dyn(int i)
    {
#pragma unroll
        for (int k = 0; k < W; k++) w[k] |= ((i >> 5) == k) ? (1u << (i & 31)) : 0u;
    }
    __device__ __forceinline__ int count() const
    {
        int c = 0;
#pragma unroll
        for (int k = 0; k < W; k++) c += __popc(w[k]);
        return c;
    }
    __device__ __forceinline__ ExpSet operator|(const ExpSet& o) const
    {
        ExpSet r;
#pragma unroll
        for (int k = 0; k < W; k++) r.w[k] = w[k] | o.w[k];
        return r;
    }
};

template <int NB>
struct VoxelResultN {
    double mean;
    double sigma;
    int kept;
    ExpSet<NB> rejected;   // exposure valid but clipped
};

template <int NB>
struct WarpSmem {
    float k[NB][32];   // sorted keys, one column per lane
};

constexpr int kWT = 128;   // block size of the deep-stack kernel (sigclip_runs.cuh)

// Block size of the warp kernel.  One block per SM holding all its resident warps (1024 threads at 64 registers
// for depths <= 24, 768 at 80 for depths <= 48): the per-tile barrier then keeps every warp of the SM in the same part of the ~45 KB loop
// body, which exceeds the 32 KB instruction cache (measured +8 % over 128-thread blocks at N = 48, +11 % at
// N = 12).  The MAD variant is shared-memory bound and stays at 128 threads.
template <int NB, bool MAD>
struct WarpThreads {
    static constexpr int value = MB_WARP_THREADS ? MB_WARP_THREADS
                                   : (MAD ? 128 : (NB <= 24 ? 1024 : (NB <= 48 ? 768 : (NB <= 64 ? 512 : (NB <= 96 ? 384 : 256)))));
};

// 128-byte opaque image of a CUtensorMap (filled on the host by cuTensorMapEncodeTiled)
struct alignas(64) TensorMap2D { unsigned long long opaque[16]; };
struct TmaMaps { TensorMap2D data, stat; };

__device__ __forceinline__ void tma_load_2d(void* dst, const TensorMap2D* map, int c0, int c1, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d(const TensorMap2D* map, int c0, int c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(map), "r"(c0), "r"(c1) : "memory");
}

template <int NB, bool MAD = false>
struct BlockSmem {
    static constexpr int kWT = WarpThreads<NB, MAD>::value;
    WarpSmem<NB> w[kWT / 32];
    double scale[NB];
    double offset[NB];
    double rcp[NB + 1];   // 1.0 / m
    const float* data[NB];   // exposure base pointers (lane-dependent lookups of rejected samples, prefetch)
    const float* var[NB];
    double sw[MAD ? kWT / 32 : 1][MAD ? NB : 1][32];   // MAD: exact w in ascending order, one column per lane
#if MB_TIMING
    unsigned long long tph[kWT / 32][12];
#endif
};

// compare-exchange variants: the max as a + b - min on the ALU (IADD3) or on the FMA pipe (2 IMAD)
__device__ __forceinline__ void key_ce_alu(float& a, float& b)
{
    const float lo = fminf(a, b);
    const unsigned hi = __float_as_uint(a) + __float_as_uint(b) - __float_as_uint(lo);
    a = lo;
    b = __uint_as_float(hi);
}
__device__ __forceinline__ void key_ce_fma(float& a, float& b, unsigned one, unsigned minus_one)
{
    const float lo = fminf(a, b);
    unsigned t, hi;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(t) : "r"(__float_as_uint(a)), "r"(one), "r"(__float_as_uint(b)));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(__float_as_uint(lo)), "r"(minus_one), "r"(t));
    a = lo;
    b = __uint_as_float(hi);
}

template <int NB, int FMA_EVERY>
__device__ __forceinline__ void sort_keys_mixed(float (&a)[NB], unsigned one, unsigned minus_one)
{
    int ce = 0;
#define MB_CE(i, j)                                                              \
    if (FMA_EVERY > 0 && (ce % FMA_EVERY) == 0) key_ce_fma(a[i], a[j], one, minus_one); \
    else key_ce_alu(a[i], a[j]);                                                 \
    ce++;
    if constexpr (NB == 4) { MB_SORTNET_4(MB_CE) }
    else if constexpr (NB == 8) { MB_SORTNET_8(MB_CE) }
    else if constexpr (NB == 12) { MB_SORTNET_12(MB_CE) }
    else if constexpr (NB == 16) { MB_SORTNET_16(MB_CE) }
    else if constexpr (NB == 24) { MB_SORTNET_24(MB_CE) }
    else if constexpr (NB == 32) { MB_SORTNET_32(MB_CE) }
    else if constexpr (NB == 48) { MB_SORTNET_48(MB_CE) }
    else if constexpr (NB == 64) { MB_SORTNET_64(MB_CE) }
    else if constexpr (NB == 96) { MB_SORTNET_96(MB_CE) }
    else if constexpr (NB == 128) { MB_SORTNET_128(MB_CE) }
#undef MB_CE
}

// The reference algorithm restated literally for one voxel (rare fallback path).
template <int NB, bool UNIT>
__device__ __noinline__ VoxelResultN<NB> exact_voxel_w(const float* const* data, long long v, const double* scale,
                                                       const double* offset, int nfiles, int nmax, double lo, double up,
                                                       int nstop)
{
    GenericScratch<NB> g;
    int n = 0;
    for (int i = 0; i < nfiles; i++) {
        const float x = __ldg(data[i] + v);   // the tile's lines are still in L1 / L2
        if (x == x) {
            g.w[n] = UNIT ? (double)x : (offset[i] + (double)x) * scale[i];
            g.file[n] = (unsigned short)i;
            g.idx[n] = (unsigned short)n;
            n++;
        }
    }
    VoxelResultN<NB> r;
    r.mean = CUDART_NAN;
    r.sigma = 0.0;
    r.kept = n;
    if (n == 1) r.mean = g.w[0];
    if (n >= 2) {
        r.kept = clip_generic<NB>(g, n, nmax, lo, up, nstop, 0, r.mean, r.sigma);
        ExpSet<NB> keep, valid;
        for (int k = 0; k < r.kept; k++) keep.set_dyn(g.file[g.idx[k]]);
        for (int k = 0; k < n; k++) valid.set_dyn(g.file[k]);
#pragma unroll
        for (int k = 0; k < ExpSet<NB>::W; k++) r.rejected.w[k] = valid.w[k] & ~keep.w[k];
    }
    return r;
}

// Branch-free lower bound over the lane's sorted column: number of leading keys with
// key <= x (LE) or key < x (!LE).  Every probe index is in range for any NB.
template <int NB, bool LE>
__device__ __forceinline__ int count_col(const float (*col)[32], int lane, float x)
{
    int base = 0;
#pragma unroll
    for (int len = NB; len > 1; len -= len / 2) {
        const int half = len / 2;
        const float kv = col[base + half - 1][lane];
        const bool take = LE ? (kv <= x) : (kv < x);
        base += take ? half : 0;
    }
    const float kv = col[base][lane];
    return base + ((LE ? (kv <= x) : (kv < x)) ? 1 : 0);
}

// Rare path: a key lies within the key-space uncertainty of a clip bound.  Recover the exact
// median from the exposure indices carried by the median keys (only possible when their
// truncated-key buckets hold a single sample), recompute the bounds, and classify the few band
// keys from their exact float64 values.  Returns false when even that cannot decide (then the
// voxel goes to exact_voxel_w).  On success lo_in / up_in are the exact counts of keys at or
// below the low bound / strictly below the high bound (positions in the whole column).
template <int NB, bool UNIT>
__device__ __noinline__ bool refine_band(const float (*sk)[32], int lane, const float* const* data, long long v,
                                         const double* scale, const double* offset, int nfiles, int ih, int il, int m,
                                         double sigma, double nlo, double nup, double clo, double cup, double T,
                                         int& lo_in, int& up_in)
{
    auto bucket = [](float k) { return __float_as_uint(k) & ~idx_mask<NB>(); };
    auto exact_w = [&](float k) {
        const int i = min((int)(__float_as_uint(k) & idx_mask<NB>()), nfiles - 1);
        const double dd = (double)__ldg(data[i] + v);
        return UNIT ? dd : (offset[i] + dd) * scale[i];
    };
    const float kh = sk[ih][lane], kl = sk[il][lane];
    if (ih + 1 < NB && bucket(sk[ih + 1][lane]) == bucket(kh)) return false;
    if (il > 0 && bucket(sk[il - 1][lane]) == bucket(kl)) return false;
    if (il != ih && bucket(kl) == bucket(kh)) return false;
    const double wh = exact_w(kh), wl = exact_w(kl);
    const double med = (m & 1) ? wh : (wh + wl) * 0.5;          // tools.c:45-48
    const double clo_x = med - (nlo * sigma);
    const double cup_x = med + (nup * sigma);
    // what is left is the sigma uncertainty (2^-28 relative) and the final roundings
    const double T2 = 0x1p-26 * (nlo + nup) * sigma + 0x1p-44 * (fabs(clo_x) + fabs(cup_x)) + 0x1p-300;
    const int lo_out = count_col<NB, true>(sk, lane, __double2float_rd(clo - T));
    const int lo_in0 = count_col<NB, true>(sk, lane, __double2float_ru(clo + T));
    const int up_in0 = count_col<NB, false>(sk, lane, __double2float_rd(cup - T));
    const int up_out = count_col<NB, false>(sk, lane, __double2float_ru(cup + T));
    if (lo_in0 > up_in0) return false;                           // bands overlap: degenerate window
    int first_inside = lo_in0;
    bool seen = false;
    for (int q = lo_out; q < lo_in0; q++) {                      // low band: outside ... then inside ...
        const double w = exact_w(sk[q][lane]);
        if (fabs(w - clo_x) <= T2) return false;
        const bool inside = w > clo_x;
        if (inside && !seen) { first_inside = q; seen = true; }
        if (!inside && seen) return false;
    }
    int first_outside = up_out;
    seen = false;
    for (int q = up_in0; q < up_out; q++) {                      // high band: inside ... then outside ...
        const double w = exact_w(sk[q][lane]);
        if (fabs(w - cup_x) <= T2) return false;
        const bool inside = w < cup_x;
        if (!inside && !seen) { first_outside = q; seen = true; }
        if (inside && seen) return false;
    }
    lo_in = first_inside;
    up_in = first_outside;
    return true;
}

// 32x32 bit-matrix transpose across the warp: on return bit l of lane i = bit i of lane l's x.
struct TransposeConsts {
    unsigned sel[5];   // per-stage bit-select mask for this lane
    unsigned rot[5];   // per-stage left-rotate amount for this lane
    __device__ __forceinline__ void init(int lane)
    {
        const unsigned m[5] = {0x0000FFFFu, 0x00FF00FFu, 0x0F0F0F0Fu, 0x33333333u, 0x55555555u};
#pragma unroll
        for (int s = 0; s < 5; s++) {
            const int j = 16 >> s;
            const bool upper = (lane & j) != 0;
            sel[s] = upper ? ~m[s] : m[s];
            rot[s] = upper ? (unsigned)(32 - j) : (unsigned)j;
        }
    }
    __device__ __forceinline__ unsigned apply(unsigned x) const
    {
#pragma unroll
        for (int s = 0; s < 5; s++) {
            const unsigned y = __shfl_xor_sync(0xffffffffu, x, 16 >> s);
            const unsigned t = __funnelshift_l(y, y, rot[s]);   // rotate left
            x = (x & sel[s]) | (t & ~sel[s]);
        }
        return x;
    }
};

// resident 128-thread blocks per SM each depth is compiled for (register budget 65536 / (128 * blocks))
template <int NB, int THREADS> struct WarpBlocks {
    static constexpr int raw = (NB <= 48 ? MB_WARP_BLOCKS : (NB <= 64 ? 4 : (NB <= 96 ? 3 : 2))) * 128 / THREADS;
    static constexpr int value = raw < 1 ? 1 : raw;
};
// FULL: every exposure slot is in use (nfiles == NB), so the per-row `i < nfiles` predicates and
// selects compile away (the common case: depths 4, 8, 12, 16, 24, 32, 48, 64, 96, 128).
//
// MAD (tools.c:122-158, :54-77): sigma = 1.4826 * median(|w - median(w)|).  The sorted keys give the
// order of the samples up to key-bucket ties; the exact float64 values are then written in that
// order to a second shared-memory column (ties repaired by a local insertion), and from there the
// reference algorithm runs literally: exact medians, the MAD as the middle of the two sorted runs
// of absolute deviations left and right of the median (outward merge), exact strict window counts
// by binary search on the float64 column, ascending-order sums for the later means.  No error
// bounds are involved in this mode: every output is the reference's, bit for bit.
// STRIDED: the rows of the stack are equally spaced in memory (one [N][V] allocation, the layout of the
// slab slots and of any stacked tensor): row i+1 = row i + S.dstride elements, so an address costs one
// IMAD.WIDE on the FMA pipe instead of a constant-bank load plus a 64-bit add on the ALU pipe.
template <int NB, bool UNIT, bool FULL, bool MAD = false, bool STRIDED = false>
__global__ void __launch_bounds__((WarpThreads<NB, MAD>::value), (WarpBlocks<NB, WarpThreads<NB, MAD>::value>::value))
sigclip_warp_kernel(const TileArgs<NB> S, const ClipArgs A, const long long tile_begin, const long long tile_end,
                    const int flags, unsigned long long* slow_total)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    BlockSmem<NB, MAD>& sm = *reinterpret_cast<BlockSmem<NB, MAD>*>(smem_raw);
    constexpr int kWT = WarpThreads<NB, MAD>::value;
    constexpr int W = ExpSet<NB>::W;
    constexpr unsigned IDX = (1u << KeyBits<NB>::idx_bits) - 1u;
    // Invalid samples become sentinel keys below every real key (above, for the MAD variant, whose code
    // addresses the valid samples as [0, n)): the valid samples then end at the fixed slot NB - 1, so the
    // largest samples -- the usual rejects -- are known by register name right after the sort.
    constexpr bool SENT_LOW = !MAD;
    constexpr unsigned SENT = (SENT_LOW ? 0xFF7FFFFFu : 0x7F7FFFFFu) & ~IDX;
    // key-space error model: float32 rounding + truncation of idx_bits mantissa bits is below
    // 2^-(23 - idx_bits) * 1.01 relative; the margins use four times that
    constexpr double KEPS = NB <= 64 ? 0x1p-15 : 0x1p-14;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wib = tid >> 5;
    const int nfiles = FULL ? NB : A.nfiles;
    const int nstop = A.nstop < 1 ? 1 : A.nstop;
    const bool want_var = (A.typ_var == 0);
    const unsigned one = (unsigned)(A.nmax > -1000000);    // opaque 1 / -1 for the IMAD compare-exchange form
    const unsigned minus_one = 0u - one;
    // ~IDX in a register the compiler cannot fold: (key & mask) | index is then ONE three-input LOP3
    const unsigned keep_mask = ~IDX ^ (unsigned)(A.nmax < -1000000);

    for (int i = tid; i < NB; i += kWT) {
        sm.scale[i] = S.os[i].y;
        sm.offset[i] = S.os[i].x;
        sm.data[i] = S.data[i];
        sm.var[i] = S.var[i];
    }
    for (int i = tid; i <= NB; i += kWT) sm.rcp[i] = i > 0 ? 1.0 / (double)i : 0.0;
#if MB_TIMING
    for (int i = tid; i < (kWT / 32) * 12; i += kWT) sm.tph[i / 12][i % 12] = 0ull;
    long long tprev = 0;
#endif
    __syncthreads();   // the only block-wide barrier of the kernel (unless flags & 1)
#if MB_TIMING
    tprev = clock64();
#endif

    float (*sk)[32] = sm.w[wib].k;
    TransposeConsts tc;
    tc.init(lane);

    // per-lane exposure counters: lane l counts exposure 32*k + l (invalid / dropped samples)
    unsigned cnt_inval[W], cnt_drop[W];
#pragma unroll
    for (int k = 0; k < W; k++) { cnt_inval[k] = 0u; cnt_drop[k] = 0u; }
    unsigned tiles_done = 0, slow_count = 0;

    const long long ntiles = tile_end;
    const long long wstride = (long long)gridDim.x * (kWT / 32);
    long long tile0 = tile_begin + (long long)blockIdx.x * (kWT / 32);

    // prefetch of a tile's rows: lane l touches rows l, l + 32, ... (one 128-byte line each);
    // into L1 by default (measured 12 % faster than L2 only; flags & 2 selects L2)
    const bool l1 = (flags & 2) == 0;
    // A prefetch fetches 64 bytes (the default L2 fetch granularity: measured as a 50 % L2 hit rate of the
    // gather when one address per line was prefetched), so each 128-byte line takes two: slot j = row * 2 + half.
    auto prefetch_rows = [&](bool stat, long long t) {
        const long long v0 = t << 5;
        if (v0 + 31 >= A.nvox) return;
        constexpr int HALVES = MB_PREFETCH_FULL_LINE ? 2 : 1;
#pragma unroll
        for (int r = 0; r < (HALVES * NB + 31) / 32; r++) {
            const int j = lane + 32 * r;
            const int i = HALVES == 2 ? (j >> 1) : j;
            if (i < nfiles) {
                // lane-dependent row: never index the constant bank with it (a divergent LDC replays once per
                // distinct address); equally spaced rows are computed, the others come from shared memory
                const float* row;
                if constexpr (STRIDED) row = stat ? S.var[0] + (size_t)i * S.vstride : S.data[0] + (size_t)i * S.dstride;
                else row = stat ? sm.var[i] : sm.data[i];
                row += v0 + (HALVES == 2 ? (j & 1) * 16 : 0);
                if (l1) asm volatile("prefetch.global.L1 [%0];" ::"l"(row));
                else asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
            }
        }
    };
    // (prefetch.global.L1 only reaches L2 on this part -- tools/microbench/l1prefetch.cu: a later load still
    // takes the ~480-cycle L2 latency.  Real L1 fills, cp.async.ca of one word per sector into a dummy
    // shared-memory word, make the later load a 33-cycle L1 hit but measured 10-18 % slower here.)
    {
        if (tile0 + wib < ntiles) prefetch_rows(false, tile0 + wib);
        if (want_var && (flags & 16) && tile0 + wib < ntiles) prefetch_rows(true, tile0 + wib);
    }

    // The trip count is uniform per block so that the optional per-tile barrier is legal; a warp
    // whose last tile lies beyond the range runs it with every lane idle.
    for (; tile0 < ntiles; tile0 += wstride) {
        // flags & 1 (experiment, slower): the four warps of a block start every tile together
        if (flags & 1) __syncthreads();
        MB_STAMP(0);   // barrier wait
        const long long tile = tile0 + wib;
        const long long v = (tile << 5) + lane;
        const bool act = tile < ntiles && v < A.nvox;
        if (want_var && (flags & 8)) prefetch_rows(true, tile);   // flags & 8: STAT prefetch at tile start (3 % slower)

        float key[NB];
        ExpSet<NB> inval;
        double sum = 0.0, p = 0.0, T1 = 0.0, S2 = 0.0;
        {
            // ---- gather: one coalesced load per exposure ------------------------------------------
            // STRIDED: idle lanes (ragged last tile, tiles past the end) re-read the last voxel and are voided
            // after pass 1, so these loads need neither a predicate nor a NaN initialisation per row
            const long long vc = (STRIDED && !act) ? A.nvox - 1 : v;
#if MB_ABLATE   // tuning builds: flags & 32 gathers from an L1-resident window (LSU work without L2 / DRAM)
            const float* rp = (flags & 32) ? S.data[0] + (vc & 1023) : S.data[0] + vc;
            const unsigned gstride = (flags & 32) ? 32u : S.dstride;
#else
            const float* rp = S.data[0] + vc;
            const unsigned gstride = S.dstride;
#endif
#pragma unroll
            for (int i = 0; i < NB; i++) {
                // predicated loads: ptxas keeps them batched at the top of the tile (unconditional
                // ones get sunk next to their first use and their latency serialises)
                if constexpr (STRIDED) {
                    asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(key[i]) : "l"(rp));
                    asm("mad.wide.u32 %0, %1, 4, %0;" : "+l"(rp) : "r"(gstride));
                } else {
                    key[i] = CUDART_NAN_F;
                    if (act && i < nfiles) asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(key[i]) : "l"(S.data[i] + v));
                }
            }
            asm volatile("" ::: "memory");
            // next tile's DATA rows into L1: issued behind the gather so that its address loads overlap the gather latency
            if (tile + wstride < ntiles && !(MB_ABLATE && (flags & 128))) prefetch_rows(false, tile + wstride);
            if (want_var && (flags & 16) && tile + wstride < ntiles) prefetch_rows(true, tile + wstride);   // flags & 16: STAT a tile ahead as well
            MB_STAMP(1);   // gather + prefetch issue
#if MB_TIMING
#pragma unroll
            for (int i = 0; i < NB; i++) MB_TOUCH(key[i]);
#endif
            MB_STAMP(2);   // DATA arrival

            // ---- pass 1: exposure order.  w, shifted moments, keys ----------------------------------
            // Straight-line, 11-12 instructions per exposure: one conversion, (offset + d) * scale exactly as
            // merging.c:426, t = w - p, and one PTX block that does everything depending on "d is not NaN"
            // under a predicate (moment updates, invalid bit, sentinel select) -- no branch per exposure.
            // MB_P1_CHAINS independent accumulator chains: the 48-deep dependent chains of float64 adds were the
            // critical path of this pass (the pipe has a long latency and only six warps share a scheduler)
            constexpr int NCH = MB_P1_CHAINS;
            double T1c[NCH], S2c[NCH];
#pragma unroll
            for (int c = 0; c < NCH; c++) { T1c[c] = 0.0; S2c[c] = 0.0; }
            bool have_p = false;
#pragma unroll
            for (int i = 0; i < NB; i++) {
                const float d = key[i];
                const unsigned sentinel = SENT | (unsigned)i;
                double w;
                float kf;
                if constexpr (UNIT) {
                    w = (double)d;
                    kf = d;
                } else {
                    w = (S.os[i].x + (double)d) * S.os[i].y;
#if MB_KEY_FFMA
                    kf = __fmaf_rn(d, S.kscale[i], S.koffset[i]);
#else
                    kf = __double2float_rn(w);
#endif
                }
                if (i < 4) {                       // pivot = first valid sample among the first four
                    const bool ok = (d == d);
                    p = (ok && !have_p) ? w : p;
                    have_p = have_p || ok;
                }
                const bool ok = (d == d);
                // invalid sample: t becomes a denormal (high word cleared, at most 2^-1042 in magnitude) with ONE
                // select; the unconditional moment updates below then add nothing that survives a rounding
                double t = w - p;
                t = __hiloint2double(ok ? __double2hiint(t) : 0, __double2loint(t));
#if MB_EXACT_SUM
                sum += ok ? w : 0.0;               // tools.c:13-16, exposure order, no FMA
#endif
                T1c[i % NCH] += t;
                S2c[i % NCH] = __fma_rn(t, t, S2c[i % NCH]);
                if (!ok) inval.set(i);
                const unsigned kb = (__float_as_uint(kf) & keep_mask) | (unsigned)i;
                key[i] = __uint_as_float(ok ? kb : sentinel);
            }
            if (STRIDED && !act) {                         // idle lane: every sample invalid
#pragma unroll
                for (int k = 0; k < W; k++) inval.w[k] = (k == W - 1 && (NB & 31)) ? ((1u << (NB & 31)) - 1u) : 0xFFFFFFFFu;
            }
            T1 = T1c[0]; S2 = S2c[0];
#pragma unroll
            for (int c = 1; c < NCH; c++) { T1 += T1c[c]; S2 += S2c[c]; }
        }
        const int n = NB - inval.count();
        float dspec[2] = {0.f, 0.f};
        MB_STAMP(3);   // pass 1

        if (!(flags & 4)) {   // flags & 4: timing experiment only (skips the sort and the clip loop; wrong results)
        sort_keys_mixed<NB, MB_SORT_FMA_EVERY>(key, one, minus_one);
        // speculative exact lookups of the two largest samples (L1 / L2 hits), in flight while sigma,
        // median and bounds are computed: a rejected sample's exact value is then usually already here
        if constexpr (SENT_LOW) {
            if (n >= 2 && act) {
#pragma unroll
                for (int j = 0; j < 2; j++) {
                    // (a key built from a non-finite w is a NaN pattern and the network may leave the canonical NaN in
                    // this slot: such a voxel is recomputed by the exact path, but its index bits must not address memory)
                    const int i = min((int)(__float_as_uint(key[NB - 1 - j]) & IDX), nfiles - 1);
                    const float* src;
                    if constexpr (STRIDED) src = S.data[0] + (size_t)i * S.dstride + v;
                    else src = sm.data[i] + v;
                    dspec[j] = __ldg(src);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < NB; q++) sk[q][lane] = key[q];
        }
        // the STAT rows are prefetched into L1 only now: close enough to their use (after the clip
        // loop) that they are still there, early enough that the latency is covered
        if (want_var && !(flags & (8 | 16)) && !(MB_ABLATE && (flags & 128))) prefetch_rows(true, tile);
        MB_STAMP(4);   // sort + column store + STAT prefetch

#if !MB_LATE_STAT
        // STAT rows: issued now, into the registers the keys just vacated, consumed after the clip
        // loop -- their latency hides behind it.  Unconditional loads: rows beyond nfiles alias
        // row 0 and idle lanes re-read the last voxel (both are masked out by `dropped`).
        float sv[NB];
        if (want_var) {
            const long long vs = act ? v : A.nvox - 1;
#pragma unroll
            for (int i = 0; i < NB; i++)
                asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(sv[i]) : "l"(S.var[i < nfiles ? i : 0] + vs));
        }
#endif

        // ---- clip iterations on the sorted column --------------------------------------------
#if MB_EXACT_SUM
        double mean = sum / (double)n;             // n == 0 -> NaN (merging.c:439)
#else
        double mean = n > 0 ? p + T1 * sm.rcp[n] : CUDART_NAN;
#endif
        double sigma = 0.0;
        int a = SENT_LOW ? NB - n : 0, b = a + n;   // window of the valid samples in the sorted column
        ExpSet<NB> rejected;
        bool slow = false;
        if constexpr (MAD) {
            double (*sw)[32] = sm.sw[wib];
            // exact values in key order (their d re-read through L1 / L2; four lookups in flight)
#pragma unroll 4
            for (int q = 0; q < n; q++) {
                const int i = min((int)(__float_as_uint(sk[q][lane]) & IDX), nfiles - 1);   // NaN-pattern keys of non-finite samples
                const double dd = (double)__ldg(sm.data[i] + v);
                sw[q][lane] = UNIT ? dd : (sm.offset[i] + dd) * sm.scale[i];
            }
            // samples sharing a key bucket are in arbitrary order: local insertion by exact value
            for (int q = 1; q < n; q++) {
                if (((__float_as_uint(sk[q][lane]) ^ __float_as_uint(sk[q - 1][lane])) & ~IDX) == 0u) {
                    for (int r = q; r > 0 && sw[r - 1][lane] > sw[r][lane]; r--) {
                        const double tw = sw[r][lane]; sw[r][lane] = sw[r - 1][lane]; sw[r - 1][lane] = tw;
                        const float tk = sk[r][lane]; sk[r][lane] = sk[r - 1][lane]; sk[r - 1][lane] = tk;
                    }
                }
            }
            if (n >= 2) {
                int it = A.nmax;
                for (bool first = true;; first = false) {
                    const int m = b - a;
                    if (!first) {                                  // tools.c:61-65 on the compacted, ascending list
                        double acc = 0.0;
#pragma unroll 4
                        for (int q = a; q < b; q++) acc += sw[q][lane];
                        mean = acc / (double)m;
                    }
                    const int ih = a + (m >> 1);
                    const double med = (m & 1) ? sw[ih][lane] : (sw[ih][lane] + sw[ih - 1][lane]) / 2;   // tools.c:45-48
                    // middle of |w - med| by merging the two sorted runs outward from the median
                    int il = (m & 1) ? ih : ih - 1, ir = il + 1;
                    double cur = 0.0, prev = 0.0;
                    for (int k = 0; k <= (m >> 1); k++) {
                        const double dl = il >= a ? fabs(sw[il][lane] - med) : CUDART_INF;
                        const double dr = ir < b ? fabs(sw[ir][lane] - med) : CUDART_INF;
                        prev = cur;
                        if (dl <= dr) { cur = dl; il--; } else { cur = dr; ir++; }
                    }
                    const double madv = (m & 1) ? cur : (cur + prev) / 2;
                    sigma = madv * 1.4826;                         // tools.c:74
                    const double clo = med - (A.nclip_low * sigma);
                    const double cup = med + (A.nclip_up * sigma);
                    // strict window on the ascending exact column: first index with w > clo, first with w >= cup
                    int lo = a, hi = b;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sw[mid][lane] > clo) hi = mid; else lo = mid + 1; }
                    const int na = lo;
                    lo = na; hi = b;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (sw[mid][lane] < cup) lo = mid + 1; else hi = mid; }
                    const int nb2 = lo;
                    // NaN bounds (non-finite input): every comparison is false, nothing is inside
                    const int ni = (clo == clo && cup == cup) ? nb2 - na : 0;
                    if (ni < nstop || ni == m) break;              // tools.c:140-143
                    if (it <= 0) break;                            // tools.c:144
                    it--;
                    a = na;
                    b = nb2;
                }
                for (int q = 0; q < n; q++)
                    if (q < a || q >= b) rejected.set_dyn(min((int)(__float_as_uint(sk[q][lane]) & IDX), nfiles - 1));
            }
        } else
        if (n >= 2 && !(flags & 4)) {
            const double S2_0 = S2;
            int it = A.nmax;
            // values beyond float range or an absurd pivot defeat the key precision model
            if (!(S2_0 < 0x1p200)) slow = true;
            // S2_0 == 0: every valid sample equals the pivot, sigma == 0, nothing lies strictly
            // inside the empty window: the reference stops at once with all samples kept (tools.c:100)
            while (!slow && S2_0 != 0.0) {
                const int m = b - a;
                const double rm = sm.rcp[m];
                const double dm1 = T1 * rm;
                double var_k = __fma_rn(S2, rm, -(dm1 * dm1));
                // conditioning: the moment variance must dominate its rounding error
                if (!(var_k >= 0x1p-13 * (S2_0 * rm))) { slow = true; break; }
                sigma = sqrt(var_k);
                const int ih = a + (m >> 1);
                const int il = ih - 1 + (m & 1);
                const float kh = __uint_as_float(__float_as_uint(sk[ih][lane]) & ~IDX);
                const float kl = __uint_as_float(__float_as_uint(sk[il][lane]) & ~IDX);
                const double med = ((double)kh + (double)kl) * 0.5;
                const double amax = fmax(fabs((double)kh), fabs((double)kl));
                const double clo = med - (A.nclip_low * sigma);
                const double cup = med + (A.nclip_up * sigma);
                // total uncertainty of a key against a bound: median from truncated keys,
                // sigma from reordered sums, key truncation of the sample itself
                const double spread = (A.nclip_low + A.nclip_up) * sigma;
                const double T = KEPS * amax + 0x1p-26 * spread + KEPS * (fabs(clo) + fabs(cup)) + S.koff;
                if (!(T < 0x1p120)) { slow = true; break; }   // NaN / inf bounds: let the exact path decide
                // keys certainly at or below the low bound / certainly below the high bound
                const float f_lo_in = __double2float_ru(clo + T), f_lo_out = __double2float_rd(clo - T);
                const float f_up_in = __double2float_rd(cup - T), f_up_out = __double2float_ru(cup + T);
                // The rejected samples are almost always the last few keys at either end of the window: probe
                // four keys per end (eight independent loads) instead of two dependent binary searches; only a
                // lane with four or more rejections on one side falls back to the search.
                float kl4[4], ku4[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int ql = a + j, qu = b - 1 - j;
                    kl4[j] = ql < b ? sk[ql < NB ? ql : NB - 1][lane] : CUDART_INF_F;
                    ku4[j] = qu >= a ? sk[qu >= 0 ? qu : 0][lane] : -CUDART_INF_F;
                }
                int jl = 0, ju = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    jl += (jl == j && kl4[j] <= f_lo_in) ? 1 : 0;        // leading keys possibly at or below the low bound
                    ju += (ju == j && !(ku4[j] < f_up_in)) ? 1 : 0;      // trailing keys not certainly below the high bound
                }
                int lo_in, up_in;
                bool lo_band, up_band;
                // band checks: no key in (clo - T, clo + T] and none in [cup - T, cup + T)
                if (jl < 4) {
                    lo_in = a + jl;
                    const float last = jl == 3 ? kl4[2] : (jl == 2 ? kl4[1] : kl4[0]);
                    lo_band = jl > 0 && !(last <= f_lo_out);
                } else {
                    lo_in = count_col<NB, true>(sk, lane, f_lo_in);
                    lo_band = lo_in > 0 && !(sk[lo_in - 1][lane] <= f_lo_out);
                }
                if (ju < 4) {
                    up_in = b - ju;
                    const float first = ju == 3 ? ku4[2] : (ju == 2 ? ku4[1] : ku4[0]);
                    up_band = ju > 0 && (first < f_up_out);
                } else {
                    up_in = count_col<NB, false>(sk, lane, f_up_in);
                    up_band = up_in < NB && (sk[up_in < NB ? up_in : NB - 1][lane] < f_up_out);
                }
                if (lo_band || up_band) {
                    // ~1e-4 of the voxels: settle the band keys from their exact values
                    if (!refine_band<NB, UNIT>(sk, lane, sm.data, v, sm.scale, sm.offset, nfiles, ih, il, m, sigma, A.nclip_low,
                                               A.nclip_up, clo, cup, T, lo_in, up_in)) { slow = true; break; }
                }
                lo_in = min(max(lo_in, a), b);
                up_in = min(max(up_in, a), b);
                const int na = lo_in;                                     // first kept position
                const int nb2 = max(up_in, na);                           // one past the last kept
                const int ni = nb2 - na;
                if (ni < nstop || ni == m) break;                         // tools.c:100-103
                if (it <= 0) break;                                       // tools.c:104
                it--;
                // drop the rejected samples [a,na) and [nb2,b): exact w recomputed from the samples, which are
                // re-read through L1 / L2 (this warp loaded them a moment ago) four at a time, so that a
                // voxel's lookups share one memory latency instead of queueing up behind each other
                const int r1 = na - a, r = r1 + (b - nb2);
                for (int k0 = 0; k0 < r; k0 += 4) {
                    float dv[4];
                    int iv[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int k = k0 + j;
                        const int q = k < r ? (k < r1 ? a + k : nb2 + (k - r1)) : a;
                        iv[j] = min((int)(__float_as_uint(sk[q][lane]) & IDX), nfiles - 1);
                        const float* src;
                        if constexpr (STRIDED) src = S.data[0] + (size_t)iv[j] * S.dstride + v;
                        else src = sm.data[iv[j]] + v;
                        const bool spec = SENT_LOW && q >= NB - 2;     // one of the two speculative lookups
                        dv[j] = spec ? (q == NB - 1 ? dspec[0] : dspec[1]) : 0.f;
                        if (k < r && !spec) dv[j] = __ldg(src);
                    }
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (k0 + j < r) {
                            const double dd = (double)dv[j];
                            const double w = UNIT ? dd : (sm.offset[iv[j]] + dd) * sm.scale[iv[j]];
                            const double t = w - p;
                            T1 -= t;
                            S2 -= t * t;
                            rejected.set_dyn(iv[j]);
                        }
                    }
                }
                a = na;
                b = nb2;
                mean = p + T1 * sm.rcp[ni];
            }
        }
        int kept = b - a;
        if (slow) {
            const VoxelResultN<NB> r = exact_voxel_w<NB, UNIT>(sm.data, v, sm.scale, sm.offset, nfiles, A.nmax,
                                                               A.nclip_low, A.nclip_up, nstop);
            mean = r.mean;
            sigma = r.sigma;
            kept = r.kept;
            rejected = r.rejected;
            slow_count++;
        }
        __syncwarp();
        MB_STAMP(5);   // clip loop

        // ---- propagated variance over the kept set --------------------------------------------
        const ExpSet<NB> dropped = inval | rejected;
        double vsum = 0.0;
#if MB_LATE_STAT
        // STAT rows loaded only now (prefetched into L1 after the sort): fewer live registers across the
        // clip loop, more resident blocks per SM.  Two half-batches: the second is in flight while the
        // first is consumed.
        {
        float sv[NB];
        if (want_var) {
            const long long vs = act ? v : A.nvox - 1;
            constexpr int HB = NB / 2;
#if MB_ABLATE   // flags & 64: STAT from an L1-resident window
            const float* vp = (flags & 64) ? S.var[0] + (vs & 1023) : S.var[0] + vs;
            const unsigned sstride = (flags & 64) ? 32u : S.vstride;
#else
            const float* vp = S.var[0] + vs;
            const unsigned sstride = S.vstride;
#endif
#pragma unroll
            for (int i = 0; i < HB; i++) {
                if constexpr (STRIDED) {
                    asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(sv[i]) : "l"(vp));
                    asm("mad.wide.u32 %0, %1, 4, %0;" : "+l"(vp) : "r"(sstride));
                } else {
                    asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(sv[i]) : "l"(S.var[i < nfiles ? i : 0] + vs));
                }
            }
#pragma unroll
            for (int i = HB; i < NB; i++) {
                if constexpr (STRIDED) {
                    asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(sv[i]) : "l"(vp));
                    asm("mad.wide.u32 %0, %1, 4, %0;" : "+l"(vp) : "r"(sstride));
                } else {
                    asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(sv[i]) : "l"(S.var[i < nfiles ? i : 0] + vs));
                }
            }
            asm volatile("" ::: "memory");
            MB_STAMP(6);   // STAT issue
#if MB_TIMING
#pragma unroll
            for (int i = 0; i < NB; i++) MB_TOUCH(sv[i]);
#endif
            MB_STAMP(7);   // STAT arrival
#if MB_VAR_F32
            if constexpr (!MAD) {
            // float32 FMAs in NACC interleaved chains (each <= 8 terms for N <= 64), chains summed in
            // float64: <= (NB / NACC + 2) * 2^-24 relative for non-negative STAT, inside the 1e-6 contract
            constexpr int NACC = NB <= 8 ? 1 : (NB <= 16 ? 2 : (NB <= 32 ? 4 : (NB <= 64 ? 8 : 16)));
            float vacc[NACC];
#pragma unroll
            for (int k = 0; k < NACC; k++) vacc[k] = 0.f;
#pragma unroll
            for (int i = 0; i < HB; i++)
                if (!dropped.test(i)) vacc[i % NACC] = __fmaf_rn(sv[i], S.scale2f[i], vacc[i % NACC]);   // merging.c:431,462
            asm volatile("" ::: "memory");
#pragma unroll
            for (int i = HB; i < NB; i++)
                if (!dropped.test(i)) vacc[i % NACC] = __fmaf_rn(sv[i], S.scale2f[i], vacc[i % NACC]);
#pragma unroll
            for (int k = 0; k < NACC; k++) vsum += (double)vacc[k];
            // the bound needs non-negative terms: a sign bit among the kept STAT samples sends the lane to float64 products
            unsigned neg = 0u;
#pragma unroll
            for (int i = 0; i < NB; i++) neg |= dropped.test(i) ? 0u : __float_as_uint(sv[i]);
            if (neg & 0x80000000u) {
                vsum = 0.0;
#pragma unroll
                for (int i = 0; i < NB; i++)
                    if (!dropped.test(i)) vsum += (double)sv[i] * S.scale2[i];   // merging.c:431,462
            }
            } else
#endif
            {   // float64 products in exposure order (the MAD variant stays bit-identical to the reference)
#pragma unroll
            for (int i = 0; i < HB; i++)
                if (!dropped.test(i)) vsum += (double)sv[i] * S.scale2[i];   // merging.c:431,462
            asm volatile("" ::: "memory");
#pragma unroll
            for (int i = HB; i < NB; i++)
                if (!dropped.test(i)) vsum += (double)sv[i] * S.scale2[i];
            }
        }
        }
#else
        if (want_var) {
#pragma unroll
            for (int i = 0; i < NB; i++)
                if (!dropped.test(i)) vsum += (double)sv[i] * S.scale2[i];   // merging.c:431,462
        }
#endif

        // ---- per-exposure counters: transpose the lanes' masks, one POPC per lane --------------
        if constexpr (NB > 32 && NB <= 48) {
            // exposures 32..47: invalid in the low half-word, dropped in the high half-word
            cnt_inval[0] += __popc(tc.apply(inval.w[0]));
            cnt_drop[0] += __popc(tc.apply(dropped.w[0]));
            cnt_inval[1] += __popc(tc.apply((inval.w[1] & 0xFFFFu) | (dropped.w[1] << 16)));
        } else {
#pragma unroll
            for (int k = 0; k < W; k++) {
                cnt_inval[k] += __popc(tc.apply(inval.w[k]));
                cnt_drop[k] += __popc(tc.apply(dropped.w[k]));
            }
        }
        tiles_done++;

        if (act) {
            double ov;
            if (want_var) {
                ov = vsum / ((double)kept * (double)kept);                // merging.c:462
            } else if (kept > 1) {
                ov = sigma * sigma;                                       // merging.c:465
                if (A.typ_var == 1) ov /= (double)(kept - 1);             // merging.c:467
            } else {
                ov = CUDART_NAN;
            }
            A.out[v] = mean;
            A.outvar[v] = ov;
            A.expmap[v] = kept;
        }
        MB_STAMP(8);   // variance, counters, stores
    }
#if MB_TIMING
    __syncthreads();
    if ((blockIdx.x == 0 || blockIdx.x == 400) && lane == 0) {
        printf("block %d warp %d tiles %u cycles: barrier %llu issue %llu data_wait %llu pass1 %llu sort %llu clip %llu stat_issue %llu stat_wait %llu rest %llu\n",
               (int)blockIdx.x, wib, tiles_done, sm.tph[wib][0], sm.tph[wib][1], sm.tph[wib][2], sm.tph[wib][3], sm.tph[wib][4],
               sm.tph[wib][5], sm.tph[wib][6], sm.tph[wib][7], sm.tph[wib][8]);
    }
#endif

    // ---- flush: valid = lanes seen - invalid, selected = lanes seen - dropped -----------------
    const unsigned seen = tiles_done * 32u;
    if constexpr (NB > 32 && NB <= 48) {
        if (lane < nfiles) {
            atomicAdd(A.valid_cnt + lane, (unsigned long long)(seen - cnt_inval[0]));
            atomicAdd(A.select_cnt + lane, (unsigned long long)(seen - cnt_drop[0]));
        }
        const int e = 32 + (lane & 15);
        if (e < nfiles) {
            if (lane < 16) atomicAdd(A.valid_cnt + e, (unsigned long long)(seen - cnt_inval[1]));
            else atomicAdd(A.select_cnt + e, (unsigned long long)(seen - cnt_inval[1]));
        }
    } else {
#pragma unroll
        for (int k = 0; k < W; k++) {
            const int e = 32 * k + lane;
            if (e < nfiles) {
                atomicAdd(A.valid_cnt + e, (unsigned long long)(seen - cnt_inval[k]));
                atomicAdd(A.select_cnt + e, (unsigned long long)(seen - cnt_drop[k]));
            }
        }
    }
    if (slow_total != nullptr && slow_count) atomicAdd(slow_total, (unsigned long long)slow_count);
}

}  // namespace mb

namespace mb {

// ---------------------------------------------------------------------------------------------
// Median (merging.c:232-253, tools.c:39-51), float32 stacks <= 64 deep.
//
// One lane owns one voxel.  Invalid samples (NaN, unused slots) are replaced alternately by
// -inf and +inf, so that the valid samples stay centred in the sorted stack: the median of the
// n valid samples is then always at the two fixed middle slots NB/2-1 and NB/2 (odd n: slot
// NB/2-1 alone), and no data-dependent selection is needed.  The network's compare-exchange is
// FMNMX + integer a+b-min with the integer part on the FMA pipe (the ALU pipe is the bottleneck
// of this kernel -- measured: FMNMX 64 lanes/clk/SM, IADD3 more than that, IMAD 64, so the plain
// FMNMX + IADD3 form sustains ~60 compare-exchanges/clk/SM, twice the 2xIMAD form); per-exposure valid counts come from one 32x32 bit transpose per 32 exposures.
// ---------------------------------------------------------------------------------------------
#ifndef MB_MEDIAN_BLOCKS
#define MB_MEDIAN_BLOCKS 0
#endif
// resident 128-thread blocks per SM the median kernel is compiled for: it is bound by memory latency
// (long_scoreboard leads its stalls), so shallow stacks trade registers for more warps in flight
template <int NB> struct MedianBlocks { static constexpr int value = MB_MEDIAN_BLOCKS ? MB_MEDIAN_BLOCKS : 1; };
template <int NB, bool FULL>
__global__ void __launch_bounds__(kThreads, MedianBlocks<NB>::value) median_warp_kernel(const TileArgs<NB> S, const ClipArgs A)
{
    constexpr int W = ExpSet<NB>::W;
    const int lane = threadIdx.x & 31;
    const int nfiles = FULL ? NB : A.nfiles;
    const unsigned one = (unsigned)(A.nfiles > 0);
    const unsigned minus_one = 0u - one;
    TransposeConsts tc;
    tc.init(lane);
    unsigned cnt[W];
#pragma unroll
    for (int k = 0; k < W; k++) cnt[k] = 0u;
    unsigned tiles_done = 0;

    const long long ntiles = (A.nvox + 31) >> 5;
    const long long wstride = (long long)gridDim.x * (kThreads / 32);
    // (no per-tile phase lock here: measured 8 % slower for the median, unlike the sigma clip)
    for (long long tile = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); tile < ntiles; tile += wstride) {
        const long long v = (tile << 5) + lane;
        const bool act = v < A.nvox;
        float key[NB];
#pragma unroll
        for (int i = 0; i < NB; i++) {
            key[i] = CUDART_NAN_F;
            if (act && i < nfiles) asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(key[i]) : "l"(S.data[i] + v));
        }
        asm volatile("" ::: "memory");
        ExpSet<NB> inval;
        bool flip = false;   // parity of the invalid samples seen so far
#pragma unroll
        for (int i = 0; i < NB; i++) {
            const float d = key[i];
            const bool ok = (d == d);
            key[i] = ok ? d : (flip ? -CUDART_INF_F : CUDART_INF_F);
            if (!ok) {
                flip = !flip;
                inval.set(i);
            }
        }
        const int n = NB - inval.count();
        sort_keys_mixed<NB, 1>(key, one, minus_one);   // 2xIMAD form: measured a few % faster here
        // first invalid sample went to +inf: with an odd number of them there is one more +inf than
        // -inf, so the valid samples occupy slots [a, a+n) with a = floor((NB-n)/2) and the middle
        // of the valid run is at slot NB/2-1 (odd n) or slots NB/2-1, NB/2 (even n)
        const double lo = (double)key[NB / 2 - 1];
        const double hi = (double)key[NB / 2];
#pragma unroll
        for (int k = 0; k < W; k++) cnt[k] += __popc(tc.apply(inval.w[k]));
        tiles_done++;
        if (act) {
            double med = (n & 1) ? lo : (lo + hi) * 0.5;
            if (n == 0) med = CUDART_NAN;
            A.out[v] = med;
            A.expmap[v] = n;
        }
    }
    const unsigned seen = tiles_done * 32u;
#pragma unroll
    for (int k = 0; k < W; k++) {
        const int e = 32 * k + lane;
        if (e < nfiles) atomicAdd(A.valid_cnt + e, (unsigned long long)(seen - cnt[k]));
    }
}

}  // namespace mb
