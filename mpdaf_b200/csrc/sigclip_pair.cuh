// Pair-packed, TMA-pipelined sigma-clip kernel for sm_100a: the main combine kernel of round 2
// (float32 stacks up to 64 deep, float64 stacks up to 48 deep, rows equally pitched and 16-byte aligned).
//
// Reference behaviour restated (never copied): src/merging.c:423-477 (gather, n == 0 / 1, outputs,
// counters) and src/tools.c:82-118 (median-centred iterative clip, strict window, stop rules).
//
// Why a new formulation.  The round-1 warp kernel (sigclip_warp.cuh) is bound by the 16-lane ALU pipe: a
// compare-exchange on 32-bit keys costs two ALU-pipe instructions per voxel, its exposure rows are gathered
// into registers by the computing warps themselves, and its 45 KB loop body forces all warps of an SM into the
// same phase.  Here
//  * a lane owns TWO neighbouring voxels and their 16-bit keys share a register, so one HMNMX2.BF16 min plus
//    one max performs the compare-exchange of both voxels: half the sort instructions per voxel;
//  * the keys are bfloat16(w - pivot) with no exposure index.  Which exposures fall outside the clip window is
//    recovered afterwards by ONE packed comparison pass (one HSETP2 per exposure and side for both voxels) of the
//    unsorted keys, kept in registers, against the window's boundary keys -- and only when something is rejected;
//  * every warp owns one (or two) shared-memory stages and keeps them filled itself: lane 0 requests the [NB][64]
//    tile of the warp's next-but-one round with cp.async.bulk.tensor (TMA, mbarrier completion) as soon as the
//    clip loop has released the stage; the load lands while the warp computes the variance and stores.  There is
//    no producer warp: at 168 registers per thread the register file holds exactly twelve warps (N = 48), three
//    per scheduler, and all of them compute;
//  * STAT rows (var = 'propagate') are read straight from L2 at the end of the tile (one TMA prefetch per tile
//    at its start) into the registers the keys vacated;
//  * the clip step is branch-free per voxel pair (both voxels' chains interleave), probes the ranks next to the
//    window's ends without clamps, and ends a voxel as soon as no clip is left.
// Decisions stay CERTIFIED: a key differs from its sample's w - pivot by at most KREL * |key| + E0, an order
// statistic of the keys from the same order statistic of the samples by the same bound (both maps are
// monotone), and sigma comes from float64 shifted moments (2^-28 relative, conditioning checked).  A voxel is
// classified from its keys only if no key lies within that uncertainty of a clip bound; the few keys outside
// the certain zone are candidates whose exact float64 values decide; a voxel with an undecidable candidate goes
// to a queue and is recomputed by the reference algorithm restated literally (sigclip_pair_slow_kernel: one
// WARP per queued voxel, full warps instead of one slow lane stalling 63 fast ones).  expmap and the
// per-exposure counters are therefore bit-exact, means and variances within ~1e-15 / 3e-7 relative.
#pragma once

#include "sigclip_warp.cuh"

#ifndef MB_PAIR_LOOKAHEAD
#define MB_PAIR_LOOKAHEAD 0   // 0: as many free stages as shared memory allows (at least 2, at most 8)
#endif
#ifndef MB_PAIR_CONSUMERS
#define MB_PAIR_CONSUMERS 0   // 0: per-depth default (PairCfg)
#endif

namespace mb {

template <int NB>
struct PairArgs {
    const void* data0;                     // row 0 of the DATA stack (float32 or float64 samples: template parameter ET)
    const void* var0;                      // row 0 of the STAT stack (null unless var = 'propagate')
    unsigned long long dstride, vstride;   // elements between consecutive rows
    double2 os[NB];                        // (offset, scale)
    double osp[NB];                        // offset * scale
    float kscale[NB];                      // float32(scale)
    float koffset[NB];                     // float32(offset * scale)
    float scale2f[NB];                     // float32(scale^2)
    double scale2[NB];
    double osmax;                          // max |offset * scale|: part of the absolute key error
    unsigned long long* hist;              // exposure-map histogram [>= NB + 1] the kernels add to (null: not wanted)
};

template <typename ET> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

template <int NB, int ES = 4>
struct PairCfg {
    static constexpr int stage_bytes = NB * 64 * ES;   // one [NB][64] tile of float32 (ES = 4) or float64 (8) samples
    static constexpr int col_bytes = NB * 32 * 4;      // sorted packed keys of one warp
    static constexpr int misc_bytes = 4096;            // barriers, reciprocal table, per-exposure constants
    static constexpr int budget = 227 * 1024 - misc_bytes;
    // Every warp of the block is a consumer that owns its stage(s) and refills them itself (no producer warp: with 168
    // registers per thread the register file holds exactly 12 warps, three per scheduler).  Their number is bounded
    // by the register file (about 2 NB + 72 registers per thread: sorted and unsorted packed keys plus the per-voxel
    // state), by 16 (four warps per scheduler), and by shared memory (a stage is held until the clip loop is over).
    static constexpr int c_regs = 2048 / (2 * NB + 72);
    static constexpr int c_smem = budget / (stage_bytes + col_bytes);
    static constexpr int c_cap = c_regs < 16 ? c_regs : 16;
    static constexpr int c_fit = c_cap < c_smem ? c_cap : c_smem;
    // the warps of a block run in phase (round barrier): 4 k + 1 of them leave three schedulers waiting for the fourth's
    // extra warp in every round (N = 64 without STAT: 8 warps 0.688 ms, 9 warps 0.785 ms)
    static constexpr int c_auto = (c_fit > 4 && c_fit % 4 == 1) ? c_fit - 1 : c_fit;
    static constexpr int C = MB_PAIR_CONSUMERS ? MB_PAIR_CONSUMERS : c_auto;
    // stages per warp: two where shared memory allows (the next tile is then requested a whole tile ahead)
    static constexpr int SPW = (budget - C * col_bytes) / (C * stage_bytes) >= 2 ? 2 : 1;
    static constexpr int ST = SPW * C;
    static constexpr int threads = C * 32;
    static_assert(C >= 2, "stack too deep for the pair kernel");
};

template <int NB, typename ET = float>
struct PairSmem {
    ET stage[PairCfg<NB, sizeof(ET)>::ST][NB][64];        // DATA tiles: stages w * SPW .. w * SPW + SPW - 1 belong to warp w
    unsigned col[PairCfg<NB, sizeof(ET)>::C][NB][32];     // sorted keys: low half voxel 2l, high half voxel 2l+1
    double rcp[NB + 1];
    double2 os[NB];                                       // (offset, scale) for lane-dependent lookups
    double osp[NB];                                       // offset * scale exactly as pass 1 uses it
    unsigned long long full[PairCfg<NB, sizeof(ET)>::ST]; // TMA completion of a stage's DATA tile
    unsigned hist[NB + 1];                                // voxels of this block by number of exposures kept
    unsigned pad[64];                                     // the clip loop may read up to five rows beyond the last column
};
static_assert(sizeof(PairSmem<48, float>) <= 227 * 1024 && sizeof(PairSmem<64, float>) <= 227 * 1024 && sizeof(PairSmem<48, double>) <= 227 * 1024 &&
                  sizeof(PairSmem<64, double>) <= 227 * 1024, "shared memory");

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// key error model: bfloat16 rounding of t32 (2^-8 relative) + float32 evaluation of w - pivot
constexpr double kPairKrel = 0x1.1p-8;
constexpr double kPairC1 = (1.0 / (1.0 + kPairKrel)) * (1.0 - 0x1p-40);   // <= 1 / (1 + KREL)
constexpr double kPairC2 = (1.0 / (1.0 - kPairKrel)) * (1.0 + 0x1p-40);   // >= 1 / (1 - KREL)
constexpr unsigned kPairSent = 0xFF7F0000u;   // float32 image of the most negative finite bfloat16: key of an invalid sample

// The reference algorithm restated literally for one voxel of a strided stack; stores the outputs and corrects
// the per-exposure `selected` counters (the main kernel counted every valid sample of a queued voxel as kept).
template <int NB, bool UNIT, typename ET = float>
__device__ __noinline__ void exact_voxel_store(const PairArgs<NB>& S, const ClipArgs& A, long long v)
{
    const ET* data0 = static_cast<const ET*>(S.data0);
    const ET* var0 = static_cast<const ET*>(S.var0);
    GenericScratch<NB> g;
    const int nfiles = A.nfiles;
    int n = 0;
    for (int i = 0; i < nfiles; i++) {
        const ET x = __ldg(data0 + (size_t)i * S.dstride + v);
        if (x == x) {
            g.w[n] = UNIT ? (double)x : (S.os[i].x + (double)x) * S.os[i].y;
            g.file[n] = (unsigned short)i;
            g.idx[n] = (unsigned short)n;
            n++;
        }
    }
    double mean = CUDART_NAN, sigma = 0.0, ov = CUDART_NAN;
    int kept = n;
    if (n == 1) mean = g.w[0];
    if (n >= 2) kept = clip_generic<NB>(g, n, A.nmax, A.nclip_low, A.nclip_up, A.nstop < 1 ? 1 : A.nstop, A.mad, mean, sigma);
    if (n >= 1) {
        if (A.typ_var == 0) {
            double vs = 0.0;
            for (int k = 0; k < kept; k++) {                                       // merging.c:461-462
                const int i = g.file[g.idx[k]];
                vs += (double)__ldg(var0 + (size_t)i * S.vstride + v) * S.scale2[i];
            }
            ov = vs / ((double)kept * (double)kept);
        } else if (kept > 1) {
            ov = sigma * sigma;                                                    // merging.c:465
            if (A.typ_var == 1) ov /= (double)(kept - 1);                          // merging.c:467
        }
    }
    if (n >= 2 && kept < n) {
        // rejected = valid but not among the survivors g.idx[0 .. kept)
        unsigned keep[(NB + 31) / 32];
        for (int k = 0; k < (NB + 31) / 32; k++) keep[k] = 0u;
        for (int k = 0; k < kept; k++) {
            const int i = g.file[g.idx[k]];
            keep[i >> 5] |= 1u << (i & 31);
        }
        for (int k = 0; k < n; k++) {
            const int i = g.file[k];
            if (!((keep[i >> 5] >> (i & 31)) & 1u)) atomicAdd(A.select_cnt + i, ~0ull);   // -1
        }
    }
    A.out[v] = mean;
    A.outvar[v] = ov;
    A.expmap[v] = kept;
    if (S.hist != nullptr) atomicAdd(S.hist + kept, 1ull);
}

// Queued voxels: the reference algorithm restated literally (merging.c:423-477, tools.c:9-24, :39-51, :82-118), one
// WARP per voxel.  The samples (<= 64: two per lane) are compacted in file order into shared memory, ranked
// against each other (ties by position, so the ranks are a permutation) and scattered into an ascending copy; the
// sums the decisions depend on run sequentially in the reference's order (file order first, ascending afterwards)
// and redundantly in every lane, the strict window counts are ballots.  Survivors of a clip are a contiguous range
// of the ascending copy (verified restatement).  Also corrects the per-exposure `selected` counters:
// the main kernel counted every valid sample of a queued voxel as kept.
template <int NB, bool UNIT, typename ET = float>
__global__ void __launch_bounds__(128) sigclip_pair_slow_kernel(const PairArgs<NB> S, const ClipArgs A, const unsigned* __restrict__ queue,
                                                                const unsigned* __restrict__ count, unsigned cap,
                                                                unsigned long long* slow_total)
{
    constexpr int NS = (NB + 31) / 32;       // samples per lane (up to 128 exposures: the duo kernel's queue comes here too)
    constexpr int NSLOT = NS * 32;
    __shared__ double s_w[4][NSLOT];         // valid samples in file order (later: the kept samples' variances)
    __shared__ double s_s[4][NSLOT];         // ascending
    __shared__ unsigned char s_f[4][NSLOT];  // exposure of the ascending samples
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const ET* data0 = static_cast<const ET*>(S.data0);
    const ET* var0 = static_cast<const ET*>(S.var0);
    const unsigned nq = min(*count, cap);
    if (blockIdx.x == 0 && threadIdx.x == 0 && slow_total != nullptr) atomicAdd(slow_total, (unsigned long long)*count);
    const int nfiles = A.nfiles;
    const int nstop = A.nstop < 1 ? 1 : A.nstop;
    const unsigned lt = (1u << lane) - 1u;
    double* fw = s_w[wp];
    double* sw = s_s[wp];
    unsigned char* sf = s_f[wp];
    for (unsigned q = blockIdx.x * 4 + wp; q < nq; q += gridDim.x * 4) {
        const long long v = (long long)queue[q];
        double w[NS];
        bool ok[NS];
#pragma unroll
        for (int s = 0; s < NS; s++) {
            const int i = lane + 32 * s;
            const ET x = i < nfiles ? __ldg(data0 + (size_t)i * S.dstride + v) : (ET)CUDART_NAN;
            ok[s] = (x == x);
            const double2 os = S.os[i < NB ? i : 0];
            w[s] = UNIT ? (double)x : (os.x + (double)x) * os.y;                 // merging.c:426
        }
        unsigned bal[NS];
        int pos[NS];
        int n = 0, first_valid = 0;
#pragma unroll
        for (int s = 0; s < NS; s++) {
            bal[s] = __ballot_sync(0xffffffffu, ok[s]);
            pos[s] = n + __popc(bal[s] & lt);
            if (n == 0 && bal[s]) first_valid = 32 * s + __ffs(bal[s]) - 1;
            n += __popc(bal[s]);
        }
        __syncwarp();
#pragma unroll
        for (int s = 0; s < NS; s++)
            if (ok[s]) fw[pos[s]] = w[s];
        __syncwarp();
        double mean = CUDART_NAN, sigma = 0.0;
        int a = 0, b = n;
        if (n == 1) mean = fw[0];
        if (n >= 2) {
            // ascending copy by ranking (equal values keep their file order: any order of equal values gives the same sums)
#pragma unroll
            for (int s = 0; s < NS; s++) {
                if (!ok[s]) continue;
                int r = 0;
                for (int j = 0; j < n; j++) {
                    const double wj = fw[j];
                    r += (wj < w[s] || (wj == w[s] && j < pos[s])) ? 1 : 0;
                }
                sw[r] = w[s];
                sf[r] = (unsigned char)(lane + 32 * s);
            }
            __syncwarp();
            // first statistics in file order (tools.c:13-22)
            double acc = 0.0;
            for (int j = 0; j < n; j++) acc += fw[j];
            mean = acc / n;
            double dev = 0.0;
            for (int j = 0; j < n; j++) dev += (fw[j] - mean) * (fw[j] - mean);
            sigma = sqrt(dev / n);
            // MAD variant (tools.c:54-77, :122-158): sigma = 1.4826 * median(|w - median|); the deviations of an ascending
            // run are two sorted runs left and right of the median, merged outward (every lane computes the same value)
            auto mad_sigma = [&](int lo, int hi) -> double {
                const int m = hi - lo;
                const int ihh = lo + (m >> 1);
                const double med = (m & 1) ? sw[ihh] : (sw[ihh] + sw[ihh - 1]) / 2;
                int jl = (m & 1) ? ihh : ihh - 1, jr = jl + 1;
                double cur = 0.0, prev = 0.0;
                for (int k = 0; k <= (m >> 1); k++) {
                    const double dl = jl >= lo ? fabs(sw[jl] - med) : CUDART_INF;
                    const double dr = jr < hi ? fabs(sw[jr] - med) : CUDART_INF;
                    prev = cur;
                    if (dl <= dr) { cur = dl; jl--; } else { cur = dr; jr++; }
                }
                return ((m & 1) ? cur : (cur + prev) / 2) * 1.4826;
            };
            if (A.mad) sigma = mad_sigma(0, n);
            int it = A.nmax;
            for (;;) {
                const int m = b - a;
                const double med = (m & 1) ? sw[a + (m >> 1)] : (sw[a + (m >> 1)] + sw[a + (m >> 1) - 1]) / 2;   // tools.c:45-48
                const double clo = med - (A.nclip_low * sigma);
                const double cup = med + (A.nclip_up * sigma);
                int ni = 0, nle = 0;
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    const int j = a + lane + 32 * s;
                    const double x = j < b ? sw[j] : 0.0;
                    const bool gt = (j < b) && (x > clo), in = gt && (x < cup);      // strict window, tools.c:95
                    ni += __popc(__ballot_sync(0xffffffffu, in));
                    nle += __popc(__ballot_sync(0xffffffffu, (j < b) && !gt));
                }
                if (ni < nstop || ni == m) break;                                      // tools.c:100-103
                if (it <= 0) break;                                                    // tools.c:104
                it--;
                a += nle;
                b = a + ni;
                acc = 0.0;                                                             // survivors, ascending order (tools.c:116 -> :13-22)
                for (int j = a; j < b; j++) acc += sw[j];
                mean = acc / ni;
                dev = 0.0;
                for (int j = a; j < b; j++) dev += (sw[j] - mean) * (sw[j] - mean);
                sigma = A.mad ? mad_sigma(a, b) : sqrt(dev / ni);
            }
        }
        const int kept = b - a;
        double ov = CUDART_NAN;
        if (n >= 1) {
            if (A.typ_var == 0) {
                // variances of the kept samples in the order of their values (merging.c:461-462)
                __syncwarp();
#pragma unroll
                for (int s = 0; s < NS; s++) {
                    const int j = a + lane + 32 * s;
                    if (j < b) {
                        const int i = n >= 2 ? (int)sf[j] : first_valid;
                        fw[j - a] = (double)__ldg(var0 + (size_t)i * S.vstride + v) * S.scale2[i];
                    }
                }
                __syncwarp();
                double vs = 0.0;
                for (int j = 0; j < kept; j++) vs += fw[j];
                ov = vs / ((double)kept * (double)kept);
            } else if (kept > 1) {
                ov = sigma * sigma;                                                    // merging.c:465
                if (A.typ_var == 1) ov /= (double)(kept - 1);                          // merging.c:467
            }
        }
        if (n >= 2) {
#pragma unroll
            for (int s = 0; s < NS; s++) {
                const int j = lane + 32 * s;
                if (j < n && (j < a || j >= b)) atomicAdd(A.select_cnt + sf[j], ~0ull);   // -1: valid but rejected
            }
        }
        if (lane == 0) {
            A.out[v] = mean;
            A.outvar[v] = ov;
            A.expmap[v] = kept;
            if (S.hist != nullptr) atomicAdd(S.hist + kept, 1ull);
        }
        __syncwarp();
    }
}

template <int NB>
__device__ __forceinline__ void sort_pairs(unsigned (&a)[NB])
{
#define MB_CE(i, j)                                                             \
    {                                                                           \
        unsigned lo_, hi_;                                                      \
        asm("min.bf16x2 %0, %1, %2;" : "=r"(lo_) : "r"(a[i]), "r"(a[j]));       \
        asm("max.bf16x2 %0, %1, %2;" : "=r"(hi_) : "r"(a[i]), "r"(a[j]));       \
        a[i] = lo_;                                                             \
        a[j] = hi_;                                                             \
    }
    if constexpr (NB == 4) { MB_SORTNET_4(MB_CE) }
    else if constexpr (NB == 8) { MB_SORTNET_8(MB_CE) }
    else if constexpr (NB == 12) { MB_SORTNET_12(MB_CE) }
    else if constexpr (NB == 16) { MB_SORTNET_16(MB_CE) }
    else if constexpr (NB == 24) { MB_SORTNET_24(MB_CE) }
    else if constexpr (NB == 32) { MB_SORTNET_32(MB_CE) }
    else if constexpr (NB == 48) { MB_SORTNET_48(MB_CE) }
    else if constexpr (NB == 64) { MB_SORTNET_64(MB_CE) }
#undef MB_CE
}

// Grid: one block per SM (persistent).  Block: PairCfg<NB>::C warps, all of them consumers.
// Tiles of 64 consecutive voxels; block b takes tiles b, b + grid, b + 2 grid, ...; its k-th tile is
// processed by warp k % C in that warp's own stage.  flags & 2 (default): the warps start their rounds of C
// tiles together in two groups (named barriers), so that a line of the long straight-line loop body is
// fetched once per SM and round instead of once per warp; flags bits 8..11: rounds that skip the barrier
// (default 1: re-aligned every second round); flags bits 4..7: number of groups (tuning).
template <int NB, bool UNIT, bool PROP, bool MAD = false, typename ET = float>
__global__ void __launch_bounds__(PairCfg<NB, sizeof(ET)>::threads, 1)
sigclip_pair_kernel(const PairArgs<NB> S, const ClipArgs A, const long long ntiles, unsigned* __restrict__ slow_q,
                    unsigned* __restrict__ slow_n, const unsigned slow_cap, const int flags, const __grid_constant__ TmaMaps M)
{
    using Cfg = PairCfg<NB, sizeof(ET)>;
    using ET2 = typename Vec2<ET>::type;
    constexpr bool F64 = sizeof(ET) == 8;
    constexpr int C = Cfg::C, ST = Cfg::ST;
    constexpr int W = ExpSet<NB>::W;
    constexpr unsigned FULLM = 0xffffffffu;
    constexpr unsigned kTileBytes = NB * 64 * sizeof(ET);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    PairSmem<NB, ET>& sm = *reinterpret_cast<PairSmem<NB, ET>*>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int nstop = A.nstop < 1 ? 1 : A.nstop;

    for (int i = tid; i <= NB; i += Cfg::threads) sm.rcp[i] = i > 0 ? 1.0 / (double)i : 0.0;
    for (int i = tid; i < NB; i += Cfg::threads) {
        sm.os[i] = S.os[i];
        sm.osp[i] = S.osp[i];
    }
    for (int i = tid; i <= NB; i += Cfg::threads) sm.hist[i] = 0u;
    if (tid == 0) {
        for (int s = 0; s < ST; s++) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();   // the only block-wide barrier

    // ---- consumer warps -------------------------------------------------------------------------------
    const int c = warp;
    TransposeConsts tc;
    tc.init(lane);
    unsigned cnt_inval[W], cnt_drop[W];
#pragma unroll
    for (int k = 0; k < W; k++) { cnt_inval[k] = 0u; cnt_drop[k] = 0u; }
    unsigned tiles_done = 0;
    // key q of voxel h: colh[q * 64 + h].  The clip loop also reads up to five ranks beyond either end of the column
    // (the neighbouring warp's column, the stage before or the tables behind the array): whatever it finds there can
    // only add candidates beyond the window's own keys, which sends the voxel to the exact kernel.
    const unsigned short* colh = reinterpret_cast<const unsigned short*>(&sm.col[c][0][lane]);
    auto col16 = [&](int q, int h) -> unsigned { return (unsigned)colh[q * 64 + h]; };
    auto keyf = [&](int q, int h) -> float { return __uint_as_float(col16(q, h) << 16); };
    const float nlo32 = (float)A.nclip_low, nup32 = (float)A.nclip_up;
    // float32 images of the clip factors must not be smaller than the factors (conservative bounds need |error| only)
    const float krel32 = (float)kPairKrel, c1f = (float)(kPairC1 * (1.0 - 0x1p-20)), c2f = (float)(kPairC2 * (1.0 + 0x1p-20));

    // rounds of C tiles; the trip count is uniform over the consumer warps so that the round barrier is legal
    const long long my_tiles = ntiles > (long long)blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long rounds = (my_tiles + C - 1) / C;
    // this warp's tiles: the c-th of every round; its j-th tile uses stage c * SPW + j % SPW.  Request the first SPW tiles.
    auto request_tile = [&](long long j) {      // lane 0 only
        const long long kk = j * C + c;
        if (kk >= my_tiles) return;
        const unsigned st = (unsigned)(c * Cfg::SPW + (int)(j % Cfg::SPW));
        mbar_expect_tx(&sm.full[st], kTileBytes);
        tma_load_2d(&sm.stage[st][0][0], &M.data, (int)(((long long)blockIdx.x + kk * gridDim.x) << 6), 0, &sm.full[st]);
    };
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < Cfg::SPW; j++) request_tile(j);
    }
    for (long long r = 0; r < rounds; r++) {
        if ((r & ((flags >> 8) & 15)) != 0) {
            // tuning: flags bits 8..11 = mask; rounds with (r & mask) != 0 run without the phase lock
        } else if ((flags >> 4) & 15) {
            // tuning: flags >> 4 groups of consecutive warps, each phase-locked on its own barrier
            const int ng = (flags >> 4) & 15, base = C / ng, rem = C % ng;
            const int g = c < rem * (base + 1) ? c / (base + 1) : rem + (c - rem * (base + 1)) / (base > 0 ? base : 1);
            const int cnt = (g < rem ? base + 1 : base) * 32;
            asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(cnt) : "memory");
        } else if (flags & 2) {
            // default: two groups of consumer warps, each phase-locked on its own barrier and free to drift against the
            // other (measured 0 .. 3 % over one group: while one group sorts on the ALU pipe the other can be in pass 1)
            constexpr int G0 = (C + 1) / 2, G1 = C - G0;
            if (c < G0) asm volatile("bar.sync 1, %0;" ::"n"(G0 * 32) : "memory");
            else asm volatile("bar.sync 2, %0;" ::"n"((G1 > 0 ? G1 : 1) * 32) : "memory");
        } else if (flags & 1) asm volatile("bar.sync 1, %0;" ::"n"(C * 32) : "memory");
        const long long kk64 = r * C + c;
        if (kk64 >= my_tiles) continue;
        const long long tile = (long long)blockIdx.x + kk64 * gridDim.x;
        const unsigned st = (unsigned)(c * Cfg::SPW + (int)(r % Cfg::SPW)), ph = (unsigned)(r / Cfg::SPW) & 1u;
        ET (*dt)[64] = sm.stage[st];
        const long long v0 = (tile << 6) + 2 * lane;            // this lane's voxels: v0, v0 + 1
        // the STAT rows of these voxels are read at the end of the tile: start them on their way into L2 now
        if (PROP && lane == 0) tma_prefetch_2d(&M.stat, (int)(tile << 6), 0);
        const bool pair_ok = v0 + 1 < A.nvox;                   // both voxels exist: 8-byte STAT loads are in bounds
        mbar_wait(&sm.full[st], ph);

        // ---- pivot: median of the first three samples (robust against one outlier), else the fourth, else 0
        float pf[2];
        {
            const ET2 r0 = *reinterpret_cast<const ET2*>(&dt[0][2 * lane]);
            const ET2 r1 = *reinterpret_cast<const ET2*>(&dt[1][2 * lane]);
            const ET2 r2 = *reinterpret_cast<const ET2*>(&dt[2][2 * lane]);
            const ET2 r3 = *reinterpret_cast<const ET2*>(&dt[3][2 * lane]);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float x0, x1, x2, x3;
                if constexpr (F64) {
                    const double y0 = h ? r0.y : r0.x, y1 = h ? r1.y : r1.x, y2 = h ? r2.y : r2.x, y3 = h ? r3.y : r3.x;
                    x0 = __double2float_rn(UNIT ? y0 : __fma_rn(y0, S.os[0].y, S.osp[0]));
                    x1 = __double2float_rn(UNIT ? y1 : __fma_rn(y1, S.os[1].y, S.osp[1]));
                    x2 = __double2float_rn(UNIT ? y2 : __fma_rn(y2, S.os[2].y, S.osp[2]));
                    x3 = __double2float_rn(UNIT ? y3 : __fma_rn(y3, S.os[3].y, S.osp[3]));
                } else {
                    x0 = h ? r0.y : r0.x, x1 = h ? r1.y : r1.x, x2 = h ? r2.y : r2.x, x3 = h ? r3.y : r3.x;
                    if constexpr (!UNIT) {
                        x0 = __fmaf_rn(x0, S.kscale[0], S.koffset[0]);
                        x1 = __fmaf_rn(x1, S.kscale[1], S.koffset[1]);
                        x2 = __fmaf_rn(x2, S.kscale[2], S.koffset[2]);
                        x3 = __fmaf_rn(x3, S.kscale[3], S.koffset[3]);
                    }
                }
                float m3 = fmaxf(fminf(x0, x1), fminf(fmaxf(x0, x1), x2));   // fminf / fmaxf skip NaNs
                m3 = (m3 == m3) ? m3 : x3;
                m3 = (m3 == m3 && fabsf(m3) < 0x1p100f) ? m3 : 0.f;
                pf[h] = m3;
            }
        }
        const double p[2] = {(double)pf[0], (double)pf[1]};

        // ---- pass 1: exposure order; float64 shifted moments, float32 key values, invalid masks ------------
        // Per voxel and exposure: one conversion, t = w - pivot as ONE fused multiply-add on (offset * scale - pivot)
        // [w itself is never needed: |t - (w_ref - pivot)| <= 2^-52 (|w| + |pivot|), far inside the margin of the
        // bounds], two moment updates, and the float32 key value.  Per voxel PAIR: one conversion to packed
        // bfloat16, one NaN test giving both voxels' "invalid" predicates, one packed max that turns NaN keys into
        // the sentinel; an invalid sample's t gets its high word cleared (a denormal: the unconditional moment
        // updates absorb it without a trace), its mask bit is added under the predicate.
        unsigned key[NB], ukey[NB];
        ExpSet<NB> inval[2];
        double T1[2] = {0.0, 0.0}, S2[2] = {0.0, 0.0};
#pragma unroll
        for (int i = 0; i < NB; i++) {
            const ET2 dd = *reinterpret_cast<const ET2*>(&dt[i][2 * lane]);
            double t[2];
            float tf[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const ET d = h ? dd.y : dd.x;
                if constexpr (UNIT) t[h] = (double)d - p[h];
                else t[h] = __fma_rn((double)d, S.os[i].y, S.osp[i] - p[h]);     // (offset + d) * scale - pivot, merging.c:426
                if constexpr (F64) tf[h] = __double2float_rn(t[h]);               // float64 samples: the key value is float32(t)
                else if constexpr (UNIT) tf[h] = (float)d - pf[h];
                else tf[h] = __fmaf_rn((float)d, S.kscale[i], S.koffset[i]) - pf[h];
            }
            {
                asm("{\n\t.reg .pred q0, q1;\n\t.reg .b32 lo0, hi0, lo1, hi1;\n\t"
                    "cvt.rn.bf16x2.f32 %0, %7, %6;\n\t"                 // high half <- voxel 1
                    "setp.nan.bf16x2 q0|q1, %0, %0;\n\t"
                    "max.bf16x2 %1, %0, %8;\n\t"                        // NaN halves -> sentinel key
                    "mov.b64 {lo0, hi0}, %2;\n\t"
                    "mov.b64 {lo1, hi1}, %3;\n\t"
                    "selp.b32 hi0, 0, hi0, q0;\n\t"
                    "selp.b32 hi1, 0, hi1, q1;\n\t"
                    "mov.b64 %2, {lo0, hi0};\n\t"
                    "mov.b64 %3, {lo1, hi1};\n\t"
                    "@q0 add.u32 %4, %4, %9;\n\t"
                    "@q1 add.u32 %5, %5, %9;\n\t}"
                    : "=r"(ukey[i]), "=r"(key[i]), "+d"(t[0]), "+d"(t[1]), "+r"(inval[0].w[i >> 5]), "+r"(inval[1].w[i >> 5])
                    : "f"(tf[0]), "f"(tf[1]), "r"(0xFF7FFF7Fu), "r"(1u << (i & 31)));
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                T1[h] += t[h];
                S2[h] = __fma_rn(t[h], t[h], S2[h]);
            }
        }
        int n[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            n[h] = NB - inval[h].count();
            // what the cleared-high-word denormals of invalid samples may have left behind when every real term is 0
            if (fabs(T1[h]) < 0x1p-1000) T1[h] = 0.0;
        }

        // ---- order the packed keys; sorted column to shared memory ------------------------------------------
        sort_pairs<NB>(key);
#pragma unroll
        for (int q = 0; q < NB; q++) sm.col[c][q][lane] = key[q];
        __syncwarp();

        // ---- clip iterations (tools.c:82-118), both voxels in lockstep ---------------------------------------
        // The variance comes from the float64 moments; everything that only feeds DECISIONS (sigma, bounds, key-space
        // thresholds) is float32 arithmetic whose rounding is covered by the 2^-19 relative term of the margin.
        int a[2], b[2], itl[2];
        bool done[2], slow[2];
        double varl[2] = {0.0, 0.0}, S2_0[2];
        float e0[2];
        ExpSet<NB> rej[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            a[h] = NB - n[h];
            b[h] = NB;
            itl[h] = A.nmax;
            S2_0[h] = S2[h];
            slow[h] = false;
            // S2 == 0: every valid sample equals the pivot, sigma == 0, nothing lies strictly inside the empty
            // window: the reference stops at once with all samples kept (tools.c:100)
            done[h] = (n[h] < 2) || (S2[h] == 0.0);
            if (n[h] >= 2 && !(S2[h] < 0x1p200)) { slow[h] = true; done[h] = true; }   // non-finite or absurd values
            e0[h] = 0x1p-21f * (fabsf(pf[h]) + (float)S.osmax) + 0x1p-100f;
        }
        // The statistics of the set that no clip follows (tools.c:104 / :144).  Only the 'stat' variances need them; with
        // var = 'propagate' sigma feeds nothing any more and the voxel is simply finished.
        auto final_stats = [&](int h) {
            if constexpr (!PROP) {
                const int m = b[h] - a[h];
                const double rm = sm.rcp[m];
                const double dm1 = T1[h] * rm;
                const double var_k = __fma_rn(S2[h], rm, -(dm1 * dm1));
                if (!(var_k >= 0x1p-13 * (S2_0[h] * rm)) || !(var_k < 0x1p100) || !(var_k > 0x1p-100)) slow[h] = true;
                else varl[h] = var_k;
            }
            done[h] = true;
        };
#pragma unroll
        for (int h = 0; h < 2; h++)
            if (!done[h] && itl[h] <= 0) final_stats(h);
        for (;;) {
            if (!__any_sync(FULLM, !(done[0] && done[1]))) break;
            unsigned thr_lo, thr_hi;                               // packed boundary keys of the two voxels
            int ncl[2], nch[2];                                    // candidates for rejection at the low / high end
            float clo32[2], cup32[2], tb32[2];
            unsigned BLh[2], BUh[2];
            // ---- part A: sigma, bounds, and the keys that are not CERTAINLY inside the window -------------------
            // Branch-free and evaluated for both voxels of the lane whether or not they are still active (the warp issues
            // these instructions as long as ANY lane needs them; without branches the two voxels' chains interleave).
            // Ranks beyond either end of the window read rejected / invalid keys or the guard rows: never fewer candidates.
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const bool act = !done[h];
                const int m = b[h] - a[h];
                const double rm = sm.rcp[m];
                const double dm1 = T1[h] * rm;
                const double var_k = __fma_rn(S2[h], rm, -(dm1 * dm1));
                // conditioning: the moment variance must dominate its rounding error; and it must sit in float32 range
                // (MAD: the moments only give the mean; sigma comes from the keys, its uncertainty is part of Tb below)
                const bool bad = !MAD && (!(var_k >= 0x1p-13 * (S2_0[h] * rm)) || !(var_k < 0x1p100) || !(var_k > 0x1p-100));
                if constexpr (!PROP) varl[h] = (act && !bad) ? var_k : varl[h];
                const unsigned short* pm = colh + ((a[h] + (m >> 1)) * 64 + h);       // upper middle of the window
                const float kh = __uint_as_float((unsigned)pm[0] << 16);
                const float kl = __uint_as_float((unsigned)((m & 1) ? pm[0] : pm[-64]) << 16);
                const float km = 0.5f * (kh + kl);                                   // median of the keys (tools.c:45-48)
                const float emed = krel32 * fmaxf(fabsf(kh), fabsf(kl)) + e0[h];     // its distance from the samples' median
                float sg, esg = 0.f;
                if constexpr (MAD) {
                    // sigma = 1.4826 * median(|w - median|) (tools.c:67-74): the middle of the absolute deviations is found
                    // by merging the two sorted runs left and right of the median outward, on the keys.  A key deviation
                    // differs from its sample's by <= e(key) + emed, and only samples within about one MAD of the median
                    // decide the middle, so the MAD from the keys is within emad of the reference's.
                    sg = 0.f;
                    if (act) {
                        const int il = a[h] + (m >> 1) - 1 + (m & 1);
                        int jl = il, jr = il + 1;
                        float dl = fabsf(keyf(jl, h) - km), dr = jr < b[h] ? fabsf(keyf(jr, h) - km) : CUDART_INF_F;
                        float cur = 0.f, prev = 0.f;
                        for (int s2 = 0; s2 <= (m >> 1); s2++) {
                            prev = cur;
                            if (dl <= dr) {
                                cur = dl;
                                jl--;
                                dl = jl >= a[h] ? fabsf(keyf(jl, h) - km) : CUDART_INF_F;
                            } else {
                                cur = dr;
                                jr++;
                                dr = jr < b[h] ? fabsf(keyf(jr, h) - km) : CUDART_INF_F;
                            }
                        }
                        const float madk = (m & 1) ? cur : 0.5f * (cur + prev);
                        sg = madk * 1.4826f;
                        esg = 1.4826f * (krel32 * 1.02f * (fabsf(km) + madk) + e0[h] + emed) + 0x1p-20f * sg;
                    }
                } else {
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(__double2float_rn(var_k)));   // 2^-21 relative in all
                }
                const float lo_s = nlo32 * sg, up_s = nup32 * sg;
                const float clo = km - lo_s, cup = km + up_s;
                // uncertainty of a bound: key median, sigma, float32 roundings of this very arithmetic
                const float Tb = emed + fmaxf(nlo32, nup32) * esg + 0x1p-19f * (fabsf(km) + lo_s + up_s);
                const bool bad_tb = !(Tb < 0x1p100f);                                 // NaN / inf bounds
                // a key in (kL2, kU2) belongs to a sample that is certainly inside the window
                const float yl = clo + Tb + e0[h], xu = cup - Tb - e0[h];
                const float kL2 = yl * (yl >= 0.f ? c2f : c1f);
                const float kU2 = xu * (xu >= 0.f ? c1f : c2f);
                const unsigned short* plo = colh + (a[h] * 64 + h);                    // lowest rank of the window
                const unsigned short* phi = colh + ((b[h] - 1) * 64 + h);              // highest
                int jl = 0, ju = 0;
                bool ql = true, qu = true;
#pragma unroll
                for (int j = 0; j < 5; j++) {
                    ql = ql && !(__uint_as_float((unsigned)plo[64 * j] << 16) > kL2);
                    qu = qu && !(__uint_as_float((unsigned)phi[-64 * j] << 16) < kU2);
                    jl += ql ? 1 : 0;
                    ju += qu ? 1 : 0;
                }
                // at most four candidates per side (the probes' reach), and the two ends must not meet
                const bool many = jl > 4 || ju > 4 || jl + ju > m;
                const bool go_slow = bad || bad_tb || many;
                const bool cand = act && !go_slow && (jl + ju) != 0;                  // jl + ju == 0: everything inside, tools.c:100
                slow[h] = slow[h] || (act && go_slow);
                done[h] = !cand;
                ncl[h] = cand ? jl : 0;
                nch[h] = cand ? ju : 0;
                clo32[h] = clo;
                cup32[h] = cup;
                tb32[h] = Tb;
                // boundary keys: the candidates are the window's keys at or below BL / at or above BU
                const unsigned bl = (unsigned)plo[64 * (jl - 1)], bu = (unsigned)phi[-64 * (ju - 1)];
                BLh[h] = (cand && jl > 0) ? bl : 0xFF80u;                             // -inf / +inf: match nothing
                BUh[h] = (cand && ju > 0) ? bu : 0x7F80u;
            }
            thr_lo = BLh[0] | (BLh[1] << 16);
            thr_hi = BUh[0] | (BUh[1] << 16);
            if (!__any_sync(FULLM, (ncl[0] | nch[0] | ncl[1] | nch[1]) != 0)) continue;
            // ---- which exposures are those?  one packed comparison pass over the unsorted keys -------------------
            // (invalid samples are NaN in the unsorted keys and match nothing; the bits of a mask word are distinct, so
            // they are ADDED under the predicates: predicated integer adds are full rate, LOP3 is not)
            unsigned RL[2][W], RH[2][W];
#pragma unroll
            for (int kk = 0; kk < W; kk++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    asm volatile("mov.b32 %0, 0;" : "=r"(RL[h][kk]));
                    asm volatile("mov.b32 %0, 0;" : "=r"(RH[h][kk]));
                }
            }
            // Cosmic-ray-like outliers are high: most passes have no candidate at the low end of any window of the warp
            // (or none at the high end) and need only one comparison per exposure.
            const bool any_low = __any_sync(FULLM, (ncl[0] | ncl[1]) != 0), any_high = __any_sync(FULLM, (nch[0] | nch[1]) != 0);
            if (any_low && any_high) {
#pragma unroll
                for (int i = 0; i < NB; i++) {
                    asm("{\n\t.reg .pred p0, p1, p2, p3;\n\t"
                        "setp.le.bf16x2 p0|p1, %4, %5;\n\t"
                        "setp.ge.bf16x2 p2|p3, %4, %6;\n\t"
                        "@p0 add.u32 %0, %0, %7;\n\t"
                        "@p1 add.u32 %1, %1, %7;\n\t"
                        "@p2 add.u32 %2, %2, %7;\n\t"
                        "@p3 add.u32 %3, %3, %7;\n\t}"
                        : "+r"(RL[0][i >> 5]), "+r"(RL[1][i >> 5]), "+r"(RH[0][i >> 5]), "+r"(RH[1][i >> 5])
                        : "r"(ukey[i]), "r"(thr_lo), "r"(thr_hi), "r"(1u << (i & 31)));
                }
            } else if (any_high) {
#pragma unroll
                for (int i = 0; i < NB; i++) {
                    asm("{\n\t.reg .pred p2, p3;\n\t"
                        "setp.ge.bf16x2 p2|p3, %2, %3;\n\t"
                        "@p2 add.u32 %0, %0, %4;\n\t"
                        "@p3 add.u32 %1, %1, %4;\n\t}"
                        : "+r"(RH[0][i >> 5]), "+r"(RH[1][i >> 5])
                        : "r"(ukey[i]), "r"(thr_hi), "r"(1u << (i & 31)));
                }
            } else {
#pragma unroll
                for (int i = 0; i < NB; i++) {
                    asm("{\n\t.reg .pred p0, p1;\n\t"
                        "setp.le.bf16x2 p0|p1, %2, %3;\n\t"
                        "@p0 add.u32 %0, %0, %4;\n\t"
                        "@p1 add.u32 %1, %1, %4;\n\t}"
                        : "+r"(RL[0][i >> 5]), "+r"(RL[1][i >> 5])
                        : "r"(ukey[i]), "r"(thr_lo), "r"(1u << (i & 31)));
                }
            }
            // ---- part B: the candidates' exact values decide (their samples are still in the stage) ---------------------
#pragma unroll
            for (int h = 0; h < 2; h++) {
                int cl = 0, ch = 0;
#pragma unroll
                for (int kk = 0; kk < W; kk++) {
                    RL[h][kk] &= ~rej[h].w[kk];      // samples rejected earlier lie beyond the window
                    RH[h][kk] &= ~rej[h].w[kk];
                    cl += __popc(RL[h][kk]);
                    ch += __popc(RH[h][kk]);
                }
                if ((ncl[h] | nch[h]) != 0 && (cl != ncl[h] || ch != nch[h])) {   // key ties the model does not cover
                    slow[h] = done[h] = true;
                    ncl[h] = nch[h] = 0;
                }
                if ((ncl[h] | nch[h]) == 0) continue;
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if ((ncl[h] | nch[h]) == 0) continue;
                const double clo_d = (double)clo32[h], cup_d = (double)cup32[h], tb_d = (double)tb32[h];
                double dT1 = 0.0, dS2 = 0.0;
                int rl = 0, rh = 0;
                bool unc = false;
                unsigned long long nr = 0ull;                                        // candidates rejected by this evaluation
                // one loop over the candidates of all mask words (W <= 2: a 64-bit mask)
                unsigned long long mw = (unsigned long long)(RL[h][0] | RH[h][0]), lowm = (unsigned long long)RL[h][0];
                if constexpr (W > 1) {
                    mw |= (unsigned long long)(RL[h][W - 1] | RH[h][W - 1]) << 32;
                    lowm |= (unsigned long long)RL[h][W - 1] << 32;
                }
                while (mw) {
                    const int i = __ffsll((long long)mw) - 1;
                    const unsigned long long rest = mw & (mw - 1ull), bit = mw ^ rest;
                    mw = rest;
                    const bool low = (lowm >> i) & 1ull;
                    const ET dval = dt[i][2 * lane + h];
                    const double t = UNIT ? (double)dval - p[h] : __fma_rn((double)dval, sm.os[i].y, sm.osp[i] - p[h]);
                    const double diff = t - (low ? clo_d : cup_d);
                    unc = unc || !(fabs(diff) > tb_d);
                    const bool out = low ? (diff <= 0.0) : (diff >= 0.0);            // strict window, tools.c:95
                    if (out) {
                        dT1 += t;
                        dS2 += t * t;
                        nr |= bit;
                        if (low) rl++; else rh++;
                    }
                }
                // An undecidable sample sends the voxel to the exact kernel.  Otherwise every rejected candidate lies more than
                // Tb beyond its bound and every kept one more than Tb inside: their float32 key values (each within
                // 2^-22 |t| + E0 of t, and Tb >= 2^-19 |bound| + E0) are ordered the same way, so the rejected candidates hold
                // the outermost ranks of the sorted column and the surviving window is again a contiguous range of ranks.
                if (unc) { slow[h] = done[h] = true; continue; }
                const int m = b[h] - a[h];
                const int ni = m - rl - rh;
                if (ni < nstop || ni == m) { done[h] = true; continue; }             // tools.c:100-103
                itl[h]--;
                T1[h] -= dT1;
                S2[h] -= dS2;
#pragma unroll
                for (int kk = 0; kk < W; kk++) rej[h].w[kk] |= (unsigned)(nr >> (kk ? 32 : 0));
                a[h] += rl;
                b[h] -= rh;
                if (itl[h] <= 0) final_stats(h);                                     // no clip left: these statistics stand
            }
        }
        // ---- the samples are consumed: request the tile that uses this stage next ---------------------------------
        // (the warp's own reads of the stage are complete: no other warp ever touches it)
        __syncwarp();
        if (lane == 0) request_tile(r + Cfg::SPW);

        // ---- results of the two voxels ------------------------------------------------------------------------
        int kept[2];
        double mean[2];
        ExpSet<NB> drop[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            kept[h] = b[h] - a[h];
            mean[h] = n[h] > 0 ? p[h] + T1[h] * sm.rcp[kept[h]] : CUDART_NAN;        // merging.c:439
#pragma unroll
            for (int kk = 0; kk < W; kk++) {
                if (slow[h]) rej[h].w[kk] = 0u;                                     // recomputed later: count no rejection here
                drop[h].w[kk] = inval[h].w[kk] | rej[h].w[kk];
            }
        }

        // ---- propagated variance over the kept set (merging.c:431,461-462) ---------------------------------------
        // STAT rows straight from L2 (prefetched at the start of the tile) into the registers the keys vacated, two
        // half-batches; float32 FMAs in chains of <= 3 terms, chains summed in float64: <= 4 * 2^-24 relative for
        // non-negative STAT.  A sign bit among the loaded values (a negative or negative-NaN STAT, kept or not)
        // sends the lane to float64 products in exposure order.
        // ---- per-exposure counters: transpose the lanes' masks, one POPC per lane ----------------------------------
        auto count_tile = [&]() {
#pragma unroll
            for (int h = 0; h < 2; h++) {
#pragma unroll
                for (int kk = 0; kk < W; kk++) {
                    cnt_inval[kk] += __popc(tc.apply(inval[h].w[kk]));
                    cnt_drop[kk] += __popc(tc.apply(drop[h].w[kk]));
                }
            }
            tiles_done++;
        };
        bool counted = false;
        double vsum[2] = {0.0, 0.0};
        if constexpr (PROP) {
            constexpr int NACC = (NB + 2) / 3;
            float vacc[2][NACC];
#pragma unroll
            for (int j = 0; j < NACC; j++) { vacc[0][j] = 0.f; vacc[1][j] = 0.f; }
            unsigned neg = 0u;
            if constexpr (F64) {
                // float64 STAT: float64 products (the reference's arithmetic, merging.c:431), loads in batches of 8 rows
                if (pair_ok) {
                    const double* vp = static_cast<const double*>(S.var0) + v0;
                    constexpr int VB = 8;
#pragma unroll
                    for (int i0 = 0; i0 < NB; i0 += VB) {
                        double2 sv[VB];
#pragma unroll
                        for (int j = 0; j < VB; j++) {
                            if (i0 + j < NB) {
                                asm volatile("ld.global.cs.v2.f64 {%0, %1}, [%2];" : "=d"(sv[j].x), "=d"(sv[j].y) : "l"(vp));
                                if (i0 + j + 1 < A.nfiles) vp += S.vstride;       // slots beyond the stack re-read its last row (dropped anyway)
                            }
                        }
                        asm volatile("" ::: "memory");
#pragma unroll
                        for (int j = 0; j < VB; j++) {
                            if (i0 + j < NB) {
                                if (!drop[0].test(i0 + j)) vsum[0] += sv[j].x * S.scale2[i0 + j];
                                if (!drop[1].test(i0 + j)) vsum[1] += sv[j].y * S.scale2[i0 + j];
                            }
                        }
                    }
                }
            } else if (pair_ok) {
                const float* vp = static_cast<const float*>(S.var0) + v0;
                float2 sv[NB];
                // Slots beyond a stack that does not fill its bucket must not be addressed (their samples are invalid -- NaN-filled
                // by the tensor map -- hence dropped): they re-read the stack's last row.  Two forms, chosen by measurement: a
                // pointer chain with a separate clamped copy for partial buckets (fastest without MAD), and row = min(slot,
                // nfiles - 1) computed per load (the chain form costs the larger MAD kernels 25 .. 40 %: N = 48, 4.29 against 3.09 ms)
                if constexpr (MAD) {
                    const int last = A.nfiles - 1;
#pragma unroll
                    for (int i = 0; i < NB; i++) {
                        const float* rp = vp + (size_t)(i < last ? i : last) * S.vstride;
                        asm volatile("ld.global.cs.v2.f32 {%0, %1}, [%2];" : "=f"(sv[i].x), "=f"(sv[i].y) : "l"(rp));
                    }
                } else if (A.nfiles == NB) {
#pragma unroll
                    for (int i = 0; i < NB; i++) {
                        asm volatile("ld.global.cs.v2.f32 {%0, %1}, [%2];" : "=f"(sv[i].x), "=f"(sv[i].y) : "l"(vp));
                        vp += S.vstride;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < NB; i++) {
                        asm volatile("ld.global.cs.v2.f32 {%0, %1}, [%2];" : "=f"(sv[i].x), "=f"(sv[i].y) : "l"(vp));
                        if (i + 1 < A.nfiles) vp += S.vstride;
                    }
                }
                asm volatile("" ::: "memory");
                count_tile();                     // the counter transposes run while the STAT rows are on their way
                counted = true;
#pragma unroll
                for (int i = 0; i < NB; i++) {
                    if (!drop[0].test(i)) vacc[0][i % NACC] = __fmaf_rn(sv[i].x, S.scale2f[i], vacc[0][i % NACC]);
                    if (!drop[1].test(i)) vacc[1][i % NACC] = __fmaf_rn(sv[i].y, S.scale2f[i], vacc[1][i % NACC]);
                    neg |= __float_as_uint(sv[i].x) | __float_as_uint(sv[i].y);
                }
#pragma unroll
                for (int j = 0; j < NACC; j++) { vsum[0] += (double)vacc[0][j]; vsum[1] += (double)vacc[1][j]; }
            }
            if (!pair_ok || (neg & 0x80000000u)) {
                // ragged last pair, or a sign bit was seen: float64 products in exposure order (merging.c:431)
                const ET* var0 = static_cast<const ET*>(S.var0);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    vsum[h] = 0.0;
                    if (v0 + h >= A.nvox) continue;
                    for (int i = 0; i < A.nfiles; i++)
                        if (!drop[h].test_dyn(i)) vsum[h] += (double)__ldg(var0 + (size_t)i * S.vstride + (v0 + h)) * (sm.os[i].y * sm.os[i].y);
                }
            }
        }

        if (!counted) count_tile();

        // ---- stores ---------------------------------------------------------------------------------------------
        // (divisions by the small integers as multiplications by their correctly rounded reciprocals: <= 2 ulp)
        double ov[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const double rk = sm.rcp[kept[h]];
            if (n[h] == 0) ov[h] = CUDART_NAN;
            else if (PROP) ov[h] = vsum[h] * rk * rk;                                       // merging.c:462
            else if (kept[h] > 1) {
                ov[h] = varl[h];                                                            // sigma^2, merging.c:465
                if (A.typ_var == 1) ov[h] *= sm.rcp[kept[h] - 1];                           // merging.c:467
            } else ov[h] = CUDART_NAN;
        }
        // exposure-map histogram (expnb, cubelist.py:69-74): the lanes of a warp mostly share one value
        if (S.hist != nullptr) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int val = (slow[h] || v0 + h >= A.nvox) ? -1 : kept[h];             // queued voxels are counted by the exact kernel
                const unsigned peers = __match_any_sync(FULLM, val);
                if (val >= 0 && lane == __ffs(peers) - 1) atomicAdd(&sm.hist[val], (unsigned)__popc(peers));
            }
        }
        if (v0 + 1 < A.nvox && !slow[0] && !slow[1]) {
            *reinterpret_cast<double2*>(A.out + v0) = make_double2(mean[0], mean[1]);
            *reinterpret_cast<double2*>(A.outvar + v0) = make_double2(ov[0], ov[1]);
            *reinterpret_cast<int2*>(A.expmap + v0) = make_int2(kept[0], kept[1]);
        } else {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const long long v = v0 + h;
                if (v >= A.nvox) continue;
                if (!slow[h]) {
                    A.out[v] = mean[h];
                    A.outvar[v] = ov[h];
                    A.expmap[v] = kept[h];
                } else {
                    const unsigned pos = atomicAdd(slow_n, 1u);
                    if (pos < slow_cap) slow_q[pos] = (unsigned)v;
                    else exact_voxel_store<NB, UNIT, ET>(S, A, v);   // queue full: recompute in place
                }
            }
        }
    }

    // ---- flush: valid = voxels seen - invalid, selected = voxels seen - dropped -------------------------------------
    const unsigned seen = tiles_done * 64u;
#pragma unroll
    for (int kk = 0; kk < W; kk++) {
        const int e = 32 * kk + lane;
        if (e < A.nfiles) {
            atomicAdd(A.valid_cnt + e, (unsigned long long)(seen - cnt_inval[kk]));
            atomicAdd(A.select_cnt + e, (unsigned long long)(seen - cnt_drop[kk]));
        }
    }
    if (S.hist != nullptr) {
        __syncthreads();      // every warp runs the same number of rounds and none has returned
        for (int i = tid; i <= NB; i += Cfg::threads)
            if (sm.hist[i]) atomicAdd(S.hist + i, (unsigned long long)sm.hist[i]);
    }
}

}  // namespace mb
