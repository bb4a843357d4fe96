// Tile-staged sigma-clip kernel for sm_100a: the main combine kernel.
//
// Reference behaviour restated (never copied): src/merging.c:423-477 (gather, n == 0 / 1,
// outputs, counters) and src/tools.c:82-118 (median-centred iterative clip).
//
// Data movement.  A block owns tiles of kThreads consecutive voxels.  For every tile the
// N exposure rows (DATA and STAT) are brought into shared memory as [N][kThreads] float32
// by 1-D TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx, one 512-byte copy per
// exposure row, issued by the lanes of warp 0).  Several blocks are resident per SM, so the
// copies of one block land while the others compute; no register is tied up by a load in
// flight.  Rows that are not 16-byte aligned (or the ragged last tile) are filled by plain
// coalesced loads instead.
//
// Arithmetic.  One thread owns one voxel.
//  * exposure order, float64, no FMA: w = (offset + d) * scale and sum(w) -> the first
//    mean is bit-identical to the reference (tools.c:13-17);
//  * the stack is ordered by a register sorting network on 32-bit keys
//    key = (float32(w) with the low 6 mantissa bits replaced by the exposure index);
//    a compare-exchange is one FMNMX plus one integer a+b-min (the keys are distinct bit
//    patterns, so the max is recovered exactly);
//  * the sorted keys go to this thread's shared-memory column; the median, the number of
//    samples inside the clip window and the rejected samples are then found by direct /
//    binary-search reads of that column, never by another pass over the stack;
//  * later means and sigmas are obtained from shifted moments sum(w-p), sum((w-p)^2) by
//    subtracting the (exactly recomputed) rejected samples.
// Decisions are CERTIFIED: the median taken from the keys and the moment-based sigma are
// within rigorous bounds of the reference's values (key truncation 2^-17, reordered
// float64 sums), so a sample is classified from its key only when it is farther from the
// clip bound than the accumulated uncertainty; any voxel with a sample inside that band,
// or with an ill-conditioned variance, is recomputed by `exact_voxel` -- the reference
// algorithm restated literally on per-thread arrays -- so expmap and the per-exposure
// counters are always bit-exact, means/variances within 1e-15 relative of the reference.
#pragma once

#include "combine_kernels.cuh"

namespace mb {

template <int NB>
struct TileArgs {
    const float* data[NB];
    const float* var[NB];   // null rows when var is not propagated
    double scale[NB];
    double offset[NB];
    double scale2[NB];
    unsigned dstride, vstride;   // elements between consecutive DATA / STAT rows when equally spaced, else 0
    // float32 images of the per-exposure constants (warp kernel): key = d * kscale + koffset, STAT * scale2f
    double2 os[NB];              // (offset, scale) side by side: one 128-bit uniform load per exposure
    float kscale[NB];
    float koffset[NB];
    float scale2f[NB];
    double koff;                 // bound on |key-before-truncation - float32(w)| from the offsets, part of the margins
};

template <int NB>
struct TileSmem {
    float d[NB][kThreads];
    float v[NB][kThreads];
    float k[NB][kThreads];
    double scale[NB];
    double offset[NB];
    unsigned long long bar[2];   // mbarriers: [0] DATA tile, [1] STAT tile
    unsigned long long slow_count;
};

// ---- mbarrier / bulk-copy PTX ---------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// compare-exchange on distinct-bit-pattern float keys: min by FMNMX, max as a + b - min
__device__ __forceinline__ void key_cmpxchg(float& a, float& b)
{
    const float lo = fminf(a, b);
    const unsigned hi = __float_as_uint(a) + __float_as_uint(b) - __float_as_uint(lo);
    a = lo;
    b = __uint_as_float(hi);
}

template <int NB>
__device__ __forceinline__ void sort_keys(float (&a)[NB])
{
#define MB_CE(i, j) key_cmpxchg(a[i], a[j]);
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

struct VoxelResult {
    double mean;
    double sigma;
    int kept;
    unsigned long long rejected;   // bit i: exposure i valid but clipped
};

// The reference algorithm restated literally for one voxel (rare fallback path).
template <int NB, bool UNIT>
__device__ __noinline__ VoxelResult exact_voxel(const TileSmem<NB>& S, int tid, int nfiles, int nmax, double lo,
                                                double up, int nstop)
{
    GenericScratch<NB> g;
    int n = 0;
    for (int i = 0; i < nfiles; i++) {
        const float d = S.d[i][tid];
        if (d == d) {
            g.w[n] = UNIT ? (double)d : (S.offset[i] + (double)d) * S.scale[i];
            g.file[n] = (unsigned short)i;
            g.idx[n] = (unsigned short)n;
            n++;
        }
    }
    VoxelResult r;
    r.mean = CUDART_NAN;
    r.sigma = 0.0;
    r.kept = n;
    r.rejected = 0ull;
    if (n == 1) r.mean = g.w[0];
    if (n >= 2) {
        r.kept = clip_generic<NB>(g, n, nmax, lo, up, nstop, 0, r.mean, r.sigma);
        unsigned long long keep = 0ull;
        for (int k = 0; k < r.kept; k++) keep |= 1ull << g.file[g.idx[k]];
        unsigned long long valid = 0ull;
        for (int k = 0; k < n; k++) valid |= 1ull << g.file[k];
        r.rejected = valid & ~keep;
    }
    return r;
}

// number of keys of this thread's sorted column that are <= x (LE) or < x (!LE), all NB slots
template <int NB, bool LE>
__device__ __forceinline__ int count_keys(const float (*col)[kThreads], int tid, float x)
{
    // branch-free binary search over a power-of-two padded range
    constexpr int P = NB <= 4 ? 4 : NB <= 8 ? 8 : NB <= 16 ? 16 : NB <= 32 ? 32 : 64;
    int pos = 0;
#pragma unroll
    for (int step = P >> 1; step >= 1; step >>= 1) {
        const int probe = pos + step - 1;
        const bool in = probe < NB;
        const float kv = col[in ? probe : NB - 1][tid];
        const bool take = in && (LE ? (kv <= x) : (kv < x));
        pos += take ? step : 0;
    }
    // pos = number of leading keys satisfying the predicate, capped at P-1; finish the last slot
    if (pos == P - 1 && pos < NB) {
        const float kv = col[pos][tid];
        pos += (LE ? (kv <= x) : (kv < x)) ? 1 : 0;
    }
    return pos;
}

template <int NB, bool UNIT>
__global__ void __launch_bounds__(kThreads, (NB <= 48 ? 3 : 2))
sigclip_tile_kernel(const TileArgs<NB> S, const ClipArgs A, const int use_tma, unsigned long long* slow_total)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    TileSmem<NB>& sm = *reinterpret_cast<TileSmem<NB>*>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int nfiles = A.nfiles;
    const int nstop = A.nstop < 1 ? 1 : A.nstop;
    const bool want_var = (A.typ_var == 0);

    // one-time setup: rows beyond nfiles stay NaN for ever (invalid samples)
    for (int i = 0; i < NB; i++) {
        sm.d[i][tid] = CUDART_NAN_F;
        sm.v[i][tid] = 0.f;
    }
    if (tid < NB) {
        sm.scale[tid] = S.scale[tid];
        sm.offset[tid] = S.offset[tid];
    }
    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        sm.slow_count = 0ull;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    LaneCounters<NB> valid_cnt, select_cnt;
    const long long ntiles = (A.nvox + kThreads - 1) / kThreads;
    unsigned phase = 0;

    // issue the bulk copies of one full tile (warp 0, one exposure row per lane and round)
    auto issue_tile = [&](long long tile) {
        const long long v0 = tile * kThreads;
        if (tid < 32) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic reads -> async writes
            if (lane == 0) {
                mbar_expect_tx(&sm.bar[0], (unsigned)(nfiles * kThreads * sizeof(float)));
                if (want_var) mbar_expect_tx(&sm.bar[1], (unsigned)(nfiles * kThreads * sizeof(float)));
            }
            __syncwarp();
            for (int i = lane; i < nfiles; i += 32) {
                bulk_g2s(&sm.d[i][0], S.data[i] + v0, kThreads * sizeof(float), &sm.bar[0]);
                if (want_var) bulk_g2s(&sm.v[i][0], S.var[i] + v0, kThreads * sizeof(float), &sm.bar[1]);
            }
        }
    };

    long long tile = blockIdx.x;
    bool tma_pending = false;
    if (tile < ntiles && use_tma && (tile + 1) * (long long)kThreads <= A.nvox) {
        issue_tile(tile);
        tma_pending = true;
    }

    for (; tile < ntiles; tile += gridDim.x) {
        const long long v = tile * kThreads + tid;
        const bool act = v < A.nvox;
        if (tma_pending) {
            mbar_wait(&sm.bar[0], phase);
        } else {
            // plain coalesced fill (misaligned rows or the ragged last tile)
            for (int i = 0; i < nfiles; i++) {
                sm.d[i][tid] = act ? ld_stream(S.data[i] + v) : CUDART_NAN_F;
                if (want_var) sm.v[i][tid] = act ? ld_stream(S.var[i] + v) : 0.f;
            }
        }

        // ---- pass 1: exposure order.  w, exact sum, shifted second moment, keys ----------
        float key[NB];
        unsigned long long inval = 0ull;
        double sum = 0.0, T1 = 0.0, S2 = 0.0, p = 0.0;
        bool have_p = false;
#pragma unroll
        for (int i = 0; i < NB; i++) {
            const float d = sm.d[i][tid];
            const bool ok = (d == d);
            double w;
            if constexpr (UNIT) w = (double)d;
            else w = (S.offset[i] + (double)d) * S.scale[i];
            if (i < 4) {                           // pivot = first valid sample among the first four
                p = (ok && !have_p) ? w : p;
                have_p = have_p || ok;
            }
            if (ok) {
                sum += w;                          // tools.c:13-16, exposure order, no FMA
                const double t = w - p;
                T1 += t;
                S2 = __fma_rn(t, t, S2);
            } else {
                inval |= 1ull << i;
            }
            float kf;
            if constexpr (UNIT) kf = d;
            else kf = __double2float_rn(w);
            const unsigned kb = (__float_as_uint(kf) & ~63u) | (unsigned)i;
            key[i] = fminf(__uint_as_float(kb), __uint_as_float(0x7F7FFFC0u | (unsigned)i));  // NaN -> sentinel
        }
        const int n = NB - __popcll(inval);

        sort_keys<NB>(key);
#pragma unroll
        for (int q = 0; q < NB; q++) sm.k[q][tid] = key[q];

        // ---- clip iterations on the sorted column ------------------------------------------
        double mean = sum / (double)n;             // n == 0 -> NaN (merging.c:439)
        double sigma = 0.0;
        int a = 0, b = n;
        unsigned long long rejected = 0ull;
        bool slow = false;
        if (n >= 2) {
            // shifted moments of the current set: T1 = sum(w - p), S2 = sum((w - p)^2)
            const double S2_0 = S2;
            int it = A.nmax;
            // values beyond float range or an absurd pivot defeat the key precision model
            if (!(S2_0 < 0x1p200)) slow = true;
            // every valid sample equals the pivot: sigma == 0, nothing lies strictly inside
            // the (empty) window, the reference stops at once with all samples kept (tools.c:100)
            while (!slow && S2_0 != 0.0) {
                const int m = b - a;
                const double dm1 = T1 / (double)m;
                double var_k = S2 / (double)m - dm1 * dm1;
                // conditioning: the moment variance must dominate its rounding error
                if (!(var_k >= 0x1p-13 * (S2_0 / (double)m))) { slow = true; break; }
                var_k = var_k > 0.0 ? var_k : 0.0;
                sigma = sqrt(var_k);
                const int ih = a + (m >> 1);
                const int il = ih - 1 + (m & 1);
                const float kh = __uint_as_float(__float_as_uint(sm.k[ih][tid]) & ~63u);
                const float kl = __uint_as_float(__float_as_uint(sm.k[il][tid]) & ~63u);
                const double med = ((double)kh + (double)kl) * 0.5;
                const double amax = fmax(fabs((double)kh), fabs((double)kl));
                const double clo = med - (A.nclip_low * sigma);
                const double cup = med + (A.nclip_up * sigma);
                // total uncertainty of a key against a bound: median from truncated keys,
                // sigma from reordered sums, key truncation of the sample itself
                const double spread = (A.nclip_low + A.nclip_up) * sigma;
                const double T = 0x1p-15 * amax + 0x1p-26 * spread + 0x1p-15 * (fabs(clo) + fabs(cup)) + 0x1p-120;
                if (!(T < 0x1p120)) { slow = true; break; }   // NaN / inf bounds: let the exact path decide
                const int lo_out = count_keys<NB, true>(sm.k, tid, __double2float_rd(clo - T));
                const int lo_in = count_keys<NB, true>(sm.k, tid, __double2float_ru(clo + T));
                const int up_in = count_keys<NB, false>(sm.k, tid, __double2float_rd(cup - T));
                const int up_out = count_keys<NB, false>(sm.k, tid, __double2float_ru(cup + T));
                // clamp to the current window [a, b)
                const int c_lo_out = min(max(lo_out, a), b), c_lo_in = min(max(lo_in, a), b);
                const int c_up_in = min(max(up_in, a), b), c_up_out = min(max(up_out, a), b);
                if (c_lo_out != c_lo_in || c_up_in != c_up_out) { slow = true; break; }   // a sample in the band
                const int na = c_lo_in;                                   // first kept position
                const int nb2 = max(c_up_in, na);                         // one past the last kept
                const int ni = nb2 - na;
                if (ni < nstop || ni == m) break;                         // tools.c:100-103
                if (it <= 0) break;                                       // tools.c:104
                it--;
                // drop the rejected samples [a,na) and [nb2,b): exact w recomputed from the tile
                for (int q = a; q < b; q++) {
                    if (q == na) { q = nb2; if (q >= b) break; }
                    const int i = (int)(__float_as_uint(sm.k[q][tid]) & 63u);
                    const double dd = (double)sm.d[i][tid];
                    const double w = UNIT ? dd : (sm.offset[i] + dd) * sm.scale[i];
                    const double t = w - p;
                    T1 -= t;
                    S2 -= t * t;
                    rejected |= 1ull << i;
                }
                a = na;
                b = nb2;
                mean = p + T1 / (double)ni;
            }
        }
        int kept = b - a;
        if (slow) {
            const VoxelResult r = exact_voxel<NB, UNIT>(sm, tid, nfiles, A.nmax, A.nclip_low, A.nclip_up, nstop);
            mean = r.mean;
            sigma = r.sigma;
            kept = r.kept;
            rejected = r.rejected;
            atomicAdd(&sm.slow_count, 1ull);
        }

        // ---- propagated variance over the kept set, counters, outputs ---------------------
        const unsigned long long dropped = inval | rejected;
        double vsum = 0.0;
        if (want_var) {
            if (tma_pending) mbar_wait(&sm.bar[1], phase);
#pragma unroll
            for (int i = 0; i < NB; i++) {
                const float sv = sm.v[i][tid];
                if (!((dropped >> i) & 1ull)) vsum += (double)sv * S.scale2[i];   // merging.c:431,462
            }
        }
#pragma unroll
        for (int i = 0; i < NB; i++) {
            valid_cnt.add(i, !((inval >> i) & 1ull), lane);
            select_cnt.add(i, !((dropped >> i) & 1ull), lane);
        }
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

        // ---- next tile -------------------------------------------------------------------------
        __syncthreads();                       // every thread is done with this tile's rows
        if (tma_pending) phase ^= 1u;
        const long long next = tile + gridDim.x;
        tma_pending = false;
        if (next < ntiles && use_tma && (next + 1) * (long long)kThreads <= A.nvox) {
            issue_tile(next);
            tma_pending = true;
        }
    }
    valid_cnt.flush(A.valid_cnt, nfiles, lane);
    select_cnt.flush(A.select_cnt, nfiles, lane);
    __syncthreads();
    if (tid == 0 && slow_total != nullptr && sm.slow_count) atomicAdd(slow_total, sm.slow_count);
}

}  // namespace mb
