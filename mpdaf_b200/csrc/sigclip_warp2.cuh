// Two-run warp-tiled sigma clip for deep stacks (64 < N <= 128), float32 samples, no MAD.
//
// Same algorithm, same certified decisions and the same exact fallback as sigclip_warp.cuh; what
// changes is how the stack is ordered.  A single register network over 96 or 128 keys needs 170-255
// registers (2-3 blocks per SM, spills) and 1007 / 1471 compare-exchanges.  Here the exposures
// are split into two runs of H = NB/2 (48 or 64): each run is gathered, taken through pass 1 and
// sorted by the H-network in the SAME registers, one after the other (a rolled loop: the code of
// pass 1 and of the network is fetched once and executed twice), and written to its own sorted
// shared-memory column.  Everything downstream works on the two sorted columns together:
//   * counts below / above a clip bound   = sum of one binary search per column;
//   * rejected samples                    = a prefix and a suffix of each column;
//   * the median of the current window    = k-th smallest of two sorted arrays (partition search).
// 2 x 367 = 734 compare-exchanges instead of 1007 for N = 96, ~85 registers, 4 blocks per SM.
#pragma once

#include "sigclip_warp.cuh"

namespace mb {

template <int H>
struct Warp2Smem {
    float k[2][H][32];   // sorted keys of run 0 / run 1, one column per lane
};

template <int H>
struct Block2Smem {
    Warp2Smem<H> w[kWT / 32];
    double scale[2 * H];
    double offset[2 * H];
    double scale2[2 * H];
    double rcp[2 * H + 1];
    const float* data[2 * H];
    const float* var[2 * H];
};

// k-th smallest (0-based) key of the union of two sorted windows A[a0,b0) and B[a1,b1) of the
// lane's columns, together with its predecessor (the (k-1)-th, or the same when k == 0).
template <int H>
__device__ __forceinline__ void kth_of_two(const float (*A)[32], const float (*B)[32], int lane, int a0, int b0, int a1,
                                           int b1, int k, float& kth, float& pred)
{
    const int la = b0 - a0, lb = b1 - a1;
    // i = number of elements taken from A among the first k: smallest i with A[i] >= B[k - i - 1]
    int lo = max(0, k - lb), hi = min(k, la);
    while (lo < hi) {
        const int i = (lo + hi) >> 1;
        // take more from A if A[i] < B[k-i-1]
        if (A[a0 + i][lane] < B[a1 + k - i - 1][lane]) lo = i + 1;
        else hi = i;
    }
    const int i = lo, j = k - lo;
    const float ai = i < la ? A[a0 + i][lane] : CUDART_INF_F;
    const float bj = j < lb ? B[a1 + j][lane] : CUDART_INF_F;
    kth = fminf(ai, bj);
    const float ap = i > 0 ? A[a0 + i - 1][lane] : -CUDART_INF_F;
    const float bp = j > 0 ? B[a1 + j - 1][lane] : -CUDART_INF_F;
    pred = (i > 0 || j > 0) ? fmaxf(ap, bp) : kth;
}

template <int H, bool UNIT>
__global__ void __launch_bounds__(kWT, (H <= 48 ? 4 : 3) * 128 / kWT)
sigclip_warp2_kernel(const ClipArgs A, const StackTable* __restrict__ tab, const int flags, unsigned long long* slow_total)
{
    constexpr int NB = 2 * H;
    constexpr int W = ExpSet<NB>::W;
    constexpr unsigned IDX = 127u;          // 7-bit exposure index in the key
    constexpr double KEPS = 0x1p-14;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Block2Smem<H>& sm = *reinterpret_cast<Block2Smem<H>*>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int wib = tid >> 5;
    const int nfiles = A.nfiles;
    const int nstop = A.nstop < 1 ? 1 : A.nstop;
    const bool want_var = (A.typ_var == 0);

    for (int i = tid; i < NB; i += kWT) {
        const bool in = i < nfiles;
        sm.scale[i] = in ? tab->scale[i] : 1.0;
        sm.offset[i] = in ? tab->offset[i] : 0.0;
        sm.scale2[i] = in ? tab->scale2[i] : 1.0;
        sm.data[i] = (const float*)tab->data[in ? i : 0];
        sm.var[i] = (const float*)tab->var[in ? i : 0];
    }
    for (int i = tid; i <= NB; i += kWT) sm.rcp[i] = i > 0 ? 1.0 / (double)i : 0.0;
    __syncthreads();

    float (*sk0)[32] = sm.w[wib].k[0];
    float (*sk1)[32] = sm.w[wib].k[1];
    TransposeConsts tc;
    tc.init(lane);
    unsigned cnt_inval[W], cnt_drop[W];
#pragma unroll
    for (int k = 0; k < W; k++) { cnt_inval[k] = 0u; cnt_drop[k] = 0u; }
    unsigned tiles_done = 0, slow_count = 0;

    const long long ntiles = (A.nvox + 31) >> 5;
    const long long wstride = (long long)gridDim.x * (kWT / 32);
    auto prefetch_rows = [&](const float* const* rows, long long t) {
        const long long v0 = t << 5;
        if (v0 + 31 >= A.nvox) return;
        for (int i = lane; i < nfiles; i += 32) asm volatile("prefetch.global.L1 [%0];" ::"l"(rows[i] + v0));
    };
    long long tile0 = (long long)blockIdx.x * (kWT / 32);
    if (tile0 + wib < ntiles) prefetch_rows(sm.data, tile0 + wib);

    for (; tile0 < ntiles; tile0 += wstride) {
        if (flags & 1) __syncthreads();
        const long long tile = tile0 + wib;
        const long long v = (tile << 5) + lane;
        const bool act = tile < ntiles && v < A.nvox;
        if (tile + wstride < ntiles) prefetch_rows(sm.data, tile + wstride);

        // ---- two runs through the same registers: gather, pass 1, sort, store ------------------
        ExpSet<NB> inval;
        double sum = 0.0, T1 = 0.0, S2 = 0.0, p = 0.0;
        bool have_p = false;
        int n0 = 0, n1 = 0;
#pragma unroll 1
        for (int r = 0; r < 2; r++) {
            const int e0 = r * H;
            float key[H];
#pragma unroll
            for (int i = 0; i < H; i++) {
                key[i] = CUDART_NAN_F;
                if (act && e0 + i < nfiles) asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(key[i]) : "l"(sm.data[e0 + i] + v));
            }
            asm volatile("" ::: "memory");
            unsigned invw[(H + 31) / 32];
#pragma unroll
            for (int k = 0; k < (H + 31) / 32; k++) invw[k] = 0u;
            int nvalid = 0;
#pragma unroll
            for (int i = 0; i < H; i++) {
                const float d = key[i];
                const bool ok = (d == d);
                double w;
                if constexpr (UNIT) w = (double)d;
                else w = (sm.offset[e0 + i] + (double)d) * sm.scale[e0 + i];
                if (i < 4 && r == 0) {         // pivot: first valid sample among the first four of run 0
                    p = (ok && !have_p) ? w : p;
                    have_p = have_p || ok;
                }
                if (ok) {
                    sum += w;                  // tools.c:13-16: exposure order continues across the runs
                    const double t = w - p;
                    T1 += t;
                    S2 = __fma_rn(t, t, S2);
                    nvalid++;
                } else {
                    invw[i >> 5] |= 1u << (i & 31);
                }
                float kf;
                if constexpr (UNIT) kf = d;
                else kf = __double2float_rn(w);
                const unsigned kb = (__float_as_uint(kf) & ~IDX) | (unsigned)(e0 + i);
                key[i] = fminf(__uint_as_float(kb), __uint_as_float((0x7F7FFFFFu & ~IDX) | (unsigned)(e0 + i)));
            }
            // this run's count and invalid bits into the stack-wide set (runs start at bit 0 of a word,
            // or at bit 16 of word 1 for the second run of 48)
            if (r == 0) {
                n0 = nvalid;
                inval.w[0] = invw[0];
                inval.w[1] = invw[1];
            } else {
                n1 = nvalid;
                if constexpr (H == 64) {
                    inval.w[2] = invw[0];
                    inval.w[3] = invw[1];
                } else {
                    inval.w[1] |= invw[0] << 16;
                    inval.w[2] = (invw[0] >> 16) | (invw[1] << 16);
                }
            }
            sort_keys_mixed<H, 0>(key, 0u, 0u);
            float (*sk)[32] = r == 0 ? sk0 : sk1;
#pragma unroll
            for (int q = 0; q < H; q++) sk[q][lane] = key[q];
        }
        const int n = n0 + n1;
        if (want_var) prefetch_rows(sm.var, tile);

        // ---- clip iterations on the two sorted columns -----------------------------------------
        double mean = sum / (double)n;
        double sigma = 0.0;
        int a0 = 0, b0 = n0, a1 = 0, b1 = n1;
        ExpSet<NB> rejected;
        bool slow = false;
        if (n >= 2) {
            const double S2_0 = S2;
            int it = A.nmax;
            if (!(S2_0 < 0x1p200)) slow = true;
            while (!slow && S2_0 != 0.0) {
                const int m = (b0 - a0) + (b1 - a1);
                const double rm = sm.rcp[m];
                const double dm1 = T1 * rm;
                double var_k = __fma_rn(S2, rm, -(dm1 * dm1));
                if (!(var_k >= 0x1p-13 * (S2_0 * rm))) { slow = true; break; }
                sigma = sqrt(var_k);
                float kh, kl;
                kth_of_two<H>(sk0, sk1, lane, a0, b0, a1, b1, m >> 1, kh, kl);
                if (m & 1) kl = kh;
                kh = __uint_as_float(__float_as_uint(kh) & ~IDX);
                kl = __uint_as_float(__float_as_uint(kl) & ~IDX);
                const double med = ((double)kh + (double)kl) * 0.5;
                const double amax = fmax(fabs((double)kh), fabs((double)kl));
                const double clo = med - (A.nclip_low * sigma);
                const double cup = med + (A.nclip_up * sigma);
                const double spread = (A.nclip_low + A.nclip_up) * sigma;
                const double T = KEPS * amax + 0x1p-26 * spread + KEPS * (fabs(clo) + fabs(cup)) + 0x1p-120;
                if (!(T < 0x1p120)) { slow = true; break; }
                const float f_lo_in = __double2float_ru(clo + T), f_lo_out = __double2float_rd(clo - T);
                const float f_up_in = __double2float_rd(cup - T), f_up_out = __double2float_ru(cup + T);
                int lo0 = count_col<H, true>(sk0, lane, f_lo_in), lo1 = count_col<H, true>(sk1, lane, f_lo_in);
                int up0 = count_col<H, false>(sk0, lane, f_up_in), up1 = count_col<H, false>(sk1, lane, f_up_in);
                // band checks on both columns; any key inside a band defers to the exact path
                const bool band = (lo0 > 0 && !(sk0[lo0 - 1][lane] <= f_lo_out)) || (lo1 > 0 && !(sk1[lo1 - 1][lane] <= f_lo_out)) ||
                                  (up0 < H && (sk0[up0 < H ? up0 : H - 1][lane] < f_up_out)) ||
                                  (up1 < H && (sk1[up1 < H ? up1 : H - 1][lane] < f_up_out));
                if (band) { slow = true; break; }
                lo0 = min(max(lo0, a0), b0); up0 = max(min(max(up0, a0), b0), lo0);
                lo1 = min(max(lo1, a1), b1); up1 = max(min(max(up1, a1), b1), lo1);
                const int ni = (up0 - lo0) + (up1 - lo1);
                if (ni < nstop || ni == m) break;
                if (it <= 0) break;
                it--;
                // rejected samples: prefix and suffix of each column
#pragma unroll 1
                for (int r = 0; r < 2; r++) {
                    float (*sk)[32] = r == 0 ? sk0 : sk1;
                    const int a = r == 0 ? a0 : a1, b = r == 0 ? b0 : b1, na = r == 0 ? lo0 : lo1, nb2 = r == 0 ? up0 : up1;
                    for (int q = a; q < b; q++) {
                        if (q == na) { q = nb2; if (q >= b) break; }
                        const int i = (int)(__float_as_uint(sk[q][lane]) & IDX);
                        const double dd = (double)__ldg(sm.data[i] + v);
                        const double w = UNIT ? dd : (sm.offset[i] + dd) * sm.scale[i];
                        const double t = w - p;
                        T1 -= t;
                        S2 -= t * t;
                        rejected.set_dyn(i);
                    }
                }
                a0 = lo0; b0 = up0; a1 = lo1; b1 = up1;
                mean = p + T1 * sm.rcp[ni];
            }
        }
        int kept = (b0 - a0) + (b1 - a1);
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

        // ---- propagated variance over the kept set (one run of STAT rows at a time) ---------------
        const ExpSet<NB> dropped = inval | rejected;
        double vsum = 0.0;
        if (want_var) {
            const long long vs = act ? v : A.nvox - 1;
#pragma unroll 1
            for (int r = 0; r < 2; r++) {
                const int e0 = r * H;
                float sv[H];
#pragma unroll
                for (int i = 0; i < H; i++)
                    asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(sv[i]) : "l"(sm.var[e0 + i] + vs));
                // dropped bits of this run, aligned to bit 0
                unsigned long long bits = 0ull;
                if (H == 64) bits = ((unsigned long long)dropped.w[2 * r + 1] << 32) | dropped.w[2 * r];
                else bits = r == 0 ? (((unsigned long long)(dropped.w[1] & 0xFFFFu) << 32) | dropped.w[0])
                                   : (((unsigned long long)dropped.w[2] << 16) | (dropped.w[1] >> 16));
#pragma unroll
                for (int i = 0; i < H; i++)
                    if (!((bits >> i) & 1ull)) vsum += (double)sv[i] * sm.scale2[e0 + i];
            }
        }

#pragma unroll
        for (int k = 0; k < W; k++) {
            cnt_inval[k] += __popc(tc.apply(inval.w[k]));
            cnt_drop[k] += __popc(tc.apply(dropped.w[k]));
        }
        tiles_done++;

        if (act) {
            double ov;
            if (want_var) {
                ov = vsum / ((double)kept * (double)kept);
            } else if (kept > 1) {
                ov = sigma * sigma;
                if (A.typ_var == 1) ov /= (double)(kept - 1);
            } else {
                ov = CUDART_NAN;
            }
            A.out[v] = mean;
            A.outvar[v] = ov;
            A.expmap[v] = kept;
        }
    }

    const unsigned seen = tiles_done * 32u;
#pragma unroll
    for (int k = 0; k < W; k++) {
        const int e = 32 * k + lane;
        if (e < nfiles) {
            atomicAdd(A.valid_cnt + e, (unsigned long long)(seen - cnt_inval[k]));
            atomicAdd(A.select_cnt + e, (unsigned long long)(seen - cnt_drop[k]));
        }
    }
    if (slow_total != nullptr && slow_count) atomicAdd(slow_total, (unsigned long long)slow_count);
}

}  // namespace mb
