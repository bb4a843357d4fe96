// Multi-run warp-tiled sigma clip for very deep stacks (128 < N <= 512), float32 samples, no MAD.
//
// The [N][32-voxel tile] key matrix of a warp lives in shared memory as R sorted runs of 64 keys
// per voxel (R = 4: N <= 256, R = 8: N <= 512).  Each run is gathered, taken through pass 1 and
// sorted by the 64-key register network -- one run after the other through the same registers --
// then stored as a sorted column.  The clip iterations work on the R columns together:
//   * counts against a clip bound  = sum of one binary search per column;
//   * rejected samples             = a prefix and a suffix of every column;
//   * median of the current window = k-th smallest key of R sorted windows, found by bisection on
//     the (order-preserving) bit pattern of the keys, 32 steps of R binary searches.
// Decisions are certified exactly as in sigclip_warp.cuh (9-bit exposure index in the key, so the
// key resolution is 2^-14); uncertain voxels go to the literal restatement `exact_voxel_w`.
// Occupancy is low (32-64 KB of keys per warp): this kernel exists for completeness of the
// stack-depth range, it replaces the thread-per-voxel generic kernel (local-memory shell sort).
#pragma once

#include "sigclip_warp.cuh"

namespace mb {

__device__ __forceinline__ unsigned key_sortable(float k)
{
    const unsigned u = __float_as_uint(k);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_unsortable(unsigned s)
{
    return __uint_as_float((s & 0x80000000u) ? (s & 0x7FFFFFFFu) : ~s);
}

__device__ __forceinline__ const float* ld_row(const float* const* rows, int e)
{
    return reinterpret_cast<const float*>(__ldg(reinterpret_cast<const unsigned long long*>(rows) + e));
}

// Row pointers: from the device table (any layout), or base + e * stride when the rows are equally spaced (STRIDED).
// The table lookup puts a dependent L2 round trip in front of every sample load (shared memory leaves the L1 too
// small to keep the table): 64 serialized lookups per run, 0.05 of the HBM peak; the strided form has none.
struct RunRows {
    const float* data0;
    const float* var0;
    unsigned long long dstride, vstride;   // elements between consecutive rows
    unsigned* next_tile;                   // device counter (zeroed before the launch): tiles are handed out dynamically,
                                           // a warp held up by an exact-fallback voxel simply takes fewer of them
    long long* phase_clk;                  // tuning: per-phase clock totals (null in production)
};

// Per-exposure constants in the kernel's parameter space (constant bank): a lookup in the device table costs an L2
// round trip in front of every dependent use (measured: 33 % of the kernel on the scale^2 lookups of the variance sum).
template <int NB>
struct RunConsts {
    double scale[NB], offset[NB], scale2[NB];
};

// warps per block: each warp owns R sorted columns of 8 KB; the block shares one copy of the per-exposure constants
template <int R>
struct RunsCfg {
    static constexpr int NB = R * 64;
    static constexpr int col_bytes = R * 64 * 32 * 4;
    static constexpr int const_bytes = NB * 3 * 8;
    static constexpr int WPB = (227 * 1024 - const_bytes) / col_bytes;
    static constexpr int smem = WPB * col_bytes + const_bytes;
};

template <int R, bool UNIT, bool STRIDED = false>
__global__ void __launch_bounds__(32 * RunsCfg<R>::WPB)
sigclip_runs_kernel(const ClipArgs A, const StackTable* __restrict__ tab, unsigned long long* slow_total, const RunRows RR,
                    const __grid_constant__ RunConsts<R * 64> K)
{
    constexpr int H = 64;
    constexpr int NB = R * H;
    constexpr int W = ExpSet<NB>::W;          // 2 words per run
    constexpr unsigned IDX = 511u;            // 9-bit exposure index in the key
    constexpr double KEPS = 0x1p-12;          // 4x the key resolution 2^-14
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float (*sk)[H][32] = reinterpret_cast<float (*)[H][32]>(smem_raw + (threadIdx.x >> 5) * RunsCfg<R>::col_bytes);   // sk[r][q][lane], this warp's
    // per-exposure constants: shared memory (the constant bank serves a run-dependent index at ~100 cycles per lookup)
    double* const c_scale = reinterpret_cast<double*>(smem_raw + RunsCfg<R>::WPB * RunsCfg<R>::col_bytes);
    double* const c_offset = c_scale + NB;
    double* const c_scale2 = c_offset + NB;
    for (int i = threadIdx.x; i < NB; i += blockDim.x) {
        c_scale[i] = K.scale[i];
        c_offset[i] = K.offset[i];
        c_scale2[i] = K.scale2[i];
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int nfiles = A.nfiles;
    const int nstop = A.nstop < 1 ? 1 : A.nstop;
    const bool want_var = (A.typ_var == 0);
    const float* const* rows = reinterpret_cast<const float* const*>(tab->data);
    const float* const* vrows = reinterpret_cast<const float* const*>(tab->var);
    auto data_row = [&](int e) -> const float* { return STRIDED ? RR.data0 + (size_t)e * RR.dstride : ld_row(rows, e); };
    auto stat_row = [&](int e) -> const float* { return STRIDED ? RR.var0 + (size_t)e * RR.vstride : ld_row(vrows, e); };

    TransposeConsts tc;
    tc.init(lane);
    unsigned cnt_inval[W], cnt_drop[W];
#pragma unroll
    for (int k = 0; k < W; k++) { cnt_inval[k] = 0u; cnt_drop[k] = 0u; }
    unsigned tiles_done = 0, slow_count = 0;

    const long long ntiles = (A.nvox + 31) >> 5;
    long long clk[5] = {0, 0, 0, 0, 0};
    auto grab_tile = [&]() -> long long {
        long long t = 0;
        if (lane == 0) t = (long long)atomicAdd(RR.next_tile, 1u);
        return __shfl_sync(0xffffffffu, t, 0);
    };
    for (;;) {
        const long long tile = grab_tile();
        if (tile >= ntiles) break;
        const long long tc0 = clock64();
        const long long v = (tile << 5) + lane;
        const bool act = v < A.nvox;

        // ---- R runs through the same registers: gather, pass 1, sort, store -------------------
        ExpSet<NB> inval;
        double sum = 0.0, T1 = 0.0, S2 = 0.0, p = 0.0;
        bool have_p = false;
        int nr[R];
#pragma unroll
        for (int s = 0; s < R; s++) nr[s] = 0;
#pragma unroll 1
        for (int r = 0; r < R; r++) {
            const int e0 = r * H;
            float key[H];
#pragma unroll
            for (int i = 0; i < H; i++) {
                key[i] = CUDART_NAN_F;
                if (act && e0 + i < nfiles) asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(key[i]) : "l"(data_row(e0 + i) + v));
            }
            asm volatile("" ::: "memory");
            unsigned inv0 = 0u, inv1 = 0u;
            int nvalid = 0;
#pragma unroll
            for (int i = 0; i < H; i++) {
                const float d = key[i];
                const bool ok = (d == d);
                double w;
                if constexpr (UNIT) w = (double)d;
                else w = (c_offset[e0 + i] + (double)d) * c_scale[e0 + i];
                if (i < 4 && r == 0) {         // pivot: first valid sample among the first four exposures
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
                    if (i < 32) inv0 |= 1u << (i & 31);
                    else inv1 |= 1u << (i & 31);
                }
                float kf;
                if constexpr (UNIT) kf = d;
                else kf = __double2float_rn(w);
                const unsigned kb = (__float_as_uint(kf) & ~IDX) | (unsigned)(e0 + i);
                key[i] = fminf(__uint_as_float(kb), __uint_as_float((0x7F7FFFFFu & ~IDX) | (unsigned)(e0 + i)));
            }
#pragma unroll
            for (int s = 0; s < R; s++) {
                if (s == r) {
                    nr[s] = nvalid;
                    inval.w[2 * s] = inv0;
                    inval.w[2 * s + 1] = inv1;
                }
            }
            sort_keys_mixed<H, 0>(key, 0u, 0u);
#pragma unroll
            for (int q = 0; q < H; q++) sk[r][q][lane] = key[q];
        }
        int n = 0;
#pragma unroll
        for (int s = 0; s < R; s++) n += nr[s];
        const long long tc1 = clock64();

        // ---- clip iterations on the R sorted columns ---------------------------------------------
        double mean = sum / (double)n;
        double sigma = 0.0;
        int a[R], b[R];
#pragma unroll
        for (int s = 0; s < R; s++) { a[s] = 0; b[s] = nr[s]; }
        ExpSet<NB> rejected;
        bool slow = false;
        if (n >= 2) {
            const double S2_0 = S2;
            int it = A.nmax;
            if (!(S2_0 < 0x1p200)) slow = true;
            while (!slow && S2_0 != 0.0) {
                int m = 0;
#pragma unroll
                for (int s = 0; s < R; s++) m += b[s] - a[s];
                const double rm = 1.0 / (double)m;
                const double dm1 = T1 * rm;
                double var_k = __fma_rn(S2, rm, -(dm1 * dm1));
                if (!(var_k >= 0x1p-13 * (S2_0 * rm))) { slow = true; break; }
                sigma = sqrt(var_k);
                // Median of the R windows (tools.c:45-48) by elimination: to skip `need` keys, look at the key t = ceil(need / R)
                // places into every window; the smallest of those has at most R (t - 1) < need keys below it, so it and the
                // t - 1 keys before it in its window can all be skipped.  ~log(need) / log(R / (R - 1)) rounds of R loads.
                int pos[R];
#pragma unroll
                for (int s = 0; s < R; s++) pos[s] = a[s];
                int need = (m & 1) ? (m >> 1) : (m >> 1) - 1;              // ranks below the lower middle
                while (need > 0) {
                    const int t = (need + R - 1) / R;
                    float best = CUDART_INF_F;
                    int bs = 0;
#pragma unroll
                    for (int s = 0; s < R; s++) {
                        const int q = pos[s] + t - 1;
                        const float x = q < b[s] ? sk[s][q][lane] : CUDART_INF_F;
                        if (x < best) { best = x; bs = s; }
                    }
#pragma unroll
                    for (int s = 0; s < R; s++) pos[s] += (s == bs) ? t : 0;
                    need -= t;
                }
                float kl = CUDART_INF_F, kh;
                int ls = 0;
#pragma unroll
                for (int s = 0; s < R; s++) {
                    const float x = pos[s] < b[s] ? sk[s][pos[s]][lane] : CUDART_INF_F;
                    if (x < kl) { kl = x; ls = s; }
                }
                kh = kl;
                if (!(m & 1)) {                                                // upper middle: the next key
                    kh = CUDART_INF_F;
#pragma unroll
                    for (int s = 0; s < R; s++) {
                        const int q = pos[s] + (s == ls ? 1 : 0);
                        const float x = q < b[s] ? sk[s][q][lane] : CUDART_INF_F;
                        kh = fminf(kh, x);
                    }
                }
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
                int na[R], nb2[R];
                bool band = false;
                int ni = 0;
#pragma unroll
                for (int s = 0; s < R; s++) {
                    int l = count_col<H, true>(sk[s], lane, f_lo_in);
                    int u = count_col<H, false>(sk[s], lane, f_up_in);
                    band = band || (l > 0 && !(sk[s][l - 1][lane] <= f_lo_out)) ||
                           (u < H && (sk[s][u < H ? u : H - 1][lane] < f_up_out));
                    l = min(max(l, a[s]), b[s]);
                    u = max(min(max(u, a[s]), b[s]), l);
                    na[s] = l;
                    nb2[s] = u;
                    ni += u - l;
                }
                if (band) { slow = true; break; }
                if (ni < nstop || ni == m) break;
                if (it <= 0) break;
                it--;
#pragma unroll
                for (int s = 0; s < R; s++) {                  // rejected: prefix and suffix of each column
                    for (int q = a[s]; q < b[s]; q++) {
                        if (q == na[s]) { q = nb2[s]; if (q >= b[s]) break; }
                        const int i = (int)(__float_as_uint(sk[s][q][lane]) & IDX);
                        const double dd = (double)__ldg(data_row(i) + v);
                        const double w = UNIT ? dd : (c_offset[i] + dd) * c_scale[i];
                        const double t = w - p;
                        T1 -= t;
                        S2 -= t * t;
                        rejected.set_dyn(i);
                    }
                    a[s] = na[s];
                    b[s] = nb2[s];
                }
                mean = p + T1 / (double)ni;
            }
        }
        int kept = 0;
#pragma unroll
        for (int s = 0; s < R; s++) kept += b[s] - a[s];
        const long long tc2 = clock64();
        if (slow) {
            const VoxelResultN<NB> r = exact_voxel_w<NB, UNIT>(rows, v, tab->scale, tab->offset, nfiles, A.nmax, A.nclip_low,
                                                               A.nclip_up, nstop);
            mean = r.mean;
            sigma = r.sigma;
            kept = r.kept;
            rejected = r.rejected;
            slow_count++;
        }
        __syncwarp();
        const long long tc3 = clock64();

        // ---- propagated variance over the kept set, one run of STAT rows at a time ----------------
        const ExpSet<NB> dropped = inval | rejected;
        double vsum = 0.0;
        if (want_var) {
            const long long vs = act ? v : A.nvox - 1;
#pragma unroll 1
            for (int r = 0; r < R; r++) {
                const int e0 = r * H;
                float sv[H];
#pragma unroll
                for (int i = 0; i < H; i++)
                    asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(sv[i]) : "l"(stat_row(e0 + i < nfiles ? e0 + i : 0) + vs));
                unsigned d0 = 0u, d1 = 0u;
#pragma unroll
                for (int s = 0; s < R; s++) {
                    if (s == r) { d0 = dropped.w[2 * s]; d1 = dropped.w[2 * s + 1]; }
                }
                const unsigned long long bits = ((unsigned long long)d1 << 32) | d0;
#pragma unroll
                for (int i = 0; i < H; i++)
                    if (!((bits >> i) & 1ull)) vsum += UNIT ? (double)sv[i] : (double)sv[i] * c_scale2[e0 + i];
            }
        }

#pragma unroll
        for (int k = 0; k < W; k++) {
            cnt_inval[k] += __popc(tc.apply(inval.w[k]));
            cnt_drop[k] += __popc(tc.apply(dropped.w[k]));
        }
        tiles_done++;
        const long long tc4 = clock64();
        clk[0] += tc1 - tc0; clk[1] += tc2 - tc1; clk[2] += tc3 - tc2; clk[3] += tc4 - tc3; clk[4]++;

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
    if (RR.phase_clk != nullptr && lane == 0) {
#pragma unroll
        for (int k = 0; k < 5; k++) atomicAdd(reinterpret_cast<unsigned long long*>(RR.phase_clk) + k, (unsigned long long)clk[k]);
    }
}

// Median of 129 .. 512 float32 exposures (src/merging.c:157-287, tools.c:39-51): the same sorted runs of 64 on the raw
// sample values (invalid and absent samples become +inf at the top of their run), the two middle values of the union
// found by elimination over the runs.  Bit-identical to the reference; replaces the thread-per-voxel generic kernel
// (local-memory shell sort, 12 GB/s) for deep stacks.
template <int R, bool STRIDED>
__global__ void __launch_bounds__(32 * RunsCfg<R>::WPB)
median_runs_kernel(const ClipArgs A, const StackTable* __restrict__ tab, const RunRows RR)
{
    constexpr int H = 64;
    constexpr int NB = R * H;
    constexpr int W = ExpSet<NB>::W;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float (*sk)[H][32] = reinterpret_cast<float (*)[H][32]>(smem_raw + (threadIdx.x >> 5) * RunsCfg<R>::col_bytes);
    const int lane = threadIdx.x & 31;
    const int nfiles = A.nfiles;
    const float* const* rows = reinterpret_cast<const float* const*>(tab->data);
    auto data_row = [&](int e) -> const float* { return STRIDED ? RR.data0 + (size_t)e * RR.dstride : ld_row(rows, e); };
    TransposeConsts tc;
    tc.init(lane);
    unsigned cnt_inval[W];
#pragma unroll
    for (int k = 0; k < W; k++) cnt_inval[k] = 0u;
    unsigned tiles_done = 0;
    const long long ntiles = (A.nvox + 31) >> 5;
    for (;;) {
        long long tile = 0;
        if (lane == 0) tile = (long long)atomicAdd(RR.next_tile, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= ntiles) break;
        const long long v = (tile << 5) + lane;
        const bool act = v < A.nvox;
        ExpSet<NB> inval;
        int nr[R];
#pragma unroll
        for (int s = 0; s < R; s++) nr[s] = 0;
#pragma unroll 1
        for (int r = 0; r < R; r++) {
            const int e0 = r * H;
            float key[H];
#pragma unroll
            for (int i = 0; i < H; i++) {
                key[i] = CUDART_NAN_F;
                if (act && e0 + i < nfiles) asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(key[i]) : "l"(data_row(e0 + i) + v));
            }
            asm volatile("" ::: "memory");
            unsigned inv0 = 0u, inv1 = 0u;
            int nvalid = 0;
#pragma unroll
            for (int i = 0; i < H; i++) {
                const bool ok = (key[i] == key[i]);
                if (ok) nvalid++;
                else {
                    key[i] = CUDART_INF_F;
                    if (i < 32) inv0 |= 1u << (i & 31);
                    else inv1 |= 1u << (i & 31);
                }
            }
#pragma unroll
            for (int s = 0; s < R; s++) {
                if (s == r) {
                    nr[s] = nvalid;
                    inval.w[2 * s] = inv0;
                    inval.w[2 * s + 1] = inv1;
                }
            }
            sort_keys_mixed<H, 0>(key, 0u, 0u);
#pragma unroll
            for (int q = 0; q < H; q++) sk[r][q][lane] = key[q];
        }
        int n = 0;
#pragma unroll
        for (int s = 0; s < R; s++) n += nr[s];
        // the two middle values of the union of the runs' valid prefixes, by elimination (see sigclip_runs_kernel)
        int pos[R];
#pragma unroll
        for (int s = 0; s < R; s++) pos[s] = 0;
        int need = (n & 1) ? (n >> 1) : (n >> 1) - 1;
        while (need > 0) {
            const int t = (need + R - 1) / R;
            float best = CUDART_INF_F;
            int bs = -1;
#pragma unroll
            for (int s = 0; s < R; s++) {
                const int q = pos[s] + t - 1;
                if (q < nr[s]) {
                    const float x = sk[s][q][lane];
                    if (bs < 0 || x < best) { best = x; bs = s; }      // (a valid sample may itself be +inf)
                }
            }
#pragma unroll
            for (int s = 0; s < R; s++) pos[s] += (s == bs) ? t : 0;
            need -= t;
        }
        float lo = 0.f, hi = 0.f;
        int ls = -1;
#pragma unroll
        for (int s = 0; s < R; s++) {
            if (pos[s] < nr[s]) {
                const float x = sk[s][pos[s]][lane];
                if (ls < 0 || x < lo) { lo = x; ls = s; }
            }
        }
        hi = lo;
        if (!(n & 1)) {
            int hs = -1;
#pragma unroll
            for (int s = 0; s < R; s++) {
                const int q = pos[s] + (s == ls ? 1 : 0);
                if (q < nr[s]) {
                    const float x = sk[s][q][lane];
                    if (hs < 0 || x < hi) { hi = x; hs = s; }
                }
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < W; k++) cnt_inval[k] += __popc(tc.apply(inval.w[k]));
        tiles_done++;
        if (act) {
            double med = (n & 1) ? (double)lo : ((double)lo + (double)hi) * 0.5;      // tools.c:45-48
            if (n == 0) med = CUDART_NAN;
            A.out[v] = med;
            A.expmap[v] = n;
        }
    }
    const unsigned seen = tiles_done * 32u;
#pragma unroll
    for (int k = 0; k < W; k++) {
        const int e = 32 * k + lane;
        if (e < nfiles) atomicAdd(A.valid_cnt + e, (unsigned long long)(seen - cnt_inval[k]));
    }
}

}  // namespace mb
