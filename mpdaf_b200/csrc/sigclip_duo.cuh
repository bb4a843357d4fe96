// Pair-packed sigma-clip kernel for 65 .. 128 float32 exposures on sm_100a: the pair kernel (sigclip_pair.cuh) with the
// stack of a voxel pair split between TWO lanes.
//
// A 96-deep stack does not fit the pair kernel's formulation twice over: 96 packed sorted keys plus 96 unsorted ones
// exceed the register file, and a [96][64] stage plus a 96-row sorted column leave six warps per SM.  Here a tile is 32
// voxels; lanes l and l + 16 share the voxel pair (2 (l & 15), 2 (l & 15) + 1): the lower lane owns exposure slots
// 0 .. HB - 1, the upper one HB .. 2 HB - 1 (HB = 48 or 64).  Per lane everything is the pair kernel at depth HB -- the
// same registers (so the same twelve / eight warps per SM), a [2 HB][32] stage of the same 12 / 16 KB, its own sorted
// column of HB packed keys -- and per tile the warp processes the same number of samples.  What changes:
//  * the float64 moments, valid counts, rejected counts and the propagated variance of a voxel are sums over the two
//    lanes (one butterfly shuffle each);
//  * the clip step of a voxel works on TWO sorted columns (its lane's and the partner's, both in shared memory): the
//    middle of their union by elimination, the five keys next to either end of the window by a two-way merge.  The
//    lower lane evaluates voxel 0, the upper lane voxel 1, and they exchange the outcome (flags, bounds, boundary
//    keys, expected candidate counts per column);
//  * the comparison pass and the exact classification of the candidates run per lane on its own exposures, exactly as in
//    the pair kernel; a rejected candidate of a lane's exposures leaves that lane's column.
// Reference behaviour restated (never copied): src/merging.c:423-477, src/tools.c:82-118.  Decisions are certified as in
// sigclip_pair.cuh (same key error model, same undecidable band, same queue for the exact recomputation).
#pragma once

#include "sigclip_pair.cuh"

namespace mb {

template <int HB>
struct DuoCfg {
    static constexpr int NB = 2 * HB;
    static constexpr int stage_bytes = NB * 32 * 4;   // one [NB][32] float32 tile
    static constexpr int col_bytes = HB * 32 * 4;     // sorted packed keys: one column per lane
    static constexpr int misc_bytes = 8192;           // per-exposure constants, reciprocals, histogram, barriers
    static constexpr int budget = 227 * 1024 - misc_bytes;
    static constexpr int c_regs = 2048 / (2 * HB + 72);
    static constexpr int c_smem = budget / (stage_bytes + col_bytes);
    static constexpr int c_cap = c_regs < 16 ? c_regs : 16;
    static constexpr int c_fit = c_cap < c_smem ? c_cap : c_smem;
    static constexpr int C = (c_fit > 4 && c_fit % 4 == 1) ? c_fit - 1 : c_fit;
    static constexpr int SPW = (budget - C * col_bytes) / (C * stage_bytes) >= 2 ? 2 : 1;
    static constexpr int ST = SPW * C;
    static constexpr int threads = C * 32;
    static_assert(C >= 2, "stack too deep for the duo kernel");
};

template <int HB>
struct DuoSmem {
    float stage[DuoCfg<HB>::ST][2 * HB][32];
    unsigned col[DuoCfg<HB>::C][HB][32];
    double rcp[2 * HB + 1];
    double2 os[2 * HB];            // (offset, scale)
    double osp[2 * HB];            // offset * scale
    float2 kso[2 * HB];            // float32 (scale, offset * scale)
    float scale2f[2 * HB];
    unsigned long long full[DuoCfg<HB>::ST];
    unsigned hist[2 * HB + 1];
};
static_assert(sizeof(DuoSmem<48>) <= 227 * 1024 && sizeof(DuoSmem<64>) <= 227 * 1024, "shared memory");

// (queued voxels: sigclip_pair_slow_kernel<2 * HB, UNIT>, the warp-per-voxel restatement, handles up to 128 samples)

template <int HB, bool UNIT, bool PROP, bool MAD = false>
__global__ void __launch_bounds__(DuoCfg<HB>::threads, 1)
sigclip_duo_kernel(const PairArgs<2 * HB> S, const ClipArgs A, const long long ntiles, unsigned* __restrict__ slow_q,
                   unsigned* __restrict__ slow_n, const unsigned slow_cap, const int flags, const __grid_constant__ TmaMaps M)
{
    using Cfg = DuoCfg<HB>;
    constexpr int NB = 2 * HB;
    constexpr int C = Cfg::C, ST = Cfg::ST;
    constexpr int W = ExpSet<HB>::W;
    constexpr unsigned FULLM = 0xffffffffu;
    constexpr unsigned kTileBytes = NB * 32 * sizeof(float);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    DuoSmem<HB>& sm = *reinterpret_cast<DuoSmem<HB>*>(smem_raw);
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int half = lane >> 4;          // which half of the exposure slots this lane owns; also the voxel it evaluates
    const int pr = lane & 15;            // voxel pair of the tile
    const int e_base = half * HB;
    const int nstop = A.nstop < 1 ? 1 : A.nstop;

    for (int i = tid; i <= NB; i += Cfg::threads) sm.rcp[i] = i > 0 ? 1.0 / (double)i : 0.0;
    for (int i = tid; i < NB; i += Cfg::threads) {
        sm.os[i] = S.os[i];
        sm.osp[i] = S.osp[i];
        sm.kso[i] = make_float2(S.kscale[i], S.koffset[i]);
        sm.scale2f[i] = S.scale2f[i];
    }
    for (int i = tid; i <= NB; i += Cfg::threads) sm.hist[i] = 0u;
    if (tid == 0) {
        for (int s = 0; s < ST; s++) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int c = warp;
    TransposeConsts tc;
    tc.init(lane);
    unsigned cnt_inval[2][W], cnt_drop[2][W];      // [half of the slots][mask word]: lane = bit = exposure slot
#pragma unroll
    for (int g = 0; g < 2; g++) {
#pragma unroll
        for (int k = 0; k < W; k++) { cnt_inval[g][k] = 0u; cnt_drop[g][k] = 0u; }
    }
    unsigned tiles_done = 0;
    // sorted columns: this lane's and the partner's; key q of voxel h at col[q * 64 + h] (16-bit units)
    const unsigned short* col_own = reinterpret_cast<const unsigned short*>(&sm.col[c][0][lane]);
    const unsigned short* col_par = reinterpret_cast<const unsigned short*>(&sm.col[c][0][lane ^ 16]);
    const double2* os_l = sm.os + e_base;
    const double* osp_l = sm.osp + e_base;
    const float2* kso_l = sm.kso + e_base;
    const float* s2f_l = sm.scale2f + e_base;
    const float nlo32 = (float)A.nclip_low, nup32 = (float)A.nclip_up;
    const float krel32 = (float)kPairKrel, c1f = (float)(kPairC1 * (1.0 - 0x1p-20)), c2f = (float)(kPairC2 * (1.0 + 0x1p-20));
    const float e0c = (float)S.osmax;

    const long long my_tiles = ntiles > (long long)blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const long long rounds = (my_tiles + C - 1) / C;
    auto request_tile = [&](long long j) {      // lane 0 only
        const long long kk = j * C + c;
        if (kk >= my_tiles) return;
        const unsigned st = (unsigned)(c * Cfg::SPW + (int)(j % Cfg::SPW));
        mbar_expect_tx(&sm.full[st], kTileBytes);
        tma_load_2d(&sm.stage[st][0][0], &M.data, (int)(((long long)blockIdx.x + kk * gridDim.x) << 5), 0, &sm.full[st]);
    };
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < Cfg::SPW; j++) request_tile(j);
    }
    for (long long r = 0; r < rounds; r++) {
        if ((r & ((flags >> 8) & 15)) == 0 && (flags & 2)) {
            constexpr int G0 = (C + 1) / 2, G1 = C - G0;
            if (c < G0) asm volatile("bar.sync 1, %0;" ::"n"(G0 * 32) : "memory");
            else asm volatile("bar.sync 2, %0;" ::"n"((G1 > 0 ? G1 : 1) * 32) : "memory");
        }
        const long long kk64 = r * C + c;
        if (kk64 >= my_tiles) continue;
        const long long tile = (long long)blockIdx.x + kk64 * gridDim.x;
        const unsigned st = (unsigned)(c * Cfg::SPW + (int)(r % Cfg::SPW)), ph = (unsigned)(r / Cfg::SPW) & 1u;
        float (*dt)[32] = sm.stage[st];
        const long long v0 = (tile << 5) + 2 * pr;                 // the pair's voxels: v0, v0 + 1
        if (PROP && lane == 0) tma_prefetch_2d(&M.stat, (int)(tile << 5), 0);
        const bool pair_ok = v0 + 1 < A.nvox;
        mbar_wait(&sm.full[st], ph);

        // ---- pivot (from the first slots: the lower lane computes, the upper one receives) ------------------------
        float pf[2];
        {
            const float2 r0 = *reinterpret_cast<const float2*>(&dt[0][2 * pr]);
            const float2 r1 = *reinterpret_cast<const float2*>(&dt[1][2 * pr]);
            const float2 r2 = *reinterpret_cast<const float2*>(&dt[2][2 * pr]);
            const float2 r3 = *reinterpret_cast<const float2*>(&dt[3][2 * pr]);
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float x0 = h ? r0.y : r0.x, x1 = h ? r1.y : r1.x, x2 = h ? r2.y : r2.x, x3 = h ? r3.y : r3.x;
                if constexpr (!UNIT) {
                    x0 = __fmaf_rn(x0, S.kscale[0], S.koffset[0]);
                    x1 = __fmaf_rn(x1, S.kscale[1], S.koffset[1]);
                    x2 = __fmaf_rn(x2, S.kscale[2], S.koffset[2]);
                    x3 = __fmaf_rn(x3, S.kscale[3], S.koffset[3]);
                }
                float m3 = fmaxf(fminf(x0, x1), fminf(fmaxf(x0, x1), x2));
                m3 = (m3 == m3) ? m3 : x3;
                m3 = (m3 == m3 && fabsf(m3) < 0x1p100f) ? m3 : 0.f;
                pf[h] = m3;                                          // both lanes read the same rows: identical values
            }
        }
        const double p[2] = {(double)pf[0], (double)pf[1]};

        // ---- pass 1 over this lane's HB slots (see sigclip_pair.cuh) ------------------------------------------------
        unsigned key[HB], ukey[HB];
        ExpSet<HB> inval[2];
        double T1[2] = {0.0, 0.0}, S2[2] = {0.0, 0.0};
#pragma unroll
        for (int i = 0; i < HB; i++) {
            const float2 dd = *reinterpret_cast<const float2*>(&dt[e_base + i][2 * pr]);
            double t[2];
            float tf[2];
            double2 osv = make_double2(0.0, 1.0);
            double ospv = 0.0;
            float2 ksv = make_float2(1.f, 0.f);
            if constexpr (!UNIT) {
                osv = os_l[i];
                ospv = osp_l[i];
                ksv = kso_l[i];
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float d = h ? dd.y : dd.x;
                if constexpr (UNIT) {
                    t[h] = (double)d - p[h];
                    tf[h] = d - pf[h];
                } else {
                    t[h] = __fma_rn((double)d, osv.y, ospv - p[h]);
                    tf[h] = __fmaf_rn(d, ksv.x, ksv.y) - pf[h];
                }
            }
            asm("{\n\t.reg .pred q0, q1;\n\t.reg .b32 lo0, hi0, lo1, hi1;\n\t"
                "cvt.rn.bf16x2.f32 %0, %7, %6;\n\t"
                "setp.nan.bf16x2 q0|q1, %0, %0;\n\t"
                "max.bf16x2 %1, %0, %8;\n\t"
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
#pragma unroll
            for (int h = 0; h < 2; h++) {
                T1[h] += t[h];
                S2[h] = __fma_rn(t[h], t[h], S2[h]);
            }
        }
        int n_own[2], n[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            n_own[h] = HB - inval[h].count();
            n[h] = n_own[h] + __shfl_xor_sync(FULLM, n_own[h], 16);
            T1[h] += __shfl_xor_sync(FULLM, T1[h], 16);          // (a + b == b + a: both lanes hold the same sums)
            S2[h] += __shfl_xor_sync(FULLM, S2[h], 16);
            if (fabs(T1[h]) < 0x1p-1000) T1[h] = 0.0;
        }

        sort_pairs<HB>(key);
#pragma unroll
        for (int q = 0; q < HB; q++) sm.col[c][q][lane] = key[q];
        __syncwarp();

        // ---- clip iterations --------------------------------------------------------------------------------------------
        int a[2], b[2], mt[2], itl[2];      // a, b: the window of THIS lane's column; mt: samples of the voxel's window (both columns)
        bool done[2], slow[2];
        double varl[2] = {0.0, 0.0}, S2_0[2];
        float e0[2];
        ExpSet<HB> rej[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            a[h] = HB - n_own[h];
            b[h] = HB;
            mt[h] = n[h];
            itl[h] = A.nmax;
            S2_0[h] = S2[h];
            slow[h] = false;
            done[h] = (n[h] < 2) || (S2[h] == 0.0);
            if (n[h] >= 2 && !(S2[h] < 0x1p200)) { slow[h] = true; done[h] = true; }
            e0[h] = 0x1p-21f * (fabsf(pf[h]) + e0c) + 0x1p-100f;
        }
        auto final_stats = [&](int h) {
            if constexpr (!PROP) {
                const double rm = sm.rcp[mt[h]];
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
            // ---- part A: this lane evaluates voxel `half` over the two columns ----------------------------------------
            // window of my column for the voxel the partner evaluates -> partner; partner's window for my voxel <- partner
            const int send_ab = half ? (a[0] | (b[0] << 8)) : (a[1] | (b[1] << 8));
            const int recv_ab = __shfl_xor_sync(FULLM, send_ab, 16);
            const int ao = half ? a[1] : a[0], bo = half ? b[1] : b[0];
            const int ap = recv_ab & 0xFF, bp = (recv_ab >> 8) & 0xFF;
            const bool act = half ? !done[1] : !done[0];
            const double T1m = half ? T1[1] : T1[0], S2m = half ? S2[1] : S2[0], S20m = half ? S2_0[1] : S2_0[0];
            const float e0m = half ? e0[1] : e0[0];
            const int m = (bo - ao) + (bp - ap);
            const unsigned short* cA = col_own + half;       // key q of my voxel in my column: cA[q * 64]
            const unsigned short* cB = col_par + half;
            auto fkey = [](unsigned u) -> float { return __uint_as_float(u << 16); };
            bool go_slow = false, cand = false;
            int jl_o = 0, jl_p = 0, ju_o = 0, ju_p = 0;
            float clo = 0.f, cup = 0.f, Tb = 0.f;
            unsigned BL = 0xFF80u, BU = 0x7F80u;
            double var_k = 0.0;
            if (act) {
                const double rm = sm.rcp[m];
                const double dm1 = T1m * rm;
                var_k = __fma_rn(S2m, rm, -(dm1 * dm1));
                // (MAD: the moments only give the mean; sigma comes from the keys, its uncertainty is part of Tb below)
                const bool bad = !MAD && (!(var_k >= 0x1p-13 * (S20m * rm)) || !(var_k < 0x1p100) || !(var_k > 0x1p-100));
                // the two middle keys of the union of the two windows, by elimination (sigclip_runs.cuh)
                int pa = ao, pb = ap;
                int need = (m - 1) >> 1;                             // ranks below the lower middle
                while (need > 0) {
                    const int t = (need + 1) >> 1;
                    const float xa = pa + t - 1 < bo ? fkey(cA[(pa + t - 1) * 64]) : CUDART_INF_F;
                    const float xb = pb + t - 1 < bp ? fkey(cB[(pb + t - 1) * 64]) : CUDART_INF_F;
                    if (xa <= xb) pa += t; else pb += t;
                    need -= t;
                }
                float ha = pa < bo ? fkey(cA[pa * 64]) : CUDART_INF_F;
                float hb = pb < bp ? fkey(cB[pb * 64]) : CUDART_INF_F;
                const float kl = fminf(ha, hb);
                float kh = kl;
                // (MAD) the lower middle and everything below it: ranks up to la in my column, up to lb in the partner's
                const int la0 = (ha <= hb) ? pa : pa - 1, lb0 = (ha <= hb) ? pb - 1 : pb;
                if (!(m & 1)) {
                    if (ha <= hb) { pa++; ha = pa < bo ? fkey(cA[pa * 64]) : CUDART_INF_F; }
                    else { pb++; hb = pb < bp ? fkey(cB[pb * 64]) : CUDART_INF_F; }
                    kh = fminf(ha, hb);
                }
                const float km = 0.5f * (kh + kl);
                const float emed = krel32 * fmaxf(fabsf(kh), fabsf(kl)) + e0m;
                float sg, esg = 0.f;
                if constexpr (MAD) {
                    // sigma = 1.4826 * median(|w - median|) (tools.c:67-74) on the keys: the absolute deviations form four
                    // sorted runs (below / above the median in either column), merged outward from the median; the error
                    // model is the pair kernel's (sigclip_pair.cuh)
                    int la = la0, lb = lb0, ra = la0 + 1, rb = lb0 + 1;
                    float dla = la >= ao ? km - fkey(cA[la * 64]) : CUDART_INF_F;
                    float dlb = lb >= ap ? km - fkey(cB[lb * 64]) : CUDART_INF_F;
                    float dra = ra < bo ? fkey(cA[ra * 64]) - km : CUDART_INF_F;
                    float drb = rb < bp ? fkey(cB[rb * 64]) - km : CUDART_INF_F;
                    float cur = 0.f, prev = 0.f;
                    for (int s2 = 0; s2 <= (m >> 1); s2++) {
                        prev = cur;
                        if (dla <= dlb && dla <= dra && dla <= drb) {
                            cur = dla; la--;
                            dla = la >= ao ? km - fkey(cA[la * 64]) : CUDART_INF_F;
                        } else if (dlb <= dra && dlb <= drb) {
                            cur = dlb; lb--;
                            dlb = lb >= ap ? km - fkey(cB[lb * 64]) : CUDART_INF_F;
                        } else if (dra <= drb) {
                            cur = dra; ra++;
                            dra = ra < bo ? fkey(cA[ra * 64]) - km : CUDART_INF_F;
                        } else {
                            cur = drb; rb++;
                            drb = rb < bp ? fkey(cB[rb * 64]) - km : CUDART_INF_F;
                        }
                    }
                    cur = fabsf(cur);
                    prev = fabsf(prev);
                    const float madk = (m & 1) ? cur : 0.5f * (cur + prev);
                    sg = madk * 1.4826f;
                    esg = 1.4826f * (krel32 * 1.02f * (fabsf(km) + madk) + e0m + emed) + 0x1p-20f * sg;
                } else {
                    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(sg) : "f"(__double2float_rn(var_k)));
                }
                const float lo_s = nlo32 * sg, up_s = nup32 * sg;
                clo = km - lo_s;
                cup = km + up_s;
                Tb = emed + fmaxf(nlo32, nup32) * esg + 0x1p-19f * (fabsf(km) + lo_s + up_s);
                const bool bad_tb = !(Tb < 0x1p100f);
                const float yl = clo + Tb + e0m, xu = cup - Tb - e0m;
                const float kL2 = yl * (yl >= 0.f ? c2f : c1f);
                const float kU2 = xu * (xu >= 0.f ? c1f : c2f);
                // Candidates: the keys at or below kL2 / at or above kU2.  Each column is sorted, so they are a prefix / suffix of
                // EACH column's window: five independent probes per column and end (no merge of the two columns is needed;
                // the boundary key of the union is the larger / smaller of the columns' last candidates).
                unsigned BLo = 0xFF80u, BLp = 0xFF80u, BUo = 0x7F80u, BUp = 0x7F80u;
#pragma unroll
                for (int j = 0; j < 5; j++) {
                    const bool ia = ao + j < bo, ib = ap + j < bp;
                    const unsigned ua = ia ? (unsigned)cA[(ao + j) * 64] : 0x7F80u;
                    const unsigned ub = ib ? (unsigned)cB[(ap + j) * 64] : 0x7F80u;
                    const bool ca = ia && !(fkey(ua) > kL2), cb = ib && !(fkey(ub) > kL2);
                    jl_o += ca ? 1 : 0;
                    jl_p += cb ? 1 : 0;
                    BLo = ca ? ua : BLo;
                    BLp = cb ? ub : BLp;
                }
#pragma unroll
                for (int j = 0; j < 5; j++) {
                    const bool ia = bo - 1 - j >= ao, ib = bp - 1 - j >= ap;
                    const unsigned ua = ia ? (unsigned)cA[(bo - 1 - j) * 64] : 0xFF80u;
                    const unsigned ub = ib ? (unsigned)cB[(bp - 1 - j) * 64] : 0xFF80u;
                    const bool ca = ia && !(fkey(ua) < kU2), cb = ib && !(fkey(ub) < kU2);
                    ju_o += ca ? 1 : 0;
                    ju_p += cb ? 1 : 0;
                    BUo = ca ? ua : BUo;
                    BUp = cb ? ub : BUp;
                }
                BL = jl_o == 0 ? BLp : (jl_p == 0 ? BLo : (fkey(BLo) >= fkey(BLp) ? BLo : BLp));
                BU = ju_o == 0 ? BUp : (ju_p == 0 ? BUo : (fkey(BUo) <= fkey(BUp) ? BUo : BUp));
                const int jl = jl_o + jl_p, ju = ju_o + ju_p;
                const bool many = jl > 4 || ju > 4 || jl + ju > m;
                go_slow = bad || bad_tb || many;
                cand = !go_slow && (jl + ju) != 0;
                if (!cand) { jl_o = jl_p = ju_o = ju_p = 0; BL = 0xFF80u; BU = 0x7F80u; }
            }
            // ---- exchange: (flags, the partner column's expected counts), bounds, boundary keys --------------------------
            const unsigned w1 = (go_slow ? 1u : 0u) | (cand ? 2u : 0u) | ((unsigned)jl_p << 4) | ((unsigned)ju_p << 8);
            const unsigned w2 = BL | (BU << 16);
            const unsigned r1 = __shfl_xor_sync(FULLM, w1, 16);
            const unsigned r2 = __shfl_xor_sync(FULLM, w2, 16);
            const float rclo = __shfl_xor_sync(FULLM, clo, 16), rcup = __shfl_xor_sync(FULLM, cup, 16), rtb = __shfl_xor_sync(FULLM, Tb, 16);
            const double rvar = __shfl_xor_sync(FULLM, var_k, 16);
            int ncl[2], nch[2];
            float clo32[2], cup32[2], tb32[2];
            unsigned BLh[2], BUh[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const bool mine = (h == half);
                const bool was_act = !done[h];
                const bool gs = mine ? go_slow : (r1 & 1u) != 0;
                const bool cd = mine ? cand : (r1 & 2u) != 0;
                ncl[h] = mine ? jl_o : (int)((r1 >> 4) & 15u);
                nch[h] = mine ? ju_o : (int)((r1 >> 8) & 15u);
                clo32[h] = mine ? clo : rclo;
                cup32[h] = mine ? cup : rcup;
                tb32[h] = mine ? Tb : rtb;
                BLh[h] = mine ? BL : (r2 & 0xFFFFu);
                BUh[h] = mine ? BU : (r2 >> 16);
                if (was_act) {
                    if constexpr (!PROP) {
                        const double vk = mine ? var_k : rvar;
                        if (!gs || cd) varl[h] = vk;
                    }
                    slow[h] = slow[h] || gs;
                    done[h] = !cd;
                }
                if (!was_act || !cd) { ncl[h] = 0; nch[h] = 0; BLh[h] = 0xFF80u; BUh[h] = 0x7F80u; }
            }
            const unsigned thr_lo = BLh[0] | (BLh[1] << 16);
            const unsigned thr_hi = BUh[0] | (BUh[1] << 16);
            // a voxel is a candidate voxel if EITHER column expects candidates: both lanes must go through part B together
            const bool cv0 = !done[0], cv1 = !done[1];
            if (!__any_sync(FULLM, cv0 || cv1)) continue;
            // ---- comparison pass over this lane's unsorted keys (see sigclip_pair.cuh) --------------------------------------
            unsigned RL[2][W], RH[2][W];
#pragma unroll
            for (int kk = 0; kk < W; kk++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    asm volatile("mov.b32 %0, 0;" : "=r"(RL[h][kk]));
                    asm volatile("mov.b32 %0, 0;" : "=r"(RH[h][kk]));
                }
            }
            const bool any_low = __any_sync(FULLM, (ncl[0] | ncl[1]) != 0), any_high = __any_sync(FULLM, (nch[0] | nch[1]) != 0);
            if (any_low) {
#pragma unroll
                for (int i = 0; i < HB; i++) {
                    asm("{\n\t.reg .pred p0, p1;\n\t"
                        "setp.le.bf16x2 p0|p1, %2, %3;\n\t"
                        "@p0 add.u32 %0, %0, %4;\n\t"
                        "@p1 add.u32 %1, %1, %4;\n\t}"
                        : "+r"(RL[0][i >> 5]), "+r"(RL[1][i >> 5])
                        : "r"(ukey[i]), "r"(thr_lo), "r"(1u << (i & 31)));
                }
            }
            if (any_high) {
#pragma unroll
                for (int i = 0; i < HB; i++) {
                    asm("{\n\t.reg .pred p2, p3;\n\t"
                        "setp.ge.bf16x2 p2|p3, %2, %3;\n\t"
                        "@p2 add.u32 %0, %0, %4;\n\t"
                        "@p3 add.u32 %1, %1, %4;\n\t}"
                        : "+r"(RH[0][i >> 5]), "+r"(RH[1][i >> 5])
                        : "r"(ukey[i]), "r"(thr_hi), "r"(1u << (i & 31)));
                }
            }
            // ---- part B: this lane's candidates, decided from their exact values; sums over the two lanes ------------------
            double dT1[2] = {0.0, 0.0}, dS2[2] = {0.0, 0.0};
            int rl[2] = {0, 0}, rh[2] = {0, 0};
            unsigned bad[2] = {0u, 0u};
            unsigned long long nrj[2] = {0ull, 0ull};
#pragma unroll
            for (int h = 0; h < 2; h++) {
                if (done[h]) continue;                                              // (identical in both lanes of the pair)
                int cl = 0, ch = 0;
#pragma unroll
                for (int kk = 0; kk < W; kk++) {
                    RL[h][kk] &= ~rej[h].w[kk];
                    RH[h][kk] &= ~rej[h].w[kk];
                    cl += __popc(RL[h][kk]);
                    ch += __popc(RH[h][kk]);
                }
                if (cl != ncl[h] || ch != nch[h]) { bad[h] = 1u; continue; }        // key ties the model does not cover
                const double clo_d = (double)clo32[h], cup_d = (double)cup32[h], tb_d = (double)tb32[h];
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
                    const float dval = dt[e_base + i][2 * pr + h];
                    const double t = UNIT ? (double)dval - p[h] : __fma_rn((double)dval, os_l[i].y, osp_l[i] - p[h]);
                    const double diff = t - (low ? clo_d : cup_d);
                    if (!(fabs(diff) > tb_d)) bad[h] = 1u;
                    const bool out = low ? (diff <= 0.0) : (diff >= 0.0);
                    if (out) {
                        dT1[h] += t;
                        dS2[h] += t * t;
                        nrj[h] |= bit;
                        if (low) rl[h]++; else rh[h]++;
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                // sums over the pair (executed by every lane: the shuffles must not sit in divergent code)
                const unsigned pk = (unsigned)rl[h] | ((unsigned)rh[h] << 8) | (bad[h] << 16);
                const unsigned pko = __shfl_xor_sync(FULLM, pk, 16);
                const double dT1o = __shfl_xor_sync(FULLM, dT1[h], 16), dS2o = __shfl_xor_sync(FULLM, dS2[h], 16);
                if (done[h]) continue;
                const int rlt = rl[h] + (int)(pko & 0xFFu), rht = rh[h] + (int)((pko >> 8) & 0xFFu);
                if (bad[h] || (pko >> 16)) { slow[h] = done[h] = true; continue; }
                const int ni = mt[h] - rlt - rht;
                if (ni < nstop || ni == mt[h]) { done[h] = true; continue; }
                itl[h]--;
                T1[h] -= dT1[h] + dT1o;
                S2[h] -= dS2[h] + dS2o;
#pragma unroll
                for (int kk = 0; kk < W; kk++) rej[h].w[kk] |= (unsigned)(nrj[h] >> (kk ? 32 : 0));
                a[h] += rl[h];
                b[h] -= rh[h];
                mt[h] = ni;
                if (itl[h] <= 0) final_stats(h);
            }
        }
        __syncwarp();
        if (lane == 0) request_tile(r + Cfg::SPW);

        // ---- results -------------------------------------------------------------------------------------------------------
        int kept[2];
        double mean[2];
        ExpSet<HB> drop[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            kept[h] = mt[h];
            mean[h] = n[h] > 0 ? p[h] + T1[h] * sm.rcp[kept[h]] : CUDART_NAN;
#pragma unroll
            for (int kk = 0; kk < W; kk++) {
                if (slow[h]) rej[h].w[kk] = 0u;
                drop[h].w[kk] = inval[h].w[kk] | rej[h].w[kk];
            }
        }
        auto count_tile = [&]() {
#pragma unroll
            for (int g = 0; g < 2; g++) {
#pragma unroll
                for (int h = 0; h < 2; h++) {
#pragma unroll
                    for (int kk = 0; kk < W; kk++) {
                        cnt_inval[g][kk] += __popc(tc.apply(half == g ? inval[h].w[kk] : 0u));
                        cnt_drop[g][kk] += __popc(tc.apply(half == g ? drop[h].w[kk] : 0u));
                    }
                }
            }
            tiles_done++;
        };
        double vsum[2] = {0.0, 0.0};
        if constexpr (PROP) {
            constexpr int NACC = (HB + 2) / 3;
            float vacc[2][NACC];
#pragma unroll
            for (int j = 0; j < NACC; j++) { vacc[0][j] = 0.f; vacc[1][j] = 0.f; }
            unsigned neg = 0u;
            if (pair_ok) {
                // this lane's rows: slots e_base .. e_base + HB - 1; slots beyond the stack re-read its last row (dropped anyway)
                const int last = A.nfiles - 1;
                const float* vp = static_cast<const float*>(S.var0) + v0;
                float2 sv[HB];
#pragma unroll
                for (int i = 0; i < HB; i++) {
                    const int e = e_base + i;
                    const float* rp = vp + (size_t)(e < last ? e : last) * S.vstride;      // row min(slot, nfiles - 1): never beyond the stack
                    asm volatile("ld.global.cs.v2.f32 {%0, %1}, [%2];" : "=f"(sv[i].x), "=f"(sv[i].y) : "l"(rp));
                }
                asm volatile("" ::: "memory");
                count_tile();
#pragma unroll
                for (int i = 0; i < HB; i++) {
                    const float s2 = UNIT ? 1.f : s2f_l[i];
                    if (!drop[0].test(i)) vacc[0][i % NACC] = __fmaf_rn(sv[i].x, s2, vacc[0][i % NACC]);
                    if (!drop[1].test(i)) vacc[1][i % NACC] = __fmaf_rn(sv[i].y, s2, vacc[1][i % NACC]);
                    neg |= __float_as_uint(sv[i].x) | __float_as_uint(sv[i].y);
                }
#pragma unroll
                for (int j = 0; j < NACC; j++) { vsum[0] += (double)vacc[0][j]; vsum[1] += (double)vacc[1][j]; }
            } else {
                count_tile();
            }
            neg |= __shfl_xor_sync(FULLM, neg, 16);
            if (!pair_ok || (neg & 0x80000000u)) {
                const float* var0 = static_cast<const float*>(S.var0);
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    vsum[h] = 0.0;
                    if (v0 + h >= A.nvox) continue;
                    for (int i = 0; i < HB && e_base + i < A.nfiles; i++)
                        if (!drop[h].test_dyn(i))
                            vsum[h] += (double)__ldg(var0 + (size_t)(e_base + i) * S.vstride + (v0 + h)) * (os_l[i].y * os_l[i].y);
                }
            }
#pragma unroll
            for (int h = 0; h < 2; h++) vsum[h] += __shfl_xor_sync(FULLM, vsum[h], 16);
        } else {
            count_tile();
        }

        // ---- stores (the lower lane of a pair) -------------------------------------------------------------------------------
        double ov[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const double rk = sm.rcp[kept[h]];
            if (n[h] == 0) ov[h] = CUDART_NAN;
            else if (PROP) ov[h] = vsum[h] * rk * rk;
            else if (kept[h] > 1) {
                ov[h] = varl[h];
                if (A.typ_var == 1) ov[h] *= sm.rcp[kept[h] - 1];
            } else ov[h] = CUDART_NAN;
        }
        if (S.hist != nullptr) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int val = (half != 0 || slow[h] || v0 + h >= A.nvox) ? -1 : kept[h];
                const unsigned peers = __match_any_sync(FULLM, val);
                if (val >= 0 && lane == __ffs(peers) - 1) atomicAdd(&sm.hist[val], (unsigned)__popc(peers));
            }
        }
        if (half == 0) {
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
                        else exact_voxel_store<NB, UNIT, float>(S, A, v);
                    }
                }
            }
        }
    }

    // ---- flush: valid = voxels seen - invalid, selected = voxels seen - dropped -----------------------------------------------
    const unsigned seen = tiles_done * 32u;
#pragma unroll
    for (int g = 0; g < 2; g++) {
#pragma unroll
        for (int kk = 0; kk < W; kk++) {
            const int e = g * HB + 32 * kk + lane;
            if (32 * kk + lane < HB && e < A.nfiles) {
                atomicAdd(A.valid_cnt + e, (unsigned long long)(seen - cnt_inval[g][kk]));
                atomicAdd(A.select_cnt + e, (unsigned long long)(seen - cnt_drop[g][kk]));
            }
        }
    }
    if (S.hist != nullptr) {
        __syncthreads();
        for (int i = tid; i <= NB; i += Cfg::threads)
            if (sm.hist[i]) atomicAdd(S.hist + i, (unsigned long long)sm.hist[i]);
    }
}

}  // namespace mb
