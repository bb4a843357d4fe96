// TMA-pipelined median kernel for sm_100a (float32 stacks up to 96 deep whose rows are equally pitched and
// 16-byte aligned): CubeList.median / mpdaf_merging_median.
//
// Reference behaviour restated (never copied): src/merging.c:232-253 (gather of the non-NaN samples, n == 0 -> NaN,
// valid counters) and src/tools.c:39-51 (median of the sorted samples, mean of the two middle ones for even n).
//
// The round-1 kernel (median_warp_kernel, sigclip_warp.cuh) gathers the exposure rows into registers itself and is
// bound by the latency of those loads (0.68 of the measured HBM peak at N = 12).  Here every warp keeps up to four
// shared-memory stages of its own filled with [NB][64] tiles by cp.async.bulk.tensor (TMA, mbarrier completion) and
// only ever reads shared memory: a lane takes TWO neighbouring voxels of a tile (two independent register sorting
// networks: twice the instruction-level parallelism, 16-byte output stores), requests the stage's next tile as soon
// as the samples are in registers, and the kernel runs at the rate the stages can be refilled.
// Invalid samples are replaced alternately by +inf / -inf, so that the valid ones stay centred in the sorted stack
// and the median sits in two fixed registers; results are bit-identical to the reference.
#pragma once

#include "sigclip_pair.cuh"

namespace mb {

template <int NB>
struct MedianOwnCfg {
    // shallow stacks take wider tiles (fewer, larger TMA copies), processed by the warp as SUB sub-tiles of 64 voxels
    static constexpr int SUB = NB <= 16 ? 2 : 1;   // (256-voxel boxes faulted on the B200: not pursued)
    static constexpr int TV = 64 * SUB;                                          // voxels per tile
    static constexpr int stage_bytes = NB * TV * 4;
    // Every warp is a consumer that owns SPW stages and refills them itself (no producer warp, as in sigclip_pair.cuh):
    // its j-th tile uses its stage j % SPW, whose next tile (j + SPW) is requested the moment the samples are in registers.
    static constexpr int c_regs = 2048 / (2 * NB + 40);                         // two voxels of NB samples per lane
    static constexpr int C = c_regs < 16 ? (c_regs < 2 ? 2 : c_regs) : 16;       // four warps per scheduler at most
    static constexpr int st_fit = (200 * 1024) / stage_bytes;
    static constexpr int SPW = st_fit / C > 4 ? 4 : st_fit / C;
    static constexpr int ST = SPW * C;
    static constexpr int threads = C * 32;
    static_assert(SPW >= 1, "stack too deep for the pipelined median kernel");
};

template <int NB>
struct MedianOwnSmem {
    float stage[MedianOwnCfg<NB>::ST][NB][MedianOwnCfg<NB>::TV];
    unsigned long long full[MedianOwnCfg<NB>::ST];
};

template <int NB>
__global__ void __launch_bounds__(MedianOwnCfg<NB>::threads, 1)
median_own_kernel(const ClipArgs A, const long long ntiles, const __grid_constant__ TmaMaps M)
{
    using Cfg = MedianOwnCfg<NB>;
    constexpr int C = Cfg::C, ST = Cfg::ST, SPW = Cfg::SPW;
    constexpr int W = ExpSet<NB>::W;
    constexpr int SUB = Cfg::SUB, TV = Cfg::TV;
    constexpr unsigned kTileBytes = NB * TV * sizeof(float);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    MedianOwnSmem<NB>& sm = *reinterpret_cast<MedianOwnSmem<NB>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < ST; s++) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    const int c = warp;
    const unsigned one = (unsigned)(A.nfiles > 0);     // opaque 1 / -1 for the IMAD form of the compare-exchange
    const unsigned minus_one = 0u - one;
    TransposeConsts tc;
    tc.init(lane);
    unsigned cnt[W];
#pragma unroll
    for (int k = 0; k < W; k++) cnt[k] = 0u;
    unsigned tiles_done = 0;
    // this warp's j-th tile is the block's tile j * C + c, i.e. tile blockIdx.x + (j * C + c) * gridDim.x of the slab
    auto request_tile = [&](long long j) {             // lane 0 only
        const long long t = (long long)blockIdx.x + (j * C + c) * (long long)gridDim.x;
        if (t >= ntiles) return;
        const unsigned st = (unsigned)(c * SPW + (int)(j % SPW));
        mbar_expect_tx(&sm.full[st], kTileBytes);
        tma_load_2d(&sm.stage[st][0][0], &M.data, (int)(t * TV), 0, &sm.full[st]);
    };
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < SPW; j++) request_tile(j);
    }
    long long tile = (long long)blockIdx.x + (long long)c * gridDim.x;
    for (long long j = 0; tile < ntiles; j++, tile += (long long)C * gridDim.x) {
        const unsigned st = (unsigned)(c * SPW + (int)(j % SPW)), ph = (unsigned)(j / SPW) & 1u;
        const float (*dt)[TV] = sm.stage[st];
        mbar_wait(&sm.full[st], ph);
#pragma unroll
        for (int sub = 0; sub < SUB; sub++) {
            const long long v0 = tile * TV + 64 * sub + 2 * lane;
            float key[2][NB];
#pragma unroll
            for (int i = 0; i < NB; i++) {
                const float2 dd = *reinterpret_cast<const float2*>(&dt[i][64 * sub + 2 * lane]);
                key[0][i] = dd.x;
                key[1][i] = dd.y;
            }
            if (sub == SUB - 1) {
                __syncwarp();
                if (lane == 0) request_tile(j + SPW);          // the samples are in registers: the stage is refilled at once
            }
            double med[2];
            int n[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                ExpSet<NB> inval;
                bool flip = false;                         // parity of the invalid samples seen so far
#pragma unroll
                for (int i = 0; i < NB; i++) {
                    const float d = key[h][i];
                    const bool ok = (d == d);              // rows beyond nfiles and voxels beyond nvox arrive as NaN (TMA fill)
                    key[h][i] = ok ? d : (flip ? -CUDART_INF_F : CUDART_INF_F);
                    if (!ok) {
                        flip = !flip;
                        inval.set(i);
                    }
                }
                n[h] = NB - inval.count();
                sort_keys_mixed<NB, 1>(key[h], one, minus_one);
                // the first invalid sample went to +inf: the valid ones occupy slots [floor((NB - n) / 2), ...) and their
                // middle is slot NB/2 - 1 (odd n) or slots NB/2 - 1 and NB/2 (even n)
                const double lo = (double)key[h][NB / 2 - 1], hi = (double)key[h][NB / 2];
                med[h] = (n[h] & 1) ? lo : (lo + hi) * 0.5;                         // tools.c:45-48
                if (n[h] == 0) med[h] = CUDART_NAN;                                 // merging.c:239
#pragma unroll
                for (int kk = 0; kk < W; kk++) cnt[kk] += __popc(tc.apply(inval.w[kk]));
            }
            if (v0 + 1 < A.nvox) {
                *reinterpret_cast<double2*>(A.out + v0) = make_double2(med[0], med[1]);
                *reinterpret_cast<int2*>(A.expmap + v0) = make_int2(n[0], n[1]);
            } else if (v0 < A.nvox) {
                A.out[v0] = med[0];
                A.expmap[v0] = n[0];
            }
        }
        tiles_done++;
    }
    const unsigned seen = tiles_done * (unsigned)TV;
#pragma unroll
    for (int kk = 0; kk < W; kk++) {
        const int e = 32 * kk + lane;
        if (e < A.nfiles) atomicAdd(A.valid_cnt + e, (unsigned long long)(seen - cnt[kk]));
    }
}

// ---- the producer / consumer form (first round-2 version): one producer lane keeps a ring of ST stages full, fifteen
// consumer warps.  Still the faster one at N <= 12 on a full-size cube (cfg2: 5707 GB/s against 5090 for the form
// above, which wins by 9 .. 10 % from N = 16 on); the launcher chooses (warp_inst.cuh).
template <int NB>
struct MedianPipeCfg {
    // shallow stacks take wider tiles (one lane issues the TMA copies: about one per 270 cycles), processed by the
    // consumer warp as SUB sub-tiles of 64 voxels
    static constexpr int SUB = NB <= 16 ? 2 : 1;   // (256-voxel boxes faulted on the B200: not pursued)
    static constexpr int TV = 64 * SUB;                                          // voxels per tile
    static constexpr int stage_bytes = NB * TV * 4;
    static constexpr int c_regs = 2048 / (2 * NB + 40) - 1;                    // two voxels of NB samples per lane
    static constexpr int C = c_regs < 15 ? (c_regs < 2 ? 2 : c_regs) : 15;       // consumer warps (+ 1 producer warp)
    static constexpr int st_fit = (200 * 1024) / stage_bytes;
    // Ring stages: a multiple of C, so that stage s is only ever consumed by warp s % C and its tiles are taken in
    // order.  (With an arbitrary ring length the warp waiting for a stage's tile k + 2 ST can overtake the warp that
    // still owes tile k + ST: an mbarrier wait only knows the PARITY of the phase it waits for, the earlier phase
    // of equal parity has completed, the wait returns at once, and the kernel reads a stale tile and later hangs.)
    static constexpr int st_mult = st_fit / C > 4 ? 4 : st_fit / C;
    static constexpr int ST = st_mult * C;
    static constexpr int threads = (C + 1) * 32;
    static_assert(st_mult >= 1, "stack too deep for the pipelined median kernel");
};

template <int NB>
struct MedianPipeSmem {
    float stage[MedianPipeCfg<NB>::ST][NB][MedianPipeCfg<NB>::TV];
    unsigned long long full[MedianPipeCfg<NB>::ST];
    unsigned long long empty[MedianPipeCfg<NB>::ST];
};

template <int NB>
__global__ void __launch_bounds__(MedianPipeCfg<NB>::threads, 1)
median_pipe_kernel(const ClipArgs A, const long long ntiles, const __grid_constant__ TmaMaps M)
{
    using Cfg = MedianPipeCfg<NB>;
    constexpr int C = Cfg::C, ST = Cfg::ST;
    constexpr int W = ExpSet<NB>::W;
    constexpr int SUB = Cfg::SUB, TV = Cfg::TV;
    constexpr unsigned kTileBytes = NB * TV * sizeof(float);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    MedianPipeSmem<NB>& sm = *reinterpret_cast<MedianPipeSmem<NB>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        for (int s = 0; s < ST; s++) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    if (warp == C) {                                   // producer: one lane keeps the ring full
        if (lane == 0) {
            long long tile = blockIdx.x;
            for (unsigned k = 0; tile < ntiles; k++, tile += gridDim.x) {
                const unsigned st = k % ST, ph = (k / ST) & 1u;
                // plain spin: at N = 12 a tile is due every ~240 cycles and a suspended wait wakes up too late
                mbar_wait(&sm.empty[st], ph ^ 1u);
                mbar_expect_tx(&sm.full[st], kTileBytes);
                tma_load_2d(&sm.stage[st][0][0], &M.data, (int)(tile * TV), 0, &sm.full[st]);
            }
        }
        return;
    }

    const int c = warp;
    const unsigned one = (unsigned)(A.nfiles > 0);     // opaque 1 / -1 for the IMAD form of the compare-exchange
    const unsigned minus_one = 0u - one;
    TransposeConsts tc;
    tc.init(lane);
    unsigned cnt[W];
#pragma unroll
    for (int k = 0; k < W; k++) cnt[k] = 0u;
    unsigned tiles_done = 0;
    long long tile = (long long)blockIdx.x + (long long)c * gridDim.x;
    for (unsigned k = c; tile < ntiles; k += C, tile += (long long)C * gridDim.x) {
        const unsigned st = k % ST, ph = (k / ST) & 1u;
        const float (*dt)[TV] = sm.stage[st];
        mbar_wait(&sm.full[st], ph);
#pragma unroll
        for (int sub = 0; sub < SUB; sub++) {
            const long long v0 = tile * TV + 64 * sub + 2 * lane;
            float key[2][NB];
#pragma unroll
            for (int i = 0; i < NB; i++) {
                const float2 dd = *reinterpret_cast<const float2*>(&dt[i][64 * sub + 2 * lane]);
                key[0][i] = dd.x;
                key[1][i] = dd.y;
            }
            if (sub == SUB - 1) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.empty[st]);     // the samples are in registers: the stage can be refilled
            }
            double med[2];
            int n[2];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                ExpSet<NB> inval;
                bool flip = false;                         // parity of the invalid samples seen so far
#pragma unroll
                for (int i = 0; i < NB; i++) {
                    const float d = key[h][i];
                    const bool ok = (d == d);              // rows beyond nfiles and voxels beyond nvox arrive as NaN (TMA fill)
                    key[h][i] = ok ? d : (flip ? -CUDART_INF_F : CUDART_INF_F);
                    if (!ok) {
                        flip = !flip;
                        inval.set(i);
                    }
                }
                n[h] = NB - inval.count();
                sort_keys_mixed<NB, 1>(key[h], one, minus_one);
                // the first invalid sample went to +inf: the valid ones occupy slots [floor((NB - n) / 2), ...) and their
                // middle is slot NB/2 - 1 (odd n) or slots NB/2 - 1 and NB/2 (even n)
                const double lo = (double)key[h][NB / 2 - 1], hi = (double)key[h][NB / 2];
                med[h] = (n[h] & 1) ? lo : (lo + hi) * 0.5;                         // tools.c:45-48
                if (n[h] == 0) med[h] = CUDART_NAN;                                 // merging.c:239
#pragma unroll
                for (int kk = 0; kk < W; kk++) cnt[kk] += __popc(tc.apply(inval.w[kk]));
            }
            if (v0 + 1 < A.nvox) {
                *reinterpret_cast<double2*>(A.out + v0) = make_double2(med[0], med[1]);
                *reinterpret_cast<int2*>(A.expmap + v0) = make_int2(n[0], n[1]);
            } else if (v0 < A.nvox) {
                A.out[v0] = med[0];
                A.expmap[v0] = n[0];
            }
        }
        tiles_done++;
    }
    const unsigned seen = tiles_done * (unsigned)TV;
#pragma unroll
    for (int kk = 0; kk < W; kk++) {
        const int e = 32 * kk + lane;
        if (e < A.nfiles) atomicAdd(A.valid_cnt + e, (unsigned long long)(seen - cnt[kk]));
    }
}

// ---- 49 .. 128 exposures: the stack of a voxel pair split between two lanes (as in sigclip_duo.cuh) ------------------------
// A tile is 32 voxels; lanes l and l + 16 share the voxel pair (2 (l & 15), 2 (l & 15) + 1) and each takes HB of the 2 HB
// exposure slots for BOTH voxels into registers (invalid samples become +inf: they sort to the top), sorts its two
// HB-deep stacks with the float32 network and stores them as sorted columns.  The lower lane then finds the two middle
// values of voxel 0 in the union of the two lanes' columns by elimination (sigclip_runs.cuh), the upper lane those of
// voxel 1.  Two sorts of HB per voxel instead of one of 2 HB (734 against ~1000 comparators at 96 exposures), a TMA
// pipeline instead of register gathers.  Bit-identical to the reference (tools.c:39-51).
template <int HB>
struct MedianDuoCfg {
    static constexpr int NB = 2 * HB;
    static constexpr int stage_bytes = NB * 32 * 4;
    static constexpr int col_bytes = HB * 32 * 8;       // sorted float32 columns: (voxel 0, voxel 1) per lane and rank
    static constexpr int c_regs = 2048 / (2 * HB + 48);
    static constexpr int c_smem = (220 * 1024) / (stage_bytes + col_bytes);
    static constexpr int c_min = c_regs < c_smem ? c_regs : c_smem;
    static constexpr int c_cap = c_min < 16 ? c_min : 16;
    static constexpr int C = (c_cap > 4 && c_cap % 4 == 1) ? c_cap - 1 : c_cap;
    static constexpr int threads = C * 32;
    static_assert(C >= 2, "stack too deep for the two-lane median kernel");
};

template <int HB>
struct MedianDuoSmem {
    float stage[MedianDuoCfg<HB>::C][2 * HB][32];
    float2 col[MedianDuoCfg<HB>::C][HB][32];
    unsigned long long full[MedianDuoCfg<HB>::C];
};

template <int HB>
__global__ void __launch_bounds__(MedianDuoCfg<HB>::threads, 1)
median_duo_kernel(const ClipArgs A, const long long ntiles, const __grid_constant__ TmaMaps M)
{
    using Cfg = MedianDuoCfg<HB>;
    constexpr int NB = 2 * HB, C = Cfg::C;
    constexpr int W = ExpSet<HB>::W;
    constexpr unsigned FULLM = 0xffffffffu;
    constexpr unsigned kTileBytes = NB * 32 * sizeof(float);
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    MedianDuoSmem<HB>& sm = *reinterpret_cast<MedianDuoSmem<HB>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, c = tid >> 5;
    const int half = lane >> 4, pr = lane & 15, e_base = half * HB;
    if (tid == 0) {
        for (int s = 0; s < C; s++) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const unsigned one = (unsigned)(A.nfiles > 0);
    const unsigned minus_one = 0u - one;
    TransposeConsts tc;
    tc.init(lane);
    unsigned cnt[2][W];
#pragma unroll
    for (int g = 0; g < 2; g++) {
#pragma unroll
        for (int k = 0; k < W; k++) cnt[g][k] = 0u;
    }
    unsigned tiles_done = 0;
    auto request_tile = [&](long long j) {             // lane 0 only: this warp's j-th tile into its stage
        const long long t = (long long)blockIdx.x + (j * C + c) * (long long)gridDim.x;
        if (t >= ntiles) return;
        mbar_expect_tx(&sm.full[c], kTileBytes);
        tma_load_2d(&sm.stage[c][0][0], &M.data, (int)(t << 5), 0, &sm.full[c]);
    };
    if (lane == 0) request_tile(0);
    const float* colf_own = reinterpret_cast<const float*>(&sm.col[c][0][lane]) + half;        // rank q of my voxel: [q * 64]
    const float* colf_par = reinterpret_cast<const float*>(&sm.col[c][0][lane ^ 16]) + half;
    long long tile = (long long)blockIdx.x + (long long)c * gridDim.x;
    for (long long j = 0; tile < ntiles; j++, tile += (long long)C * gridDim.x) {
        const float (*dt)[32] = sm.stage[c];
        mbar_wait(&sm.full[c], (unsigned)j & 1u);
        const long long v0 = (tile << 5) + 2 * pr;
        float key[2][HB];
#pragma unroll
        for (int i = 0; i < HB; i++) {
            const float2 dd = *reinterpret_cast<const float2*>(&dt[e_base + i][2 * pr]);
            key[0][i] = dd.x;
            key[1][i] = dd.y;
        }
        __syncwarp();
        if (lane == 0) request_tile(j + 1);            // the samples are in registers: the stage is refilled at once
        int n_own[2];
        ExpSet<HB> inval[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int i = 0; i < HB; i++) {
                const float d = key[h][i];
                const bool ok = (d == d);              // slots beyond nfiles and voxels beyond nvox arrive as NaN (TMA fill)
                key[h][i] = ok ? d : CUDART_INF_F;
                if (!ok) inval[h].set(i);
            }
            n_own[h] = HB - inval[h].count();
            sort_keys_mixed<HB, 1>(key[h], one, minus_one);
        }
#pragma unroll
        for (int q = 0; q < HB; q++) sm.col[c][q][lane] = make_float2(key[0][q], key[1][q]);
        __syncwarp();
        // my voxel: `half`; its valid samples are the first nA ranks of my column and the first nB of the partner's
        const int nA = half ? n_own[1] : n_own[0];
        const int nB = __shfl_xor_sync(FULLM, half ? n_own[0] : n_own[1], 16);
        const int n = nA + nB;
        double med = CUDART_NAN;
        if (n > 0) {
            int pa = 0, pb = 0;
            int need = (n - 1) >> 1;                   // ranks below the lower middle
            while (need > 0) {
                const int t = (need + 1) >> 1;
                const bool ia = pa + t - 1 < nA, ib = pb + t - 1 < nB;
                const float xa = ia ? colf_own[(pa + t - 1) * 64] : 0.f;
                const float xb = ib ? colf_par[(pb + t - 1) * 64] : 0.f;
                if (ia && (!ib || xa <= xb)) pa += t; else pb += t;     // (a valid sample may itself be +inf)
                need -= t;
            }
            bool ia = pa < nA, ib = pb < nB;
            float xa = ia ? colf_own[pa * 64] : 0.f, xb = ib ? colf_par[pb * 64] : 0.f;
            const bool takeA = ia && (!ib || xa <= xb);
            const float lo = takeA ? xa : xb;
            float hi = lo;
            if (!(n & 1)) {
                if (takeA) { pa++; ia = pa < nA; xa = ia ? colf_own[pa * 64] : 0.f; }
                else { pb++; ib = pb < nB; xb = ib ? colf_par[pb * 64] : 0.f; }
                hi = (ia && (!ib || xa <= xb)) ? xa : xb;
            }
            med = (n & 1) ? (double)lo : ((double)lo + (double)hi) * 0.5;          // tools.c:45-48
        }
        const double med_o = __shfl_xor_sync(FULLM, med, 16);
        const int n_o = __shfl_xor_sync(FULLM, n, 16);
#pragma unroll
        for (int g = 0; g < 2; g++) {
#pragma unroll
            for (int h = 0; h < 2; h++) {
#pragma unroll
                for (int kk = 0; kk < W; kk++) cnt[g][kk] += __popc(tc.apply(half == g ? inval[h].w[kk] : 0u));
            }
        }
        tiles_done++;
        if (half == 0) {                               // the lower lane holds voxel 0's result and receives voxel 1's
            if (v0 + 1 < A.nvox) {
                *reinterpret_cast<double2*>(A.out + v0) = make_double2(med, med_o);
                *reinterpret_cast<int2*>(A.expmap + v0) = make_int2(n, n_o);
            } else if (v0 < A.nvox) {
                A.out[v0] = med;
                A.expmap[v0] = n;
            }
        }
        __syncwarp();                                  // the columns are rewritten by the next tile
    }
    const unsigned seen = tiles_done * 32u;
#pragma unroll
    for (int g = 0; g < 2; g++) {
#pragma unroll
        for (int kk = 0; kk < W; kk++) {
            const int e = g * HB + 32 * kk + lane;
            if (32 * kk + lane < HB && e < A.nfiles) atomicAdd(A.valid_cnt + e, (unsigned long long)(seen - cnt[g][kk]));
        }
    }
}

}  // namespace mb
