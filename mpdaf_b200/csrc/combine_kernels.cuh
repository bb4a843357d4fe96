// Per-voxel combination kernels for sm_100a (B200).
//
// Reference behaviour restated here (never copied):
//   src/merging.c:423-477   per-voxel gather / n==0 / n==1 / clip / outputs
//   src/merging.c:232-253   per-voxel median
//   src/tools.c:9-24        sequential mean + population sigma
//   src/tools.c:82-118      median-centred iterative clip (strict window, stop rules)
//   src/tools.c:122-158     MAD variant
//
// Layout: exposure i is one contiguous array of nvox samples (x fastest, then y,
// then wavelength), outputs are flat arrays of nvox elements.  One thread owns one
// voxel; the 32 lanes of a warp read 32 consecutive samples of each exposure, so
// every exposure load is one fully coalesced 128-byte line per warp.
//
// Two kernel families:
//   *_reg<NB,...>  the exposure stack of a voxel lives in registers (NB <= 64
//                  compile-time slots, unused slots hold +inf) and is ordered by a
//                  straight-line merge-exchange network (sort_networks.cuh);
//   *_generic      any depth up to 512, any element type, MAD: per-thread arrays
//                  in local memory, direct restatement of the reference loops.
//
// Arithmetic contract: every float64 statistic is computed without FMA
// contraction (file is compiled with -fmad=false) and, where it feeds a
// rejection decision, in the reference's summation order: exposure order on the
// first evaluation, ascending value order afterwards.
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "sort_networks.cuh"

namespace mb {

constexpr int kThreads = 128;      // threads per block for every kernel in this file
constexpr int kGenericMax = 512;   // deepest stack the generic kernels accept

// Kernel-parameter block describing one exposure stack (lives in the constant
// bank, so unrolled per-exposure accesses become uniform loads).
template <int NB>
struct StackArgs {
    const void* data[NB];
    const void* var[NB];
    double scale[NB];
    double offset[NB];
    double scale2[NB];
};

// Same content for the generic kernels, stored in device memory.
struct StackTable {
    const void* data[kGenericMax];
    const void* var[kGenericMax];
    double scale[kGenericMax];
    double offset[kGenericMax];
    double scale2[kGenericMax];
};

struct ClipArgs {
    int nfiles;
    long long nvox;
    double* out;
    double* outvar;
    int* expmap;
    unsigned long long* valid_cnt;   // [nfiles], added to
    unsigned long long* select_cnt;  // [nfiles], added to (null for median)
    int nmax;
    double nclip_low;
    double nclip_up;
    int nstop;
    int typ_var;  // 0 propagate, 1 stat_mean, 2 stat_one
    int mad;
};

// ---------------------------------------------------------------------------
// small helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ double ld_stream(const double* p) { return __ldcs(p); }

template <typename T>
__device__ __forceinline__ void cmpxchg(T& a, T& b)
{
    const T lo = a < b ? a : b;
    const T hi = a < b ? b : a;
    a = lo;
    b = hi;
}
template <>
__device__ __forceinline__ void cmpxchg<float>(float& a, float& b)
{
    const float lo = fminf(a, b);
    const float hi = fmaxf(a, b);
    a = lo;
    b = hi;
}

template <int NB, typename T>
__device__ __forceinline__ void sort_stack(T (&a)[NB])
{
#define MB_CE(i, j) cmpxchg(a[i], a[j]);
    if constexpr (NB == 4) { MB_SORTNET_4(MB_CE) }
    else if constexpr (NB == 8) { MB_SORTNET_8(MB_CE) }
    else if constexpr (NB == 12) { MB_SORTNET_12(MB_CE) }
    else if constexpr (NB == 16) { MB_SORTNET_16(MB_CE) }
    else if constexpr (NB == 24) { MB_SORTNET_24(MB_CE) }
    else if constexpr (NB == 32) { MB_SORTNET_32(MB_CE) }
    else if constexpr (NB == 48) { MB_SORTNET_48(MB_CE) }
    else if constexpr (NB == 64) { MB_SORTNET_64(MB_CE) }
    else { static_assert(NB == 4, "no sorting network for this stack depth"); }
#undef MB_CE
}

// Per-exposure counters: one ballot per exposure, lane (i % 32) keeps the running
// count of exposure i in register slot i / 32.  Flushed once per thread at the end.
template <int NB>
struct LaneCounters {
    unsigned c[(NB + 31) / 32];
    __device__ __forceinline__ LaneCounters()
    {
#pragma unroll
        for (int k = 0; k < (NB + 31) / 32; k++) c[k] = 0;
    }
    // call only from fully unrolled loops: i must fold to a constant so that
    // c[i >> 5] stays a register
    __device__ __forceinline__ void add(int i, bool flag, int lane)
    {
        const unsigned b = __ballot_sync(0xffffffffu, flag);
        if (lane == (i & 31)) c[i >> 5] += __popc(b);
    }
    __device__ __forceinline__ void flush(unsigned long long* dst, int nfiles, int lane) const
    {
        if (dst == nullptr) return;
#pragma unroll
        for (int k = 0; k < (NB + 31) / 32; k++) {
            const int i = k * 32 + lane;
            if (i < nfiles && c[k] != 0) atomicAdd(dst + i, (unsigned long long)c[k]);
        }
    }
};

// ---------------------------------------------------------------------------
// register-resident median  (merging.c:232-253)
// ---------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(kThreads) median_reg_kernel(const StackArgs<NB> S, const ClipArgs A)
{
    const int lane = threadIdx.x & 31;
    LaneCounters<NB> valid;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long base = (long long)blockIdx.x * kThreads; base < A.nvox; base += stride) {
        const long long v = base + threadIdx.x;
        const bool act = v < A.nvox;
        float k[NB];
#pragma unroll
        for (int i = 0; i < NB; i++)
            k[i] = (act && i < A.nfiles) ? ld_stream((const float*)S.data[i] + v) : CUDART_NAN_F;
        int n = 0;
#pragma unroll
        for (int i = 0; i < NB; i++) {
            const bool ok = (k[i] == k[i]);
            n += ok;
            k[i] = ok ? k[i] : CUDART_INF_F;
            // counters need warp-wide ballots: done by every lane, active or not
            valid.add(i, ok, lane);
        }

        sort_stack<NB>(k);
        const int ih = n >> 1;
        const int il = ih - 1 + (n & 1);
        float xh = 0.f, xl = 0.f;
#pragma unroll
        for (int p = 0; p < NB; p++) {
            xh = (p == ih) ? k[p] : xh;
            xl = (p == il) ? k[p] : xl;
        }
        if (act) {
            double med = (n & 1) ? (double)xh : ((double)xh + (double)xl) * 0.5;
            if (n == 0) med = CUDART_NAN;
            A.out[v] = med;
            A.expmap[v] = n;
        }
    }
    valid.flush(A.valid_cnt, A.nfiles, lane);
}

// ---------------------------------------------------------------------------
// register-resident sigma clip  (merging.c:423-477 + tools.c:82-118)
//
// UNIT = every scale is 1 and every offset 0: w == (double)d exactly, so the
// sort keys are the float32 samples themselves and every comparison against a
// float64 bound is done exactly in float32 with the bound rounded towards the
// safe side (k > c  <=>  k > rd(c);  k < c  <=>  k < ru(c) for float k).
// Otherwise the keys are the float64 values w = (offset + d) * scale.
// ---------------------------------------------------------------------------
template <bool UNIT> struct KeyOf { using type = double; };
template <> struct KeyOf<true> { using type = float; };

template <bool UNIT>
struct Bound {  // a strict clip window in key space
    typename KeyOf<UNIT>::type lo, up;
    __device__ __forceinline__ Bound(double clo, double cup)
    {
        if constexpr (UNIT) {
            lo = __double2float_rd(clo);
            up = __double2float_ru(cup);
        } else {
            lo = clo;
            up = cup;
        }
    }
};

template <int NB, bool UNIT>
__global__ void __launch_bounds__(kThreads) sigclip_reg_kernel(const StackArgs<NB> S, const ClipArgs A)
{
    using Key = typename KeyOf<UNIT>::type;
    const int lane = threadIdx.x & 31;
    LaneCounters<NB> valid, select;
    const int nstop = A.nstop < 1 ? 1 : A.nstop;
    const long long stride = (long long)gridDim.x * kThreads;

    for (long long base = (long long)blockIdx.x * kThreads; base < A.nvox; base += stride) {
        const long long v = base + threadIdx.x;
        const bool act = v < A.nvox;

        // ---- gather in exposure order (merging.c:425-436)
        float d[NB];
#pragma unroll
        for (int i = 0; i < NB; i++)
            d[i] = (act && i < A.nfiles) ? ld_stream((const float*)S.data[i] + v) : CUDART_NAN_F;

        Key ws[NB];
        int n = 0;
        double sum = 0.0;
#pragma unroll
        for (int i = 0; i < NB; i++) {
            const bool ok = (d[i] == d[i]);
            double w;
            if constexpr (UNIT) w = (double)d[i];
            else w = (S.offset[i] + (double)d[i]) * S.scale[i];
            n += ok;
            if (ok) sum += w;                                   // tools.c:13-16, exposure order
            if constexpr (UNIT) ws[i] = ok ? d[i] : CUDART_INF_F;
            else ws[i] = ok ? w : CUDART_INF;
        }
        double mean = sum / (double)n;                          // n == 0 -> NaN (merging.c:439)
        double dev = 0.0;
#pragma unroll
        for (int i = 0; i < NB; i++) {
            const bool ok = (d[i] == d[i]);
            const double t = (double)ws[i] - mean;
            if (ok) dev += t * t;                               // tools.c:18-21 (no FMA)
        }
        double sigma = sqrt(dev / (double)n);

        // ---- order the stack once; later sets are windows [a,b) of it
        sort_stack<NB>(ws);
        int a = 0, b = n;
        Key keepLo, keepUp;                                     // strict bounds of the kept set
        if constexpr (UNIT) { keepLo = -CUDART_INF_F; keepUp = CUDART_INF_F; }
        else { keepLo = -CUDART_INF; keepUp = CUDART_INF; }

        if (n >= 2) {
            int it = A.nmax;
            for (;;) {
                const int m = b - a;
                const int ih = a + (m >> 1);
                const int il = ih - 1 + (m & 1);
                Key xh = 0, xl = 0;
#pragma unroll
                for (int p = 0; p < NB; p++) {
                    xh = (p == ih) ? ws[p] : xh;
                    xl = (p == il) ? ws[p] : xl;
                }
                const double med = (m & 1) ? (double)xh : ((double)xh + (double)xl) * 0.5;
                const double clo = med - (A.nclip_low * sigma);  // tools.c:89-90
                const double cup = med + (A.nclip_up * sigma);
                const Bound<UNIT> bd(clo, cup);
                int ni = 0, nle = 0;
#pragma unroll
                for (int p = 0; p < NB; p++) {
                    const bool inw = (p >= a) && (p < b);
                    const bool gt = ws[p] > bd.lo;
                    const bool lt = ws[p] < bd.up;
                    ni += (inw && gt && lt);
                    nle += (inw && !gt);
                }
                if (ni < nstop || ni == m) break;                // tools.c:100-103
                if (it <= 0) break;                              // tools.c:104
                it--;
                a += nle;
                b = a + ni;
                keepLo = bd.lo > keepLo ? bd.lo : keepLo;
                keepUp = bd.up < keepUp ? bd.up : keepUp;
                // statistics of the survivors, ascending order (tools.c:116 -> :9-24)
                double s1 = 0.0;
#pragma unroll
                for (int p = 0; p < NB; p++)
                    if (p >= a && p < b) s1 += (double)ws[p];
                mean = s1 / (double)ni;
                double s2 = 0.0;
#pragma unroll
                for (int p = 0; p < NB; p++) {
                    const double t = (double)ws[p] - mean;
                    if (p >= a && p < b) s2 += t * t;
                }
                sigma = sqrt(s2 / (double)ni);
            }
        }
        const int kept = b - a;

        // ---- kept set per exposure, propagated variance, counters
        double vsum = 0.0;
#pragma unroll
        for (int i = 0; i < NB; i++) {
            const bool ok = (d[i] == d[i]);
            bool in;
            if constexpr (UNIT) {
                in = ok && (d[i] > keepLo) && (d[i] < keepUp);
            } else {
                const double w = (S.offset[i] + (double)d[i]) * S.scale[i];
                in = ok && (w > keepLo) && (w < keepUp);
            }
            if (A.typ_var == 0 && i < A.nfiles) {
                const float sv = act ? ld_stream((const float*)S.var[i] + v) : 0.f;
                if (in) vsum += (double)sv * S.scale2[i];       // merging.c:431,462
            }
            valid.add(i, ok, lane);
            select.add(i, in, lane);
        }

        if (act) {
            double ov;
            if (A.typ_var == 0) {
                ov = vsum / ((double)kept * (double)kept);       // merging.c:462
            } else if (kept > 1) {
                ov = sigma * sigma;                              // merging.c:465
                if (A.typ_var == 1) ov /= (double)(kept - 1);    // merging.c:467
            } else {
                ov = CUDART_NAN;
            }
            A.out[v] = mean;
            A.outvar[v] = ov;
            A.expmap[v] = kept;
        }
    }
    valid.flush(A.valid_cnt, A.nfiles, lane);
    select.flush(A.select_cnt, A.nfiles, lane);
}

// ---------------------------------------------------------------------------
// generic kernels: any depth <= kGenericMax, float32 or float64 samples, MAD.
// Per-thread arrays in local memory; the reference loops restated directly.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void shell_argsort(const double* w, unsigned short* idx, int n)
{
    // ascending; ties keep no particular order (equal keys are equal values)
    const int gaps[6] = {301, 132, 57, 23, 10, 4};
    for (int g = 0; g < 7; g++) {
        const int gap = g < 6 ? gaps[g] : 1;
        if (gap >= n && gap != 1) continue;
        for (int j = gap; j < n; j++) {
            const unsigned short t = idx[j];
            const double key = w[t];
            int i = j - gap;
            while (i >= 0 && w[idx[i]] > key) {
                idx[i + gap] = idx[i];
                i -= gap;
            }
            idx[i + gap] = t;
        }
    }
}

__device__ __forceinline__ double middle_of(const double* w, const unsigned short* idx, int n)
{
    if ((n & 1) == 0) return (w[idx[n >> 1]] + w[idx[(n >> 1) - 1]]) * 0.5;
    return w[idx[n >> 1]];
}

template <int MAXN>
struct GenericScratch {
    double w[MAXN];
    double work[MAXN];
    unsigned short idx[MAXN];
    unsigned short widx[MAXN];
    unsigned short file[MAXN];
};

// tools.c:82-118 / :122-158 as a loop; returns kept count, idx[0..kept) = survivors
template <int MAXN>
__device__ int clip_generic(GenericScratch<MAXN>& g, int n, int nmax, double lo, double up,
                            int nstop, int mad, double& mean_out, double& sigma_out)
{
    double mean = 0.0, sigma = 0.0;
    for (;;) {
        double acc = 0.0;
        for (int i = 0; i < n; i++) acc += g.w[g.idx[i]];
        mean = acc / (double)n;
        double med;
        if (!mad) {
            double dv = 0.0;
            for (int i = 0; i < n; i++) {
                const double t = g.w[g.idx[i]] - mean;
                dv += t * t;
            }
            sigma = sqrt(dv / (double)n);
            shell_argsort(g.w, g.idx, n);
            med = middle_of(g.w, g.idx, n);
        } else {
            shell_argsort(g.w, g.idx, n);
            med = middle_of(g.w, g.idx, n);
            for (int i = 0; i < n; i++) {
                g.widx[i] = (unsigned short)i;
                g.work[i] = fabs(g.w[g.idx[i]] - med);
            }
            shell_argsort(g.work, g.widx, n);
            sigma = middle_of(g.work, g.widx, n) * 1.4826;
        }
        const double clo = med - (lo * sigma);
        const double cup = med + (up * sigma);
        int ni = 0;
        for (int i = 0; i < n; i++) {
            const double x = g.w[g.idx[i]];
            ni += (x < cup) && (x > clo);
        }
        if (ni < nstop || ni == n) break;
        if (nmax <= 0) break;
        int k = 0;
        for (int i = 0; i < n; i++) {
            const double x = g.w[g.idx[i]];
            if ((x < cup) && (x > clo)) g.idx[k++] = g.idx[i];
        }
        n = k;
        nmax--;
    }
    mean_out = mean;
    sigma_out = sigma;
    return n;
}

template <int MAXN, typename ElemT>
__global__ void __launch_bounds__(kThreads)
sigclip_generic_kernel(const StackTable* __restrict__ T, const ClipArgs A)
{
    __shared__ unsigned s_valid[MAXN];
    __shared__ unsigned s_select[MAXN];
    for (int i = threadIdx.x; i < MAXN; i += kThreads) { s_valid[i] = 0; s_select[i] = 0; }
    __syncthreads();
    const int nstop = A.nstop < 1 ? 1 : A.nstop;
    GenericScratch<MAXN> g;
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < A.nvox; v += stride) {
        int n = 0;
        for (int i = 0; i < A.nfiles; i++) {
            const double dd = (double)ld_stream((const ElemT*)T->data[i] + v);
            if (dd == dd) {
                g.w[n] = (T->offset[i] + dd) * T->scale[i];
                g.file[n] = (unsigned short)i;
                g.idx[n] = (unsigned short)n;
                n++;
                atomicAdd(&s_valid[i], 1u);
            }
        }
        double mean = CUDART_NAN, sigma = 0.0, ov = CUDART_NAN;
        int kept = n;
        if (n == 1) {
            mean = g.w[0];
        } else if (n >= 2) {
            kept = clip_generic<MAXN>(g, n, A.nmax, A.nclip_low, A.nclip_up, nstop, A.mad, mean, sigma);
        }
        if (n >= 1) {
            if (A.typ_var == 0) {
                double vs = 0.0;
                for (int k = 0; k < kept; k++) {
                    const int i = g.file[g.idx[k]];
                    vs += (double)ld_stream((const ElemT*)T->var[i] + v) * T->scale2[i];
                }
                ov = vs / ((double)kept * (double)kept);
            } else if (n >= 2 && kept > 1) {
                ov = sigma * sigma;
                if (A.typ_var == 1) ov /= (double)(kept - 1);
            }
            for (int k = 0; k < kept; k++) atomicAdd(&s_select[g.file[g.idx[k]]], 1u);
        }
        A.out[v] = mean;
        A.outvar[v] = ov;
        A.expmap[v] = kept;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < A.nfiles; i += kThreads) {
        if (s_valid[i]) atomicAdd(A.valid_cnt + i, (unsigned long long)s_valid[i]);
        if (s_select[i]) atomicAdd(A.select_cnt + i, (unsigned long long)s_select[i]);
    }
}

template <int MAXN, typename ElemT>
__global__ void __launch_bounds__(kThreads)
median_generic_kernel(const StackTable* __restrict__ T, const ClipArgs A)
{
    __shared__ unsigned s_valid[MAXN];
    for (int i = threadIdx.x; i < MAXN; i += kThreads) s_valid[i] = 0;
    __syncthreads();
    double w[MAXN];
    unsigned short idx[MAXN];
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long v = (long long)blockIdx.x * kThreads + threadIdx.x; v < A.nvox; v += stride) {
        int n = 0;
        for (int i = 0; i < A.nfiles; i++) {
            const double dd = (double)ld_stream((const ElemT*)T->data[i] + v);
            if (dd == dd) {
                w[n] = dd;
                idx[n] = (unsigned short)n;
                n++;
                atomicAdd(&s_valid[i], 1u);
            }
        }
        double med = CUDART_NAN;
        if (n == 1) med = w[0];
        else if (n >= 2) {
            shell_argsort(w, idx, n);
            med = middle_of(w, idx, n);
        }
        A.out[v] = med;
        A.expmap[v] = n;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < A.nfiles; i += kThreads)
        if (s_valid[i]) atomicAdd(A.valid_cnt + i, (unsigned long long)s_valid[i]);
}

// Histogram of the exposure map (how many voxels were built from 0, 1, ... N exposures): what
// CubeList._compute_expnb (cubelist.py:69-74) needs, without a host pass over the whole map.
static __global__ void expmap_hist_kernel(const int* __restrict__ expmap, long long n, unsigned long long* __restrict__ hist)
{
    __shared__ unsigned s_hist[kGenericMax + 1];
    for (int i = threadIdx.x; i <= kGenericMax; i += blockDim.x) s_hist[i] = 0u;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const int e = __ldcs(expmap + i);
        // most voxels of a warp share one value: aggregate before touching shared memory
        const unsigned peers = __match_any_sync(__activemask(), e);
        if ((int)(__ffs(peers) - 1) == (int)(threadIdx.x & 31) && e >= 0 && e <= kGenericMax)
            atomicAdd(&s_hist[e], (unsigned)__popc(peers));
    }
    __syncthreads();
    for (int i = threadIdx.x; i <= kGenericMax; i += blockDim.x)
        if (s_hist[i]) atomicAdd(hist + i, (unsigned long long)s_hist[i]);
}

// In-place byte swap of big-endian float32 samples staged from a FITS HDU.
static __global__ void bswap32_kernel(uint32_t* __restrict__ p, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        p[i] = __byte_perm(p[i], 0, 0x0123);
}

static __global__ void bswap64_kernel(unsigned long long* __restrict__ p, long long n)
{
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long x = p[i];
        const unsigned lo = __byte_perm((unsigned)x, 0, 0x0123);
        const unsigned hi = __byte_perm((unsigned)(x >> 32), 0, 0x0123);
        p[i] = ((unsigned long long)lo << 32) | hi;
    }
}

}  // namespace mb
