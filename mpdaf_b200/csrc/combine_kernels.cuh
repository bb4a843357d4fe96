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

// Mask of the voxels that hold no finite value (what the reference builds on the host when it wraps the combined
// arrays in a Cube, lib/mpdaf/obj/cubelist.py:399-400 -> data.py mask = ~isfinite): one byte per voxel, 1 = masked.
// `var` may be null; when given, a voxel with a non-finite variance is NOT masked (the reference masks on data only).
static __global__ void finite_mask_kernel(const double* __restrict__ data, long long n, unsigned char* __restrict__ mask,
                                          unsigned long long* __restrict__ nmasked)
{
    const long long stride = (long long)gridDim.x * blockDim.x * 4;
    unsigned local = 0;
    for (long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
        unsigned packed = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const long long j = i + k;
            const unsigned long long bits = j < n ? (unsigned long long)__double_as_longlong(__ldcs(data + j)) : 0ull;
            const unsigned bad = ((bits >> 52) & 0x7FFull) == 0x7FFull ? 1u : 0u;   // exponent all ones: inf or NaN
            packed |= bad << (8 * k);
            local += bad;
        }
        if (i + 3 < n && (reinterpret_cast<uintptr_t>(mask + i) & 3u) == 0) *reinterpret_cast<unsigned*>(mask + i) = packed;
        else for (int k = 0; k < 4 && i + k < n; k++) mask[i + k] = (unsigned char)((packed >> (8 * k)) & 1u);
    }
    local = __reduce_add_sync(0xffffffffu, local);
    if ((threadIdx.x & 31) == 0 && local && nmasked) atomicAdd(nmasked, (unsigned long long)local);
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
