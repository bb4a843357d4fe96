// How fast can the 48-key register sorting network of the warp kernel run on its own?  (sm_100a)
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../mpdaf_b200/csrc -o sort48 sort48.cu
// Prints SM cycles per sorted warp-column set for several block shapes and the three compare-exchange forms:
//   alu : FMNMX + IADD3 (a + b - min)         -- the form the kernel uses
//   mnmx: FMNMX + FMNMX (min, max)
//   fma : FMNMX + 2 IMAD
#include <cstdio>
#include <cuda_runtime.h>

#include "sigclip_warp.cuh"

using namespace mb;

template <int MODE>
__device__ __forceinline__ void ce(float& a, float& b, unsigned one, unsigned minus_one)
{
    if constexpr (MODE == 0) key_ce_alu(a, b);
    else if constexpr (MODE == 1) { const float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }
    else key_ce_fma(a, b, one, minus_one);
}

template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS, 768 / THREADS) sort_bench(float* out, int rounds, unsigned seed, unsigned long long* cyc)
{
    float key[48];
    const unsigned one = seed ? 1u : 0u, minus_one = 0u - one;
    const unsigned tid = blockIdx.x * THREADS + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 48; i++) key[i] = __uint_as_float((((tid * 2654435761u) ^ (i * 40503u + seed)) & 0x3FFFFFC0u) | (unsigned)i | 0x3F000000u);
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; r++) {
#define MB_CE(i, j) ce<MODE>(key[i], key[j], one, minus_one);
        MB_SORTNET_48(MB_CE)
#undef MB_CE
        // scramble again (cheap, order-destroying): rotate the index bits into the value bits of a few keys
#pragma unroll
        for (int i = 0; i < 48; i += 2) {
            const float t = key[i];
            key[i] = key[47 - i];
            key[47 - i] = t;
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 48; i++) acc += key[i];
    if (acc == 1.2345f) out[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = (unsigned long long)(t1 - t0);
}

template <int MODE, int THREADS>
void run(const char* name, int sms)
{
    float* out;
    unsigned long long* cyc;
    cudaMalloc(&out, 4);
    cudaMalloc(&cyc, 8);
    const int rounds = 200;
    const int grid = sms * (768 / THREADS);
    sort_bench<MODE, THREADS><<<grid, THREADS>>>(out, 10, 1u, cyc);
    cudaDeviceSynchronize();
    sort_bench<MODE, THREADS><<<grid, THREADS>>>(out, rounds, 1u, cyc);
    cudaDeviceSynchronize();
    unsigned long long c = 0;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // 24 warps per SM = 6 per scheduler: cycles per sort per scheduler = c / rounds / 6
    printf("%-5s threads %4d: %8.1f cycles per round of 6 warps/scheduler -> %6.1f cycles per sorted column set, %5.2f cycles per compare-exchange  (%s)\n",
           name, THREADS, (double)c / rounds, (double)c / rounds / 6.0, (double)c / rounds / 6.0 / 367.0, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    run<0, 768>("alu", p.multiProcessorCount);
    run<1, 768>("mnmx", p.multiProcessorCount);
    run<2, 768>("fma", p.multiProcessorCount);
    run<0, 128>("alu", p.multiProcessorCount);
    run<1, 128>("mnmx", p.multiProcessorCount);
    return 0;
}
