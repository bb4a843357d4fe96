// Does prefetch.global.L1 make a later ld.global(.cs) an L1 hit on B200?  One warp, one line at a time.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1prefetch l1prefetch.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ long long clk() { long long c; asm volatile("mov.u64 %0, %%clock64;" : "=l"(c)); return c; }

template <int MODE>   // 0: none, 1: prefetch L1, 2: prefetch L2, 3: ld.ca into a register, 4: cp.async.ca to a dummy smem word, 5: ld.ca, result unused
__global__ void probe(const float* buf, long long stride, int delay, int reps, long long* out, float* sink)
{
    long long total = 0;
    float acc = 0.f;
    for (int r = 0; r < reps; r++) {
        const float* p = buf + (long long)r * stride + threadIdx.x;
        if (MODE == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
        if (MODE == 2) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        if (MODE == 4) {
            __shared__ float dummy[32];
            const unsigned sa = (unsigned)__cvta_generic_to_shared(&dummy[threadIdx.x]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(p));
        }
        if (MODE == 5) { float t; asm volatile("ld.global.ca.f32 %0, [%1];" : "=f"(t) : "l"(p)); }
        if (MODE == 3) { float t; asm volatile("ld.global.ca.f32 %0, [%1];" : "=f"(t) : "l"(p)); acc += t; }
        const long long t0 = clk();
        while (clk() - t0 < delay) { }
        __syncwarp();
        const long long a = clk();
        float v;
        asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p));
        acc += v;
        asm volatile("" ::"f"(acc));
        const long long b = clk();
        total += b - a;
    }
    if (threadIdx.x == 0) out[0] = total / reps;
    sink[threadIdx.x] = acc;
}

int main()
{
    const long long n = 1ll << 28;   // 1 GiB of floats: every probed line is cold in L2
    float* buf; long long* out; float* sink;
    cudaMalloc(&buf, n * 4); cudaMemset(buf, 0, n * 4);
    cudaMalloc(&out, 8); cudaMalloc(&sink, 128);
    const int reps = 200;
    const long long stride = 1 << 18;   // 1 MiB apart
    for (int delay : {0, 2000, 20000}) {
        long long h[6];
        probe<0><<<1, 32>>>(buf, stride, delay, reps, out, sink); cudaMemcpy(&h[0], out, 8, cudaMemcpyDeviceToHost);
        cudaMemset(buf, 0, n * 4);
        probe<1><<<1, 32>>>(buf, stride, delay, reps, out, sink); cudaMemcpy(&h[1], out, 8, cudaMemcpyDeviceToHost);
        cudaMemset(buf, 0, n * 4);
        probe<2><<<1, 32>>>(buf, stride, delay, reps, out, sink); cudaMemcpy(&h[2], out, 8, cudaMemcpyDeviceToHost);
        cudaMemset(buf, 0, n * 4);
        probe<3><<<1, 32>>>(buf, stride, delay, reps, out, sink); cudaMemcpy(&h[3], out, 8, cudaMemcpyDeviceToHost);
        cudaMemset(buf, 0, n * 4);
        probe<4><<<1, 32>>>(buf, stride, delay, reps, out, sink); cudaMemcpy(&h[4], out, 8, cudaMemcpyDeviceToHost);
        cudaMemset(buf, 0, n * 4);
        probe<5><<<1, 32>>>(buf, stride, delay, reps, out, sink); cudaMemcpy(&h[5], out, 8, cudaMemcpyDeviceToHost);
        cudaMemset(buf, 0, n * 4);
        printf("delay %6d clk: load latency  none %lld  prefetch.L1 %lld  prefetch.L2 %lld  ld.ca %lld  cp.async.ca %lld  ld.ca(unused) %lld  (%s)\n", delay, h[0], h[1], h[2],
               h[3], h[4], h[5], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
