// Round-2 design probe (sm_100a): how fast is the 48-key sorting network when TWO voxels share a lane and their
// 16-bit keys are packed in one register (one packed min + one packed max per compare-exchange PAIR), and which
// pipes do the candidate instructions occupy?  Not part of the product.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../mpdaf_b200/csrc -o packsort packsort.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "sort_networks.cuh"

// ---- compare-exchange forms --------------------------------------------------------------------
// 0: f32 keys, FMNMX + IADD3 (round-1 form)            1: f32 integer-valued keys, FMNMX + FADD + FADD
// 2: bf16x2, HMNMX2.BF16 min + max                      3: f16x2, HMNMX2 min + max
// 4: u16x2, VIMNMX.U16x2 min + max                      5: VIMNMX.U16x2 min + HMNMX2 (f16) max
// 6: f32 FMNMX min + FMNMX max
template <int MODE>
__device__ __forceinline__ void ce(unsigned& a, unsigned& b)
{
    unsigned lo, hi;
    if constexpr (MODE == 0) {
        asm("min.f32 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b));
        hi = a + b - lo;
    } else if constexpr (MODE == 1) {
        unsigned s;
        asm("min.f32 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b));
        asm("add.rn.f32 %0, %1, %2;" : "=r"(s) : "r"(a), "r"(b));
        asm("sub.rn.f32 %0, %1, %2;" : "=r"(hi) : "r"(s), "r"(lo));
    } else if constexpr (MODE == 2) {
        asm("min.bf16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b));
        asm("max.bf16x2 %0, %1, %2;" : "=r"(hi) : "r"(a), "r"(b));
    } else if constexpr (MODE == 3) {
        asm("min.f16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b));
        asm("max.f16x2 %0, %1, %2;" : "=r"(hi) : "r"(a), "r"(b));
    } else if constexpr (MODE == 4) {
        asm("min.u16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b));
        asm("max.u16x2 %0, %1, %2;" : "=r"(hi) : "r"(a), "r"(b));
    } else if constexpr (MODE == 5) {
        asm("min.u16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b));
        asm("max.f16x2 %0, %1, %2;" : "=r"(hi) : "r"(a), "r"(b));
    } else if constexpr (MODE == 7) {          // packed min + three-input XOR: max = a ^ b ^ min (bit patterns)
        asm("min.bf16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b));
        asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(hi) : "r"(a), "r"(b), "r"(lo));
    } else if constexpr (MODE == 8) {          // packed min + integer a + b - min (ptxas fuses the two adds into one IADD3)
        unsigned t;
        asm("min.bf16x2 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b));
        asm volatile("sub.u32 %0, %1, %2;" : "=r"(t) : "r"(a), "r"(lo));
        asm volatile("add.u32 %0, %1, %2;" : "=r"(hi) : "r"(t), "r"(b));
    } else {
        asm("min.f32 %0, %1, %2;" : "=r"(lo) : "r"(a), "r"(b));
        asm("max.f32 %0, %1, %2;" : "=r"(hi) : "r"(a), "r"(b));
    }
    a = lo;
    b = hi;
}

template <int MODE, int THREADS, int BLOCKS>
__global__ void __launch_bounds__(THREADS, BLOCKS) sort_bench(unsigned* out, int rounds, unsigned seed, unsigned long long* cyc)
{
    unsigned key[48];
    const unsigned tid = blockIdx.x * THREADS + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 48; i++) {
        const unsigned h = (tid * 2654435761u) ^ (i * 40503u + seed);
        if (MODE == 0 || MODE == 6) key[i] = (h & 0x3FFFFFC0u) | (unsigned)i | 0x3F000000u;
        else if (MODE == 1) key[i] = __float_as_uint((float)((h & 0x1FFFFu) * 64u + i));
        else key[i] = (h & 0x3BFF3BFFu) | 0x04000400u;   // two positive normal f16 / bf16 patterns
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int r = 0; r < rounds; r++) {
#define MB_CE(i, j) ce<MODE>(key[i], key[j]);
        MB_SORTNET_48(MB_CE)
#undef MB_CE
#pragma unroll
        for (int i = 0; i < 48; i += 2) {
            const unsigned t = key[i];
            key[i] = key[47 - i];
            key[47 - i] = t;
        }
    }
    const long long t1 = clock64();
    unsigned acc = 0;
#pragma unroll
    for (int i = 0; i < 48; i++) acc += key[i];
    if (acc == 12345u) out[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = (unsigned long long)(t1 - t0);
}

template <int MODE, int THREADS, int BLOCKS>
void run_sort(const char* name, int sms, int voxels_per_lane)
{
    unsigned* out;
    unsigned long long* cyc;
    cudaMalloc(&out, 4);
    cudaMalloc(&cyc, 8);
    const int rounds = 200;
    const int grid = sms * BLOCKS;
    sort_bench<MODE, THREADS, BLOCKS><<<grid, THREADS>>>(out, 10, 1u, cyc);
    cudaDeviceSynchronize();
    sort_bench<MODE, THREADS, BLOCKS><<<grid, THREADS>>>(out, rounds, 1u, cyc);
    cudaDeviceSynchronize();
    unsigned long long c = 0;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    const double wps = THREADS * BLOCKS / 32 / 4.0;   // warps per scheduler
    const double per_sort = (double)c / rounds / wps;   // scheduler cycles per warp-sort
    printf("%-22s %4d thr x %d: %7.1f cyc per warp-sort, %5.2f cyc per CE instr-group, %6.1f cyc per 32 voxels  (%s)\n", name, THREADS, BLOCKS,
           per_sort, per_sort / 367.0, per_sort / voxels_per_lane, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
    cudaFree(cyc);
}

// ---- pipe probes -------------------------------------------------------------------------------
constexpr int CH = 8;
constexpr int IT = 2048;
constexpr int TPB = 256;

#define KERNEL(name, DECL, INIT, BODY, SINK)                                              \
    __global__ void name(unsigned long long* cyc, float* out, float seedf, double seedd, int seedi) \
    {                                                                                     \
        unsigned smid;                                                                    \
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));                                 \
        DECL;                                                                             \
        _Pragma("unroll") for (int c = 0; c < CH; c++) { INIT; }                        \
        __syncthreads();                                                                  \
        long long t0 = clock64();                                                         \
        for (int it = 0; it < IT; it++) {                                                 \
            _Pragma("unroll") for (int c = 0; c < CH; c++) { BODY; }                    \
        }                                                                                 \
        long long t1 = clock64();                                                         \
        float acc = 0;                                                                    \
        _Pragma("unroll") for (int c = 0; c < CH; c++) { SINK; }                        \
        if (acc == 123.456f) out[0] = acc;                                                \
        if ((threadIdx.x & 31) == 0) {                                                    \
            atomicMin(&cyc[smid], (unsigned long long)t0);                                \
            atomicMax(&cyc[256 + smid], (unsigned long long)t1);                          \
            atomicAdd(&cyc[512 + smid], 1ull);                                            \
        }                                                                                 \
    }

KERNEL(k_hmnmx2, unsigned a[CH]; unsigned b = seedi * 77; unsigned d = seedi * 1031,
       a[c] = seedi + c * 131 + threadIdx.x * 7, asm volatile("min.f16x2 %0, %0, %1; max.f16x2 %0, %0, %2;" : "+r"(a[c]) : "r"(b), "r"(d)), acc += a[c])   // 2 HMNMX2
KERNEL(k_hmnmx2bf, unsigned a[CH]; unsigned b = seedi * 77; unsigned d = seedi * 1031,
       a[c] = seedi + c * 131 + threadIdx.x * 7, asm volatile("min.bf16x2 %0, %0, %1; max.bf16x2 %0, %0, %2;" : "+r"(a[c]) : "r"(b), "r"(d)), acc += a[c]) // 2 HMNMX2.BF16
KERNEL(k_vimnmx2, unsigned a[CH]; unsigned b = seedi * 77; unsigned d = seedi * 1031,
       a[c] = seedi + c * 131 + threadIdx.x * 7, asm volatile("min.u16x2 %0, %0, %1; max.u16x2 %0, %0, %2;" : "+r"(a[c]) : "r"(b), "r"(d)), acc += a[c])   // 2 VIMNMX.U16x2
KERNEL(k_hmn_vim, unsigned a[CH]; unsigned b = seedi * 77; unsigned d = seedi * 1031,
       a[c] = seedi + c * 131 + threadIdx.x * 7, asm volatile("min.u16x2 %0, %0, %1; max.f16x2 %0, %0, %2;" : "+r"(a[c]) : "r"(b), "r"(d)), acc += a[c])   // VIMNMX + HMNMX2
KERNEL(k_hmn_fmn, unsigned a[CH]; float f[CH]; unsigned b = seedi * 77; float d = seedf * 3,
       a[c] = seedi + c * 131 + threadIdx.x * 7; f[c] = seedf + c, asm volatile("min.f16x2 %0, %0, %2; min.f32 %1, %1, %3;" : "+r"(a[c]), "+f"(f[c]) : "r"(b), "f"(d)), acc += a[c] + f[c])   // HMNMX2 + FMNMX
KERNEL(k_hmn_ffma, unsigned a[CH]; float f[CH]; unsigned b = seedi * 77; float d = seedf * 3,
       a[c] = seedi + c * 131 + threadIdx.x * 7; f[c] = seedf + c, asm volatile("min.f16x2 %0, %0, %2; fma.rn.f32 %1, %1, %3, %3;" : "+r"(a[c]), "+f"(f[c]) : "r"(b), "f"(d)), acc += a[c] + f[c])   // HMNMX2 + FFMA
KERNEL(k_hmn_dfma, unsigned a[CH]; double f[CH]; unsigned b = seedi * 77; double d = seedd * 3,
       a[c] = seedi + c * 131 + threadIdx.x * 7; f[c] = seedd + c, asm volatile("min.f16x2 %0, %0, %2; fma.rn.f64 %1, %1, %3, %3;" : "+r"(a[c]), "+d"(f[c]) : "r"(b), "d"(d)), acc += a[c] + (float)f[c])   // HMNMX2 + DFMA
KERNEL(k_vim_ffma, unsigned a[CH]; float f[CH]; unsigned b = seedi * 77; float d = seedf * 3,
       a[c] = seedi + c * 131 + threadIdx.x * 7; f[c] = seedf + c, asm volatile("min.u16x2 %0, %0, %2; fma.rn.f32 %1, %1, %3, %3;" : "+r"(a[c]), "+f"(f[c]) : "r"(b), "f"(d)), acc += a[c] + f[c])   // VIMNMX + FFMA
KERNEL(k_fmn_ffma, float a[CH]; float f[CH]; float b = seedf * 7; float d = seedf * 3,
       a[c] = seedf + c * 131; f[c] = seedf + c, asm volatile("min.f32 %0, %0, %2; fma.rn.f32 %1, %1, %3, %3;" : "+f"(a[c]), "+f"(f[c]) : "f"(b), "f"(d)), acc += a[c] + f[c])   // FMNMX + FFMA
KERNEL(k_fmn_ffma2, float a[CH]; float f[CH]; float g[CH]; float b = seedf * 7; float d = seedf * 3,
       a[c] = seedf + c * 131; f[c] = seedf + c; g[c] = seedf - c, asm volatile("min.f32 %0, %0, %3; fma.rn.f32 %1, %1, %4, %4; fma.rn.f32 %2, %2, %4, %4;" : "+f"(a[c]), "+f"(f[c]), "+f"(g[c]) : "f"(b), "f"(d)), acc += a[c] + f[c] + g[c])   // FMNMX + 2 FFMA
KERNEL(k_f2fp, float a[CH]; unsigned r[CH]; float b = seedf,
       a[c] = seedf + c; r[c] = c, asm volatile("{.reg .b32 t; cvt.rn.bf16x2.f32 t, %1, %2; add.u32 %0, %0, t;}" : "+r"(r[c]) : "f"(a[c]), "f"(b)), acc += r[c])   // F2FP.BF16 + IADD
KERNEL(k_hsetp2, unsigned a[CH]; unsigned r[CH]; unsigned b = seedi * 77,
       a[c] = seedi + c * 131 + threadIdx.x * 7; r[c] = c, asm volatile("{.reg .pred p, q; setp.ge.bf16x2 p|q, %1, %2; @p add.u32 %0, %0, 1; @q add.u32 %0, %0, 2;}" : "+r"(r[c]) : "r"(a[c]), "r"(b)), acc += r[c])   // HSETP2 + 2 predicated VIADD
KERNEL(k_prmt, unsigned a[CH]; unsigned b = seedi * 77,
       a[c] = seedi + c * 131 + threadIdx.x * 7, asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(a[c]) : "r"(b)), acc += a[c])   // PRMT
KERNEL(k_f2f_up, float a[CH]; double d[CH],
       a[c] = seedf + c; d[c] = 0, asm volatile("{.reg .f64 t; cvt.f64.f32 t, %1; add.rn.f64 %0, %0, t;}" : "+d"(d[c]) : "f"(a[c])), acc += (float)d[c])   // F2F.F64.F32 + DADD
KERNEL(k_f2f_up2, float a[CH]; double d[CH]; double e[CH],
       a[c] = seedf + c; d[c] = 0; e[c] = 1, asm volatile("{.reg .f64 t; cvt.f64.f32 t, %2; add.rn.f64 %0, %0, t; fma.rn.f64 %1, t, t, %1; add.rn.f32 %2, %2, %2;}" : "+d"(d[c]), "+d"(e[c]), "+f"(a[c])), acc += (float)(d[c] + e[c]))   // F2F + DADD + DFMA + FADD
KERNEL(k_pass1, float a[CH]; double d[CH]; double e[CH]; double sc = seedd; double of = seedd * 0.5; double p = seedd * 3,
       a[c] = seedf + c; d[c] = 0; e[c] = 1,
       asm volatile("{.reg .f64 t; .reg .pred q; cvt.f64.f32 t, %2; add.rn.f64 t, t, %3; mul.rn.f64 t, t, %4; sub.rn.f64 t, t, %5; setp.eq.f32 q, %2, %2; @q add.rn.f64 %0, %0, t; @q fma.rn.f64 %1, t, t, %1; add.rn.f32 %2, %2, %2;}"
                    : "+d"(d[c]), "+d"(e[c]), "+f"(a[c]) : "d"(of), "d"(sc), "d"(p)), acc += (float)(d[c] + e[c]))   // pass-1 body: F2F DADD DMUL DADD FSETP DADD DFMA (+FADD)
KERNEL(k_lds64, int a[CH]; __shared__ int sm[TPB * 8]; sm[threadIdx.x] = seedi; sm[threadIdx.x + TPB] = 0,
       a[c] = threadIdx.x * 8, asm volatile("{.reg .b32 t, u; ld.shared.v2.b32 {t, u}, [%0]; add.s32 %0, %0, u;}" : "+r"(a[c])), acc += a[c])   // LDS.64 + IADD (dependent)
KERNEL(k_viaddp, unsigned a[CH]; float f[CH]; float b = seedf * 3,
       a[c] = seedi + c + threadIdx.x; f[c] = seedf + c, asm volatile("{.reg .pred q; setp.eq.f32 q, %1, %1; @q add.u32 %0, %0, 256; add.rn.f32 %1, %1, %2;}" : "+r"(a[c]), "+f"(f[c]) : "f"(b)), acc += a[c] + f[c])   // FSETP + predicated VIADD + FADD

template <typename K>
void run(const char* name, K kernel, double ops_per_body)
{
    int sms = 0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    sms = prop.multiProcessorCount;
    const int grid = sms * 4;
    unsigned long long* cyc;
    float* out;
    cudaMalloc(&cyc, 768 * sizeof(unsigned long long));
    cudaMalloc(&out, 4);
    unsigned long long init[768];
    for (int i = 0; i < 768; i++) init[i] = i < 256 ? ~0ull : 0ull;
    for (int rep = 0; rep < 2; rep++) kernel<<<grid, TPB>>>(cyc, out, 1.5f, 1.25, 3);
    cudaMemcpy(cyc, init, sizeof(init), cudaMemcpyHostToDevice);
    kernel<<<grid, TPB>>>(cyc, out, 1.5f, 1.25, 3);
    cudaDeviceSynchronize();
    unsigned long long h[768];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double rate = 0;
    int used = 0;
    for (int s = 0; s < 256; s++) {
        if (h[512 + s] == 0) continue;
        rate += (double)h[512 + s] * 32 * CH * IT / (double)(h[256 + s] - h[s]);
        used++;
    }
    rate /= used;
    // scheduler cycles per warp-level body = 128 lanes per SM-clock / rate
    printf("%-28s %7.1f lane-bodies/clk/SM = %5.2f scheduler cycles per warp body (%g instr)  (%s)\n", name, rate, 128.0 / rate, ops_per_body,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(cyc);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("-- 48-key network, 24 warps/SM (768 threads), one voxel per lane (f32) or two (packed) --\n");
    run_sort<0, 768, 1>("f32 FMNMX+IADD3", sms, 1);
    run_sort<6, 768, 1>("f32 FMNMX+FMNMX", sms, 1);
    run_sort<1, 768, 1>("f32 FMNMX+FADD+FADD", sms, 1);
    run_sort<2, 768, 1>("bf16x2 HMNMX2", sms, 2);
    run_sort<3, 768, 1>("f16x2 HMNMX2", sms, 2);
    run_sort<4, 768, 1>("u16x2 VIMNMX", sms, 2);
    run_sort<5, 768, 1>("VIMNMX min + HMNMX2 max", sms, 2);
    run_sort<7, 768, 1>("bf16x2 HMNMX2 + LOP3", sms, 2);
    run_sort<8, 768, 1>("bf16x2 HMNMX2 + IADD3", sms, 2);
    printf("-- 12 warps/SM (384 threads) --\n");
    run_sort<2, 384, 1>("bf16x2 HMNMX2", sms, 2);
    run_sort<7, 384, 1>("bf16x2 HMNMX2 + LOP3", sms, 2);
    run_sort<8, 384, 1>("bf16x2 HMNMX2 + IADD3", sms, 2);
    printf("-- same, 8 warps/SM (256 threads) --\n");
    run_sort<7, 256, 1>("bf16x2 HMNMX2 + LOP3", sms, 2);
    run_sort<2, 256, 1>("bf16x2 HMNMX2", sms, 2);
    run_sort<4, 256, 1>("u16x2 VIMNMX", sms, 2);
    run_sort<5, 256, 1>("VIMNMX min + HMNMX2 max", sms, 2);
    run_sort<0, 256, 1>("f32 FMNMX+IADD3", sms, 1);
    printf("-- pipes --\n");
    run("2 HMNMX2.f16", k_hmnmx2, 2);
    run("2 HMNMX2.bf16", k_hmnmx2bf, 2);
    run("2 VIMNMX.U16x2", k_vimnmx2, 2);
    run("VIMNMX + HMNMX2", k_hmn_vim, 2);
    run("HMNMX2 + FMNMX", k_hmn_fmn, 2);
    run("HMNMX2 + FFMA", k_hmn_ffma, 2);
    run("HMNMX2 + DFMA", k_hmn_dfma, 2);
    run("VIMNMX + FFMA", k_vim_ffma, 2);
    run("FMNMX + FFMA", k_fmn_ffma, 2);
    run("FMNMX + 2 FFMA", k_fmn_ffma2, 3);
    run("F2FP.BF16 + IADD", k_f2fp, 2);
    run("HSETP2 + 2 @VIADD", k_hsetp2, 3);
    run("PRMT", k_prmt, 1);
    run("F2F.F64.F32 + DADD", k_f2f_up, 2);
    run("F2F + DADD + DFMA + FADD", k_f2f_up2, 4);
    run("pass-1 body (8 instr)", k_pass1, 8);
    run("LDS.64 + IADD", k_lds64, 2);
    run("FSETP + @VIADD + FADD", k_viaddp, 3);
    return 0;
}
