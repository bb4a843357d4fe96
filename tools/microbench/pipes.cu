// Instruction-throughput probe for the pipes the combine kernels lean on (sm_100a).
// Prints lane-ops per clock per SM for each pattern; used to size the kernel design
// (DESIGN.md "instruction budget").  Not part of the product.
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>

constexpr int CH = 8;      // independent chains per thread
constexpr int IT = 2048;    // loop iterations
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

// ops per BODY instance noted beside each kernel
KERNEL(k_fmnmx, float a[CH]; float b = seedf; float d = seedf * 3,
       a[c] = seedf + c, asm volatile("min.f32 %0, %0, %1; max.f32 %0, %0, %2;" : "+f"(a[c]) : "f"(b), "f"(d)), acc += a[c])  // 2
KERNEL(k_imnmx, int a[CH]; int b = seedi; int d = seedi * 3,
       a[c] = seedi + c, asm volatile("min.s32 %0, %0, %1; max.s32 %0, %0, %2;" : "+r"(a[c]) : "r"(b), "r"(d)), acc += a[c])  // 2
KERNEL(k_lop3, int a[CH]; int b = seedi; int d = seedi * 3,
       a[c] = seedi + c, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[c]) : "r"(b), "r"(d)), acc += a[c])  // 1
KERNEL(k_iadd, int a[CH]; int b = seedi,
       a[c] = seedi + c, asm volatile("add.s32 %0, %0, %1;" : "+r"(a[c]) : "r"(b)), acc += a[c])  // 1
KERNEL(k_fadd, float a[CH]; float b = seedf,
       a[c] = seedf + c, asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b)), acc += a[c])  // 1
KERNEL(k_ffma, float a[CH]; float b = seedf,
       a[c] = seedf + c, asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[c]) : "f"(b)), acc += a[c])  // 1
KERNEL(k_fsetp_sel, float a[CH]; float b = seedf; float d = seedf * 3,
       a[c] = seedf + c, asm volatile("{.reg .pred p; setp.lt.f32 p, %0, %1; selp.f32 %0, %2, %0, p;}" : "+f"(a[c]) : "f"(b), "f"(d)), acc += a[c])  // 2
KERNEL(k_dadd, double a[CH]; double b = seedd,
       a[c] = seedd + c, asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(a[c]) : "d"(b)), acc += (float)a[c])  // 1
KERNEL(k_dmul, double a[CH]; double b = seedd,
       a[c] = seedd + c, asm volatile("mul.rn.f64 %0, %0, %1;" : "+d"(a[c]) : "d"(b)), acc += (float)a[c])  // 1
KERNEL(k_dfma, double a[CH]; double b = seedd,
       a[c] = seedd + c, asm volatile("fma.rn.f64 %0, %0, %1, %1;" : "+d"(a[c]) : "d"(b)), acc += (float)a[c])  // 1
KERNEL(k_dsetp_sel, double a[CH]; int r[CH]; double b = seedd,
       a[c] = seedd + c; r[c] = c, asm volatile("{.reg .pred p; setp.lt.f64 p, %1, %2; selp.b32 %0, %0, 7, p;}" : "+r"(r[c]) : "d"(a[c]), "d"(b)), acc += r[c])  // 1 DSETP + 1 SEL
KERNEL(k_dce, double a[CH]; double b[CH],
       a[c] = seedd + c; b[c] = seedd * c,
       asm volatile("{.reg .pred p; .reg .f64 t; setp.lt.f64 p, %0, %1; selp.f64 t, %0, %1, p; selp.f64 %1, %1, %0, p; mov.f64 %0, t;}" : "+d"(a[c]), "+d"(b[c])), acc += (float)(a[c] + b[c]))  // 1 f64 compare-exchange
KERNEL(k_fce, float a[CH]; float b[CH],
       a[c] = seedf + c; b[c] = seedf * c,
       asm volatile("{.reg .f32 t; min.f32 t, %0, %1; max.f32 %1, %0, %1; mov.f32 %0, t;}" : "+f"(a[c]), "+f"(b[c])), acc += (a[c] + b[c]))  // 1 f32 compare-exchange
KERNEL(k_fce_addsub, float a[CH]; float b[CH],
       a[c] = seedf + c; b[c] = seedf * c,
       asm volatile("{.reg .f32 t; .reg .b32 x, y, z; min.f32 t, %0, %1; mov.b32 x, %0; mov.b32 y, %1; mov.b32 z, t; add.s32 x, x, y; sub.s32 x, x, z; mov.b32 %1, x; mov.f32 %0, t;}" : "+f"(a[c]), "+f"(b[c])), acc += (a[c] + b[c]))  // CE = FMNMX + integer a+b-lo
KERNEL(k_shfl, int a[CH]; int b = seedi,
       a[c] = seedi + c + threadIdx.x, asm volatile("{.reg .b32 t; shfl.sync.bfly.b32 t, %0, 1, 0x1f, 0xffffffff; xor.b32 %0, %0, t;}" : "+r"(a[c])), acc += a[c])  // SHFL + LOP
KERNEL(k_popc, int a[CH]; int b = seedi,
       a[c] = seedi + c, asm volatile("{.reg .b32 t; popc.b32 t, %0; add.s32 %0, %0, t;}" : "+r"(a[c])), acc += a[c])  // POPC + IADD
KERNEL(k_cvt_up, float a[CH]; int r[CH],
       a[c] = seedf + c; r[c] = c, asm volatile("{.reg .f64 t; .reg .b32 lo, hi; cvt.f64.f32 t, %1; mov.b64 {lo, hi}, t; xor.b32 %0, %0, hi;}" : "+r"(r[c]) : "f"(a[c])), acc += r[c])  // F2F.F64.F32 + LOP
KERNEL(k_cvt_down, double a[CH]; int r[CH],
       a[c] = seedd + c; r[c] = c, asm volatile("{.reg .f32 t; .reg .b32 u; cvt.rn.f32.f64 t, %1; mov.b32 u, t; xor.b32 %0, %0, u;}" : "+r"(r[c]) : "d"(a[c])), acc += r[c])  // F2F.F32.F64 + LOP
KERNEL(k_lds, int a[CH]; __shared__ int sm[TPB * 4]; sm[threadIdx.x] = seedi; sm[threadIdx.x + TPB] = 0,
       a[c] = threadIdx.x * 4, asm volatile("{.reg .b32 t; ld.shared.b32 t, [%0]; add.s32 %0, %0, t;}" : "+r"(a[c])), acc += a[c])  // LDS + IADD (dependent)
KERNEL(k_imad, int a[CH]; int b = seedi; int d = seedi * 5,
       a[c] = seedi + c + threadIdx.x, asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a[c]) : "r"(b), "r"(d)), acc += a[c])  // IMAD
KERNEL(k_iadd3, int a[CH]; int b = seedi + threadIdx.x; int d = seedi * 5,
       a[c] = seedi + c + threadIdx.x, asm volatile("{.reg .b32 t; add.s32 t, %0, %1; sub.s32 %0, t, %2; xor.b32 %0, %0, %1;}" : "+r"(a[c]) : "r"(b), "r"(d)), acc += a[c])  // IADD3 + LOP3
KERNEL(k_fce_mixed, float a[CH]; float b[CH]; int one = seedi / 3; int mone = -one,
       a[c] = seedf + c; b[c] = seedf * c,
       asm volatile("{.reg .f32 t; .reg .b32 x, y, z; min.f32 t, %0, %1; mov.b32 x, %0; mov.b32 y, %1; mov.b32 z, t; mad.lo.s32 x, x, %2, y; mad.lo.s32 x, z, %3, x; mov.b32 %1, x; mov.f32 %0, t;}" : "+f"(a[c]), "+f"(b[c]) : "r"(one), "r"(mone)), acc += (a[c] + b[c]))  // CE = FMNMX + 2 IMAD
KERNEL(k_f2f_up, float a[CH]; double d[CH],
       a[c] = seedf + c; d[c] = 0, asm volatile("{.reg .f64 t; cvt.f64.f32 t, %1; add.rn.f64 %0, %0, t;}" : "+d"(d[c]) : "f"(a[c])), acc += (float)d[c])  // 1 F2F + 1 DADD
KERNEL(k_f2f_down, double a[CH]; float f[CH],
       a[c] = seedd + c; f[c] = 0, asm volatile("{.reg .f32 t; cvt.rn.f32.f64 t, %1; add.rn.f32 %0, %0, t;}" : "+f"(f[c]) : "d"(a[c])), acc += f[c])  // 1 F2F + 1 FADD
KERNEL(k_mix_fmnmx_dadd, float a[CH]; double d[CH]; float b = seedf; double e = seedd,
       a[c] = seedf + c; d[c] = seedd, asm volatile("min.f32 %0, %0, %2; add.rn.f64 %1, %1, %3;" : "+f"(a[c]), "+d"(d[c]) : "f"(b), "d"(e)), acc += a[c] + (float)d[c])  // 1 FMNMX + 1 DADD
KERNEL(k_mix_fmnmx_ffma, float a[CH]; float d[CH]; float b = seedf,
       a[c] = seedf + c; d[c] = seedf, asm volatile("min.f32 %0, %0, %2; fma.rn.f32 %1, %1, %2, %2;" : "+f"(a[c]), "+f"(d[c]) : "f"(b)), acc += a[c] + d[c])  // 1 FMNMX + 1 FFMA
KERNEL(k_vote, int a[CH]; int b = seedi,
       a[c] = seedi + c, asm volatile("{.reg .pred p; .reg .b32 t; setp.lt.s32 p, %0, %1; vote.sync.ballot.b32 t, p, 0xffffffff; popc.b32 t, t; add.s32 %0, %0, t;}" : "+r"(a[c]) : "r"(b)), acc += a[c])  // ISETP+VOTE+POPC+IADD

template <typename K>
void run(const char* name, K kernel, double ops_per_body)
{
    int dev = 0, sms = 0;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    sms = prop.multiProcessorCount;
    const int bps = 4;
    const int grid = sms * bps;
    unsigned long long* cyc;
    float* out;
    cudaMalloc(&cyc, 768 * sizeof(unsigned long long));
    cudaMalloc(&out, 4);
    std::vector<unsigned long long> init(768, 0ull);
    for (int i = 0; i < 256; i++) init[i] = ~0ull;
    for (int rep = 0; rep < 2; rep++) kernel<<<grid, TPB>>>(cyc, out, 1.5f, 1.25, 3);
    cudaMemcpy(cyc, init.data(), 768 * sizeof(unsigned long long), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kernel<<<grid, TPB>>>(cyc, out, 1.5f, 1.25, 3);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<unsigned long long> h(768);
    cudaMemcpy(h.data(), cyc, 768 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    // per SM: lane-bodies executed / (last warp end - first warp start)
    double rate = 0, cycles = 0;
    int used = 0;
    for (int s = 0; s < 256; s++) {
        if (h[512 + s] == 0) continue;
        const double c = (double)(h[256 + s] - h[s]);
        rate += (double)h[512 + s] * 32 * CH * IT / c;
        cycles += c;
        used++;
    }
    rate /= used;
    cycles /= used;
    printf("%-22s %7.1f lane-bodies/clk/SM (%4.1f ops/body -> %6.1f lane-ops/clk/SM)  SMs=%d cycles=%.0f  %.3f ms -> %.2f GHz  %s\n",
           name, rate, ops_per_body, rate * ops_per_body, used, cycles, ms, cycles / (ms * 1e6),
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(cyc);
    cudaFree(out);
}

int main()
{
    run("fmnmx", k_fmnmx, 2);
    run("imnmx", k_imnmx, 2);
    run("lop3", k_lop3, 1);
    run("iadd", k_iadd, 1);
    run("fadd", k_fadd, 1);
    run("ffma", k_ffma, 1);
    run("fsetp+sel", k_fsetp_sel, 2);
    run("dadd", k_dadd, 1);
    run("dmul", k_dmul, 1);
    run("dfma", k_dfma, 1);
    run("dsetp+sel", k_dsetp_sel, 2);
    run("f64 cmpxchg", k_dce, 1);
    run("f32 cmpxchg", k_fce, 1);
    run("imad", k_imad, 1);
    run("iadd3+lop3", k_iadd3, 2);
    run("f32 cmpxchg 2xIMAD", k_fce_mixed, 1);
    run("f32 cmpxchg addsub", k_fce_addsub, 1);
    run("shfl+lop", k_shfl, 2);
    run("popc+iadd", k_popc, 2);
    run("cvt.f64.f32+lop", k_cvt_up, 2);
    run("cvt.f32.f64+lop", k_cvt_down, 2);
    run("lds+iadd", k_lds, 2);
    run("f2f.f64.f32+dadd", k_f2f_up, 2);
    run("f2f.f32.f64+fadd", k_f2f_down, 2);
    run("fmnmx+dadd", k_mix_fmnmx_dadd, 2);
    run("fmnmx+ffma", k_mix_fmnmx_ffma, 2);
    run("isetp+vote+popc+iadd", k_vote, 4);
    return 0;
}
