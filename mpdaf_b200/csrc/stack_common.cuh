// Declarations shared by the sigma-clip kernel families (sigclip_pair.cuh, sigclip_warp.cuh, sigclip_runs.cuh):
// the kernel-parameter block of a float32 stack and the mbarrier helpers of the TMA pipelines.
#pragma once

#include "combine_kernels.cuh"

namespace mb {

// Kernel-parameter block describing one float32 exposure stack (constant bank: unrolled per-exposure
// accesses become uniform loads).
template <int NB>
struct TileArgs {
    const float* data[NB];
    const float* var[NB];   // null rows when var is not propagated
    double scale[NB];
    double offset[NB];
    double scale2[NB];
    unsigned dstride, vstride;   // elements between consecutive DATA / STAT rows when equally spaced, else 0
    // float32 images of the per-exposure constants (warp kernel): key = d * kscale + koffset, STAT * scale2f
    double2 os[NB];              // (offset, scale) side by side: one 128-bit uniform load per exposure
    float kscale[NB];
    float koffset[NB];
    float scale2f[NB];
    double koff;                 // bound on |key-before-truncation - float32(w)| from the offsets, part of the margins
};

// ---- mbarrier / bulk-copy PTX ---------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

}  // namespace mb
