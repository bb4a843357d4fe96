// Shared declarations between the C-ABI translation unit (ctools_abi.cu) and the per-depth
// kernel translation units (warp_<NB>.cu), which are compiled in parallel by build.py.
#pragma once
#include <cstddef>
#include <string>

#include <cuda_runtime.h>

#include "../../include/mpdaf_b200.h"
#include "combine_kernels.cuh"

namespace mb {

struct Job {
    int nfiles = 0;
    int elem = MPDAF_B200_F32;
    bool median = false;
    const void* const* data = nullptr;  // host table of DEVICE pointers
    const void* const* var = nullptr;
    const double* scale = nullptr;       // host, may be null
    const double* offset = nullptr;
    StackTable* table_dev = nullptr;     // device scratch for the generic kernels' stack table
    unsigned* slow_q = nullptr;          // per-stream queue for the pair kernel's exact pass (null: the device's own)
    unsigned* slow_n = nullptr;
    unsigned slow_cap = 0;
    ClipArgs a{};
};

// context a kernel launcher needs from the ABI layer
struct LaunchCtx {
    int sm_count = 0;
    unsigned long long* slow = nullptr;  // device counter of exact-fallback voxels
    unsigned* slow_q = nullptr;          // device queue of voxels the pair kernel could not certify
    unsigned* slow_n = nullptr;          // its fill count
    unsigned slow_cap = 0;
    cudaStream_t stream = nullptr;
    char* name = nullptr;                // receives the kernel family name
    size_t name_len = 0;
    int launches = 0;                    // incremented per kernel launch
    unsigned long long* hist = nullptr;  // device exposure-map histogram a kernel may add to itself (null: not wanted)
    bool hist_fused = false;             // set by a launcher whose kernels filled `hist`: no separate histogram pass
};

int fail(int code, const std::string& msg);   // records the message (ctools_abi.cu)

}  // namespace mb

#define MB_CUDA(call)                                                                     \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return mb::fail(MPDAF_B200_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

// one pair of entry points per stack-depth bucket, each defined in its own translation unit
#define MB_FOR_EACH_DEPTH(X) X(4) X(8) X(12) X(16) X(24) X(32) X(48) X(64) X(96) X(128)
#define MB_DECLARE_DEPTH(NB)                                                        \
    int mb_launch_warp_##NB(const mb::Job& j, bool unit, mb::LaunchCtx& ctx);       \
    int mb_launch_median_warp_##NB(const mb::Job& j, mb::LaunchCtx& ctx);
MB_FOR_EACH_DEPTH(MB_DECLARE_DEPTH)
#undef MB_DECLARE_DEPTH
// float64 samples: pair kernel in four depth buckets (pair64_<NB>.cu)
int mb_launch_pair64_12(const mb::Job& j, bool unit, mb::LaunchCtx& ctx, bool& taken);
int mb_launch_pair64_24(const mb::Job& j, bool unit, mb::LaunchCtx& ctx, bool& taken);
int mb_launch_pair64_48(const mb::Job& j, bool unit, mb::LaunchCtx& ctx, bool& taken);
int mb_launch_pair64_64(const mb::Job& j, bool unit, mb::LaunchCtx& ctx, bool& taken);
// multi-run kernel for 129..512 exposures (sigclip_runs.cuh); the stack table must already be on the device
int mb_launch_runs(const mb::Job& j, bool unit, mb::LaunchCtx& ctx);
