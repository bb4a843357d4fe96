// sigclip_pair_kernel for float64 samples, up to 64 exposures (see warp_inst.cuh).
#include "warp_inst.cuh"
MB_DEFINE_PAIR64(64)
