// Warp-tiled kernels for stacks of up to 96 exposures (one translation unit per depth bucket).
#include "warp_inst.cuh"
MB_DEFINE_DEPTH(96)
