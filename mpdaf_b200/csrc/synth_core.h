// Counter-based synthetic MUSE-like exposure samples.
//
// One sample is a pure function of (seed, exposure, global voxel index), so any
// wavelength slab can be regenerated bit-identically on the host (for the
// oracle) and on the device (for full-size runs that do not fit host RAM).
// Only integer arithmetic and single float32 operations with one rounding each
// are used -- no transcendental functions, no FMA -- so host and device agree.
//
//   DATA  = base(exposure) + g            g ~ Irwin-Hall(6 x u16), sigma = 1
//           (+ 50 for ~1 % cosmic-ray-like outliers), NaN for ~5 % of the
//           samples, for the two outermost spaxel columns on each side (field
//           edge, all exposures) and for the first (exposure % 3) rows (dither).
//   STAT  = 0.5 + 1.5 * u24               uniform in [0.5, 2)
//   base  = 10 + 0.02 * ((exposure % 7) - 3)   per-exposure flux bias
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define MB_HD __host__ __device__ __forceinline__
#else
#define MB_HD static inline
#endif

namespace mbsynth {

MB_HD uint64_t mix64(uint64_t z)
{
    z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
    z ^= z >> 27; z *= 0x94D049BB133111EBull;
    z ^= z >> 31;
    return z;
}

MB_HD uint64_t sample_key(uint64_t seed, int exposure, long long voxel)
{
    return seed * 0x9E3779B97F4A7C15ull + (uint64_t)(exposure + 1) * 0xD1B54A32D192ED03ull +
           (uint64_t)voxel;
}

MB_HD float f_mul(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}

MB_HD float f_add(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}

MB_HD float exposure_base(int exposure)
{
    return f_add(10.0f, f_mul(0.02f, (float)((exposure % 7) - 3)));
}

MB_HD uint32_t quiet_nan_bits() { return 0x7FC00000u; }

// DATA sample as raw float32 bits (so NaN payloads are identical everywhere).
MB_HD uint32_t data_bits(uint64_t seed, int exposure, long long voxel, int nx, int ny)
{
    const uint64_t key = sample_key(seed, exposure, voxel);
    const uint64_t z1 = mix64(key);
    const uint64_t z2 = mix64(key ^ 0xA0761D6478BD642Full);
    const int x = (int)(voxel % nx);
    const int y = (int)((voxel / nx) % ny);
    const bool edge = (x < 2) || (x >= nx - 2) || (y < (exposure % 3));
    const bool hole = ((z2 >> 32) & 0xFFFFu) < 3277u;              // 5.0 %
    if (edge || hole) return quiet_nan_bits();
    int s = (int)(z1 & 0xFFFFu) + (int)((z1 >> 16) & 0xFFFFu) + (int)((z1 >> 32) & 0xFFFFu) +
            (int)((z1 >> 48) & 0xFFFFu) + (int)(z2 & 0xFFFFu) + (int)((z2 >> 16) & 0xFFFFu);
    float g = f_mul((float)(s - 196605), 2.1579186e-5f);           // 1 / (65536 * sqrt(0.5))
    float v = f_add(exposure_base(exposure), g);
    if (((z2 >> 48) & 0xFFFFu) < 655u) v = f_add(v, 50.0f);         // 1.0 % outliers
    union { float f; uint32_t u; } cvt;
    cvt.f = v;
    return cvt.u;
}

MB_HD uint32_t stat_bits(uint64_t seed, int exposure, long long voxel)
{
    const uint64_t z3 = mix64(sample_key(seed, exposure, voxel) ^ 0xE7037ED1A0B428DBull);
    float v = f_add(0.5f, f_mul((float)(uint32_t)(z3 >> 40), 8.9406967e-8f));  // 1.5 / 2^24
    union { float f; uint32_t u; } cvt;
    cvt.f = v;
    return cvt.u;
}

}  // namespace mbsynth
