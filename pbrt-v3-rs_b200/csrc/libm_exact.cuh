// sinf / cosf bit-compatible with the host libm the reference resolves to.
//
// Why: Rust's f32::sin / f32::cos lower to the platform libm's sinf / cosf (glibc on x86-64 Linux).
// Sampled directions (cosine_sample_hemisphere, Trowbridge-Reitz, infinite-light sampling) feed
// straight into ray origins/directions, so a 1-ulp difference in sin/cos makes a later bounce hit a
// neighbouring triangle once in ~10^4 rays and the two renders drift apart pixel by pixel.  CUDA's sinf
// differs from glibc's in roughly one call out of ten.  These functions re-implement glibc's algorithm
// (sysdeps/ieee754/flt-32/s_sincosf.h, the ARM "optimized routines" sincosf: double-precision range
// reduction by pi/2 and degree-7/8 minimax polynomials) with the exact operation grouping and FMA
// placement of the x86-64 FMA ifunc variant glibc selects on every AVX2+FMA host, so the GPU path produces the
// same bits as the CPU path.  tests/test_libm_exact.py checks host-vs-glibc equality on 10^8 inputs and
// tests/test_render_gpu.py checks device-vs-host equality.
//
// Domain: |x| < 120 (all call sites pass angles in [-pi, 2*pi]); outside it the CUDA routine is used.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LMX_HD __host__ __device__ __forceinline__
#else
#define LMX_HD inline
#endif

namespace lmx {

// __sincosf_table[0] / [1] of glibc 2.39 (cosine coefficients negated in the second table).
struct SinCosTab {
    double c0, c1, c2, c3, c4, s1, s2, s3;
};
LMX_HD SinCosTab tab(int neg) {
    SinCosTab t;
    t.c0 = 0x1p0; t.c1 = -0x1.ffffffd0c621cp-2; t.c2 = 0x1.55553e1068f19p-5; t.c3 = -0x1.6c087e89a359dp-10; t.c4 = 0x1.99343027bf8c3p-16;
    t.s1 = -0x1.555545995a603p-3; t.s2 = 0x1.1107605230bc4p-7; t.s3 = -0x1.994eb3774cf24p-13;
    if (neg) { t.c0 = -t.c0; t.c1 = -t.c1; t.c2 = -t.c2; t.c3 = -t.c3; t.c4 = -t.c4; }
    return t;
}
LMX_HD uint32_t fbits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
LMX_HD double dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
LMX_HD double dmul(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);  // never contracted
#else
    return a * b;
#endif
}
// x + x^3 s1 + x^5 (s2 + x^2 s3)
LMX_HD float sin_poly(double x, double x2, const SinCosTab& p) {
    double t = dfma(x2, p.s3, p.s2);
    double x3 = dmul(x2, x);
    double x5 = dmul(x2, x3);
    double s = dfma(x3, p.s1, x);
    return (float)dfma(t, x5, s);
}
// (c0 + x^2 c1) + x^4 c2 + x^6 (c3 + x^2 c4)
LMX_HD float cos_poly(double x2, const SinCosTab& p) {
    double x4 = dmul(x2, x2);
    double c1 = dfma(x2, p.c1, p.c0);
    double c2 = dfma(x2, p.c4, p.c3);
    double x6 = dmul(x2, x4);
    double c = dfma(x4, p.c2, c1);
    return (float)dfma(c2, x6, c);
}
LMX_HD int reduce(double x, double* xr) {
    double r = dmul(x, 0x1.45f306dc9c883p+23);  // 2/pi * 2^24
    int n = ((int32_t)r + 0x800000) >> 24;
    *xr = dfma(-(double)n, 0x1.921fb54442d18p+0, x);
    return n;
}

LMX_HD float sinf_glibc(float y) {
    uint32_t top = (fbits(y) >> 20) & 0x7ff;
    double x = (double)y;
    if (top <= 0x3f3) {                 // |y| < pi/4
        if (top <= 0x397) return y;     // |y| < 2^-12
        return sin_poly(x, dmul(x, x), tab(0));
    }
    if (top <= 0x42e) {                 // |y| < 120
        double xr;
        int n = reduce(x, &xr);
        SinCosTab p = tab((n & 2) != 0);
        double x2 = dmul(xr, xr);
        if ((n & 1) == 0) {
            double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
            return sin_poly(dmul(xr, sgn), x2, p);
        }
        return cos_poly(x2, p);
    }
    return sinf(y);
}

LMX_HD float cosf_glibc(float y) {
    uint32_t top = (fbits(y) >> 20) & 0x7ff;
    double x = (double)y;
    if (top <= 0x3f3) {
        if (top <= 0x397) return 1.0f;
        return cos_poly(dmul(x, x), tab(0));
    }
    if (top <= 0x42e) {
        double xr;
        int n = reduce(x, &xr);
        SinCosTab p = tab((n & 2) != 0);
        double x2 = dmul(xr, xr);
        if (n & 1) {
            double sgn = ((n & 3) == 1 || (n & 3) == 2) ? -1.0 : 1.0;
            return sin_poly(dmul(xr, sgn), x2, p);
        }
        return cos_poly(x2, p);
    }
    return cosf(y);
}

}  // namespace lmx
