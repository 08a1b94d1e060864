// cc_math.cuh -- float transcendentals that are BIT-IDENTICAL to the host libm the reference links against.
//
// The reference computes the range-image column of every point from std::atan2(float, float)
// (continuous_clustering.cpp:142), inclinations and association windows from std::asin(float)
// (cpp:232, 805) and one ignore rule from std::atan2 (cpp:598).  A one-ulp difference moves a point into
// the neighbouring column, so CUDA's own atan2f/asinf (different algorithms) cannot be used.  These are
// re-implementations of the algorithms glibc 2.39 ships for x86-64 (the fdlibm-derived
// sysdeps/ieee754/flt-32/{e_atan2f,s_atanf,e_asinf}.c; no FMA ifunc variants exist for them), with the
// constants and thresholds read back from the libm.so.6 of this image.  Every operation is a plain IEEE
// binary32 add/sub/mul/div/sqrt in the same order, so compiled with -fmad=false (device) or without FMA
// contraction (host) the results are bit-equal; tests/test_math.py sweeps this against the host libm
// (exhaustively for asinf/atanf) and tests/test_gpu_math.py does the same for the device build.
#ifndef CC_MATH_CUH
#define CC_MATH_CUH

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define CC_HD __host__ __device__ __forceinline__
#else
#define CC_HD inline
#endif

namespace ccm
{

CC_HD uint32_t f2u(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}

CC_HD float u2f(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

CC_HD float sqrt_rn(float x)
{
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(x);
#else
    return __builtin_sqrtf(x);
#endif
}

CC_HD float div_rn(float a, float b)
{
#if defined(__CUDA_ARCH__)
    return __fdiv_rn(a, b);
#else
    return a / b;
#endif
}

CC_HD float fabs_f(float x)
{
    return u2f(f2u(x) & 0x7fffffffu);
}

// atanf: argument reduction to [0, 7/16) around 0.5, 1, 1.5, inf + odd/even split degree-11 polynomial.
CC_HD float atanf_glibc(float x)
{
    const float atanhi0 = u2f(0x3eed6338u), atanhi1 = u2f(0x3f490fdau), atanhi2 = u2f(0x3f7b985eu),
                atanhi3 = u2f(0x3fc90fdau);
    const float atanlo0 = u2f(0x31ac3769u), atanlo1 = u2f(0x33222168u), atanlo2 = u2f(0x33140fb4u),
                atanlo3 = u2f(0x33a22168u);
    const float aT0 = u2f(0x3eaaaaabu), aT1 = u2f(0xbe4ccccdu), aT2 = u2f(0x3e124925u), aT3 = u2f(0xbde38e38u),
                aT4 = u2f(0x3dba2e6eu), aT5 = u2f(0xbd9d8795u), aT6 = u2f(0x3d886b35u), aT7 = u2f(0xbd6ef16bu),
                aT8 = u2f(0x3d4bda59u), aT9 = u2f(0xbd15a221u), aT10 = u2f(0x3c8569d7u);
    const int32_t hx = (int32_t)f2u(x);
    const int32_t ix = hx & 0x7fffffff;
    int id;
    float hi = 0.f, lo = 0.f;
    if (ix >= 0x4c000000)
    { // |x| >= 2^25
        if (ix > 0x7f800000)
            return x + x; // NaN
        if (hx > 0)
            return atanhi3 + atanlo3;
        return -atanhi3 - atanlo3;
    }
    if (ix < 0x3ee00000)
    { // |x| < 0.4375
        if (ix < 0x31000000)
            return x; // |x| < 2^-29
        id = -1;
    }
    else
    {
        x = fabs_f(x);
        if (ix < 0x3f980000)
        { // |x| < 1.1875
            if (ix < 0x3f300000)
            { // 7/16 <= |x| < 11/16
                id = 0;
                hi = atanhi0;
                lo = atanlo0;
                x = div_rn((x + x) - 1.0f, 2.0f + x);
            }
            else
            { // 11/16 <= |x| < 19/16
                id = 1;
                hi = atanhi1;
                lo = atanlo1;
                x = div_rn(x - 1.0f, x + 1.0f);
            }
        }
        else
        {
            if (ix < 0x401c0000)
            { // |x| < 2.4375
                id = 2;
                hi = atanhi2;
                lo = atanlo2;
                x = div_rn(x - 1.5f, 1.0f + 1.5f * x);
            }
            else
            { // 2.4375 <= |x| < 2^25
                id = 3;
                hi = atanhi3;
                lo = atanlo3;
                x = div_rn(-1.0f, x);
            }
        }
    }
    const float z = x * x;
    const float w = z * z;
    const float s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
    const float s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
    if (id < 0)
        return x - x * (s1 + s2);
    const float r = hi - ((x * (s1 + s2) - lo) - x);
    return (hx < 0) ? -r : r;
}

CC_HD float atan2f_glibc(float y, float x)
{
    const float tiny = u2f(0x0da24260u); // 1e-30
    const float pi_o_4 = u2f(0x3f490fdbu), pi_o_2 = u2f(0x3fc90fdbu), pi = u2f(0x40490fdbu),
                pi_lo = u2f(0xb3bbbd2eu); // -8.7422776573e-08
    const int32_t hx = (int32_t)f2u(x), hy = (int32_t)f2u(y);
    const int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
    if (ix > 0x7f800000 || iy > 0x7f800000)
        return x + y; // NaN
    if (hx == 0x3f800000)
        return atanf_glibc(y); // x == 1.0
    const int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
    if (iy == 0)
    {
        switch (m)
        {
            case 0:
            case 1:
                return y;
            case 2:
                return pi + tiny;
            default:
                return -pi - tiny;
        }
    }
    if (ix == 0)
        return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    if (ix == 0x7f800000)
    {
        if (iy == 0x7f800000)
        {
            switch (m)
            {
                case 0:
                    return pi_o_4 + tiny;
                case 1:
                    return -pi_o_4 - tiny;
                case 2:
                    return 3.0f * pi_o_4 + tiny;
                default:
                    return -3.0f * pi_o_4 - tiny;
            }
        }
        switch (m)
        {
            case 0:
                return 0.0f;
            case 1:
                return -0.0f;
            case 2:
                return pi + tiny;
            default:
                return -pi - tiny;
        }
    }
    if (iy == 0x7f800000)
        return (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    const int32_t k = (iy - ix) >> 23;
    float z;
    if (k > 60)
        z = pi_o_2 + 0.5f * pi_lo;
    else if (hx < 0 && k < -60)
        z = 0.0f;
    else
        z = atanf_glibc(fabs_f(div_rn(y, x)));
    switch (m)
    {
        case 0:
            return z;
        case 1:
            return u2f(f2u(z) ^ 0x80000000u);
        case 2:
            return pi - (z - pi_lo);
        default:
            return (z - pi_lo) - pi;
    }
}

CC_HD float asinf_glibc(float x)
{
    const float pio2_hi = u2f(0x3fc90fdbu), pio2_lo = u2f(0xb33bbd2eu), pio4_hi = u2f(0x3f490fdbu);
    const float p0 = u2f(0x3e2aaae4u), p1 = u2f(0x3d9980f2u), p2 = u2f(0x3d3a3f25u), p3 = u2f(0x3cc6141eu),
                p4 = u2f(0x3d2cb694u);
    const int32_t hx = (int32_t)f2u(x);
    const int32_t ix = hx & 0x7fffffff;
    if (ix == 0x3f800000)
        return x * pio2_hi + x * pio2_lo; // asin(+-1)
    if (ix > 0x3f800000)
        return div_rn(x - x, x - x); // |x| > 1 or NaN -> NaN
    if (ix < 0x3f000000)
    { // |x| < 0.5
        if (ix < 0x32000000)
            return x; // |x| < 2^-27
        const float t = x * x;
        const float w = t * (p0 + t * (p1 + t * (p2 + t * (p3 + t * p4))));
        return x + x * w;
    }
    float w = 1.0f - fabs_f(x);
    float t = w * 0.5f;
    const float p = t * (p0 + t * (p1 + t * (p2 + t * (p3 + t * p4))));
    const float s = sqrt_rn(t);
    if (ix >= 0x3f79999a)
    { // |x| > 0.975
        t = pio2_hi - (2.0f * (s + s * p) - pio2_lo);
    }
    else
    {
        w = u2f(f2u(s) & 0xfffff000u);
        const float c = div_rn(t - w * w, s + w);
        const float r = p;
        const float s2 = 2.0f * s * r - (pio2_lo - 2.0f * c);
        const float q = pio4_hi - 2.0f * w;
        t = pio4_hi - (s2 - q);
    }
    return (hx > 0) ? t : -t;
}

} // namespace ccm

#endif
