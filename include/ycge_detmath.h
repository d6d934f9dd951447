/*
 * ycge_detmath.h — deterministic transcendental functions for host and device.
 *
 * Why this exists: the reference calls MathF.Sin/Cos/Tan/SinCos/Exp/Log/Pow
 * (RaytraceRenderer.cs:415-416,429,694-697,754; RaytraceSampler.cs:88; ToneMapper.cs:77,83,87,214-216),
 * which forward to the platform C runtime and are not bit-reproducible even between two .NET
 * installations.  To make CPU-oracle <-> GPU comparisons exact *through* those call sites, both sides
 * evaluate the same function below, built only from exactly defined IEEE-754 operations (+ - * /
 * and, for exp, fused multiply-add written out explicitly) and integer bit manipulation, evaluated in
 * a fixed order.  sin / cos / tan / log / pow are evaluated in binary64 and rounded once to binary32
 * (within ~1e-13 relative before the final rounding, i.e. correctly rounded except on roughly one
 * input in 10^6); exp -- 300 calls per pixel and frame -- is evaluated in binary32 (within 1 ulp).
 * The oracle can be switched to glibc libm to measure how many cells that changes (the "documented
 * float ties" of BASELINE.json).
 *
 * Requirements on the build: host code must be compiled with -ffp-contract=off (no implicit FMA
 * contraction) and without -ffast-math, and should enable the hardware FMA (-mfma) so that the explicit
 * fmaf of ycge_expf is one instruction; device code uses the explicit round-to-nearest intrinsics,
 * which nvcc never contracts.
 */
#ifndef YCGE_DETMATH_H
#define YCGE_DETMATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define YDM_HD __host__ __device__ __forceinline__
#else
#define YDM_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define YDM_MUL(a, b) __dmul_rn((a), (b))
#define YDM_ADD(a, b) __dadd_rn((a), (b))
#define YDM_SUB(a, b) __dadd_rn((a), -(b))
#define YDM_DIV(a, b) __ddiv_rn((a), (b))
#else
#define YDM_MUL(a, b) ((a) * (b))
#define YDM_ADD(a, b) ((a) + (b))
#define YDM_SUB(a, b) ((a) - (b))
#define YDM_DIV(a, b) ((a) / (b))
#endif

/* 2^(i/32), i = 0..31, correctly rounded to binary64 (generated with mpmath at 200 bits). */
#define YDM_T32_INIT                                                                               \
    {0x3FF0000000000000ULL, 0x3FF059B0D3158574ULL, 0x3FF0B5586CF9890FULL, 0x3FF11301D0125B51ULL,    \
     0x3FF172B83C7D517BULL, 0x3FF1D4873168B9AAULL, 0x3FF2387A6E756238ULL, 0x3FF29E9DF51FDEE1ULL,    \
     0x3FF306FE0A31B715ULL, 0x3FF371A7373AA9CBULL, 0x3FF3DEA64C123422ULL, 0x3FF44E086061892DULL,    \
     0x3FF4BFDAD5362A27ULL, 0x3FF5342B569D4F82ULL, 0x3FF5AB07DD485429ULL, 0x3FF6247EB03A5585ULL,    \
     0x3FF6A09E667F3BCDULL, 0x3FF71F75E8EC5F74ULL, 0x3FF7A11473EB0187ULL, 0x3FF82589994CCE13ULL,    \
     0x3FF8ACE5422AA0DBULL, 0x3FF93737B0CDC5E5ULL, 0x3FF9C49182A3F090ULL, 0x3FFA5503B23E255DULL,    \
     0x3FFAE89F995AD3ADULL, 0x3FFB7F76F2FB5E47ULL, 0x3FFC199BDD85529CULL, 0x3FFCB720DCEF9069ULL,    \
     0x3FFD5818DCFBA487ULL, 0x3FFDFC97337B9B5FULL, 0x3FFEA4AFA2A490DAULL, 0x3FFF50765B6E4540ULL}

#if defined(__CUDACC__)
static __device__ const unsigned long long ydm_t32_dev[32] = YDM_T32_INIT;
#endif
static const unsigned long long ydm_t32_host[32] = YDM_T32_INIT;

YDM_HD double ydm_from_bits(unsigned long long b) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, sizeof d);
    return d;
#endif
}
YDM_HD unsigned long long ydm_to_bits(double d) {
#if defined(__CUDA_ARCH__)
    return (unsigned long long)__double_as_longlong(d);
#else
    unsigned long long b;
    memcpy(&b, &d, sizeof b);
    return b;
#endif
}
YDM_HD double ydm_t32(int i) {
#if defined(__CUDA_ARCH__)
    return ydm_from_bits(ydm_t32_dev[i]);
#else
    return ydm_from_bits(ydm_t32_host[i]);
#endif
}

/* round to nearest integer (ties to even) for |z| < 2^51, without libm */
YDM_HD double ydm_rint(double z) {
    const double magic = 6755399441055744.0; /* 1.5 * 2^52 */
    return YDM_SUB(YDM_ADD(z, magic), magic);
}

/* e^x for binary64 x; caller guarantees -745 < x < 709 roughly (clamped here to the binary32 range). */
YDM_HD double ydm_exp_core(double x) {
    if (x > 89.0) x = 89.0;      /* (float) of the result is +inf above 88.7228... */
    if (x < -104.0) return 0.0;  /* (float) of the result is 0 below -103.97... */
    const double inv_ln2_32 = 46.166241308446828; /* 32/ln2 */
    const double ln2_32 = 0.021660849392498291;   /* ln2/32 */
    double z = YDM_MUL(x, inv_ln2_32);
    double kd = ydm_rint(z);
    int k = (int)kd;
    double r = YDM_SUB(z, kd);          /* |r| <= 0.5, units of 1/32 octave */
    double w = YDM_MUL(r, ln2_32);      /* |w| <= 0.01083 */
    /* e^w, Taylor to w^6 (remainder < 4e-18 relative) */
    double p = YDM_ADD(0.0083333333333333332, YDM_MUL(w, 0.0013888888888888889));
    p = YDM_ADD(0.041666666666666664, YDM_MUL(w, p));
    p = YDM_ADD(0.16666666666666666, YDM_MUL(w, p));
    p = YDM_ADD(0.5, YDM_MUL(w, p));
    p = YDM_ADD(1.0, YDM_MUL(w, p));
    p = YDM_ADD(1.0, YDM_MUL(w, p));
    double t = ydm_t32(k & 31);
    int e = k >> 5; /* arithmetic shift: floor(k/32) */
    double scale = ydm_from_bits((unsigned long long)(e + 1023) << 52);
    return YDM_MUL(YDM_MUL(t, scale), p);
}

/* ln(x) for finite normal-or-subnormal binary64 x > 0 */
YDM_HD double ydm_log_core(double x) {
    unsigned long long b = ydm_to_bits(x);
    int e = (int)((b >> 52) & 0x7FF);
    if (e == 0) { /* binary64 subnormal: never produced from a binary32 input, handled for completeness */
        x = YDM_MUL(x, 18014398509481984.0); /* 2^54 */
        b = ydm_to_bits(x);
        e = (int)((b >> 52) & 0x7FF) - 54;
    }
    e -= 1023;
    double m = ydm_from_bits((b & 0x000FFFFFFFFFFFFFULL) | 0x3FF0000000000000ULL); /* [1,2) */
    if (m > 1.4142135623730951) {
        m = YDM_MUL(m, 0.5);
        e += 1;
    }
    double f = YDM_DIV(YDM_SUB(m, 1.0), YDM_ADD(m, 1.0)); /* |f| <= 0.1716 */
    double f2 = YDM_MUL(f, f);
    /* 2*atanh(f) = f * sum 2/(2n+1) f^(2n), n = 0..10 */
    double s = 0.095238095238095233;
    s = YDM_ADD(0.10526315789473684, YDM_MUL(f2, s));
    s = YDM_ADD(0.11764705882352941, YDM_MUL(f2, s));
    s = YDM_ADD(0.13333333333333333, YDM_MUL(f2, s));
    s = YDM_ADD(0.15384615384615385, YDM_MUL(f2, s));
    s = YDM_ADD(0.18181818181818182, YDM_MUL(f2, s));
    s = YDM_ADD(0.22222222222222221, YDM_MUL(f2, s));
    s = YDM_ADD(0.2857142857142857, YDM_MUL(f2, s));
    s = YDM_ADD(0.40000000000000002, YDM_MUL(f2, s));
    s = YDM_ADD(0.66666666666666663, YDM_MUL(f2, s));
    s = YDM_ADD(2.0, YDM_MUL(f2, s));
    s = YDM_MUL(f, s);
    return YDM_ADD(YDM_MUL((double)e, 0.69314718055994529), s);
}

/* sin and cos of binary64 x, |x| < ~1e5 */
YDM_HD void ydm_sincos_core(double x, double *sn, double *cs) {
    const double two_over_pi = 0.63661977236758138;
    const double pio2_hi = 1.5707963267341256;     /* 33 significant bits */
    const double pio2_lo = 6.0771005065061922e-11; /* pi/2 - pio2_hi */
    double kd = ydm_rint(YDM_MUL(x, two_over_pi));
    int k = (int)kd;
    double r = YDM_SUB(YDM_SUB(x, YDM_MUL(kd, pio2_hi)), YDM_MUL(kd, pio2_lo)); /* |r| <= pi/4 (+eps) */
    double r2 = YDM_MUL(r, r);
    /* sin r = r * (1 - r2/3! + r2^2/5! - ... + r2^8/17!) */
    double s = 2.8114572543455206e-15;                            /*  1/17! */
    s = YDM_ADD(-7.6471637318198164e-13, YDM_MUL(r2, s));         /* -1/15! */
    s = YDM_ADD(1.6059043836821613e-10, YDM_MUL(r2, s));          /*  1/13! */
    s = YDM_ADD(-2.505210838544172e-08, YDM_MUL(r2, s));          /* -1/11! */
    s = YDM_ADD(2.7557319223985893e-06, YDM_MUL(r2, s));          /*  1/9!  */
    s = YDM_ADD(-0.00019841269841269841, YDM_MUL(r2, s));         /* -1/7!  */
    s = YDM_ADD(0.0083333333333333332, YDM_MUL(r2, s));           /*  1/5!  */
    s = YDM_ADD(-0.16666666666666666, YDM_MUL(r2, s));            /* -1/3!  */
    s = YDM_ADD(1.0, YDM_MUL(r2, s));
    s = YDM_MUL(r, s);
    /* cos r = 1 - r2/2! + r2^2/4! - ... + r2^9/18! */
    double c = -1.5619206968586225e-16;                           /* -1/18! */
    c = YDM_ADD(4.7794773323873853e-14, YDM_MUL(r2, c));          /*  1/16! */
    c = YDM_ADD(-1.1470745597729725e-11, YDM_MUL(r2, c));         /* -1/14! */
    c = YDM_ADD(2.08767569878681e-09, YDM_MUL(r2, c));            /*  1/12! */
    c = YDM_ADD(-2.7557319223985888e-07, YDM_MUL(r2, c));         /* -1/10! */
    c = YDM_ADD(2.48015873015873e-05, YDM_MUL(r2, c));            /*  1/8!  */
    c = YDM_ADD(-0.0013888888888888889, YDM_MUL(r2, c));          /* -1/6!  */
    c = YDM_ADD(0.041666666666666664, YDM_MUL(r2, c));            /*  1/4!  */
    c = YDM_ADD(-0.5, YDM_MUL(r2, c));
    c = YDM_ADD(1.0, YDM_MUL(r2, c));
    switch (k & 3) {
        case 0: *sn = s; *cs = c; break;
        case 1: *sn = c; *cs = -s; break;
        case 2: *sn = -s; *cs = -c; break;
        default: *sn = -c; *cs = s; break;
    }
}

/* ---- binary32 entry points (the MathF.* call sites) ---- */

/* Explicit binary32 operations: FMA is an exactly defined IEEE operation (one rounding), so using it is not a
 * contraction; the host needs a hardware FMA (-mfma; the oracle's Makefile) or falls back to libm's exact fmaf. */
#if defined(__CUDA_ARCH__)
#define YDM_FMAF(a, b, c) __fmaf_rn((a), (b), (c))
#define YDM_MULF(a, b) __fmul_rn((a), (b))
#define YDM_ADDF(a, b) __fadd_rn((a), (b))
#else
#define YDM_FMAF(a, b, c) __builtin_fmaf((a), (b), (c))
#define YDM_MULF(a, b) ((a) * (b))
#define YDM_ADDF(a, b) ((a) + (b))
#endif
YDM_HD float ydm_f32_from_bits(uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(b);
#else
    float f;
    memcpy(&f, &b, sizeof f);
    return f;
#endif
}
YDM_HD uint32_t ydm_f32_to_bits(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t b;
    memcpy(&b, &f, sizeof b);
    return b;
#endif
}

/* e^x in binary32 arithmetic only (the hot one: four per a-trous tap, RaytraceRenderer.cs:694-697, and the one dependent
 * chain of the in-place pass): 20 operations, 11 deep, branch-free.  k = rint(x / ln 2) by the magic-number addition,
 * r = x - k ln 2 with a two-part ln 2 (|r| <= 0.3466), e^r = 1 + (r + r^2 g(r)) with g = Taylor to r^5 in Estrin form
 * (remainder < 0.1 ulp), scaled by 2^k in two exact steps so that subnormal results are rounded once.  Within 1 ulp of
 * the true value for every input (tools/check_expf.c compares all 2^32 inputs with the binary64 evaluation above: never
 * more than 1 ulp apart, equal on 99.6 %, monotonic) -- the accuracy class of a platform expf, which is what MathF.Exp forwards to. */
YDM_HD float ydm_expf_core(float x) {
    const float t = YDM_FMAF(x, 1.44269502f, 12582912.0f);                 /* 1.5 * 2^23 + rint(x / ln 2) */
    const float kf = YDM_ADDF(t, -12582912.0f);
    const int k = (int)(ydm_f32_to_bits(t) - 0x4B400000u);
    float r = YDM_FMAF(kf, -0.693145751953125f, x);                        /* ln 2 = 0x3F317200 + 0x35BFBE8E */
    r = YDM_FMAF(kf, -1.42860676533018e-06f, r);
    const float r2 = YDM_MULF(r, r);
    const float h01 = YDM_FMAF(r, 0.166666672f, 0.5f);
    const float h23 = YDM_FMAF(r, 0.00833333377f, 0.0416666679f);
    const float h45 = YDM_FMAF(r, 0.000198412701f, 0.00138888892f);
    const float u = YDM_FMAF(r2, h45, h23);
    const float g = YDM_FMAF(r2, u, h01);
    const float q = YDM_FMAF(r2, g, r);
    const float p = YDM_ADDF(1.0f, q);
    const int k1 = k >> 1, k2 = k - k1;                                    /* both scale factors are normal numbers */
    const float s1 = ydm_f32_from_bits((uint32_t)(k1 + 127) << 23), s2 = ydm_f32_from_bits((uint32_t)(k2 + 127) << 23);
    return YDM_MULF(YDM_MULF(p, s1), s2);
}
YDM_HD float ycge_expf(float x) {
    float res = ydm_expf_core(x);
    res = (x > 88.7228317f) ? ydm_f32_from_bits(0x7F800000u) : res;        /* largest finite result: x = 0x42B17217 */
    res = (x < -104.0f) ? 0.0f : res;                                      /* e^-104 < 2^-150: rounds to 0 */
    return (x != x) ? x : res;
}
/* the same function for callers that know x <= 0 or NaN (every a-trous weight is exp(-d / phi), d >= 0): one select less;
 * tools/check_expf.c compares it with ycge_expf on all 2^31 + NaN inputs of that domain */
YDM_HD float ycge_expf_nonpos(float x) {
    float res = ydm_expf_core(x);
    res = (x < -104.0f) ? 0.0f : res;
    return (x != x) ? x : res;
}

YDM_HD float ycge_logf(float x) {
    if (x != x) return x;
    if (x < 0.0f) return (float)ydm_from_bits(0x7FF8000000000000ULL); /* NaN */
    if (x == 0.0f) return (float)ydm_from_bits(0xFFF0000000000000ULL); /* -inf */
    double xd = (double)x;
    if (ydm_to_bits(xd) == 0x7FF0000000000000ULL) return x; /* +inf */
    return (float)ydm_log_core(xd);
}

YDM_HD float ycge_powf(float x, float y) {
    if (y == 0.0f) return 1.0f;
    if (x == 1.0f) return 1.0f;
    if (x != x || y != y) return x + y;
    double ax = (double)x;
    int neg = 0;
    if (x < 0.0f || (x == 0.0f && (ydm_to_bits(ax) >> 63))) {
        /* negative base: defined only for integer exponents */
        double yd = (double)y;
        double yi = ydm_rint(yd);
        int is_int = (yd > -4503599627370496.0 && yd < 4503599627370496.0) ? (yi == yd) : 1;
        if (x < 0.0f && !is_int) return (float)ydm_from_bits(0x7FF8000000000000ULL);
        if (is_int && yd > -9007199254740992.0 && yd < 9007199254740992.0) {
            double h = YDM_MUL(yd, 0.5);
            neg = (ydm_rint(h) != h); /* odd */
        }
        ax = -ax;
    }
    float res;
    if (ax == 0.0) {
        res = (y > 0.0f) ? 0.0f : (float)ydm_from_bits(0x7FF0000000000000ULL);
    } else if (ydm_to_bits(ax) == 0x7FF0000000000000ULL) {
        res = (y > 0.0f) ? (float)ax : 0.0f;
    } else {
        double z = YDM_MUL((double)y, ydm_log_core(ax));
        res = (float)ydm_exp_core(z);
    }
    return neg ? -res : res;
}

YDM_HD void ycge_sincosf(float x, float *s, float *c) {
    double sd, cd;
    ydm_sincos_core((double)x, &sd, &cd);
    *s = (float)sd;
    *c = (float)cd;
}
YDM_HD float ycge_sinf(float x) {
    double sd, cd;
    ydm_sincos_core((double)x, &sd, &cd);
    return (float)sd;
}
YDM_HD float ycge_cosf(float x) {
    double sd, cd;
    ydm_sincos_core((double)x, &sd, &cd);
    return (float)cd;
}
YDM_HD float ycge_tanf(float x) {
    double sd, cd;
    ydm_sincos_core((double)x, &sd, &cd);
    return (float)YDM_DIV(sd, cd);
}

#endif /* YCGE_DETMATH_H */
