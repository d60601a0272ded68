// gvc_expf.h -- bit-faithful expf for the exact-mode sigmoid.
//
// The reference's sigmoid::forward (reference src/gnn_inference.cpp:49-52) calls
// glibc's expf, which is NOT correctly rounded, and CUDA's expf differs from it
// in the last bits, so neither the device expf nor (float)exp((double)x) gives
// bit-equal scores.  This is the published algorithm of glibc's expf (2.27+,
// "exp2f_data": 32-entry table of 2^(i/32), degree-3 polynomial, all in double)
// restated for host and device; SURVEY.md App. B measured 0 mismatches in 50 M
// arguments against libm.  tests/test_host_expf.py re-checks that on the host
// build of this very header.
//
// Table entries are bits(2^(i/32)) - (i << 47), derived independently with
// 60-digit arithmetic (tools/make_expf_table.py) -- not copied.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define GVC_HD __host__ __device__ __forceinline__
#else
#define GVC_HD static inline
#include <string.h>
#endif

#define GVC_EXP2F_TAB {\
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL,\
    0x3fef72b83c7d517bULL, 0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL,\
    0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL, 0x3feedea64c123422ULL, 0x3feece086061892dULL,\
    0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL, 0x3feea47eb03a5585ULL,\
    0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,\
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL,\
    0x3feee89f995ad3adULL, 0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL,\
    0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL, 0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL,}
static const uint64_t gvc_exp2f_tab_host[32] = GVC_EXP2F_TAB;
#if defined(__CUDACC__)
static __device__ __constant__ uint64_t gvc_exp2f_tab_dev[32] = GVC_EXP2F_TAB;
#endif

GVC_HD uint32_t gvc_f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}
GVC_HD uint64_t gvc_d2u(double d) {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(d);
#else
    uint64_t u; memcpy(&u, &d, 8); return u;
#endif
}
GVC_HD float gvc_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}
GVC_HD double gvc_u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double d; memcpy(&d, &u, 8); return d;
#endif
}

// expf with glibc's operation order.  Products and sums are written so that
// neither nvcc nor gcc may contract them differently from each other: the
// device path uses explicit __dmul_rn/__dadd_rn/__fma_rn, the host path must be
// compiled with -ffp-contract=off.  (Whether glibc's own build fuses the
// polynomial is immaterial: the double result is rounded to float at the end
// and SURVEY.md App. B found both variants bit-equal to libm on 50 M inputs.)
GVC_HD float gvc_expf_glibc(float x) {
    const double InvLn2N = 0x1.71547652b82fep+0 * 32.0;
    const double Shift = 0x1.8p52;
    const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0;
    const double C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0;
    const double C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
    uint32_t ux = gvc_f2u(x);
    uint32_t abstop = (ux >> 20) & 0x7ff;
    if (abstop >= 0x42b) {                       // |x| >= 88 or NaN (top12(88.0f) = 0x42b)
        if (ux == 0xff800000u) return 0.0f;      // -inf
        if (abstop >= 0x7f8) return x + x;       // inf / NaN
        if (x > 0x1.62e42ep6f) return gvc_u2f(0x7f800000u);   // overflow -> +inf
        if (x < -0x1.9fe368p6f) return 0.0f;     // underflow -> +0
    }
    double xd = (double)x;
#if defined(__CUDA_ARCH__)
    double z = __dmul_rn(InvLn2N, xd);
    double kd = __dadd_rn(z, Shift);
    uint64_t ki = gvc_d2u(kd);
    kd = __dsub_rn(kd, Shift);
    double r = __dsub_rn(z, kd);
    uint64_t t = gvc_exp2f_tab_dev[ki & 31];
    t += ki << 47;
    double s = gvc_u2d(t);
    double zz = __dadd_rn(__dmul_rn(C0, r), C1);
    double r2 = __dmul_rn(r, r);
    double y = __dadd_rn(__dmul_rn(C2, r), 1.0);
    y = __dadd_rn(__dmul_rn(zz, r2), y);
    y = __dmul_rn(y, s);
    return (float)y;
#else
    double z = InvLn2N * xd;
    volatile double kdv = z + Shift;             // must be rounded to double (glibc: math_narrow_eval)
    double kd = kdv;
    uint64_t ki = gvc_d2u(kd);
    kd -= Shift;
    double r = z - kd;
    uint64_t t = gvc_exp2f_tab_host[ki & 31];
    t += ki << 47;
    double s = gvc_u2d(t);
    double zz = C0 * r + C1;
    double r2 = r * r;
    double y = C2 * r + 1.0;
    y = zz * r2 + y;
    y = y * s;
    return (float)y;
#endif
}
