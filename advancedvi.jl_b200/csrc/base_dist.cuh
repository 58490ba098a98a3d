// Base distribution of MvLocationScale(location, scale, dist) (src/families/location_scale.jl:15-19): the univariate
// `dist` whose iid draws u give z = scale * u + location (:71-87), whose log-density gives logpdf(q, z) =
// sum_i logpdf(dist, u_i) - logdet(scale) (:59-63) and whose entropy gives entropy(q) = D * entropy(dist) + logdet(scale)
// (:52-57).  Normal(0, 1) is what MeanFieldGaussian / FullRankGaussian use; docs/src/families.md:72-101 also runs
// TDist(nu) and Laplace(0, 1).  Everything downstream of the sampler needs just three things from the base:
//   base_nl2(u)       = -2 log phi(u) - log(2 pi)   (== u^2 for the Gaussian, so sum_i base_nl2(u_i) plays |eps|^2's role
//                                                     in every log q(z) formula of the path)
//   base_negscore(u)  = -d log phi(u) / du          (== u for the Gaussian: the `eps` of the sticking-the-landing and
//                                                     score-function terms)
//   h0                = entropy(dist)
#pragma once

#include <math.h>

#include "device_utils.cuh"

#include "../../include/avi.h"   // AVI_BASE_*

enum { AVI_STREAM_EPS_B = 4 };   // second Philox stream of the Student-t draws (two variates per Philox block)

struct BaseDist {
    int kind = AVI_BASE_NORMAL;
    float nu = 0.0f;        // Student-t: degrees of freedom
    float h0 = AVI_H0;      // entropy(dist)
    float nl2_c = 0.0f;     // additive constant of base_nl2: -2 * (log normaliser) - log(2 pi)
};

#ifdef __CUDACC__
__device__ __forceinline__ float base_nl2(const BaseDist& b, float u) {
    if (b.kind == AVI_BASE_NORMAL) return u * u;
    if (b.kind == AVI_BASE_LAPLACE) return fmaf(2.0f, fabsf(u), b.nl2_c);
    return fmaf(b.nu + 1.0f, log1pf(u * u / b.nu), b.nl2_c);
}
__device__ __forceinline__ float base_negscore(const BaseDist& b, float u) {
    if (b.kind == AVI_BASE_NORMAL) return u;
    if (b.kind == AVI_BASE_LAPLACE) return u > 0.0f ? 1.0f : (u < 0.0f ? -1.0f : 0.0f);
    return (b.nu + 1.0f) * u / (b.nu + u * u);
}

// ln(u) for u in (0, 1] from a 23-bit uniform: lg2.approx has a fixed absolute error, so near 1 the series of ln(1 + t),
// t = u - 1 (exact), takes over (box_muller_fast's radius uses the same switch)
__device__ __forceinline__ float ln_unit(float u) {
    const float t = u - 1.0f;
    const float series = t * (1.0f + t * (-0.5f + t * (0.33333334f + t * (-0.25f + t * 0.2f))));
    float l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u));
    return u > 0.96875f ? series : 0.6931471805599453f * l2;
}

// Laplace(0, 1) by inversion of the CDF: p = uniform23(x) in (0, 1);  u = ln(2 p) for p < 1/2, -ln(2 (1 - p)) otherwise
// (2 p and 2 (1 - p) are exact in fp32).  oracle/philox.py: laplace_matrix.
__device__ __forceinline__ float laplace_from_word(uint32_t x) {
    const float p = uniform23(x);
    const bool lower = p < 0.5f;
    const float w = lower ? 2.0f * p : 2.0f * (1.0f - p);
    const float l = ln_unit(w);
    return lower ? l : -l;
}

// Student-t(nu) without rejection (Bailey 1994, the polar form of Box-Muller): with U, V uniform,
//   t = sqrt(nu * (U^(-2/nu) - 1)) * cos(2 pi V)   is exactly t_nu distributed.
// (The sine partner is t_nu too but NOT independent of the cosine one, so a pair of words yields one variate.)
// a = -(2/nu) ln U >= 0; exp(a) - 1 by its series below 1/4 (the square root would amplify ex2.approx's absolute error).
// oracle/philox.py: student_t_matrix.
__device__ __forceinline__ float student_t_from_words(uint32_t x0, uint32_t x1, float nu) {
    const float a = -(2.0f / nu) * ln_unit(uniform23(x0));
    const float series = a * (1.0f + a * (0.5f + a * (0.16666667f + a * (0.041666668f + a * (0.008333334f + a * 0.0013888889f)))));
    const float big = exp2f(a * 1.4426950408889634f) - 1.0f;
    const float r = sqrtf(nu * (a < 0.25f ? series : big));
    const float vsf = __uint_as_float(0x4b000000u | ((x1 >> 9) ^ 0x400000u)) - 12582912.0f;   // signed 23-bit reading (box_muller_fast)
    const float th = fmaf(vsf, 6.283185307179586f * 1.1920928955078125e-07f, 6.283185307179586f * 5.9604644775390625e-08f);
    return r * __cosf(th);
}

// four iid base draws for coordinates 4q .. 4q+3 of Monte-Carlo sample m
__device__ __forceinline__ float4 base_draw4(const BaseDist& b, uint32_t q, uint32_t m, uint32_t c2, uint32_t c3,
                                             const PhiloxKeys& pk) {
    if (b.kind == AVI_BASE_NORMAL) return normal4(q, m, c2, c3, pk);
    uint32_t x[4];
    philox4x32_10(q, m, c2, c3, pk, x);
    float4 e;
    if (b.kind == AVI_BASE_LAPLACE) {
        e.x = laplace_from_word(x[0]); e.y = laplace_from_word(x[1]);
        e.z = laplace_from_word(x[2]); e.w = laplace_from_word(x[3]);
        return e;
    }
    uint32_t y[4];
    philox4x32_10(q, m, c2, (c3 & ~0xFFu) | (uint32_t)AVI_STREAM_EPS_B, pk, y);
    e.x = student_t_from_words(x[0], x[1], b.nu); e.y = student_t_from_words(x[2], x[3], b.nu);
    e.z = student_t_from_words(y[0], y[1], b.nu); e.w = student_t_from_words(y[2], y[3], b.nu);
    return e;
}
#endif

// host: fill h0 and nl2_c for (kind, nu); false if the parameters are invalid
static inline double avi_digamma(double x) {
    double r = 0.0;
    while (x < 6.0) { r -= 1.0 / x; x += 1.0; }
    const double f = 1.0 / (x * x);
    return r + log(x) - 0.5 / x - f * (1.0 / 12.0 - f * (1.0 / 120.0 - f * (1.0 / 252.0 - f * (1.0 / 240.0 - f / 132.0))));
}
static inline bool avi_base_make(int kind, float nu, BaseDist* out) {
    const double LOG2PI = 1.8378770664093453;
    BaseDist b;
    b.kind = kind; b.nu = nu;
    if (kind == AVI_BASE_NORMAL) { b.nu = 0.0f; }
    else if (kind == AVI_BASE_LAPLACE) {
        b.nu = 0.0f;
        b.h0 = (float)(1.0 + log(2.0));                 // entropy(Laplace(0, 1))
        b.nl2_c = (float)(2.0 * log(2.0) - LOG2PI);     // log phi(u) = -|u| - log 2
    } else if (kind == AVI_BASE_STUDENT_T) {
        if (!(nu > 0.0f) || !(nu < 1e6f)) return false;
        const double v = nu;
        const double c = lgamma((v + 1.0) / 2.0) - lgamma(v / 2.0) - 0.5 * log(v * 3.141592653589793);   // log normaliser
        // entropy(TDist(nu)) = (nu+1)/2 (digamma((nu+1)/2) - digamma(nu/2)) + log(sqrt(nu) B(nu/2, 1/2))
        const double lbeta = lgamma(v / 2.0) + lgamma(0.5) - lgamma((v + 1.0) / 2.0);
        b.h0 = (float)((v + 1.0) / 2.0 * (avi_digamma((v + 1.0) / 2.0) - avi_digamma(v / 2.0)) + 0.5 * log(v) + lbeta);
        b.nl2_c = (float)(-2.0 * c - LOG2PI);
    } else {
        return false;
    }
    *out = b;
    return true;
}
