// Per-sample prior pieces of the hierarchical GLM targets, shared by the
// stand-alone k_glm_pre (glm.cu) and the sampling kernel that fuses it (family.cu).
#pragma once

#include "avi_internal.cuh"
#include "device_utils.cuh"

// {log prior, 1 / sigma^2, d log pi / d eta, |beta|^2} for theta = [beta(d); eta], sigma = exp(eta)
__device__ __forceinline__ float4 glm_prior_terms(float bsq, float eta, int d, int variant, int include_prior) {
    const float LOG3 = 1.0986122886681098f;
    const float s2 = expf(2.0f * eta), inv = 1.0f / s2;
    float lp = -0.5f * (float)d * AVI_LOG2PI - (float)d * eta - 0.5f * bsq * inv - LOG3 - 0.5f * AVI_LOG2PI;
    float ge = -(float)d + bsq * inv;
    if (variant == AVI_GLM_SUBSAMPLING) { lp -= s2 / 18.0f; ge -= s2 / 9.0f; }   // logpdf(Normal(0, 3), sigma)
    else { lp -= eta * eta / 18.0f; ge -= eta / 9.0f; }                          // LogNormal(0, 3) + log-Jacobian
    return include_prior ? make_float4(lp, inv, ge, bsq) : make_float4(0.f, 0.f, 0.f, bsq);
}
