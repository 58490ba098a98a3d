// Mean-field closed-form gradient of the value slot from the reduced sums,
// shared by the stand-alone finalize kernel (family.cu) and the fused finalize + update kernel (opt.cu).
#pragma once

#include "avi_internal.cuh"
#include "device_utils.cuh"

// sums over the GLOBAL M samples: s0 = sum logp, s1 = sum |eps|^2, s2 = sum f, s3 = sum f^2 (f shifted)
struct MfSums {
    float s0, s1, s2, s3, logdet;
    float h0 = AVI_H0;   // entropy of the base distribution (base_dist.cuh); Normal(0, 1) unless the caller overrides it
};

// gradient entries (d/d mu_i, d/d s_i) of -ELBO (RepGrad) or of the VarGrad value (ScoreGrad) from the
// four reduced sums of coordinate i (layout: avi_internal.cuh)
__device__ __forceinline__ void mf_grad_vals(float v0, float v1, float v2, float v3, float si, int M, int objective,
                                             int entropy, const MfSums& S, float& gm, float& gs) {
    const float invM = 1.0f / (float)M, inv = 1.0f / si;
    if (objective == AVI_REPGRAD) {
        float sg = v0, sge = v1;
        if (entropy == AVI_ENT_STL || entropy == AVI_ENT_STL_ZEROGRAD) {   // w = g + eps / s
            sg = fmaf(v2, inv, sg);
            sge = fmaf(v3, inv, sge);
        }
        gm = -sg * invM; gs = -sge * invM;
        if (entropy == AVI_ENT_CLOSEDFORM || entropy == AVI_ENT_MONTECARLO) gs -= inv;
        else if (entropy == AVI_ENT_STL_ZEROGRAD) gs += inv;
    } else {
        const float fbar = S.s2 * invM;
        gm = (v0 - fbar * v2) * invM * inv;
        gs = (v1 - fbar * v3) * invM * inv;
    }
}

__device__ __forceinline__ void mf_grad_entry(const float* __restrict__ acc, int accv, float si, int i, int M,
                                              int objective, int entropy, const MfSums& S, float& gm, float& gs) {
    const bool need23 = objective == AVI_SCOREGRAD || entropy == AVI_ENT_STL || entropy == AVI_ENT_STL_ZEROGRAD;
    mf_grad_vals(acc[i], acc[accv + i], need23 ? acc[2 * (size_t)accv + i] : 0.f,
                 need23 ? acc[3 * (size_t)accv + i] : 0.f, si, M, objective, entropy, S, gm, gs);
}

// value slot, elbo and the ScoreGrad centring shift for the next call; `shift` = current out[3]
__device__ __forceinline__ void mf_outputs(int D, int M, int objective, int entropy, const MfSums& S, float shift,
                                           float& value, float& elbo, float& shift_next) {
    const float invM = 1.0f / (float)M;
    if (objective == AVI_REPGRAD) {
        const float ent = (entropy == AVI_ENT_CLOSEDFORM || entropy == AVI_ENT_CLOSEDFORM_ZEROGRAD)
                              ? (float)D * S.h0 + S.logdet
                              : 0.5f * S.s1 * invM + 0.5f * (float)D * AVI_LOG2PI + S.logdet;
        value = -(S.s0 * invM + ent);
        elbo = -value;
        shift_next = shift;
    } else {
        const float fbar = S.s2 * invM;
        value = 0.5f * (S.s3 * invM - fbar * fbar);   // VarGrad value (shift-invariant)
        elbo = -(fbar + shift);                       // mean(log pi - log q)
        shift_next = fbar + shift;                    // centre f for the next call
    }
}

// Collect the global sums.  deferred != 0: RepGrad on one rank (or row-sharded): sum logp and sum |eps|^2
// are taken here from the per-sample vectors instead of by a separate k_scalars launch.
// Every thread of the CTA receives the result; `sm` holds 33 floats.
__device__ __forceinline__ MfSums mf_collect_sums(const float* __restrict__ lambda, int D,
                                                  const float* __restrict__ scal, const float* __restrict__ logp,
                                                  const float* __restrict__ esq, int Mloc, int deferred, float* sm) {
    MfSums S;
    float part = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) part += logf(__ldg(lambda + D + i));
    S.logdet = block_sum(part, sm);
    if (deferred) {
        float a = 0.f, b = 0.f;
        for (int m = threadIdx.x; m < Mloc; m += blockDim.x) { a += logp[m]; b += esq[m]; }
        S.s0 = block_sum(a, sm); S.s1 = block_sum(b, sm); S.s2 = 0.f; S.s3 = 0.f;
    } else {
        S.s0 = scal[0]; S.s1 = scal[1]; S.s2 = scal[2]; S.s3 = scal[3];
    }
    return S;
}
