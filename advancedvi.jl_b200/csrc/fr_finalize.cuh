// Full-rank closed-form gradient of the scale entries, shared by the stand-alone
// finalize kernel (family.cu) and the fused finalize + update kernel (opt.cu).
#pragma once

#include "avi_internal.cuh"
#include "device_utils.cuh"

// entry (i, j), i >= j, of d value / d L from the two contraction entries c1 = C1[idx], c2 = C2[idx] (ScoreGrad only);
// C*[j*D + i] = sum_m W[m][i] E[m][j]  (column-major L layout)
__device__ __forceinline__ float fr_grad_value(float c1, float c2, const float* __restrict__ scal, float l_ij, bool diag,
                                               int M, int objective, int entropy) {
    const float invM = 1.0f / (float)M;
    if (objective == AVI_REPGRAD) {
        float g = -c1 * invM;
        if (diag) {
            const float inv = 1.0f / l_ij;
            if (entropy == AVI_ENT_CLOSEDFORM || entropy == AVI_ENT_MONTECARLO) g -= inv;
            else if (entropy == AVI_ENT_STL_ZEROGRAD) g += inv;
        }
        return g;
    }
    const float fbar = scal[2] * invM;
    return (c1 - fbar * c2) * invM;
}
__device__ __forceinline__ float fr_grad_entry(const float* __restrict__ C1, const float* __restrict__ C2,
                                               const float* __restrict__ scal, float l_ij, size_t idx, int i, int j,
                                               int M, int objective, int entropy) {
    return fr_grad_value(C1[idx], objective == AVI_REPGRAD ? 0.0f : C2[idx], scal, l_ij, i == j, M, objective, entropy);
}

// Location block of the gradient + value / ELBO / log det of the full-rank family, one CTA (fixed summation order).
//   deferred  : RepGrad only -- sum_m logp and sum_m |eps_m|^2 are taken from the per-sample vectors here instead of from
//               a k_scalars launch (they feed only the value slot)
//   write_grad: false when the column-sum stage already wrote grad[0 .. D) itself (family_fr.cu: k_fr_outer_prep)
// `sm` must hold 33 floats.
__device__ __forceinline__ void fr_vec_finalize(const float* __restrict__ acc, int accv, const float* __restrict__ lambda,
                                                int D, int M, int objective, int entropy, float* __restrict__ grad,
                                                float* __restrict__ out, const float* __restrict__ logp,
                                                const float* __restrict__ esq, int Mloc, bool deferred, bool write_grad,
                                                float* sm, float h0 = AVI_H0) {
    float part = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) part += logf(__ldg(lambda + D + (size_t)i * (D + 1)));
    const float logdet = block_sum(part, sm);
    const float* scal = acc + 4 * (size_t)accv;
    const float invM = 1.0f / (float)M;
    if (objective == AVI_REPGRAD) {
        float s0, s1;
        if (deferred) {
            float a = 0.f, b = 0.f;
            for (int m = threadIdx.x; m < Mloc; m += blockDim.x) { a += logp[m]; b += esq[m]; }
            s0 = block_sum(a, sm); s1 = block_sum(b, sm);
        } else {
            s0 = scal[0]; s1 = scal[1];
        }
        if (write_grad)
            for (int i = threadIdx.x; i < D; i += blockDim.x) grad[i] = -acc[i] * invM;
        if (threadIdx.x == 0) {
            float ent = (entropy == AVI_ENT_CLOSEDFORM || entropy == AVI_ENT_CLOSEDFORM_ZEROGRAD)
                            ? (float)D * h0 + logdet
                            : 0.5f * s1 * invM + 0.5f * (float)D * AVI_LOG2PI + logdet;
            float value = -(s0 * invM + ent);
            out[0] = value; out[1] = -value; out[2] = logdet;
        }
    } else {
        const float fbar = scal[2] * invM;
        const float* v2 = acc + 2 * (size_t)accv;
        for (int i = threadIdx.x; i < D; i += blockDim.x) grad[i] = (acc[i] - fbar * v2[i]) * invM;
        if (threadIdx.x == 0) {
            float shift = out[3];
            out[0] = 0.5f * (scal[3] * invM - fbar * fbar);
            out[1] = -(fbar + shift);
            out[2] = logdet;
            out[3] = fbar + shift;
        }
    }
}
