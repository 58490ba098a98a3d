// Full-rank closed-form gradient of the scale entries, shared by the stand-alone
// finalize kernel (family.cu) and the fused finalize + update kernel (opt.cu).
#pragma once

#include "avi_internal.cuh"

// entry (i, j), i >= j, of d value / d L;  C*[j*D + i] = sum_m W[m][i] E[m][j]  (column-major L layout)
__device__ __forceinline__ float fr_grad_entry(const float* __restrict__ C1, const float* __restrict__ C2,
                                               const float* __restrict__ scal, float l_ij, size_t idx, int i, int j,
                                               int M, int objective, int entropy) {
    const float invM = 1.0f / (float)M;
    if (objective == AVI_REPGRAD) {
        float g = -C1[idx] * invM;
        if (i == j) {
            const float inv = 1.0f / l_ij;
            if (entropy == AVI_ENT_CLOSEDFORM || entropy == AVI_ENT_MONTECARLO) g -= inv;
            else if (entropy == AVI_ENT_STL_ZEROGRAD) g += inv;
        }
        return g;
    }
    const float fbar = scal[2] * invM;
    return (C1[idx] - fbar * C2[idx]) * invM;
}
