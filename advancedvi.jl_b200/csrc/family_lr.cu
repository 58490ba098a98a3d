// Low-rank location-scale Gaussian (src/families/location_scale_low_rank.jl):
//   z = scale_diag .* u_diag + scale_factors * u_fact + location          (:79-86)
//   lambda = [location (D); scale_diag (D); vec(scale_factors) (D x r, column-major)]   (Functors order, :26)
// RepGradELBO + ClosedFormEntropy only.  With g_m = grad log pi(z_m) the gradient of -ELBO is
//   d/d location      = -mean_m g_m
//   d/d scale_diag    = -mean_m g_m .* u_diag_m - dH/dD
//   d/d scale_factors = -mean_m g_m u_fact_m'   - dH/dU
// and the entropy (:34-43) H = D h0 + sum log D_i + logdet(B) / 2 with B = I + U' D^-2 U (r x r) gives, with
// W = D^-2 U:  dH/dU = W B^-1,  dH/dD_i = 1 / D_i - (U B^-1 U')_ii / D_i^3.  Only the r x r capacitance matrix is
// ever factored (one CTA, Gauss-Jordan in shared memory; B is symmetric positive definite: no pivoting).
#include "avi_internal.cuh"
#include "device_utils.cuh"

namespace {

constexpr int LR_MAX_RANK = 32;

// Z[m][i] = mu[i] + D[i] * u1[m][i] + sum_k U[i + D*k] * u2[m][k]; padding columns zero.  One CTA per sample.
__global__ void __launch_bounds__(256)
k_lr_affine(const float* __restrict__ lambda, int D, int r, int ld, int ldr, const float* __restrict__ E1,
            const float* __restrict__ E2, float* __restrict__ Z) {
    __shared__ float u2[LR_MAX_RANK];
    const int m = blockIdx.x;
    if (threadIdx.x < r) u2[threadIdx.x] = E2[(size_t)m * ldr + threadIdx.x];
    __syncthreads();
    const float* mu = lambda;
    const float* sd = lambda + D;
    const float* U = lambda + 2 * (size_t)D;
    for (int i = threadIdx.x; i < ld; i += blockDim.x) {
        float z = 0.0f;
        if (i < D) {
            z = fmaf(__ldg(sd + i), E1[(size_t)m * ld + i], __ldg(mu + i));
            for (int k = 0; k < r; ++k) z = fmaf(__ldg(U + (size_t)k * D + i), u2[k], z);
        }
        Z[(size_t)m * ld + i] = z;
    }
}

// ent = [H | dH/dD (D) | dH/dU (D x r, column-major)] from lambda.  One CTA of 1024 threads.
__global__ void __launch_bounds__(1024)
k_lr_entropy(const float* __restrict__ lambda, int D, int r, float* __restrict__ ent) {
    __shared__ float B[LR_MAX_RANK][LR_MAX_RANK + 1];      // capacitance matrix, overwritten by its inverse
    __shared__ float Binv[LR_MAX_RANK][LR_MAX_RANK + 1];
    __shared__ float sm[33];
    __shared__ float logdetB;
    const float* sd = lambda + D;
    const float* U = lambda + 2 * (size_t)D;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // B[k][l] = delta_kl + sum_i U[i][k] U[i][l] / D_i^2 : warp w handles the pairs w, w + 32, ...
    for (int pr = w; pr < r * r; pr += 32) {
        const int k = pr / r, l = pr % r;
        float s = 0.0f;
        for (int i = lane; i < D; i += 32) {
            const float d = __ldg(sd + i);
            s = fmaf(__ldg(U + (size_t)k * D + i) / (d * d), __ldg(U + (size_t)l * D + i), s);
        }
        s = warp_sum(s);
        if (lane == 0) { B[k][l] = s + (k == l ? 1.0f : 0.0f); Binv[k][l] = k == l ? 1.0f : 0.0f; }
    }
    __syncthreads();
    // Gauss-Jordan on [B | I] without pivoting; log det = sum of the log pivots.  Thread (row, col) = (tid / 32, lane).
    float ld_acc = 0.0f;
    for (int p = 0; p < r; ++p) {
        const float piv = B[p][p];
        __syncthreads();
        if (tid == 0) ld_acc += logf(piv);
        const int row = w, col = lane;
        float f = 0.0f, bp = 0.0f, ip = 0.0f;
        if (row < r && col < r) { f = B[row][p] / piv; bp = B[p][col]; ip = Binv[p][col]; }
        __syncthreads();
        if (row < r && col < r) {
            if (row == p) { B[row][col] = bp / piv; Binv[row][col] = ip / piv; }
            else { B[row][col] -= f * bp; Binv[row][col] -= f * ip; }
        }
        __syncthreads();
    }
    if (tid == 0) logdetB = ld_acc;
    // sum log D_i
    float part = 0.0f;
    for (int i = tid; i < D; i += blockDim.x) part += logf(__ldg(sd + i));
    const float sumlog = block_sum(part, sm);
    if (tid == 0) ent[0] = (float)D * AVI_H0 + sumlog + 0.5f * logdetB;
    // dH/dU[i][l] = sum_k (U[i][k] / D_i^2) Binv[k][l];  dH/dD_i = 1 / D_i - (sum_kl U[i][k] Binv[k][l] U[i][l]) / D_i^3
    float* gD = ent + 1;
    float* gU = ent + 1 + D;
    for (int i = tid; i < D; i += blockDim.x) {
        const float d = __ldg(sd + i), inv2 = 1.0f / (d * d);
        float quad = 0.0f;
        for (int l = 0; l < r; ++l) {
            float t = 0.0f;
            for (int k = 0; k < r; ++k) t = fmaf(__ldg(U + (size_t)k * D + i), Binv[k][l], t);
            gU[(size_t)l * D + i] = t * inv2;
            quad = fmaf(t, __ldg(U + (size_t)l * D + i), quad);
        }
        gD[i] = 1.0f / d - quad * inv2 / d;
    }
}

// grad = [-v0 / M | -v1 / M - dH/dD | -CU / M - dH/dU], out = {value = -(mean log pi + H), elbo, H, shift unchanged}
__global__ void __launch_bounds__(256)
k_lr_finalize(const float* __restrict__ acc, int accv, const float* __restrict__ CU, const float* __restrict__ ent,
              int D, int r, int M, float* __restrict__ grad, float* __restrict__ out) {
    const float invM = 1.0f / (float)M;
    const long long P = 2LL * D + (long long)D * r;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        float g;
        if (p < D) g = -acc[p] * invM;
        else if (p < 2LL * D) g = -acc[(size_t)accv + (p - D)] * invM - ent[1 + (p - D)];
        else g = -CU[p - 2LL * D] * invM - ent[1 + D + (p - 2LL * D)];
        grad[p] = g;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const float* scal = acc + 4 * (size_t)accv;
        const float elbo = scal[0] * invM + ent[0];
        out[0] = -elbo; out[1] = elbo; out[2] = ent[0];
    }
}

}  // namespace

int avi_lr_max_rank() { return LR_MAX_RANK; }

int32_t avi_lr_affine(avi_obj* o, const float* lambda, const float* E1, const float* E2, float* Z, int Mloc) {
    k_lr_affine<<<Mloc, 256, 0, o->ctx->stream>>>(lambda, o->D, o->rank, o->ld, o->ldr, E1, E2, Z);
    AVI_LAUNCHED(o->ctx);
    return AVI_OK;
}

int32_t avi_lr_entropy(avi_obj* o, const float* lambda) {
    k_lr_entropy<<<1, 1024, 0, o->ctx->stream>>>(lambda, o->D, o->rank, o->lr_ent);
    AVI_LAUNCHED(o->ctx);
    return AVI_OK;
}

int32_t avi_lr_finalize(avi_obj* o, float* grad, float* out) {
    const float* CU = o->acc + 4 * (size_t)o->accv + ACC_NSCAL;
    const unsigned nb = (unsigned)std::min<int64_t>(ceil_div(o->P, 256), 296);
    k_lr_finalize<<<nb, 256, 0, o->ctx->stream>>>(o->acc, o->accv, CU, o->lr_ent, o->D, o->rank, o->M, grad, out);
    AVI_LAUNCHED(o->ctx);
    return AVI_OK;
}
