// Low-rank location-scale Gaussian (src/families/location_scale_low_rank.jl):
//   z = scale_diag .* u_diag + scale_factors * u_fact + location          (:79-86)
//   lambda = [location (D); scale_diag (D); vec(scale_factors) (D x r, column-major)]   (Functors order, :26)
// All of entropy.jl's estimators except the STL zero-gradient variant, and ScoreGradELBO (VarGrad).
// RepGradELBO + ClosedFormEntropy: with g_m = grad log pi(z_m) the gradient of -ELBO is
//   d/d location      = -mean_m g_m
//   d/d scale_diag    = -mean_m g_m .* u_diag_m - dH/dD
//   d/d scale_factors = -mean_m g_m u_fact_m'   - dH/dU
// and the entropy (:34-43) H = D h0 + sum log D_i + logdet(B) / 2 with B = I + U' D^-2 U (r x r) gives, with
// W = D^-2 U:  dH/dU = W B^-1,  dH/dD_i = 1 / D_i - (U B^-1 U')_ii / D_i^3.  Only the r x r capacitance matrix is
// ever factored (one CTA, Gauss-Jordan in shared memory; B is symmetric positive definite: no pivoting).
// The estimators that need log q(z) go through the same r x r inverse (Woodbury): per sample, y = z - mu,
//   w = Sigma^-1 y = D^-2 y - W B^-1 (U' D^-2 y),   log q(z) = -y'w / 2 - sum log D_i - logdet(B) / 2 - D log(2 pi) / 2,
//   d log q / d mu = w,   d log q / d D = D .* (w^2 - diag Sigma^-1),   d log q / d U = w (U'w)' - Sigma^-1 U
// with D .* diag Sigma^-1 = dH/dD and Sigma^-1 U = dH/dU (the entropy gradient above), so that
//   StickingTheLandingEntropy (entropy.jl:59-65)  g_m -> g_m + w_m in the ClosedFormEntropy formulas, no dH terms
//   MonteCarloEntropy         (entropy.jl:42-46)  the STL terms + [mean w, D .* mean w^2 - dH/dD, mean w (U'w)' - dH/dU]
//   ScoreGradELBO / VarGrad (scoregradelbo.jl:87-117)  mean_m c_m [w_m, D .* w_m^2, w_m (U'w_m)'],  c_m = f_m - mean f,
//                                                    f_m = log q(z_m) - log pi(z_m)  (the dH terms cancel: sum c_m = 0)
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
    // for the log q(z) kernel: B^-1 (row pitch LR_MAX_RANK), log det B, sum log D_i
    float* ext = ent + 1 + (size_t)D + (size_t)D * r;
    if (w < r && lane < r) ext[w * LR_MAX_RANK + lane] = Binv[w][lane];
    if (tid == 0) { ext[LR_MAX_RANK * LR_MAX_RANK] = logdetB; ext[LR_MAX_RANK * LR_MAX_RANK + 1] = sumlog; }
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

// Per sample (one CTA): w = Sigma^-1 (z - mu) -> Wm[m][.], v = U'w -> V[m][.], log q(z_m) -> lq[m]; Gadd != nullptr:
// Gadd[m][.] += w (the sticking-the-landing path term).  ext = [B^-1 | log det B | sum log D] from k_lr_entropy.
__global__ void __launch_bounds__(256)
k_lr_logq(const float* __restrict__ lambda, int D, int r, int ld, int ldr, const float* __restrict__ E1,
          const float* __restrict__ E2, const float* __restrict__ ext, float* __restrict__ Wm, float* __restrict__ V,
          float* __restrict__ lq, float* __restrict__ Gadd) {
    __shared__ float u2[LR_MAX_RANK], t[LR_MAX_RANK], sv[LR_MAX_RANK], sm[33];
    const int m = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const float* sd = lambda + D;
    const float* U = lambda + 2 * (size_t)D;
    float* wrow = Wm + (size_t)m * ld;
    if (tid < r) u2[tid] = E2[(size_t)m * ldr + tid];
    __syncthreads();
    // y / D^2 (kept in the output row for the two passes over U)
    for (int i = tid; i < ld; i += blockDim.x) {
        float yd = 0.0f;
        if (i < D) {
            const float d = __ldg(sd + i);
            float y = d * E1[(size_t)m * ld + i];
            for (int k = 0; k < r; ++k) y = fmaf(__ldg(U + (size_t)k * D + i), u2[k], y);
            yd = y / (d * d);
        }
        wrow[i] = yd;
    }
    __syncthreads();
    // t = U' D^-2 y
    for (int k = w; k < r; k += 8) {
        float a = 0.0f;
        for (int i = lane; i < D; i += 32) a = fmaf(__ldg(U + (size_t)k * D + i), wrow[i], a);
        a = warp_sum(a);
        if (lane == 0) t[k] = a;
    }
    __syncthreads();
    if (tid < r) {
        float a = 0.0f;
        for (int l = 0; l < r; ++l) a = fmaf(ext[tid * LR_MAX_RANK + l], t[l], a);
        sv[tid] = a;   // B^-1 t
    }
    __syncthreads();
    float quad = 0.0f;
    for (int i = tid; i < D; i += blockDim.x) {
        const float d = __ldg(sd + i), yd = wrow[i];
        float c = 0.0f;
        for (int k = 0; k < r; ++k) c = fmaf(__ldg(U + (size_t)k * D + i), sv[k], c);
        const float wi = yd - c / (d * d);
        quad = fmaf(yd * d * d, wi, quad);
        wrow[i] = wi;
        if (Gadd) Gadd[(size_t)m * ld + i] += wi;
    }
    quad = block_sum(quad, sm);   // (contains the barrier that publishes the w row)
    for (int k = w; k < r; k += 8) {
        float a = 0.0f;
        for (int i = lane; i < D; i += 32) a = fmaf(__ldg(U + (size_t)k * D + i), wrow[i], a);
        a = warp_sum(a);
        if (lane == 0) V[(size_t)m * ldr + k] = a;
    }
    if (tid == 0)
        lq[m] = -0.5f * quad - ext[LR_MAX_RANK * LR_MAX_RANK + 1] - 0.5f * ext[LR_MAX_RANK * LR_MAX_RANK] -
                0.5f * (float)D * AVI_LOG2PI;
}

// scal = {sum log pi, sum log q, sum f, sum f^2}; ScoreGrad: fbuf[m] = c_m = f_m - mean f.  One CTA: fixed order.
__global__ void __launch_bounds__(1024)
k_lr_stats(const float* __restrict__ logp, const float* __restrict__ lq, int Mloc, int score, float* __restrict__ fbuf,
           float* __restrict__ scal) {
    __shared__ float sm[33];
    float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
    for (int m = threadIdx.x; m < Mloc; m += blockDim.x) {
        const float lp = logp[m], q = lq[m], f = q - lp;
        a += lp; b += q; c += f; d = fmaf(f, f, d);
    }
    a = block_sum(a, sm); b = block_sum(b, sm); c = block_sum(c, sm); d = block_sum(d, sm);
    if (score) {
        const float fbar = c / (float)Mloc;
        for (int m = threadIdx.x; m < Mloc; m += blockDim.x) fbuf[m] = (lq[m] - logp[m]) - fbar;
    }
    if (threadIdx.x == 0) {
        scal[0] = a; scal[1] = b; scal[2] = c; scal[3] = d;
        scal[4] = 0.f; scal[5] = 0.f; scal[6] = 0.f; scal[7] = 0.f;
    }
}

// out0[i] = sum_m c_m A[m][i], out1[i] = sum_m c_m A[m][i] B[m][i]  (c == nullptr: c_m = 1); 32 coordinates x 32 sample groups
__global__ void __launch_bounds__(1024)
k_lr_reduce2(const float* __restrict__ A, const float* __restrict__ Bm, const float* __restrict__ c, int ld, int Mloc, int D,
             float* __restrict__ out0, float* __restrict__ out1) {
    __shared__ float sm[2][32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * 32 + tx;
    float v0 = 0.f, v1 = 0.f;
    if (i < D)
        for (int m = ty; m < Mloc; m += 32) {
            const float a = (c ? c[m] : 1.0f) * A[(size_t)m * ld + i];
            v0 += a; v1 = fmaf(a, Bm[(size_t)m * ld + i], v1);
        }
    sm[0][ty][tx] = v0; sm[1][ty][tx] = v1;
    __syncthreads();
    if (ty < 2 && i < D) {
        float s = 0.f;
#pragma unroll
        for (int q = 0; q < 32; ++q) s += sm[ty][q][tx];
        (ty ? out1 : out0)[i] = s;
    }
}

// out[m][i] = c[m] * W[m][i]
__global__ void k_lr_scale_rows(const float* __restrict__ W, const float* __restrict__ c, int ld, int Mloc, float* __restrict__ out) {
    const size_t n = (size_t)Mloc * ld;
    for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (size_t)gridDim.x * blockDim.x)
        out[p] = c[p / ld] * W[p];
}

// Gradient of the value slot in destructure order [location | scale_diag | vec(scale_factors)] and the scalars:
//   RepGrad  value = -(mean log pi + H^),  H^ = H (closed forms) or -mean log q (MonteCarlo / STL)
//     ClosedFormEntropy            [-v0/M | -v1/M - dH/dD | -CU/M - dH/dU]
//     ClosedFormEntropyZeroGradient, StickingTheLandingEntropy (v0, v1, CU then come from g + w): no dH terms
//     StickingTheLandingEntropyZeroGradient: StickingTheLandingEntropy's with + dH/dD, + dH/dU
//     MonteCarloEntropy            [(v2 - v0)/M | (D .* v3 - v1)/M - dH/dD | (CU2 - CU)/M - dH/dU]
//   ScoreGrad (v0, v1, CU weighted by c_m): [v0/M | D .* v1/M | CU/M], value = (mean f^2 - (mean f)^2) / 2
__global__ void __launch_bounds__(256)
k_lr_finalize(const float* __restrict__ acc, int accv, const float* __restrict__ CU, const float* __restrict__ CU2,
              const float* __restrict__ ent, const float* __restrict__ lambda, int D, int r, int M, int objective,
              int entropy, float* __restrict__ grad, float* __restrict__ out) {
    const float invM = 1.0f / (float)M;
    const long long P = 2LL * D + (long long)D * r;
    const bool score = objective == AVI_SCOREGRAD;
    const bool dH = !score && (entropy == AVI_ENT_CLOSEDFORM || entropy == AVI_ENT_MONTECARLO);
    const bool dHplus = !score && entropy == AVI_ENT_STL_ZEROGRAD;   // STL - H(q) + H(q_stop): + grad H (entropy.jl:80-90)
    const bool mc = !score && entropy == AVI_ENT_MONTECARLO;
    const float* v0 = acc; const float* v1 = acc + accv; const float* v2 = acc + 2 * (size_t)accv; const float* v3 = acc + 3 * (size_t)accv;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        float g;
        if (p < D) {
            g = score ? v0[p] * invM : (mc ? (v2[p] - v0[p]) * invM : -v0[p] * invM);
        } else if (p < 2LL * D) {
            const long long i = p - D;
            const float d = __ldg(lambda + D + i);
            g = score ? d * v1[i] * invM : (mc ? (d * v3[i] - v1[i]) * invM : -v1[i] * invM);
            if (dH) g -= ent[1 + i];
            if (dHplus) g += ent[1 + i];
        } else {
            const long long q = p - 2LL * D;
            g = score ? CU[q] * invM : (mc ? (CU2[q] - CU[q]) * invM : -CU[q] * invM);
            if (dH) g -= ent[1 + D + q];
            if (dHplus) g += ent[1 + D + q];
        }
        grad[p] = g;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        const float* scal = acc + 4 * (size_t)accv;
        if (score) {
            const float fbar = scal[2] * invM;
            out[0] = 0.5f * (scal[3] * invM - fbar * fbar);
            out[1] = (scal[0] - scal[1]) * invM;
            out[2] = 0.0f;
        } else {
            const bool closed = entropy == AVI_ENT_CLOSEDFORM || entropy == AVI_ENT_CLOSEDFORM_ZEROGRAD;
            const float H = closed ? ent[0] : -scal[1] * invM;
            const float elbo = scal[0] * invM + H;
            out[0] = -elbo; out[1] = elbo; out[2] = H;
        }
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

// the estimators that need log q(z): everything but RepGradELBO with a closed-form entropy
bool avi_lr_needs_logq(const avi_obj* o) {
    return o->objective == AVI_SCOREGRAD || o->entropy == AVI_ENT_MONTECARLO || o->entropy == AVI_ENT_STL ||
           o->entropy == AVI_ENT_STL_ZEROGRAD;
}

// w, U'w and log q per sample (into o->U, o->V, o->esq), the scalar sums, and for ScoreGrad the centred weights (o->fbuf).
// RepGrad + STL / MonteCarlo: G += w.  Needs avi_lr_entropy (B^-1) first.
// forward_only (estimate_objective): log q per sample into o->esq, nothing else.
int32_t avi_lr_logq(avi_obj* o, const float* lambda, int Mloc, bool forward_only) {
    avi_ctx* ctx = o->ctx;
    const float* ext = o->lr_ent + 1 + (size_t)o->D + (size_t)o->D * o->rank;
    const bool rep = o->objective == AVI_REPGRAD;
    k_lr_logq<<<Mloc, 256, 0, ctx->stream>>>(lambda, o->D, o->rank, o->ld, o->ldr, o->E, o->E2, ext, o->U, o->V, o->esq,
                                             (rep && !forward_only) ? o->G : nullptr);
    AVI_LAUNCHED(ctx);
    if (forward_only) return AVI_OK;
    k_lr_stats<<<1, 1024, 0, ctx->stream>>>(o->logp, o->esq, Mloc, rep ? 0 : 1, o->fbuf, o->acc + 4 * (size_t)o->accv);
    AVI_LAUNCHED(ctx);
    return AVI_OK;
}

// the sums over the samples that the logq-based estimators add (see k_lr_finalize); CU / CU2 = the two D x r blocks of acc
int32_t avi_lr_logq_sums(avi_obj* o, int Mloc) {
    avi_ctx* ctx = o->ctx;
    const int D = o->D, accv = o->accv;
    float* CU = o->acc + 4 * (size_t)accv + ACC_NSCAL;
    float* CU2 = CU + (size_t)D * o->rank;
    const dim3 grid((unsigned)ceil_div(D, 32)), block(32, 32);
    if (o->objective == AVI_SCOREGRAD) {
        // v0 = sum c w, v1 = sum c w^2, CU = sum c w (U'w)'   (c_m w_m staged in G: the target gradient is not used here)
        k_lr_reduce2<<<grid, block, 0, ctx->stream>>>(o->U, o->U, o->fbuf, o->ld, Mloc, D, o->acc, o->acc + accv);
        AVI_LAUNCHED(ctx);
        k_lr_scale_rows<<<(unsigned)std::min<int64_t>(ceil_div((int64_t)Mloc * o->ld, 256), 1184), 256, 0, ctx->stream>>>(
            o->U, o->fbuf, o->ld, Mloc, o->G);
        AVI_LAUNCHED(ctx);
        AVI_CHECK(avi_gemm_simt(ctx, o->V, 1, o->ldr, o->G, 1, o->ld, CU, D, 1, o->rank, D, Mloc, 1.0f));
    } else if (o->entropy == AVI_ENT_MONTECARLO) {
        // v2 = sum w, v3 = sum w^2, CU2 = sum w (U'w)'
        k_lr_reduce2<<<grid, block, 0, ctx->stream>>>(o->U, o->U, nullptr, o->ld, Mloc, D, o->acc + 2 * (size_t)accv,
                                                      o->acc + 3 * (size_t)accv);
        AVI_LAUNCHED(ctx);
        AVI_CHECK(avi_gemm_simt(ctx, o->V, 1, o->ldr, o->U, 1, o->ld, CU2, D, 1, o->rank, D, Mloc, 1.0f));
    }
    return AVI_OK;
}

int32_t avi_lr_finalize(avi_obj* o, const float* lambda, float* grad, float* out) {
    const float* CU = o->acc + 4 * (size_t)o->accv + ACC_NSCAL;
    const float* CU2 = CU + (size_t)o->D * o->rank;
    const unsigned nb = (unsigned)std::min<int64_t>(ceil_div(o->P, 256), 296);
    k_lr_finalize<<<nb, 256, 0, o->ctx->stream>>>(o->acc, o->accv, CU, CU2, o->lr_ent, lambda, o->D, o->rank, o->M,
                                                  o->objective, o->entropy, grad, out);
    AVI_LAUNCHED(o->ctx);
    return AVI_OK;
}
