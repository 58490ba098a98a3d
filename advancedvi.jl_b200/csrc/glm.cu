// Hierarchical GLM targets, theta = [beta(d); eta], sigma = exp(eta), evaluated for all M
// Monte-Carlo samples at once (the reference calls logdensity once per sample,
// src/algorithms/repgradelbo.jl:84-86, i.e. M GEMVs over X):
//   variant SUBSAMPLING  docs/src/tutorials/subsampling.md:26-38 (+ subsample :99-102)
//   variant BASIC        README.md:47-58 under the exp bijector README.md:91-106
//   likelihood GAUSSIAN  builder-defined Gaussian GLM (BASELINE.json config 4)
// Arithmetic: SURVEY.md Appendix A.5 / oracle/models.py.
//
// Device data: X is kept in both K-major layouts, Xr [n][dK] (features contiguous, forward
// contraction) and Xc [d][nP] (rows contiguous, backward contraction).  In the TF32 modes the
// stored X is rounded to TF32 once (round-to-nearest) so the tensor-core reads are exact.
#include <cmath>
#include <cstdlib>

#include "avi_internal.cuh"
#include "device_utils.cuh"
#include "gemm_tc.cuh"
#include "step_fused.cuh"
#include "glm_prior.cuh"
#include "tc_common.cuh"

namespace {

// split x = hi + lo (both TF32, round to nearest) for the 3xTF32 mode; see family_fr.cu for the K-concatenation
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = tc::round_tf32(x);
    lo = tc::round_tf32(x - hi);
}

// per-sample prior pieces: pre[m] = {prior_lp, 1/sigma^2, d logp / d eta, |beta|^2}
__global__ void __launch_bounds__(256)
k_glm_pre(const float* __restrict__ Z, int ld, int d, int variant, int include_prior, float* __restrict__ Zt,
          int zt_ld, int zt_seg, float4* __restrict__ pre, unsigned long long* __restrict__ zt_owner) {
    __shared__ float sm[33];
    const int m = blockIdx.x;
    if (m == 0 && threadIdx.x == 0) *zt_owner = 0ull;   // Zt / pre no longer hold what the fused iteration kernel drew ahead
    float part = 0.f;
    const int iend = zt_seg > 0 ? max(ld, zt_seg) : ld;
    for (int i = threadIdx.x; i < iend; i += blockDim.x) {
        float z = i < d ? Z[(size_t)m * ld + i] : 0.0f;
        part = fmaf(z, z, part);
        if (Zt) {
            if (zt_seg == 0) { if (i < zt_ld) Zt[(size_t)m * zt_ld + i] = tc::round_tf32(z); }
            else if (i < zt_seg) {   // 3xTF32: [hi | hi | lo]
                float hi, lo; split_tf32(z, hi, lo);
                float* row = Zt + (size_t)m * zt_ld;
                row[i] = hi; row[zt_seg + i] = hi; row[2 * zt_seg + i] = lo;
            }
        }
    }
    const float bsq = block_sum(part, sm);
    if (threadIdx.x == 0) pre[m] = glm_prior_terms(bsq, Z[(size_t)m * ld + d], d, variant, include_prior);
}

// SIMT mode: logits in R -> weighted residual in place, per-sample log-likelihood sum
__global__ void __launch_bounds__(256)
k_glm_lik(float* __restrict__ R, int ldR, int n, const float* __restrict__ y, int likelihood, float w,
          float* __restrict__ llsum) {
    __shared__ float sm[33];
    const int m = blockIdx.x;
    float part = 0.f;
    for (int j = threadIdx.x; j < ldR; j += blockDim.x) {
        float r = 0.0f;
        if (j < n) {
            const float l = R[(size_t)m * ldR + j], yv = __ldg(y + j);
            if (likelihood == AVI_GLM_BERNOULLI_LOGIT) {
                part += yv * l - softplus_f(l);
                r = yv - sigmoid_f(l);
            } else {
                r = yv - l;
                part += -0.5f * AVI_LOG2PI - 0.5f * r * r;
            }
        }
        R[(size_t)m * ldR + j] = w * r;
    }
    part = block_sum(part, sm);
    if (threadIdx.x == 0) llsum[m] = part;
}

// fused mean-field tail: a1[i] = sum_m G[m][i], a2[i] = sum_m G[m][i] eps[m][i] where
// G = (split-K partial sums of X'R) - beta / sigma^2 for i < d and d logp / d eta for i == d;
// the trailing CTAs assemble logp[m] = w * sum(partial log-lik) + prior.
// CTA = 32 coordinates (or 32 samples) x 32 groups; every sum is combined in a fixed order.
__global__ void __launch_bounds__(1024)
k_glm_post_sums(const float* __restrict__ Z, const float* __restrict__ E, int ld, int M, int d,
                const float4* __restrict__ pre, const float* __restrict__ a1p, const float* __restrict__ a2p,
                int nslab, int ldslab, const float* __restrict__ llpart, int nparts, int ldpart, float w,
                int ncoordblk, float* __restrict__ a1, float* __restrict__ a2, float* __restrict__ logp) {
    __shared__ float sm[4][32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    pdl_trigger();
    pdl_wait();
    if ((int)blockIdx.x >= ncoordblk) {
        const int m = (blockIdx.x - ncoordblk) * 32 + tx;
        float s = 0.f;
        if (m < M) {
#pragma unroll 8
            for (int q = ty; q < nparts; q += 32) s += llpart[(size_t)q * ldpart + m];
        }
        sm[0][ty][tx] = s;
        __syncthreads();
        if (ty == 0 && m < M) {
            float t = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) t += sm[0][r][tx];
            logp[m] = fmaf(w, t, pre[m].x);
        }
    } else {
        const int i = blockIdx.x * 32 + tx;
        float t1 = 0.f, t2 = 0.f, p1 = 0.f, p2 = 0.f;
        if (i <= d) {
#pragma unroll 8
            for (int m = ty; m < M; m += 32) {
                const float4 pm = pre[m];
                const float e = E[(size_t)m * ld + i];
                const float g = i < d ? -Z[(size_t)m * ld + i] * pm.y : pm.z;
                t1 += g; t2 = fmaf(g, e, t2);
            }
            if (i < d)
#pragma unroll 4
                for (int q = ty; q < nslab; q += 32) {
                    p1 += a1p[(size_t)q * ldslab + i];
                    p2 += a2p[(size_t)q * ldslab + i];
                }
        }
        sm[0][ty][tx] = t1; sm[1][ty][tx] = t2; sm[2][ty][tx] = p1; sm[3][ty][tx] = p2;
        __syncthreads();
        if (ty < 2 && i <= d) {
            float s = 0.f, ps = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) { s += sm[ty][r][tx]; ps += sm[ty + 2][r][tx]; }
            (ty == 0 ? a1 : a2)[i] = s + ps;
        }
    }
}

// full gradient tail: G[m][i] = sum_s slab[s][m][i] - beta_i / sigma^2, G[m][d] = dlogp/deta, logp[m].
// One CTA per sample, a thread per coordinate quad (the launch brings ld / 4 threads rounded up to whole warps, so no
// thread walks a second quad): the partial sums are fetched in batches of POST_BATCH independent 16-byte loads (ld % 4
// == 0 and every buffer is 16-byte aligned) and added in slab order -- the kernel is a chain of L2 round trips, so the
// number of batches is its duration.  The LAST warp assembles logp[m] first (its loads are in flight before the slabs').
constexpr int POST_BATCH = 12;
__global__ void __launch_bounds__(512)
k_glm_post_full(const float* __restrict__ Z, int ld, int d, const float4* __restrict__ pre,
                const float* __restrict__ slabs, int nslab, long long slab_stride,
                const float* __restrict__ llpart, int nparts, int ldpart, float w, float* __restrict__ G,
                float* __restrict__ logp) {
    const int m = blockIdx.x;
    const float4 pm = pre[m];
    if (threadIdx.x >= blockDim.x - 32) {   // fixed shuffle tree over the partial log-likelihood sums
        const int lane = threadIdx.x & 31;
        float s = 0.f;
#pragma unroll 4
        for (int q = lane; q < nparts; q += 32) s += llpart[(size_t)q * ldpart + m];
        s = warp_sum(s);
        if (lane == 0) logp[m] = fmaf(w, s, pm.x);
    }
    if (G) {
        for (int q = threadIdx.x; q < ld / 4; q += blockDim.x) {
            const size_t base = (size_t)m * ld + 4 * q;
            const float4 z = *reinterpret_cast<const float4*>(Z + base);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s0 = 0; s0 < nslab; s0 += POST_BATCH) {
                float4 v[POST_BATCH];
#pragma unroll
                for (int u = 0; u < POST_BATCH; ++u)
                    v[u] = *reinterpret_cast<const float4*>(slabs + (size_t)min(s0 + u, nslab - 1) * slab_stride + base);   // (unconditional: batched)
#pragma unroll
                for (int u = 0; u < POST_BATCH; ++u)
                    if (s0 + u < nslab) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
            }
            float gv[4] = {acc.x, acc.y, acc.z, acc.w};
            const float zv[4] = {z.x, z.y, z.z, z.w};
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int i = 4 * q + c;
                gv[c] = i < d ? fmaf(-zv[c], pm.y, gv[c]) : (i == d ? pm.z : 0.0f);
            }
            *reinterpret_cast<float4*>(G + base) = make_float4(gv[0], gv[1], gv[2], gv[3]);
        }
    }
}

// column-major host layout (n x d, ldsrc = n) -> Xr [n][dK] and Xc [d][nP].
// mode 0: exact copy; 1: TF32-rounded; 2 (3xTF32): Xr rows hold [hi | lo | hi] in segments of segd (B-operand
// pattern), Xc rows hold [hi | hi | lo] in segments of segn (A-operand pattern).
__global__ void k_glm_layout(const float* __restrict__ src, long long n, int d, int dK, long long nP, int mode,
                             int segd, long long segn, float* __restrict__ Xr, float* __restrict__ Xc) {
    __shared__ float t[32][33];
    const long long r0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int yy = threadIdx.y; yy < 32; yy += blockDim.y) {
        long long r = r0 + threadIdx.x; int c = c0 + yy;
        float v = (r < n && c < d) ? src[(size_t)c * n + r] : 0.0f;
        t[yy][threadIdx.x] = v;
        if (c < d) {
            if (mode == 2) {
                if (r < segn) {
                    float hi, lo; split_tf32(v, hi, lo);
                    float* row = Xc + (size_t)c * nP;
                    row[r] = hi; row[segn + r] = hi; row[2 * segn + r] = lo;
                }
            } else if (r < nP) {
                Xc[(size_t)c * nP + r] = mode == 1 ? tc::round_tf32(v) : v;
            }
        }
    }
    __syncthreads();
    for (int yy = threadIdx.y; yy < 32; yy += blockDim.y) {
        long long r = r0 + yy; int c = c0 + threadIdx.x;
        if (r >= n) continue;
        const float v = t[threadIdx.x][yy];
        if (mode == 2) {
            if (c < segd) {
                float hi, lo; split_tf32(v, hi, lo);
                float* row = Xr + (size_t)r * dK;
                row[c] = hi; row[segd + c] = lo; row[2 * segd + c] = hi;
            }
        } else if (c < dK) {
            Xr[(size_t)r * dK + c] = mode == 1 ? tc::round_tf32(v) : v;
        }
    }
}

// minibatch gather: rows idx[cursor*batch + j] of the full data -> contiguous batch buffers (both layouts).
// 3xTF32 (segd > 0): the source row is [hi | lo | hi]; Xc_b rows are rebuilt as [hi | hi | lo] in segments of segnb.
// idx == nullptr: identity (Xr_full is then the batch buffer itself: only the column layout is rebuilt, see
// Glm::ensure_cols); write_rows == 0: Xr_b / y_b are left alone.
__global__ void k_glm_gather(const float* __restrict__ Xr_full, const float* __restrict__ y_full, int dK, int d,
                             const int32_t* __restrict__ idx, const ObjDeviceState* __restrict__ st, long long batch,
                             long long nPb, int segd, long long segnb, float* __restrict__ Xr_b,
                             float* __restrict__ Xc_b, float* __restrict__ y_b, int write_rows) {
    __shared__ float t[32][33];
    const int32_t* ix = idx ? idx + (st ? st->batch_cursor * batch : 0) : nullptr;
    const long long j0 = (long long)blockIdx.x * 32;
    const int c0 = blockIdx.y * 32;
    for (int yy = threadIdx.y; yy < 32; yy += blockDim.y) {
        long long j = j0 + yy; int c = c0 + threadIdx.x;
        float v = 0.0f;
        if (j < batch && c < dK) {
            v = Xr_full[(size_t)(ix ? (long long)ix[j] : j) * dK + c];
            if (write_rows) Xr_b[(size_t)j * dK + c] = v;
        }
        t[yy][threadIdx.x] = v;
        if (write_rows && blockIdx.y == 0 && threadIdx.x == 0 && j < batch) y_b[j] = y_full[ix ? (long long)ix[j] : j];
    }
    __syncthreads();
    for (int yy = threadIdx.y; yy < 32; yy += blockDim.y) {
        long long j = j0 + threadIdx.x; int c = c0 + yy;
        const float v = j < batch ? t[threadIdx.x][yy] : 0.0f;
        if (segd == 0) {
            if (j < nPb && c < d) Xc_b[(size_t)c * nPb + j] = v;
        } else if (j < segnb) {
            if (c < segd) {                       // hi segment of the source row
                if (c < d) { Xc_b[(size_t)c * nPb + j] = v; Xc_b[(size_t)c * nPb + segnb + j] = v; }
            } else if (c < 2 * segd) {            // lo segment
                if (c - segd < d) Xc_b[(size_t)(c - segd) * nPb + 2 * segnb + j] = v;
            }
        }
    }
}

// minibatch gather, rows only: row idx[cursor*batch + j] -> row j of the batch buffer, 16 bytes per thread.  What the
// whole-iteration kernel needs (it reads the row-major copy for both contractions); the column layout is rebuilt on
// demand for the kernels that want it (Glm::ensure_cols).
__global__ void __launch_bounds__(128)
k_glm_gather_rows(const float* __restrict__ Xr_full, const float* __restrict__ y_full, int dK,
                  const int32_t* __restrict__ idx, const ObjDeviceState* __restrict__ st, long long batch,
                  float* __restrict__ Xr_b, float* __restrict__ y_b) {
    const int32_t* ix = idx + (st ? st->batch_cursor * batch : 0);
    for (long long j = blockIdx.x; j < batch; j += gridDim.x) {
        const long long r = ix[j];
        const float4* src = reinterpret_cast<const float4*>(Xr_full + (size_t)r * dK);
        float4* dst = reinterpret_cast<float4*>(Xr_b + (size_t)j * dK);
        for (int q = threadIdx.x; q < dK / 4; q += blockDim.x) dst[q] = __ldg(src + q);
        if (threadIdx.x == 0) y_b[j] = y_full[r];
    }
}

struct Glm : avi_model {
    int d = 0, dK = 0;
    int x3 = 0, segd = 0;             // 3xTF32: operands stored as three K segments (see k_glm_layout)
    long long segn_full = 0, segn_b = 0, segn = 0;
    int zt_ld = 0;
    long long n_full = 0, nP_full = 0, n_data = 0;
    int likelihood = 0, variant = 0, mode = 0;
    int nshards = 1; long long rows_global = 0; int include_prior = 1;
    float *Xr_full = nullptr, *Xc_full = nullptr, *y_full = nullptr;
    float *Xr_b = nullptr, *Xc_b = nullptr, *y_b = nullptr;
    long long batch_cap = 0, nP_b = 0;
    int32_t* idx_own = nullptr; long long idx_own_cap = 0;
    // active view
    const float *Xr = nullptr, *Xc = nullptr, *y = nullptr;
    long long n_act = 0, nP = 0;
    bool subsampled = false;
    // work buffers
    int capM = 0, cap_ld = 0; long long cap_n = 0;
    long long ldR = 0;
    float* ldpart = nullptr;          // step_fused.cu: per-CTA partials of log det(scale) (host-resident lambda)
    float *R = nullptr, *Zt = nullptr, *llpart = nullptr, *a1p = nullptr, *slabs = nullptr, *spart = nullptr;
    float4* pre = nullptr;
    long long llpart_cap = 0, ap_cap = 0, slab_cap = 0;

    ~Glm() override {
        avi_free(Xr_full); avi_free(Xc_full); avi_free(y_full); avi_free(Xr_b); avi_free(Xc_b); avi_free(y_b);
        avi_free(idx_own); avi_free(R); avi_free(Zt); avi_free(llpart); avi_free(a1p); avi_free(spart);
        avi_free(slabs); avi_free(pre); avi_free(tickets); avi_free(gbar); avi_free(ldpart); avi_free(dummy_st);
    }
    bool hooked = false;               // the sampling kernel produced Zt / pre for exactly (hooked_Z, hooked_M)
    const float* hooked_Z = nullptr; int hooked_M = 0;
    unsigned int* tickets = nullptr;   // last-CTA election per coordinate block (fused backward post-processing)
    unsigned long long* gbar = nullptr;   // grid barrier of the fused iteration kernel: [counter, base] (+ [2]: done ticket)
    int cluster_mode = 1;   // 0: never use thread-block clusters (AVI_TC_CLUSTER=0), 1: planner decides
    bool tc_mode() const { return mode != AVI_GEMM_SIMT_FP32; }
    // Both layouts of X fit the 126 MB L2 together with R: each kernel then pulls the next kernel's copy of X
    // into L2 while it computes, so a step that starts with a cold L2 streams X from HBM behind the math.
    // AVI_L2_STREAM bits: 1 the sampling kernel pulls Xr, 2 the forward kernel pulls Xc (all requests up front),
    // 4 the forward kernel pulls Xc paced over its mainloop.  Default 0: measured slower on C2 (profiles/README.md).
    int l2_stream() const {
        static const int mode = getenv("AVI_L2_STREAM") ? atoi(getenv("AVI_L2_STREAM")) : 0;
        const double bytes = 4.0 * ((double)n_act * dK + (double)d * nP + (double)capM * ldR);
        return (tc_mode() && !subsampled && bytes < 100e6) ? mode : 0;
    }
    int64_t view_key() const override { return subsampled ? (int64_t)n_act : -1; }
    int64_t rows_full() const override { return n_full; }
    void clear_hook() override { hooked = false; hooked_Z = nullptr; hooked_M = 0; }
    bool sample_hook(const float* Z, int ld, int M, SampleHook* h) override {
        clear_hook();
        if (M <= 0 || ensure(M, ld) != AVI_OK) return false;
        h->kind = 1; h->d = d; h->variant = variant; h->include_prior = include_prior;
        h->Zt = tc_mode() ? Zt : nullptr; h->pre = pre; h->zt_ld = zt_ld; h->zt_seg = x3 ? segd : 0;
        h->tl = ctx->tl; h->zt_owner = gbar + 3;
        if (l2_stream() & 1) { h->pf_ptr = Xr; h->pf_bytes = (unsigned long long)n_act * dK * sizeof(float); }
        hooked = true; hooked_Z = Z; hooked_M = M;
        return true;
    }
    float likeadj() const {
        if (variant != AVI_GLM_SUBSAMPLING) return 1.0f;
        double rows = subsampled ? (double)n_act * nshards : (double)rows_global;
        return (float)((double)n_data / rows);
    }
    void view_full() { Xr = Xr_full; Xc = Xc_full; y = y_full; n_act = n_full; nP = nP_full; segn = segn_full; subsampled = false; }
    long long kf() const { return x3 ? 3LL * segd : d; }        // K extent of the forward contraction
    long long kb() const { return x3 ? 3LL * segn : n_act; }    // K extent of the backward contraction

    int32_t ensure(int M, int ld) {
        if (M <= capM && ld == cap_ld && n_act <= cap_n) return AVI_OK;
        avi_free(R); avi_free(Zt); avi_free(pre); avi_free(spart);
        generation++;
        capM = std::max(M, capM); cap_ld = ld; cap_n = std::max(cap_n, n_act);
        ldR = (x3 ? 3 : 1) * round_up(cap_n, 32);
        zt_ld = x3 ? 3 * segd : ld;
        AVI_CHECK(avi_alloc(ctx, &R, (size_t)round_up(capM, 32) * ldR));   // [M][ldR], or transposed [ldR][round_up(M, 4)] (fused_step)
        AVI_CHECK(avi_alloc(ctx, &Zt, (size_t)capM * zt_ld));
        AVI_CHECK(avi_alloc(ctx, &spart, (size_t)capM * (2 * (size_t)ctx->prop.multiProcessorCount + 1)));   // step_fused.cu: draw_slice
        AVI_CHECK(avi_alloc(ctx, &pre, (size_t)capM));
        return AVI_OK;
    }
    int32_t ensure_buf(float** p, long long* cap, long long need) {
        if (need <= *cap) return AVI_OK;
        generation++;
        avi_free(*p);
        AVI_CHECK(avi_alloc(ctx, p, (size_t)need));
        *cap = need;
        return AVI_OK;
    }

    // forward: R <- w * resid, llpart/nparts <- partial log-lik sums
    int32_t forward(const float* Z, int ld, int M, int* nparts, bool want_backward) {
        const float w = likeadj();
        const bool pre_done = hooked && hooked_Z == Z && hooked_M == M;
        clear_hook();
        if (pre_done) {
            // the sampling kernel already produced Zt and pre for exactly these samples
        } else {
            k_glm_pre<<<M, 256, 0, ctx->stream>>>(Z, ld, d, variant, include_prior, tc_mode() ? Zt : nullptr, zt_ld,
                                                  x3 ? segd : 0, pre, gbar + 3);
            AVI_LAUNCHED(ctx);
        }
        if (!tc_mode()) {
            AVI_CHECK(ensure_buf(&llpart, &llpart_cap, capM));
            // logits[m][j] = sum_k Z[m][k] Xr[j][k]
            AVI_CHECK(avi_gemm_simt(ctx, Z, ld, 1, Xr, dK, 1, R, ldR, 1, M, (int)n_act, d, 1.0f));
            k_glm_lik<<<M, 256, 0, ctx->stream>>>(R, (int)ldR, (int)n_act, y, likelihood, w, llpart);
            AVI_LAUNCHED(ctx);
            *nparts = 1;
            return AVI_OK;
        }
        TcParams p{};
        AVI_CHECK(avi_tc_plan(ctx, M, n_act, kf(), false, cluster_mode, &p));
        p.C = R; p.ldc = (int)ldR; p.y = y; p.w = w; p.likelihood = likelihood;
        p.r_seg = x3 ? (int)segn : 0;
        p.static_op = subsampled ? 0 : 2;   // B = X rows (a minibatch copy is rewritten every step: not static)
        p.tl = ctx->tl; p.tl_id = 1;
        if ((l2_stream() & 6) && want_backward) {
            p.pf_ptr = Xc; p.pf_bytes = (unsigned long long)d * nP * sizeof(float);
            p.pf_pace_ns = (l2_stream() & 4) ? 600u : 0u;
        }
        AVI_CHECK(ensure_buf(&llpart, &llpart_cap, (long long)p.n_bchunk * 4 * capM));
        p.part1 = llpart; p.ldpart = capM;
        CUtensorMap tmA, tmB;
        AVI_CHECK(avi_tc_make_tmap(ctx, &tmA, Zt, M, kf(), zt_ld, 128 / p.cb));
        AVI_CHECK(avi_tc_make_tmap(ctx, &tmB, Xr, n_act, kf(), dK, p.pair ? p.nt / 2 : p.nt / p.ca));
        AVI_CHECK(avi_tc_launch(ctx, EPI_GLM_FWD, tmA, tmB, p));
        *nparts = p.n_bchunk * 4;
        return AVI_OK;
    }

    // store: the plain store epilogue follows (full gradient G): the planner may trade k-splits for b-chunks
    int32_t backward_setup(int M, TcParams* p, CUtensorMap* tmA, CUtensorMap* tmB, bool store = false) {
        AVI_CHECK(ensure_cols());
        AVI_CHECK(avi_tc_plan(ctx, d, M, kb(), true, cluster_mode, p, 1, 0, store ? 1 : 0));
        p->static_op = subsampled ? 0 : 1;   // A = X columns
        p->tl = ctx->tl; p->tl_id = 2;
        AVI_CHECK(avi_tc_make_tmap(ctx, tmA, Xc, d, kb(), nP, 128 / p->cb));
        AVI_CHECK(avi_tc_make_tmap(ctx, tmB, R, M, kb(), ldR, p->pair ? p->nt / 2 : p->nt / p->ca));
        return AVI_OK;
    }

    // logdensity_and_gradient for all M samples as ONE persistent launch (the forward and backward phases of the
    // whole-iteration kernel, step_fused.cu, with a plain store epilogue): X is read once (row-major copy for both
    // contractions, the second time as an MN-major operand), the residuals travel transposed, one prologue / teardown
    // instead of two.  Leaves the split-K slabs of G and the partial log-likelihoods for k_glm_post_full.
    float* dummy_st = nullptr;   // the kernel's entry snapshot reads an ObjDeviceState; this path has none to show it
    bool fused_store_ok(int M) const {
        static const bool on = !(getenv("AVI_FUSED_EVAL") && atoi(getenv("AVI_FUSED_EVAL")) == 0);
        return on && !x3 && fused_step_ok(M) && dK % 4 == 0;
    }
    int32_t eval_fused_store(const float* Z, int ld, int M, int* nparts, const float** sl, int* nslab, long long* sstride) {
        const bool pre_done = hooked && hooked_Z == Z && hooked_M == M;
        clear_hook();
        if (!pre_done) {
            k_glm_pre<<<M, 256, 0, ctx->stream>>>(Z, ld, d, variant, include_prior, Zt, zt_ld, 0, pre, gbar + 3);
            AVI_LAUNCHED(ctx);
        }
        if (!dummy_st) AVI_CHECK(avi_alloc(ctx, &dummy_st, 32));
        StepParams sp{};
        AVI_CHECK(avi_tc_plan(ctx, M, n_act, kf(), false, 0, &sp.f, 0));
        sp.f.C = R; sp.f.y = y; sp.f.w = likeadj(); sp.f.likelihood = likelihood; sp.f.r_seg = 0;
        sp.f.static_op = subsampled ? 0 : 2;
        AVI_CHECK(ensure_buf(&llpart, &llpart_cap, (long long)sp.f.n_bchunk * 4 * capM));
        sp.f.part1 = llpart; sp.f.ldpart = capM; sp.f.post_on = 0;   // per-sample partials, as the stand-alone forward kernel
        // widest backward tile (more k-splits: 6 us instead of 10 us of mainloop at C3, twice the slabs) unless
        // AVI_FUSED_EVAL_WIDE=0 leaves the width to the planner (measured on C3: 11.35 k vs 11.17 k steps/s)
        static const bool wide = !(getenv("AVI_FUSED_EVAL_WIDE") && atoi(getenv("AVI_FUSED_EVAL_WIDE")) == 0);
        AVI_CHECK(avi_tc_plan(ctx, d, M, kb(), true, 0, &sp.b, 0, 0, /*nt_search=*/wide ? 0 : 1));
        sp.b.static_op = subsampled ? 0 : 1;
        *sstride = (long long)capM * ld;
        AVI_CHECK(ensure_buf(&slabs, &slab_cap, *sstride * sp.b.n_ksplit));
        sp.b.C = slabs; sp.b.ldc = ld; sp.b.slab_stride = *sstride; sp.b.post_on = 0;
        CUtensorMap tmZ, tmXr, tmXc, tmR;
        AVI_CHECK(avi_tc_make_tmap(ctx, &tmZ, Zt, M, kf(), zt_ld, 128));
        AVI_CHECK(avi_tc_make_tmap(ctx, &tmXr, Xr, n_act, kf(), dK, sp.f.nt));
        static const bool one_box = !(getenv("AVI_TC_MN3") && atoi(getenv("AVI_TC_MN3")) == 0);
        if (one_box && dK % 32 == 0) {   // MN-major A operand straight from the row-major copy: one 3-D box per tile
            sp.b.a_mn = 2;
            AVI_CHECK(avi_tc_make_tmap_mn3(ctx, &tmXc, Xr, n_act, dK, dK, 4));
        } else {
            sp.b.a_mn = 1;
            AVI_CHECK(avi_tc_make_tmap(ctx, &tmXc, Xr, n_act, d, dK, 32, /*atom32=*/1));
        }
        const long long ldRt = round_up(capM, 32);
        sp.f.c_mn = 1; sp.f.ldc = (int)ldRt;   // R transposed [data row][sample]: coalesced forward stores, MN-major B operand
        if (one_box && (sp.b.n_bchunk == 1 || sp.b.nt % 32 == 0)) {
            sp.b.b_mn = 2;
            AVI_CHECK(avi_tc_make_tmap_mn3(ctx, &tmR, R, kb(), round_up(M, 32), ldRt, (sp.b.nt + 31) / 32));
        } else {
            sp.b.b_mn = 1;
            AVI_CHECK(avi_tc_make_tmap(ctx, &tmR, R, kb(), M, ldRt, 32, /*atom32=*/1));
        }
        sp.bwd_store = 1; sp.do_sample = 0; sp.draw_ahead = 0;
        sp.D = d + 1; sp.ld = ld; sp.Mloc = M; sp.st = reinterpret_cast<ObjDeviceState*>(dummy_st);
        sp.d = d; sp.variant = variant; sp.include_prior = include_prior;
        sp.Zt = Zt; sp.zt_ld = zt_ld; sp.pre = reinterpret_cast<float*>(pre);
        sp.t.mode = STEP_TAIL_NONE; sp.t.comm.nranks = 1;
        sp.gbar = gbar;
        AVI_CHECK(avi_step_fused_launch(ctx, tmZ, tmXr, tmXc, tmR, sp));
        *nparts = sp.f.n_bchunk * 4; *sl = slabs; *nslab = sp.b.n_ksplit;
        return AVI_OK;
    }

    int32_t eval(const float* Z, int ld, int M, float* logp, float* G) override {
        if (M <= 0) return AVI_OK;
        AVI_CHECK(ensure(M, ld));
        if (G && tc_mode() && fused_store_ok(M)) {
            int np = 0, ns = 0; long long ss = 0; const float* s0 = nullptr;
            AVI_CHECK(eval_fused_store(Z, ld, M, &np, &s0, &ns, &ss));
            k_glm_post_full<<<M, (unsigned)std::min<int64_t>(512, round_up(ld / 4, 32)), 0, ctx->stream>>>(Z, ld, d, pre, s0, ns, ss, llpart, np,
                                                                                                   capM, likeadj(), G, logp);
            AVI_LAUNCHED(ctx);
            return AVI_OK;
        }
        int nparts = 0;
        AVI_CHECK(forward(Z, ld, M, &nparts, G != nullptr));
        const float w = likeadj();
        const float* sl = nullptr; int nslab = 0; long long sstride = 0;
        if (G) {
            if (!tc_mode()) {
                // G[m][i] = sum_j R[m][j] Xc[i][j]
                AVI_CHECK(ensure_cols());
                AVI_CHECK(avi_gemm_simt(ctx, R, ldR, 1, Xc, nP, 1, G, ld, 1, M, d, (int)n_act, 1.0f));
                sl = G; nslab = 1; sstride = 0;
            } else {
                TcParams p{}; CUtensorMap tmA, tmB;
                AVI_CHECK(backward_setup(M, &p, &tmA, &tmB, /*store=*/true));
                sstride = (long long)capM * ld;
                AVI_CHECK(ensure_buf(&slabs, &slab_cap, sstride * p.n_ksplit));
                p.C = slabs; p.ldc = ld; p.slab_stride = sstride;
                AVI_CHECK(avi_tc_launch(ctx, EPI_STORE, tmA, tmB, p));
                sl = slabs; nslab = p.n_ksplit;
            }
        }
        k_glm_post_full<<<M, (unsigned)std::min<int64_t>(512, round_up(ld / 4, 32)), 0, ctx->stream>>>(Z, ld, d, pre, sl, nslab, sstride,
                                                                                                    llpart, nparts, capM, w, G, logp);
        AVI_LAUNCHED(ctx);
        return AVI_OK;
    }

    bool has_gradsums() const override { return tc_mode(); }
    int32_t eval_gradsums(const float* Z, const float* E, int ld, int M, float* logp, float* a1, float* a2) override {
        if (M <= 0) return AVI_OK;
        AVI_CHECK(ensure(M, ld));
        int nparts = 0;
        AVI_CHECK(forward(Z, ld, M, &nparts, true));
        TcParams p{}; CUtensorMap tmA, tmB;
        AVI_CHECK(backward_setup(M, &p, &tmA, &tmB));
        const int nslab = p.n_ksplit * p.n_bchunk;
        const int ldslab = (int)round_up(d, 32);
        AVI_CHECK(ensure_buf(&a1p, &ap_cap, 2LL * nslab * ldslab));
        float* a2p = a1p + (size_t)nslab * ldslab;
        p.E = E; p.lde = ld; p.part1 = a1p; p.part2 = a2p; p.ldpart = ldslab;
        // fold the whole post-processing into the backward kernel's epilogue (one launch less per step)
        static const bool fuse_post = !(getenv("AVI_FUSE_POST") && atoi(getenv("AVI_FUSE_POST")) == 0);
        if (fuse_post && !p.pair && p.ca == 1 && p.cb == 1) {
            p.post_on = 1; p.post_Z = Z; p.post_pre = reinterpret_cast<const float*>(pre);
            p.post_llpart = llpart; p.post_nparts = nparts; p.post_ldll = capM; p.post_w = likeadj();
            p.post_logp = logp; p.post_a1 = a1; p.post_a2 = a2; p.post_tickets = tickets;
            AVI_CHECK(avi_tc_launch(ctx, EPI_GLM_BWD, tmA, tmB, p));
            return AVI_OK;
        }
        AVI_CHECK(avi_tc_launch(ctx, EPI_GLM_BWD, tmA, tmB, p));
        const int ncb = (int)ceil_div(d + 1, 32);
        const unsigned grid = (unsigned)(ncb + ceil_div(M, 32));
        {
            cudaError_t e = avi_launch_pdl(ctx, k_glm_post_sums, dim3(grid), dim3(1024), 0, Z, E, ld, M, d,
                                           (const float4*)pre, (const float*)a1p, (const float*)a2p, nslab, ldslab,
                                           (const float*)llpart, nparts, capM, likeadj(), ncb, a1, a2, logp);
            if (e != cudaSuccess) AVI_FAIL(ctx, AVI_ERR_CUDA, std::string("post_sums launch: ") + cudaGetErrorString(e));
        }
        AVI_LAUNCHED(ctx);
        return AVI_OK;
    }

    // ---- the whole iteration in one launch (step_fused.cu) ----
    // AVI_FUSED_STEP: 0 never, 1 (default) for problems where the per-launch fixed costs matter (the large ones keep the
    // CTA-pair tiles of the stand-alone kernels), 2 always
    int fused_mode = getenv("AVI_FUSED_STEP") ? atoi(getenv("AVI_FUSED_STEP")) : 1;
    int32_t set_fused_step(int m) override {
        if (m < 0 || m > 2) return AVI_ERR_INVALID;
        if (m != fused_mode) generation++;   // captured iterations were built for the other launch shape
        fused_mode = m;
        return AVI_OK;
    }
    bool fused_step_ok(int Mloc) const override {
        if (!fused_mode || !tc_mode() || Mloc <= 0) return false;
        const double flops = 4.0 * (double)n_act * d * Mloc * (x3 ? 3.0 : 1.0);
        return fused_mode >= 2 || flops <= 2e11;
    }
    bool fused_host_lambda_ok() const override {
        static const bool ahead = !(getenv("AVI_DRAW_AHEAD") && atoi(getenv("AVI_DRAW_AHEAD")) == 0);
        static const bool host_lam = !(getenv("AVI_HOST_LAMBDA") && atoi(getenv("AVI_HOST_LAMBDA")) == 0);
        return ahead && host_lam && !x3;
    }
    int32_t fused_step(const FusedStepArgs& fa) override {
        const int M = fa.Mloc, ld = fa.ld;
        AVI_CHECK(ensure(M, ld));
        clear_hook();
        StepParams sp{};
        const float w = likeadj();
        // forward: logits[m][j] = sum_k Zt[m][k] Xr[j][k]
        AVI_CHECK(avi_tc_plan(ctx, M, n_act, kf(), false, 0, &sp.f, 0));
        sp.f.C = R; sp.f.ldc = (int)ldR; sp.f.y = y; sp.f.w = w; sp.f.likelihood = likelihood;
        sp.f.r_seg = x3 ? (int)segn : 0;
        sp.f.static_op = subsampled ? 0 : 2;
        const long long units_f = (long long)sp.f.n_ablk * sp.f.n_bchunk * sp.f.n_ksplit;
        AVI_CHECK(ensure_buf(&llpart, &llpart_cap, std::max<long long>(units_f, (long long)sp.f.n_bchunk * 4 * capM)));
        sp.f.part1 = llpart; sp.f.ldpart = capM;
        sp.f.post_on = 2;   // one log-likelihood total per unit (the tail phase only needs sum_m log pi)
        // backward: G[i][m] = sum_j Xc[i][j] R[m][j], reduced against eps in the epilogue
        AVI_CHECK(avi_tc_plan(ctx, d, M, kb(), true, 0, &sp.b, 0));
        sp.b.static_op = subsampled ? 0 : 1;
        const int nslab = sp.b.n_ksplit * sp.b.n_bchunk;
        const int ldslab = (int)round_up(d, 32);
        AVI_CHECK(ensure_buf(&a1p, &ap_cap, 2LL * nslab * ldslab));
        sp.b.E = fa.E; sp.b.lde = ld; sp.b.part1 = a1p; sp.b.part2 = a1p + (size_t)nslab * ldslab; sp.b.ldpart = ldslab;
        // post_on = 2: prior gradient and the eta coordinate in the epilogue, slab rows left for the tail phase
        sp.b.post_on = 2; sp.b.post_Z = fa.Z; sp.b.post_pre = reinterpret_cast<const float*>(pre);
        sp.b.post_a1 = fa.t.acc; sp.b.post_a2 = fa.t.acc + fa.t.accv;
        CUtensorMap tmZ, tmXr, tmXc, tmR;
        AVI_CHECK(avi_tc_make_tmap(ctx, &tmZ, Zt, M, kf(), zt_ld, 128));
        AVI_CHECK(avi_tc_make_tmap(ctx, &tmXr, Xr, n_act, kf(), dK, sp.f.nt));
        // backward A operand: X columns from the transposed copy Xc (K-major) or, default, MN-major straight from the
        // row-major copy the forward phase has just pulled through L2 (AVI_TC_AMN=0 selects the former)
        static const bool a_mn = !(getenv("AVI_TC_AMN") && atoi(getenv("AVI_TC_AMN")) == 0);
        if (a_mn) {
            sp.b.a_mn = 1; sp.b.a_seg_kb = x3 ? (int)(segn / 32) : 0; sp.b.a_seg_off = x3 ? segd : 0;
            static const bool one_box = !(getenv("AVI_TC_MN3") && atoi(getenv("AVI_TC_MN3")) == 0);
            if (one_box && dK % 32 == 0) {   // one 3-D box per tile
                sp.b.a_mn = 2;
                AVI_CHECK(avi_tc_make_tmap_mn3(ctx, &tmXc, Xr, n_act, dK, dK, 4));
            } else {
                AVI_CHECK(avi_tc_make_tmap(ctx, &tmXc, Xr, n_act, x3 ? 3LL * segd : d, dK, 32, /*atom32=*/1));
            }
        } else {
            if (!fa.dry_run) AVI_CHECK(ensure_cols());
            AVI_CHECK(avi_tc_make_tmap(ctx, &tmXc, Xc, d, kb(), nP, 128));
        }
        // R between the phases: transposed (data row major) by default, so that the forward epilogue's stores coalesce; the
        // backward contraction reads it as an MN-major B operand (AVI_TC_BMN=0: sample-major R, K-major operand)
        static const bool b_mn = !(getenv("AVI_TC_BMN") && atoi(getenv("AVI_TC_BMN")) == 0);
        if (b_mn) {
            const long long ldRt = round_up(capM, 32);
            sp.f.c_mn = 1; sp.f.ldc = (int)ldRt;
            static const bool one_box = !(getenv("AVI_TC_MN3") && atoi(getenv("AVI_TC_MN3")) == 0);
            if (one_box && (sp.b.n_bchunk == 1 || sp.b.nt % 32 == 0)) {
                sp.b.b_mn = 2;
                AVI_CHECK(avi_tc_make_tmap_mn3(ctx, &tmR, R, kb(), round_up(M, 32), ldRt, (sp.b.nt + 31) / 32));
            } else {
                sp.b.b_mn = 1;
                AVI_CHECK(avi_tc_make_tmap(ctx, &tmR, R, kb(), M, ldRt, 32, /*atom32=*/1));
            }
        } else {
            AVI_CHECK(avi_tc_make_tmap(ctx, &tmR, R, M, kb(), ldR, sp.b.nt));
        }
        sp.do_sample = 1;
        sp.lambda = fa.lambda; sp.D = fa.D; sp.ld = ld; sp.m0 = fa.m0; sp.Mloc = M; sp.st = fa.st;
        sp.Z = fa.Z; sp.E = fa.E; sp.esq = fa.esq;
        sp.d = d; sp.variant = variant; sp.include_prior = include_prior;
        sp.Zt = Zt; sp.zt_ld = zt_ld; sp.zt_seg = x3 ? segd : 0; sp.pre = reinterpret_cast<float*>(pre);
        // draw the next iteration's samples in the tail phase (AVI_DRAW_AHEAD=0: sample phase at the start of every launch)
        static const bool ahead = !(getenv("AVI_DRAW_AHEAD") && atoi(getenv("AVI_DRAW_AHEAD")) == 0);
        sp.draw_ahead = (ahead && fa.t.mode == STEP_TAIL_UPDATE && !x3) ? 1 : 0;
        // estimate_gradient! boundary with lambda still in pinned host memory: slice-major sampling at the start of the launch
        // (the tail of this mode never draws ahead), every CTA fetching its own slice from the host
        if (fa.lambda_src) {
            if (x3 || fa.t.mode != STEP_TAIL_GRAD_OUT || !ahead) AVI_FAIL(ctx, AVI_ERR_STATE, "fused_step: host-resident lambda needs the slice-major sampler");
            if (!ldpart) AVI_CHECK(avi_alloc(ctx, &ldpart, 256));
            sp.draw_ahead = 1; sp.lambda_src = fa.lambda_src; sp.ldpart = ldpart;
        }
        sp.spart = spart; sp.spart_stride = ctx->prop.multiProcessorCount;
        static const int tc_dbg = getenv("AVI_TC_DBG") ? atoi(getenv("AVI_TC_DBG")) : 0;   // timing experiments (gemm_tc.cuh)
        sp.f.dbg = tc_dbg; sp.b.dbg = tc_dbg;
        sp.t = fa.t;
        sp.t.unit_ll = llpart; sp.t.n_units_f = (int)units_f; sp.t.w_lik = w;
        sp.t.part1 = sp.b.part1; sp.t.part2 = sp.b.part2; sp.t.nslab = nslab; sp.t.ldslab = ldslab;
        sp.t.done_ticket = reinterpret_cast<unsigned int*>(gbar + 2);
        sp.gbar = gbar;
        if (fa.dry_run) return AVI_OK;
        return avi_step_fused_launch(ctx, tmZ, tmXr, tmXc, tmR, sp);
    }

    int32_t ensure_batch(long long batch) {
        if (batch <= batch_cap) return AVI_OK;
        avi_free(Xr_b); avi_free(Xc_b); avi_free(y_b);
        generation++;
        batch_cap = batch; segn_b = round_up(batch, 32); nP_b = (x3 ? 3 : 1) * segn_b;
        AVI_CHECK(avi_alloc(ctx, &Xr_b, (size_t)batch_cap * dK));
        AVI_CHECK(avi_alloc(ctx, &Xc_b, (size_t)d * nP_b));
        AVI_CHECK(avi_alloc(ctx, &y_b, (size_t)batch_cap));
        return AVI_OK;
    }
    // Xc_b holds the columns of the rows currently in Xr_b?  (The device-side gather of the optimiser loop copies rows
    // only; whoever needs the column layout -- the stand-alone backward contraction, the SIMT path -- rebuilds it.)
    bool cols_valid = true;
    int32_t ensure_cols() {
        if (!subsampled || cols_valid) return AVI_OK;
        dim3 grid((unsigned)ceil_div(segn_b, 32), (unsigned)ceil_div(dK, 32));
        k_glm_gather<<<grid, dim3(32, 8), 0, ctx->stream>>>(Xr_b, y_b, dK, d, (const int32_t*)nullptr, (const ObjDeviceState*)nullptr,
                                                            n_act, nP_b, x3 ? segd : 0, segn_b, Xr_b, Xc_b, y_b, 0);
        AVI_LAUNCHED(ctx);
        cols_valid = true;
        return AVI_OK;
    }
    int32_t gather(const int32_t* idx_dev, long long batch, const ObjDeviceState* st, bool rows_only = false) {
        AVI_CHECK(ensure_batch(batch));
        static const bool lazy_cols = !(getenv("AVI_GATHER_ROWS") && atoi(getenv("AVI_GATHER_ROWS")) == 0);
        if (rows_only && lazy_cols && dK % 4 == 0) {
            k_glm_gather_rows<<<(unsigned)std::min<long long>(batch, 8192), 128, 0, ctx->stream>>>(Xr_full, y_full, dK, idx_dev, st,
                                                                                                  batch, Xr_b, y_b);
            AVI_LAUNCHED(ctx);
            cols_valid = false;
            Xr = Xr_b; Xc = Xc_b; y = y_b; n_act = batch; nP = nP_b; segn = segn_b; subsampled = true;
            return AVI_OK;
        }
        // the pitch of Xc_b follows the allocated capacity so that captured tensor maps stay valid
        dim3 grid((unsigned)ceil_div(segn_b, 32), (unsigned)ceil_div(dK, 32));
        k_glm_gather<<<grid, dim3(32, 8), 0, ctx->stream>>>(Xr_full, y_full, dK, d, idx_dev, st, batch, nP_b,
                                                            x3 ? segd : 0, segn_b, Xr_b, Xc_b, y_b, 1);
        AVI_LAUNCHED(ctx);
        cols_valid = true;
        // (buffers and pitches follow the allocated capacity: graphs captured on a view of this shape stay valid; a
        // different shape is a different view_key())
        Xr = Xr_b; Xc = Xc_b; y = y_b; n_act = batch; nP = nP_b; segn = segn_b; subsampled = true;
        return AVI_OK;
    }
    int32_t subsample(const int32_t* idx_host, int64_t batch) override {
        if (!idx_host) { view_full(); return AVI_OK; }
        if (batch <= 0) AVI_FAIL(ctx, AVI_ERR_INVALID, "empty batch");
        for (int64_t j = 0; j < batch; ++j)
            if (idx_host[j] < 0 || idx_host[j] >= n_full) AVI_FAIL(ctx, AVI_ERR_INVALID, "batch index out of range");
        if (batch > idx_own_cap) {
            avi_free(idx_own);
            AVI_CHECK(avi_alloc(ctx, &idx_own, (size_t)batch));
            idx_own_cap = batch;
        }
        AVI_CUDA(ctx, cudaMemcpyAsync(idx_own, idx_host, batch * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // idx_host is only borrowed for the call
        return gather(idx_own, batch, nullptr);
    }
    int32_t subsample_dev(const int32_t* idx_dev, int64_t batch, const ObjDeviceState* st) override {
        if (batch <= 0 || batch > n_full) AVI_FAIL(ctx, AVI_ERR_INVALID, "bad batch size");
        return gather(idx_dev, batch, st, /*rows_only=*/true);
    }
    int32_t set_gemm_mode(int m) override {
        if (m == mode) return AVI_OK;
        AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED,
                 "the arithmetic mode fixes how X is stored (exact fp32 or TF32-rounded); create a new target");
    }
};

}  // namespace

int32_t avi_model_glm_make(avi_ctx* ctx, const float* X, const float* y, int64_t n, int d, int64_t n_data,
                           int likelihood, int variant, int gemm_mode, avi_model** out) {
    if (!X || !y || n <= 0 || d <= 0 || n_data <= 0) AVI_FAIL(ctx, AVI_ERR_INVALID, "bad arguments");
    if (n > 0x7fffffffLL) AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED, "more than 2^31-1 rows per device");
    if (likelihood != AVI_GLM_BERNOULLI_LOGIT && likelihood != AVI_GLM_GAUSSIAN) AVI_FAIL(ctx, AVI_ERR_INVALID, "likelihood");
    if (variant != AVI_GLM_SUBSAMPLING && variant != AVI_GLM_BASIC) AVI_FAIL(ctx, AVI_ERR_INVALID, "variant");
    if (gemm_mode != AVI_GEMM_SIMT_FP32 && gemm_mode != AVI_GEMM_TF32 && gemm_mode != AVI_GEMM_TF32X3)
        AVI_FAIL(ctx, AVI_ERR_INVALID, "gemm_mode");
    Glm* g = new Glm();
    g->ctx = ctx; g->D = d + 1; g->capability = 1;
    g->x3 = gemm_mode == AVI_GEMM_TF32X3 ? 1 : 0;
    g->segd = (int)round_up(d, 32); g->segn_full = round_up(n, 32);
    g->d = d; g->dK = g->x3 ? 3 * g->segd : (int)round_up(d, 4);
    g->n_full = n; g->nP_full = (g->x3 ? 3 : 1) * g->segn_full; g->n_data = n_data; g->rows_global = n;
    g->likelihood = likelihood; g->variant = variant; g->mode = gemm_mode;
    if (const char* e = getenv("AVI_TC_CLUSTER")) g->cluster_mode = atoi(e);
    float* tmp = nullptr;
    int32_t rc = avi_alloc(ctx, &g->Xr_full, (size_t)n * g->dK);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &g->Xc_full, (size_t)d * g->nP_full);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &g->y_full, (size_t)n);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &g->tickets, (size_t)ceil_div(d, 128) + 1);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &g->gbar, 4);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &tmp, (size_t)n * d);
    if (rc != AVI_OK) { avi_free(tmp); delete g; return rc; }
    cudaError_t e = avi_copy(ctx, tmp, X, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = avi_copy(ctx, g->y_full, y, (size_t)n * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) {
        dim3 grid((unsigned)ceil_div(g->segn_full, 32), (unsigned)ceil_div(g->segd, 32));
        k_glm_layout<<<grid, dim3(32, 8), 0, ctx->stream>>>(tmp, n, d, g->dK, g->nP_full, g->x3 ? 2 : (g->tc_mode() ? 1 : 0),
                                                            g->segd, g->segn_full, g->Xr_full, g->Xc_full);
        ctx->launches++;
        e = cudaStreamSynchronize(ctx->stream);
    }
    avi_free(tmp);
    if (e != cudaSuccess) {
        avi_set_error(ctx, std::string("avi_model_glm_make: ") + cudaGetErrorString(e));
        delete g;
        return AVI_ERR_CUDA;
    }
    g->view_full();
    *out = g;
    return AVI_OK;
}

// declared in api.cu
int32_t avi_glm_set_data_shard(avi_model* model, int32_t nshards, int64_t rows_global, int32_t include_prior) {
    Glm* g = dynamic_cast<Glm*>(model);
    if (!g) return AVI_ERR_UNSUPPORTED;
    if (nshards < 1 || rows_global < g->n_full) return AVI_ERR_INVALID;
    g->nshards = nshards; g->rows_global = rows_global; g->include_prior = include_prior ? 1 : 0;
    return AVI_OK;
}
