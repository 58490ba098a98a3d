// Full-rank family contractions on the tensor cores (tcgen05 kind::tf32 through gemm_tc.cu):
//   Z = E * L' + mu        scale * eps of src/families/location_scale.jl:71-77         (avi_fr_affine)
//   C[j*D + i] = sum_m W[m][i] E[m][j]   pullback of the same product onto vec(L)      (avi_fr_outer)
// These two feed first-order quantities (the samples themselves and the gradient of L), so they run as 3xTF32:
// every operand is split x = hi + lo (both TF32, round-to-nearest) and the three significant products
// hi*hi + hi*lo + lo*hi are folded into ONE contraction by concatenating the splits along K:
//   A' = [A_hi | A_hi | A_lo],  B' = [B_hi | B_lo | B_hi]   =>   A' . B' = A_hi.B_hi + A_hi.B_lo + A_lo.B_hi
// (error ~2^-21 relative, i.e. fp32-grade, at 3x the tiny MMA cost).  The kernel itself is unchanged.
#include "avi_internal.cuh"
#include "device_utils.cuh"
#include "gemm_tc.cuh"
#include "fr_finalize.cuh"
#include "glm_prior.cuh"
#include "tc_common.cuh"

namespace {

constexpr int FR_MAX_ZSLABS = 8;
constexpr int FR_TC_MAX_M = 4096;

__device__ __forceinline__ void split3_store(float x, float* dst, int seg, bool b_pattern) {
    const float hi = tc::round_tf32(x);
    const float lo = tc::round_tf32(x - hi);
    dst[0] = hi;
    dst[seg] = b_pattern ? lo : hi;
    dst[2 * (size_t)seg] = b_pattern ? hi : lo;
}

// dst[r][s*seg + c] (s = 0..2) from src[r][c]; rows r < R, columns c < seg (zero beyond C)
__global__ void k_split3_rows(const float* __restrict__ src, int R, int C, int lds, float* __restrict__ dst, int seg,
                              int b_pattern) {
    const int r = blockIdx.x;
    for (int c = threadIdx.x; c < seg; c += blockDim.x) {
        const float x = c < C ? src[(size_t)r * lds + c] : 0.0f;
        split3_store(x, dst + (size_t)r * 3 * seg + c, seg, b_pattern != 0);
    }
}

// dst[c][s*seg + r] (s = 0..2) = split(src[r][c]); src has R rows, C columns (row pitch lds); r runs to seg
__global__ void k_transpose_split3(const float* __restrict__ src, int R, int C, long long lds, float* __restrict__ dst,
                                   int seg, int b_pattern) {
    __shared__ float t[32][33];
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int yy = threadIdx.y; yy < 32; yy += blockDim.y) {
        const int r = r0 + yy, c = c0 + threadIdx.x;
        t[yy][threadIdx.x] = (r < R && c < C) ? src[(size_t)r * lds + c] : 0.0f;
    }
    __syncthreads();
    for (int yy = threadIdx.y; yy < 32; yy += blockDim.y) {
        const int c = c0 + yy, r = r0 + threadIdx.x;
        if (c < C && r < seg) split3_store(t[threadIdx.x][yy], dst + (size_t)c * 3 * seg + r, seg, b_pattern != 0);
    }
}

// Z[m][i] = mu[i] + sum_s slab[s][m][i]  (i < D), 0 in the padding columns; a thread per coordinate quad (the launch
// brings one thread per quad, rounded up to whole warps).  HOOK: the target's per-sample preprocessing (SampleHook kind
// 1: TF32 image of beta for the forward contraction and the prior terms of glm_prior.cuh) in the same pass over z --
// what the mean-field sampler does for its family -- so the stand-alone k_glm_pre launch disappears.
template <bool HOOK>
__global__ void __launch_bounds__(512)
k_fr_zreduce(const float* __restrict__ slabs, int nslab, long long stride, const float* __restrict__ mu,
             int D, int ld, float* __restrict__ Z, SampleHook hk) {
    __shared__ float sm[33];
    const int m = blockIdx.x;
    if (HOOK && hk.zt_owner && m == 0 && threadIdx.x == 0) *hk.zt_owner = 0ull;
    float bsq = 0.f, eta = 0.f;
    const int qend = (HOOK && hk.Zt && hk.zt_seg > ld ? hk.zt_seg : ld) / 4;
    for (int q = threadIdx.x; q < qend; q += blockDim.x) {
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        if (4 * q < ld) {
            const size_t base = (size_t)m * ld + 4 * q;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int s0 = 0; s0 < nslab; s0 += FR_MAX_ZSLABS) {   // (one pass: nslab <= FR_MAX_ZSLABS; all loads first)
                float4 v[FR_MAX_ZSLABS];
#pragma unroll
                for (int u = 0; u < FR_MAX_ZSLABS; ++u)
                    v[u] = *reinterpret_cast<const float4*>(slabs + (size_t)min(s0 + u, nslab - 1) * stride + base);   // (unconditional: batched)
#pragma unroll
                for (int u = 0; u < FR_MAX_ZSLABS; ++u)
                    if (s0 + u < nslab) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
            }
            r[0] = acc.x; r[1] = acc.y; r[2] = acc.z; r[3] = acc.w;
#pragma unroll
            for (int c = 0; c < 4; ++c) r[c] = 4 * q + c < D ? r[c] + __ldg(mu + 4 * q + c) : 0.0f;
            *reinterpret_cast<float4*>(Z + base) = make_float4(r[0], r[1], r[2], r[3]);
        }
        if (HOOK) {
            float hi[4], lo[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int i = 4 * q + c;
                const float zb = i < hk.d ? r[c] : 0.0f;
                bsq = fmaf(zb, zb, bsq);
                if (i == hk.d) eta = r[c];
                hi[c] = tc::round_tf32(zb); lo[c] = tc::round_tf32(zb - hi[c]);
            }
            if (hk.Zt) {
                float* row = hk.Zt + (size_t)m * hk.zt_ld + 4 * q;
                if (hk.zt_seg == 0) {
                    if (4 * q < hk.zt_ld) *reinterpret_cast<float4*>(row) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                } else if (4 * q < hk.zt_seg) {   // 3xTF32: [hi | hi | lo]
                    *reinterpret_cast<float4*>(row) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(row + hk.zt_seg) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(row + 2 * hk.zt_seg) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
    }
    if (HOOK) {
        bsq = block_sum(bsq, sm);
        eta = block_sum(eta, sm);   // non-zero in exactly one thread
        if (threadIdx.x == 0) hk.pre[m] = glm_prior_terms(bsq, eta, hk.d, hk.variant, hk.include_prior);
    }
}

// One launch for everything the D x D pullback contraction needs from the sample-major buffers:
//   blockIdx.z == 0: Wt3 = transpose + split of W (A pattern);  == 1: Et3 = transpose + split of E (B pattern);
//   == 2 (first column of the grid only): v[i] = sum_m W[m][i], the location block's sums (k_colsum's arithmetic:
//   32 sample groups, combined in a fixed order).
//   pf.on: CTA (1, 0, 2) finishes the value slot (fr_finalize.cuh) and the column-sum CTAs write the location block of
//   the gradient themselves, so no finalize launch follows.
__global__ void __launch_bounds__(256)
k_fr_outer_prep(const float* __restrict__ W, const float* __restrict__ E, int Mloc, int D, int ld,
                float* __restrict__ Wt3, float* __restrict__ Et3, int seg, float* __restrict__ v, FrPrepFinalize pf) {
    __shared__ float t[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 warps
    if (blockIdx.z == 2) {
        if (blockIdx.x == 1 && blockIdx.y == 0 && pf.on) {
            fr_vec_finalize(v, pf.accv, pf.lambda, D, pf.M, pf.objective, pf.entropy, pf.grad, pf.out, pf.logp, pf.esq,
                            pf.Mloc, /*deferred=*/true, /*write_grad=*/false, &t[0][0], pf.h0);
            return;
        }
        if (blockIdx.x != 0) return;
        const int i = blockIdx.y * 32 + tx;
        float a[4] = {0.f, 0.f, 0.f, 0.f};   // sample groups ty, ty + 8, ty + 16, ty + 24 of k_colsum's 32
        if (i < D) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
                for (int m = ty + 8 * g; m < Mloc; m += 32) a[g] += W[(size_t)m * ld + i];
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) t[ty + 8 * g][tx] = a[g];
        __syncthreads();
        if (ty == 0 && i < D) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r) s += t[r][tx];
            v[i] = s;
            if (pf.on) pf.grad[i] = -s * (1.0f / (float)pf.M);   // RepGrad location block (fr_vec_finalize's formula)
        }
        return;
    }
    const float* src = blockIdx.z == 0 ? W : E;
    float* dst = blockIdx.z == 0 ? Wt3 : Et3;
    const bool b_pattern = blockIdx.z == 1;
    const int r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (int yy = ty; yy < 32; yy += 8) {
        const int r = r0 + yy, c = c0 + tx;
        t[yy][tx] = (r < Mloc && c < D) ? src[(size_t)r * ld + c] : 0.0f;
    }
    __syncthreads();
    for (int yy = ty; yy < 32; yy += 8) {
        const int c = c0 + yy, r = r0 + tx;
        if (c < D && r < seg) split3_store(t[tx][yy], dst + (size_t)c * 3 * seg + r, seg, b_pattern);
    }
}


int32_t ensure(avi_ctx* ctx, float** p, size_t* cap, size_t need) {
    if (need <= *cap) return AVI_OK;
    avi_free(*p);
    AVI_CHECK(avi_alloc(ctx, p, need));
    *cap = need;
    return AVI_OK;
}

}  // namespace

bool avi_fr_tc_ok(const avi_obj* o, int Mloc) { return Mloc > 0 && Mloc <= FR_TC_MAX_M; }

// scratch of avi_fr_affine_tc for Mloc samples (the sampler writes Er3 itself: the buffers must exist before it runs)
int32_t avi_fr_affine_prepare(avi_obj* o, int Mloc, float** Er3, int* seg) {
    avi_ctx* ctx = o->ctx;
    const int D = o->D, ld = o->ld;
    const int segD = (int)round_up(D, 32);
    FrWork& w = o->fr;
    const size_t gen0 = w.Lr3_cap + w.Er3_cap + w.zslab_cap;
    AVI_CHECK(ensure(ctx, &w.Lr3, &w.Lr3_cap, (size_t)D * 3 * segD));
    AVI_CHECK(ensure(ctx, &w.Er3, &w.Er3_cap, (size_t)Mloc * 3 * segD));
    AVI_CHECK(ensure(ctx, &w.zslab, &w.zslab_cap, (size_t)FR_MAX_ZSLABS * Mloc * ld));
    if (gen0 != w.Lr3_cap + w.Er3_cap + w.zslab_cap) { o->generation++; w.Lr3_owner = nullptr; }
    if (Er3) *Er3 = w.Er3;
    if (seg) *seg = segD;
    return AVI_OK;
}

int32_t avi_fr_refresh_split(avi_obj* o, const float* lambda) {
    avi_ctx* ctx = o->ctx;
    const int D = o->D;
    const int segD = (int)round_up(D, 32);
    FrWork& w = o->fr;
    if (!w.Lr3) AVI_CHECK(avi_fr_affine_prepare(o, std::max(o->Mloc, 1), nullptr, nullptr));
    // L is column-major: as a row-major matrix S[j][i] (pitch D) it is L'; Lr3[i][.] = split(S[.][i])
    dim3 tg((unsigned)ceil_div(segD, 32), (unsigned)ceil_div(D, 32));
    k_transpose_split3<<<tg, dim3(32, 8), 0, ctx->stream>>>(lambda + D, D, D, D, w.Lr3, segD, /*A pattern*/ 0);
    AVI_LAUNCHED(ctx);
    w.Lr3_owner = nullptr;
    return AVI_OK;
}

// Z = E * L' + mu on the tensor cores.  E: [Mloc][ld] sample-major eps; lambda = [mu; vec(L)].
// er3_done: the sampling kernel already wrote the split of eps (Er3).  hook (nullable): see k_fr_zreduce.
int32_t avi_fr_affine_tc(avi_obj* o, const float* lambda, const float* E, float* Z, int Mloc, bool er3_done,
                         const SampleHook* hook) {
    avi_ctx* ctx = o->ctx;
    const int D = o->D, ld = o->ld;
    const int segD = (int)round_up(D, 32);
    FrWork& w = o->fr;
    AVI_CHECK(avi_fr_affine_prepare(o, Mloc, nullptr, nullptr));
    if (!w.Lr3_maintained) AVI_CHECK(avi_fr_refresh_split(o, lambda));   // (else: the optimiser loop's update kernel keeps it, opt.cu)
    if (!er3_done) {
        k_split3_rows<<<Mloc, 256, 0, ctx->stream>>>(E, Mloc, D, ld, w.Er3, segD, /*B pattern*/ 1);
        AVI_LAUNCHED(ctx);
    }
    int nsl = 0;
    // a = coordinate i (rows of Lr3), b = sample m (rows of Er3): slab[ks][m * ld + i]
    AVI_CHECK(avi_tc_gemm_store(ctx, w.Lr3, D, 3LL * segD, w.Er3, Mloc, 3LL * segD, 3LL * segD, w.zslab, ld,
                                (int64_t)Mloc * ld, FR_MAX_ZSLABS, &nsl));
    const int qmax = (int)((hook && hook->Zt && hook->zt_seg > ld ? hook->zt_seg : ld) / 4);
    const unsigned threads = (unsigned)std::min<int64_t>(512, round_up(qmax, 32));
    if (hook && hook->kind == 1)
        k_fr_zreduce<true><<<Mloc, threads, 0, ctx->stream>>>(w.zslab, nsl, (long long)Mloc * ld, lambda, D, ld, Z, *hook);
    else
        k_fr_zreduce<false><<<Mloc, threads, 0, ctx->stream>>>(w.zslab, nsl, (long long)Mloc * ld, lambda, D, ld, Z, SampleHook{});
    AVI_LAUNCHED(ctx);
    return AVI_OK;
}

// C[j*D + i] = sum_m W[m][i] E[m][j]  (column-major L layout), contraction over the local samples.
// which: 0 -> C1 operands (Wt3), 1 -> C2 operands (Ut3); Et3 is shared when reuse_E.
// colsum (nullable, which == 0 && !reuse_E only): also v[i] = sum_m W[m][i], in the same preparation launch.
int32_t avi_fr_outer_tc(avi_obj* o, const float* W, const float* E, float* C, int Mloc, int which, bool reuse_E,
                        float* colsum, const FrPrepFinalize* pf) {
    avi_ctx* ctx = o->ctx;
    const int D = o->D, ld = o->ld;
    const int segM = (int)round_up(Mloc, 32);
    FrWork& w = o->fr;
    float** Wt = which == 0 ? &w.Wt3 : &w.Ut3;
    size_t* Wcap = which == 0 ? &w.Wt3_cap : &w.Ut3_cap;
    const size_t gen0 = *Wcap + w.Et3_cap;
    AVI_CHECK(ensure(ctx, Wt, Wcap, (size_t)D * 3 * segM));
    AVI_CHECK(ensure(ctx, &w.Et3, &w.Et3_cap, (size_t)D * 3 * segM));
    if (gen0 != *Wcap + w.Et3_cap) o->generation++;
    dim3 tg((unsigned)ceil_div(segM, 32), (unsigned)ceil_div(D, 32));
    if (!reuse_E) {
        tg.z = colsum ? 3 : 2;
        // (the finalize CTA sits at x == 1 of the column-sum layer: the grid is at least two tiles wide in x for that)
        if (colsum && pf && pf->on) tg.x = std::max(tg.x, 2u);
        k_fr_outer_prep<<<tg, 256, 0, ctx->stream>>>(W, E, Mloc, D, ld, *Wt, w.Et3, segM, colsum,
                                                     colsum && pf ? *pf : FrPrepFinalize{});
        AVI_LAUNCHED(ctx);
    } else {
        k_transpose_split3<<<tg, dim3(32, 8), 0, ctx->stream>>>(W, Mloc, D, ld, *Wt, segM, /*A pattern*/ 0);
        AVI_LAUNCHED(ctx);
    }
    // a = i (rows of Wt3), b = j (rows of Et3): C[b * D + a]
    AVI_CHECK(avi_tc_gemm_store(ctx, *Wt, D, 3LL * segM, w.Et3, D, 3LL * segM, 3LL * segM, C, D, 0, 1, nullptr));
    return AVI_OK;
}

void avi_fr_free(avi_obj* o) {
    FrWork& w = o->fr;
    avi_free(w.Lr3); avi_free(w.Er3); avi_free(w.Et3); avi_free(w.Wt3); avi_free(w.Ut3); avi_free(w.zslab);
    w = FrWork{};
}
