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
#include "tc_common.cuh"

namespace {

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

// Z[m][i] = mu[i] + sum_s slab[s][m][i]  (i < D), 0 in the padding columns; a thread per coordinate quad
__global__ void k_fr_zreduce(const float* __restrict__ slabs, int nslab, long long stride, const float* __restrict__ mu,
                             int D, int ld, float* __restrict__ Z) {
    const int m = blockIdx.x;
    for (int q = threadIdx.x; q < ld / 4; q += blockDim.x) {
        const size_t base = (size_t)m * ld + 4 * q;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int s = 0; s < nslab; ++s) {
            const float4 v = *reinterpret_cast<const float4*>(slabs + (size_t)s * stride + base);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        float r[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) r[c] = 4 * q + c < D ? r[c] + __ldg(mu + 4 * q + c) : 0.0f;
        *reinterpret_cast<float4*>(Z + base) = make_float4(r[0], r[1], r[2], r[3]);
    }
}

constexpr int FR_MAX_ZSLABS = 8;
constexpr int FR_TC_MAX_M = 4096;

int32_t ensure(avi_ctx* ctx, float** p, size_t* cap, size_t need) {
    if (need <= *cap) return AVI_OK;
    avi_free(*p);
    AVI_CHECK(avi_alloc(ctx, p, need));
    *cap = need;
    return AVI_OK;
}

}  // namespace

bool avi_fr_tc_ok(const avi_obj* o, int Mloc) { return Mloc > 0 && Mloc <= FR_TC_MAX_M; }

// Z = E * L' + mu on the tensor cores.  E: [Mloc][ld] sample-major eps; lambda = [mu; vec(L)].
int32_t avi_fr_affine_tc(avi_obj* o, const float* lambda, const float* E, float* Z, int Mloc) {
    avi_ctx* ctx = o->ctx;
    const int D = o->D, ld = o->ld;
    const int segD = (int)round_up(D, 32);
    FrWork& w = o->fr;
    const size_t gen0 = w.Lr3_cap + w.Er3_cap + w.zslab_cap;
    AVI_CHECK(ensure(ctx, &w.Lr3, &w.Lr3_cap, (size_t)D * 3 * segD));
    AVI_CHECK(ensure(ctx, &w.Er3, &w.Er3_cap, (size_t)Mloc * 3 * segD));
    AVI_CHECK(ensure(ctx, &w.zslab, &w.zslab_cap, (size_t)FR_MAX_ZSLABS * Mloc * ld));
    if (gen0 != w.Lr3_cap + w.Er3_cap + w.zslab_cap) o->generation++;
    // L is column-major: as a row-major matrix S[j][i] (pitch D) it is L'; Lr3[i][.] = split(S[.][i])
    dim3 tg((unsigned)ceil_div(segD, 32), (unsigned)ceil_div(D, 32));
    k_transpose_split3<<<tg, dim3(32, 8), 0, ctx->stream>>>(lambda + D, D, D, D, w.Lr3, segD, /*A pattern*/ 0);
    AVI_LAUNCHED(ctx);
    k_split3_rows<<<Mloc, 256, 0, ctx->stream>>>(E, Mloc, D, ld, w.Er3, segD, /*B pattern*/ 1);
    AVI_LAUNCHED(ctx);
    int nsl = 0;
    // a = coordinate i (rows of Lr3), b = sample m (rows of Er3): slab[ks][m * ld + i]
    AVI_CHECK(avi_tc_gemm_store(ctx, w.Lr3, D, 3LL * segD, w.Er3, Mloc, 3LL * segD, 3LL * segD, w.zslab, ld,
                                (int64_t)Mloc * ld, FR_MAX_ZSLABS, &nsl));
    k_fr_zreduce<<<Mloc, 256, 0, ctx->stream>>>(w.zslab, nsl, (long long)Mloc * ld, lambda, D, ld, Z);
    AVI_LAUNCHED(ctx);
    return AVI_OK;
}

// C[j*D + i] = sum_m W[m][i] E[m][j]  (column-major L layout), contraction over the local samples.
// which: 0 -> C1 operands (Wt3), 1 -> C2 operands (Ut3); Et3 is shared when reuse_E.
int32_t avi_fr_outer_tc(avi_obj* o, const float* W, const float* E, float* C, int Mloc, int which, bool reuse_E) {
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
    k_transpose_split3<<<tg, dim3(32, 8), 0, ctx->stream>>>(W, Mloc, D, ld, *Wt, segM, /*A pattern*/ 0);
    AVI_LAUNCHED(ctx);
    if (!reuse_E) {
        k_transpose_split3<<<tg, dim3(32, 8), 0, ctx->stream>>>(E, Mloc, D, ld, w.Et3, segM, /*B pattern*/ 1);
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
