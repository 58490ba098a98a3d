// Generic fp32 SIMT GEMM and triangular solve used by the exact-fp32 mode of the GLM targets and
// by the full-rank family (W' * eps, L^-T eps).  Arbitrary strides, bounds-checked, deterministic.
#include "avi_internal.cuh"
#include "device_utils.cuh"

namespace {

constexpr int TM = 64, TN = 64, TK = 16;

template <bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(256)
gemm_simt_kernel(const float* __restrict__ A, long long sa_r, long long sa_k,
                 const float* __restrict__ B, long long sb_r, long long sb_k,
                 float* __restrict__ C, long long sc_r, long long sc_c,
                 int Ma, int Nb, int K_in, float alpha, int tri_b) {
    __shared__ __align__(16) float As[TK][TM + 4];
    __shared__ __align__(16) float Bs[TK][TN + 4];
    const int tile_a = blockIdx.y * TM, tile_b = blockIdx.x * TN;
    // tri_b: B[b][k] == 0 for k > b (lower-triangular L): the contraction stops at the tile's last row
    const int K = tri_b ? min(K_in, tile_b + TN) : K_in;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

    for (int kt = 0; kt < K; kt += TK) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            int idx = tid + e * 256;
            int r, kk;
            if (A_KCONTIG) { kk = idx & 15; r = idx >> 4; } else { r = idx & 63; kk = idx >> 6; }
            int gr = tile_a + r, gk = kt + kk;
            As[kk][r] = (gr < Ma && gk < K) ? __ldg(A + gr * sa_r + gk * sa_k) : 0.0f;
            if (B_KCONTIG) { kk = idx & 15; r = idx >> 4; } else { r = idx & 63; kk = idx >> 6; }
            gr = tile_b + r; gk = kt + kk;
            Bs[kk][r] = (gr < Nb && gk < K) ? __ldg(B + gr * sb_r + gk * sb_k) : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < TK; ++kk) {
            float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int r = tile_a + ty * 4 + i;
        if (r >= Ma) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int c = tile_b + tx * 4 + j;
            if (c < Nb) C[r * sc_r + c * sc_c] = alpha * acc[i][j];
        }
    }
}

// ------------------------------------------------------------------------------------------
// U[m][:] = L^-T E[m][:]  (solve L' u = e; L lower triangular, column-major, D x D).
// CTA = S samples, 8 warps.  Blocks of 32 coordinates are solved last to first: phase 1 removes
// the contribution of the already-solved coordinates (dot products down the L columns, coalesced),
// phase 2 back-substitutes inside the 32 x 32 diagonal block with warp shuffles.
template <int S>
__global__ void __launch_bounds__(256)
k_trsm_lt(const float* __restrict__ L, int D, const float* __restrict__ E, float* __restrict__ U, int ld, int M, BaseDist bd) {
    extern __shared__ float smem[];
    float* u = smem;                      // [S][D]
    float* rb = u + (size_t)S * D;        // [S][32]
    float* Lb = rb + S * 32;              // [32][33]
    const int mb = blockIdx.x * S;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nblk = (D + 31) / 32;
    for (int b = nblk - 1; b >= 0; --b) {
        const int ib = b * 32;
        // diagonal block -> shared
        for (int e = threadIdx.x; e < 32 * 32; e += 256) {
            int k = e & 31, l = e >> 5;   // Lb[k][l] = L[ib + k][ib + l]
            int gk = ib + k, gl = ib + l;
            Lb[k * 33 + l] = (gk < D && gl < D && gk >= gl) ? __ldg(L + (size_t)gl * D + gk) : (gk == gl ? 1.0f : 0.0f);
        }
        // phase 1
        for (int ii = w; ii < 32; ii += 8) {
            const int i = ib + ii;
            float part[S];
#pragma unroll
            for (int s = 0; s < S; ++s) part[s] = 0.0f;
            if (i < D) {
                const float* Lc = L + (size_t)i * D;
                for (int j = ib + 32 + lane; j < D; j += 32) {
                    float l = __ldg(Lc + j);
#pragma unroll
                    for (int s = 0; s < S; ++s) part[s] = fmaf(l, u[(size_t)s * D + j], part[s]);
                }
            }
#pragma unroll
            for (int s = 0; s < S; ++s) {
                float t = warp_sum(part[s]);
                if (lane == 0) {
                    int m = mb + s;
                    float e = (i < D && m < M) ? base_negscore(bd, E[(size_t)m * ld + i]) : 0.0f;
                    rb[s * 32 + ii] = e - t;
                }
            }
        }
        __syncthreads();
        // phase 2
        for (int s = w; s < S; s += 8) {
            float r = rb[s * 32 + lane];
            float ul = 0.0f;
            for (int k = 31; k >= 0; --k) {
                float uk = __shfl_sync(0xffffffffu, r, k) / Lb[k * 33 + k];
                if (lane == k) ul = uk;
                if (lane < k) r = fmaf(-Lb[k * 33 + lane], uk, r);
            }
            if (ib + lane < D) u[(size_t)s * D + ib + lane] = ul;
        }
        __syncthreads();
    }
    for (int s = 0; s < S; ++s) {
        int m = mb + s;
        if (m >= M) break;
        for (int i = threadIdx.x; i < ld; i += 256) U[(size_t)m * ld + i] = i < D ? u[(size_t)s * D + i] : 0.0f;
    }
}

template <int S>
int32_t launch_trsm(avi_ctx* ctx, const float* L, int D, const float* E, float* U, int ld, int M, const BaseDist& bd) {
    size_t smem = ((size_t)S * D + S * 32 + 32 * 33) * sizeof(float);
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_trsm_lt<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) AVI_FAIL(ctx, AVI_ERR_CUDA, std::string("trsm smem: ") + cudaGetErrorString(e));
    }
    k_trsm_lt<S><<<(unsigned)ceil_div(M, S), 256, smem, ctx->stream>>>(L, D, E, U, ld, M, bd);
    AVI_LAUNCHED(ctx);
    return AVI_OK;
}

}  // namespace

int32_t avi_gemm_simt(avi_ctx* ctx, const float* A, long long sa_r, long long sa_k, const float* B,
                      long long sb_r, long long sb_k, float* C, long long sc_r, long long sc_c, int Ma,
                      int Nb, int K, float alpha, int tri_b) {
    if (Ma <= 0 || Nb <= 0 || K <= 0) return AVI_OK;
    dim3 grid((unsigned)ceil_div(Nb, TN), (unsigned)ceil_div(Ma, TM));
    bool ak = (sa_k == 1), bk = (sb_k == 1);
#define LAUNCH(AK, BK)                                                                              \
    gemm_simt_kernel<AK, BK><<<grid, 256, 0, ctx->stream>>>(A, sa_r, sa_k, B, sb_r, sb_k, C, sc_r, \
                                                            sc_c, Ma, Nb, K, alpha, tri_b)
    if (ak && bk) LAUNCH(true, true);
    else if (ak) LAUNCH(true, false);
    else if (bk) LAUNCH(false, true);
    else LAUNCH(false, false);
#undef LAUNCH
    AVI_LAUNCHED(ctx);
    return AVI_OK;
}

int32_t avi_trsm_lt(avi_ctx* ctx, const float* L, int D, const float* E, float* U, int ld, int M, const BaseDist& bd) {
    if (M <= 0) return AVI_OK;
    const size_t budget = 200 * 1024 - (32 * 33) * sizeof(float);
    auto fits = [&](int S) { return ((size_t)S * D + S * 32) * sizeof(float) <= budget; };
    if (fits(16) && M >= 16) return launch_trsm<16>(ctx, L, D, E, U, ld, M, bd);
    if (fits(4)) return launch_trsm<4>(ctx, L, D, E, U, ld, M, bd);
    if (fits(1)) return launch_trsm<1>(ctx, L, D, E, U, ld, M, bd);
    AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED, "full-rank dimension too large for the triangular solve");
}
