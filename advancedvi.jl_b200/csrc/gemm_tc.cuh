// Host-side interface of the tcgen05 contraction kernel (gemm_tc.cu).
#pragma once

#include <cuda.h>
#include <stdint.h>

struct avi_ctx;

enum { EPI_GLM_FWD = 0, EPI_GLM_BWD = 1, EPI_STORE = 2 };

// D[a, b] = sum_k A[a, k] B[b, k];  a < Ma (blocks of 128), b < Nb (chunks of nt), k in 32-float blocks
struct TcParams {
    int Ma, Nb;
    int n_ablk, n_bchunk, n_ksplit;
    int n_kblk, kb_per_split;
    int nt;
    int pair;            // 1: CTA-pair kernel (cta_group::2): A box = 128 rows, B box = nt/2 rows
    int ca, cb;          // cluster shape: ca a-blocks x cb b-chunks share operands by TMA multicast
    int stages;          // filled by avi_tc_launch
    int static_op;       // 1: A is static data (not written by the previous kernel), 2: B is, 0: neither
    // fused iteration kernel only: A is read MN-major, i.e. from a matrix stored [k][a] (a contiguous): the backward
    // contraction then takes X from the SAME row-major copy the forward contraction uses (X is read once per iteration).
    // The A tensor map has box {32 a-elements, 32 k-rows} and the 32-byte-atom swizzle; a 128 x 32 tile is four boxes.  3xTF32: the k range consists
    // of three segments of a_seg_kb blocks [hi | hi | lo]; the lo part of the source sits a_seg_off elements further.
    int a_mn, a_seg_kb, a_seg_off;   // a_mn / b_mn = 2: one 3-D box per tile (avi_tc_make_tmap_mn3) instead of one 2-D box per group
    // fused iteration kernel only: the forward epilogue stores R TRANSPOSED, Rt [data row][sample] (c_mn = 1, ldc = row pitch
    // of Rt), so that the 32 lanes of a warp -- 32 consecutive samples -- write 128 contiguous bytes per data row (the
    // sample-major layout made every warp store touch 32 different rows: 2.4 us of the forward epilogue); the backward
    // contraction then reads it as an MN-major B operand (b_mn = 1): boxes of {32 samples, 32 data rows}, 4 KB apart.
    int c_mn, b_mn;
    const void* pf_ptr;  // static operand of the NEXT kernel, pulled into L2 by the idle warp 3 while this one computes
    unsigned long long pf_bytes;
    unsigned pf_pace_ns;
    unsigned long long* tl;     // step timeline (diagnostic), slot id tl_id
    int tl_id;
    unsigned long long* prof;   // AVI_TC_PROF: per-CTA phase timestamps
    int dbg;             // AVI_TC_DBG timing experiments (results are then meaningless): 1 no operand loads, 2 no MMAs
    // epilogue operands
    float* C;            // FWD: R [a][ldc];  STORE: slabs [ks][b * ldc + a]
    int ldc;
    long long slab_stride;
    const float* y;      // FWD: response per data row (b)
    float w;             // FWD: likelihood adjustment folded into R
    int r_seg;           // FWD, 3xTF32: R rows hold [hi | lo | hi] in segments of r_seg floats (0: plain TF32)
    int likelihood;
    const float* E;      // BWD: eps [b][lde]
    int lde;
    float* part1;        // FWD: partial log-lik [(bc*4+cq)][ldpart];  BWD: sum g  [ks*n_bchunk+bc][ldpart]
    float* part2;        // BWD: sum g*eps
    int ldpart;
    // BWD, fused post-processing (replaces k_glm_post_sums): prior gradient, log pi assembly, slab combine
    int post_on;
    const float* post_Z;       // z [b][lde]
    const float* post_pre;     // float4 per sample: {log prior, 1/sigma^2, d log pi / d eta, |beta|^2}
    const float* post_llpart;  // partial log-likelihood sums of the forward kernel [post_nparts][post_ldll]
    int post_nparts, post_ldll;
    float post_w;
    float* post_logp;          // [Nb]
    float *post_a1, *post_a2;  // final sum_m g, sum_m g*eps  [Ma + 1]
    unsigned int* post_tickets;   // [n_ablk], zero between launches
};

// atom32 = 1: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B (MN-major TF32 operand, tc_common.cuh) instead of the 128-byte swizzle
int32_t avi_tc_make_tmap(avi_ctx* ctx, CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld,
                         int box_rows, int atom32 = 0);
// MN-major operand [k_rows][ld] (MN index contiguous, mn_cols % 32 == 0) as one 3-D box {32, 32 k rows, groups} per tile
int32_t avi_tc_make_tmap_mn3(avi_ctx* ctx, CUtensorMap* map, const float* base, int64_t k_rows, int64_t mn_cols, int64_t ld,
                             int groups);
// fills Ma, Nb, n_ablk, n_kblk, nt, n_bchunk, n_ksplit, kb_per_split, ca, cb.  force_cluster: 0 = never cluster
// allow_pair = 0: single-CTA tiles only (the fused iteration kernel)
// max_ksplit > 0 bounds the number of k-splits; nt_search: split-K plans may use tiles narrower than the widest one
int32_t avi_tc_plan(avi_ctx* ctx, int64_t Ma, int64_t Nb, int64_t K, bool split_k, int force_cluster, TcParams* p,
                    int allow_pair = 1, int max_ksplit = 0, int nt_search = 0);
int32_t avi_tc_launch(avi_ctx* ctx, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p);
// generic contraction with the plain store epilogue (full-rank family): see gemm_tc.cu
int32_t avi_tc_gemm_store(avi_ctx* ctx, const float* A, int64_t Ma, int64_t lda, const float* B, int64_t Nb, int64_t ldb,
                          int64_t K, float* C, int64_t ldc, int64_t slab_stride, int max_slabs, int* n_slabs);
