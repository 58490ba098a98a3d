// tcgen05 / TMEM / TMA contraction  D[a, b] = sum_k A[a, k] * B[b, k]  (both operands K-major fp32
// rows in HBM, read as TF32) with three fused epilogues:
//   EPI_GLM_FWD : a = Monte-Carlo sample, b = data row, k = feature.   D = logits.  Writes the
//                 weighted residual R[a][b] (TF32-rounded: it is the B operand of the backward
//                 contraction) and per-(row-chunk) partial log-likelihood sums.  The logits never
//                 reach HBM.                         (X*beta + BernoulliLogit/Normal log-likelihood,
//                 docs/src/tutorials/subsampling.md:35-36 evaluated for all M samples at once
//                 instead of the per-sample loop of src/algorithms/repgradelbo.jl:84-86)
//   EPI_GLM_BWD : a = feature, b = sample, k = data row.  D = X' R = grad_beta log-lik for every
//                 sample.  Reduced against eps in the epilogue to sum_m g and sum_m g*eps (the
//                 mean-field pullback of src/families/location_scale.jl:86); G never reaches HBM.
//   EPI_STORE   : C[ks][b * ldc + a] = D (split-K slabs), for the full-rank family.
//
// Structure (one CTA per SM, persistent over work units = (a-block, b-chunk, k-split)):
//   warp 0      TMA producer   (cp.async.bulk.tensor, 128B swizzle, mbarrier ring as deep as shared memory allows)
//   warp 1      MMA issuer     (one thread chosen by elect.sync -- NOT by `lane == 0`, which makes nvcc wrap every
//                               tcgen05.mma in a divergence loop: tcgen05.mma.kind::tf32, M = 128, N = NT)
//   warp 2      TMEM allocator (512 columns = 2 accumulator stages of up to 256 columns)
//   warp 3      idle (optionally pulls the next kernel's static operand into L2: TcParams.pf_*)
//   warps 4-19  epilogue       (tcgen05.ld 32x32b; warp w reads lanes 32*(w%4).., column quarter (w-4)/4)
// k_gemm_tc_pair is the cta_group::2 variant (UMMA M = 256 over two CTAs).  What bounds the mainloop (tensor pipe
// max(48, NT/2) cycles per MMA vs ~71 B/cycle of TMA ingest per SM) is measured in profiles/README.md and is the
// cost model of avi_tc_plan.
#include <cstdio>
#include <cstdlib>

#include "avi_internal.cuh"
#include "device_utils.cuh"
#include "gemm_tc.cuh"
#include "tc_common.cuh"
#include "gemm_tc_dev.cuh"

namespace {

// LIK: 0 Bernoulli-logit, 1 Gaussian (EPI_GLM_FWD only)
template <int EPI, int LIK>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_gemm_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int NT = p.nt;
    const int stage_bytes = A_TILE_BYTES + NT * BK * 4;
    const int stages = p.stages;
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(tiles + stages * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    TC_STAMP(0, threadIdx.x == 0);
    tl_min(p.tl, p.tl_id);

    // thread-block cluster: CA a-blocks x CB b-chunks of one k-split.  The A tile of an a-block is needed
    // by the CB CTAs of its row, the B tile of a b-chunk by the CA CTAs of its column: every CTA loads
    // 1/CB of its A tile and 1/CA of its B tile and TMA-multicasts the slice to the CTAs that share it,
    // so each operand byte crosses the L2 -> SM crossbar once per cluster instead of once per CTA.
    const int CA = p.ca, CB = p.cb, CSZ = CA * CB;
    const uint32_t crank = CSZ > 1 ? tc::cluster_ctarank() : 0u;
    const int ra = (int)crank % CA, rb = (int)crank / CA;
    uint32_t a_mask = 0, b_mask = 0;
    for (int j = 0; j < CB; ++j) a_mask |= 1u << (ra + CA * j);
    for (int j = 0; j < CA; ++j) b_mask |= 1u << (j + CA * rb);
    const uint32_t peer_mask = a_mask | b_mask;   // CTAs that write into my stages == CTAs I write into
    const int cluster_id = blockIdx.x / CSZ, n_clusters = gridDim.x / CSZ;

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmA);
        tc::tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) { tc::mbar_init(&ctl->full[s], 1); tc::mbar_init(&ctl->empty[s], CA + CB - 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&ctl->tmem_full[s], 1); tc::mbar_init(&ctl->tmem_empty[s], EPI_WARPS); }
        tc::mbar_fence_init();
    }
    if (warp == 2) tc::tmem_alloc(&ctl->tmem_base, 512);
    pdl_trigger();
    tc::fence_before_sync();
    __syncthreads();
    if (CSZ > 1) tc::cluster_sync();   // peers must see initialised barriers before any multicast lands

    // cluster units: (a-group, b-group, k-split); this CTA's tile inside the unit is (ra, rb)
    const int n_ag = (p.n_ablk + CA - 1) / CA, n_bg = (p.n_bchunk + CB - 1) / CB;
    const int units = n_ag * n_bg * p.n_ksplit;
    const uint32_t stage_tx = (uint32_t)stage_bytes;
    const int a_rows = BM / CB, b_rows = NT / CA;   // rows of the slices this CTA loads
#define UNIT_COORDS(u)                                                                       \
    const int ab = ((u) % n_ag) * CA + ra, bc = (((u) / n_ag) % n_bg) * CB + rb, ks = (u) / (n_ag * n_bg)

    // One operand is static data (X): its tiles for the first ring fill are requested BEFORE waiting for the
    // previous kernel (programmatic dependent launch), hiding the cold-pipeline latency behind that kernel.
    // p.static_op: 1 = A is static, 2 = B is static, 0 = neither.  (single-CTA, non-multicast path only)
    int pre_issued = 0;
    if (p.static_op && CSZ == 1 && warp == 0 && lane == 0 && (int)blockIdx.x < units) {
        const int u = cluster_id;
        UNIT_COORDS(u);
        const int kb0 = ks * p.kb_per_split, kb1 = min(p.n_kblk, kb0 + p.kb_per_split);
        for (int kb = kb0; kb < kb1 && pre_issued < stages; ++kb, ++pre_issued) {
            uint8_t* sa = tiles + pre_issued * stage_bytes;
            if (p.static_op == 1) {
                tc::mbar_expect_tx(&ctl->full[pre_issued], (uint32_t)A_TILE_BYTES);
                tc::tma_load_2d(sa, &tmA, &ctl->full[pre_issued], kb * BK, ab * BM);
            } else {
                tc::mbar_expect_tx(&ctl->full[pre_issued], (uint32_t)(NT * BK * 4));
                tc::tma_load_2d(sa + A_TILE_BYTES, &tmB, &ctl->full[pre_issued], kb * BK, bc * NT);
            }
        }
    }
    pdl_wait();   // everything above overlapped the previous kernel's tail; dependent operands from here on
    tl_min(p.tl, 4 + p.tl_id);
    tc::fence_after_sync();
    const uint32_t tmem_base = ctl->tmem_base;
    TC_STAMP(1, threadIdx.x == 0);

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (tc::elect_one()) {   // elect.sync: the compiler then treats the whole region as warp-uniform
            int stage = 0; uint32_t phase = 0;
            int kcount = 0;   // k-blocks issued so far by this CTA
            for (int u = cluster_id; u < units; u += n_clusters) {
                UNIT_COORDS(u);
                const int kb0 = ks * p.kb_per_split, kb1 = min(p.n_kblk, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb, ++kcount) {
                    tc::mbar_wait(&ctl->empty[stage], phase ^ 1);   // every CTA I write into has drained this stage
                    uint8_t* sa = tiles + stage * stage_bytes;
                    if (kcount < pre_issued) {
                        // the static operand of this k-block is already in flight: add the dependent one
                        if (p.static_op == 1) {
                            tc::mbar_arrive_expect_tx(&ctl->full[stage], (uint32_t)(NT * BK * 4));
                            tc::tma_load_2d(sa + A_TILE_BYTES, &tmB, &ctl->full[stage], kb * BK, bc * NT);
                        } else {
                            tc::mbar_arrive_expect_tx(&ctl->full[stage], (uint32_t)A_TILE_BYTES);
                            tc::tma_load_2d(sa, &tmA, &ctl->full[stage], kb * BK, ab * BM);
                        }
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    if (p.dbg & 1) {   // timing experiment: no operand traffic (MMA + epilogue bound)
                        tc::mbar_arrive(&ctl->full[stage]);
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    tc::mbar_arrive_expect_tx(&ctl->full[stage], stage_tx);
                    uint8_t* dst_a = sa + rb * a_rows * (BK * 4);
                    uint8_t* dst_b = sa + A_TILE_BYTES + ra * b_rows * (BK * 4);
                    if (CB > 1) tc::tma_load_2d_mc(dst_a, &tmA, &ctl->full[stage], kb * BK, ab * BM + rb * a_rows, (uint16_t)a_mask);
                    else tc::tma_load_2d(dst_a, &tmA, &ctl->full[stage], kb * BK, ab * BM);
                    if (CA > 1) tc::tma_load_2d_mc(dst_b, &tmB, &ctl->full[stage], kb * BK, bc * NT + ra * b_rows, (uint16_t)b_mask);
                    else tc::tma_load_2d(dst_b, &tmB, &ctl->full[stage], kb * BK, bc * NT);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (tc::elect_one()) {   // elect.sync: the compiler then treats the whole region as warp-uniform
            const uint32_t idesc = tc::idesc_tf32(BM, NT);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            long long prof_wait = 0, prof_t0 = 0;   // AVI_TC_PROF: cycles stalled on full[] / first operands seen
            for (int u = cluster_id; u < units; u += n_clusters) {
                const int ks = u / (n_ag * n_bg);
                const int kb0 = ks * p.kb_per_split, kb1 = min(p.n_kblk, kb0 + p.kb_per_split);
                tc::mbar_wait(&ctl->tmem_empty[as], aphase ^ 1);
                tc::fence_after_sync();
                const uint32_t tacc = tmem_base + (uint32_t)(as * 256);
                for (int kb = kb0; kb < kb1; ++kb) {
                    const long long c0 = p.prof ? clock64() : 0;
                    if (!(p.dbg & 8)) tc::mbar_wait(&ctl->full[stage], phase);   // dbg 8: free-running issue (garbage operands)
                    if (p.prof) { const long long c1 = clock64(); if (prof_t0) prof_wait += c1 - c0; else prof_t0 = c1; }
                    tc::fence_after_sync();
                    TC_STAMP(2, kb == kb0 && u == cluster_id);
                    const uint32_t sa = tc::smem_u32(tiles + stage * stage_bytes);
                    const uint64_t da = tc::smem_desc_k_sw128(sa);
                    const uint64_t db = tc::smem_desc_k_sw128(sa + A_TILE_BYTES);
                    if (!(p.dbg & 2) || kb == kb0) {   // dbg bit 1: timing experiment without the MMAs (TMA bound)
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k)   // UMMA K = 8 tf32 = 32 bytes: +2 in 16-byte units
                        tc::umma_tf32(tacc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                      (kb > kb0 || k > 0) ? 1u : 0u);
                    }
                    if (CSZ > 1) tc::umma_commit_mc(&ctl->empty[stage], (uint16_t)peer_mask);
                    else tc::umma_commit(&ctl->empty[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit(&ctl->tmem_full[as]);
                TC_STAMP(3, true);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
            // slot 7: (cycles stalled on operands after the first k-block) << 32 | mainloop cycles
            if (p.prof) p.prof[(size_t)blockIdx.x * 8 + 7] = ((unsigned long long)prof_wait << 32) |
                                                            (unsigned long long)(uint32_t)(clock64() - prof_t0);
        }
    } else if (warp == 3) {
        if (p.pf_bytes) l2_prefetch_span(p.pf_ptr, p.pf_bytes, blockIdx.x, gridDim.x, 16384, p.pf_pace_ns);
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        // warp w may only touch TMEM lanes [32 (w % 4), +32); the 4 warps of a lane quarter split the
        // NT columns into runs of 8-column groups
        const int ew = warp - 4, quarter = warp & 3, cq = ew >> 2;
        const int et = threadIdx.x - 128;   // 0..511
        const int ngroups = NT / 8;
        const int c_begin = 8 * ((ngroups * cq) / 4), c_end = 8 * ((ngroups * (cq + 1)) / 4);
        int as = 0; uint32_t aphase = 0;
        for (int u = cluster_id; u < units; u += n_clusters) {
            UNIT_COORDS(u);
            const int a = ab * BM + quarter * 32 + lane;   // this thread's accumulator row
            const bool a_ok = a < p.Ma && bc < p.n_bchunk;   // tiles past the edge of the unit grid only feed the pipeline
            const uint32_t tacc = tmem_base + (uint32_t)(as * 256) + ((uint32_t)(quarter * 32) << 16);
            float s1, s2;
            epilogue_unit<EPI, LIK>(p, ctl, as, aphase, tacc, NT, a, a_ok, bc, ks, c_begin, c_end, et, s1, s2);
            TC_STAMP(5, et == 0);
            // accumulator stage drained: hand it back to the MMA warp
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&ctl->tmem_empty[as]);
            epilogue_store_partials<EPI>(p, a_ok, a, bc, ks, cq, s1, s2);
            if (EPI == EPI_GLM_BWD) epilogue_bwd_store(p, ctl, a_ok, a, bc, ks, cq, et, s1, s2);
            if (EPI == EPI_GLM_BWD && p.post_on) epilogue_bwd_combine(p, ctl, ab, et);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }

#undef UNIT_COORDS
    tc::fence_before_sync();
    __syncthreads();
    if (CSZ > 1) tc::cluster_sync();   // no peer may still signal my barriers / write my smem after I exit
    if (warp == 2) {
        tc::fence_after_sync();
        tc::tmem_dealloc(tmem_base, 512);
    }
    TC_STAMP(6, threadIdx.x == 64);
    tl_max(p.tl, 8 + p.tl_id);
}

// ---------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05.mma.cta_group::2, UMMA M = 256): a cluster of 2 CTAs works on two a-blocks of one
// (b-chunk, k-split).  Each CTA stages its own 128 x 32 A tile and HALF of the b-chunk (NT/2 rows); the leader
// CTA's elected thread issues the MMAs for both tensor cores; accumulator rows of CTA r stay in CTA r's TMEM and
// are drained by CTA r's epilogue warps.  Per SM and k-block this moves 16 KB + NT*64 B instead of
// 16 KB + NT*128 B, which is what bounds the 1-CTA kernel (SM ingress).
template <int EPI, int LIK>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_gemm_tc_pair(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int NT = p.nt;
    const int b_half_bytes = (NT / 2) * BK * 4;
    const int stage_bytes = A_TILE_BYTES + b_half_bytes;
    const int stages = p.stages;
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>(tiles + stages * stage_bytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();   // 0 = leader
    const int pair_id = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    TC_STAMP(0, threadIdx.x == 0);

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmA);
        tc::tma_prefetch_desc(&tmB);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < stages; ++s) { tc::mbar_init(&ctl->full[s], 1); tc::mbar_init(&ctl->empty[s], 1); }
        for (int s = 0; s < 2; ++s) { tc::mbar_init(&ctl->tmem_full[s], 1); tc::mbar_init(&ctl->tmem_empty[s], 2 * EPI_WARPS); }
        tc::mbar_fence_init();
    }
    if (warp == 2) tc::tmem_alloc_cg2(&ctl->tmem_base, 512);
    pdl_trigger();
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync();
    pdl_wait();
    tc::fence_after_sync();
    const uint32_t tmem_base = ctl->tmem_base;
    TC_STAMP(1, threadIdx.x == 0);

    const int n_ag = (p.n_ablk + 1) / 2;
    const int units = n_ag * p.n_bchunk * p.n_ksplit;
#define PAIR_COORDS(u) \
    const int ab = ((u) % n_ag) * 2 + (int)rank, bc = ((u) / n_ag) % p.n_bchunk, ks = (u) / (n_ag * p.n_bchunk)

    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (tc::elect_one()) {   // elect.sync: the compiler then treats the whole region as warp-uniform
            int stage = 0; uint32_t phase = 0;
            for (int u = pair_id; u < units; u += n_pairs) {
                PAIR_COORDS(u);
                const int kb0 = ks * p.kb_per_split, kb1 = min(p.n_kblk, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb) {
                    tc::mbar_wait(&ctl->empty[stage], phase ^ 1);   // released by the leader's commit (multicast)
                    uint8_t* sa = tiles + stage * stage_bytes;
                    const uint32_t leader_full = tc::mapa_u32(tc::smem_u32(&ctl->full[stage]), 0u);
                    if (rank == 0) tc::mbar_arrive_expect_tx(&ctl->full[stage], 2u * (uint32_t)stage_bytes);
                    tc::tma_load_2d_cg2(sa, &tmA, leader_full, kb * BK, ab * BM);
                    tc::tma_load_2d_cg2(sa + A_TILE_BYTES, &tmB, leader_full, kb * BK, bc * NT + (int)rank * (NT / 2));
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0 && tc::elect_one()) {
            const uint32_t idesc = tc::idesc_tf32(2 * BM, NT);
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int u = pair_id; u < units; u += n_pairs) {
                const int ks = u / (n_ag * p.n_bchunk);
                const int kb0 = ks * p.kb_per_split, kb1 = min(p.n_kblk, kb0 + p.kb_per_split);
                tc::mbar_wait(&ctl->tmem_empty[as], aphase ^ 1);   // both CTAs' epilogues drained this stage
                tc::fence_after_sync();
                const uint32_t tacc = tmem_base + (uint32_t)(as * 256);
                for (int kb = kb0; kb < kb1; ++kb) {
                    tc::mbar_wait(&ctl->full[stage], phase);       // both CTAs' tiles have landed
                    tc::fence_after_sync();
                    TC_STAMP(2, kb == kb0 && u == pair_id);
                    const uint32_t sa = tc::smem_u32(tiles + stage * stage_bytes);
                    const uint64_t da = tc::smem_desc_k_sw128(sa);
                    const uint64_t db = tc::smem_desc_k_sw128(sa + A_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k)
                        tc::umma_tf32_cg2(tacc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc,
                                          (kb > kb0 || k > 0) ? 1u : 0u);
                    tc::umma_commit_cg2_mc(&ctl->empty[stage], (uint16_t)3);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit_cg2_mc(&ctl->tmem_full[as], (uint16_t)3);
                TC_STAMP(3, true);
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue (both CTAs, own accumulator rows) =====================
        const int ew = warp - 4, quarter = warp & 3, cq = ew >> 2;
        const int et = threadIdx.x - 128;
        const int ngroups = NT / 8;
        const int c_begin = 8 * ((ngroups * cq) / 4), c_end = 8 * ((ngroups * (cq + 1)) / 4);
        int as = 0; uint32_t aphase = 0;
        for (int u = pair_id; u < units; u += n_pairs) {
            PAIR_COORDS(u);
            const int a = ab * BM + quarter * 32 + lane;
            const bool a_ok = a < p.Ma;
            const uint32_t tacc = tmem_base + (uint32_t)(as * 256) + ((uint32_t)(quarter * 32) << 16);
            float s1, s2;
            epilogue_unit<EPI, LIK>(p, ctl, as, aphase, tacc, NT, a, a_ok, bc, ks, c_begin, c_end, et, s1, s2);
            TC_STAMP(5, et == 0);
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(tc::mapa_u32(tc::smem_u32(&ctl->tmem_empty[as]), 0u));
            epilogue_store_partials<EPI>(p, a_ok, a, bc, ks, cq, s1, s2);
            if (EPI == EPI_GLM_BWD) epilogue_bwd_store(p, ctl, a_ok, a, bc, ks, cq, et, s1, s2);
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }
#undef PAIR_COORDS
    tc::fence_before_sync();
    __syncthreads();
    tc::cluster_sync();
    if (warp == 2) {
        tc::fence_after_sync();
        tc::tmem_dealloc_cg2(tmem_base, 512);
    }
    TC_STAMP(6, threadIdx.x == 64);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

}  // namespace

// rows x cols fp32 matrix with row pitch ld (floats); box = box_rows x 32 floats, 128B swizzle,
// out-of-bounds elements read as zero.
int32_t avi_tc_make_tmap(avi_ctx* ctx, CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld,
                         int box_rows, int atom32) {
    EncodeTiledFn enc = get_encode();
    if (!enc) AVI_FAIL(ctx, AVI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 4) || box_rows < 1 || box_rows > 256 || rows < 1 || cols < 1)
        AVI_FAIL(ctx, AVI_ERR_INVALID, "tensor map: misaligned operand");
    cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t gstr[1] = {(cuuint64_t)ld * sizeof(float)};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) AVI_FAIL(ctx, AVI_ERR_CUDA, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
    return AVI_OK;
}

// MN-major TF32 operand stored [k_rows][ld] with the MN index contiguous (mn_cols of them, a multiple of 32), as ONE box
// per tile: a 3-D view {32 MN elements | k rows | groups of 32 MN elements} whose box {32, 32, groups} lands in shared
// memory as `groups` consecutive 4 KB blocks -- exactly the layout smem_desc_mn_sw128(.., 4096) describes.  (Separate
// 2-D boxes per group work too, but every TMA request costs the producer ~18 cycles of the SM's ingest time.)
int32_t avi_tc_make_tmap_mn3(avi_ctx* ctx, CUtensorMap* map, const float* base, int64_t k_rows, int64_t mn_cols, int64_t ld,
                             int groups) {
    EncodeTiledFn enc = get_encode();
    if (!enc) AVI_FAIL(ctx, AVI_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    if ((reinterpret_cast<uintptr_t>(base) & 15) || (ld % 4) || (mn_cols % 32) || mn_cols > ld || groups < 1 || groups > 8 || k_rows < 1)
        AVI_FAIL(ctx, AVI_ERR_INVALID, "tensor map (MN-major): misaligned operand");
    cuuint64_t gdim[3] = {32, (cuuint64_t)k_rows, (cuuint64_t)(mn_cols / 32)};
    cuuint64_t gstr[2] = {(cuuint64_t)ld * sizeof(float), 128};
    cuuint32_t box[3] = {32, (cuuint32_t)BK, (cuuint32_t)groups};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstr, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) AVI_FAIL(ctx, AVI_ERR_CUDA, "cuTensorMapEncodeTiled (MN-major 3-D) failed: " + std::to_string((int)r));
    return AVI_OK;
}

static int smem_bytes_for(int nt, int* stages_out, bool pair = false, bool fwd = false) {
    const int stage_bytes = A_TILE_BYTES + (pair ? nt / 2 : nt) * BK * 4;
    const int extra = 1024 + (int)sizeof(SmemCtl);
    (void)fwd;
    const int stages = std::max(2, std::min(MAX_STAGES, (SMEM_LIMIT - extra) / stage_bytes));
    if (stages_out) *stages_out = stages;
    return stages * stage_bytes + extra;
}

static int32_t set_attrs(avi_ctx* ctx) {
    static bool attr_done = false;
    if (attr_done) return AVI_OK;
    AVI_CUDA(ctx, cudaFuncSetAttribute(k_gemm_tc<EPI_GLM_FWD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    AVI_CUDA(ctx, cudaFuncSetAttribute(k_gemm_tc<EPI_GLM_FWD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    AVI_CUDA(ctx, cudaFuncSetAttribute(k_gemm_tc<EPI_GLM_BWD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    AVI_CUDA(ctx, cudaFuncSetAttribute(k_gemm_tc<EPI_STORE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    AVI_CUDA(ctx, cudaFuncSetAttribute(k_gemm_tc_pair<EPI_GLM_FWD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    AVI_CUDA(ctx, cudaFuncSetAttribute(k_gemm_tc_pair<EPI_GLM_FWD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    AVI_CUDA(ctx, cudaFuncSetAttribute(k_gemm_tc_pair<EPI_GLM_BWD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    AVI_CUDA(ctx, cudaFuncSetAttribute(k_gemm_tc_pair<EPI_STORE, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
    attr_done = true;
    return AVI_OK;
}

static int env_int(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}

// how many clusters of `csz` CTAs (one CTA per SM, full shared memory) can be resident at once
static int max_active_clusters(avi_ctx* ctx, int csz) {
    static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (csz <= 1) return ctx->prop.multiProcessorCount;
    if (cache[csz]) return cache[csz];
    if (set_attrs(ctx) != AVI_OK) return 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(csz * 64)); cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem_bytes_for(256, nullptr); cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, (const void*)k_gemm_tc<EPI_GLM_BWD, 0>, &cfg) != cudaSuccess) { cudaGetLastError(); n = 0; }
    cache[csz] = n > 0 ? n : -1;
    return cache[csz];
}

// Tiling plan.  split_k = false: many b-chunks, full K per unit (the epilogue is non-linear: GLM forward);
// split_k = true: few output tiles, K split across the chip (linear epilogues).
// Cost model (cycles; scripts/ubench/mma_rate.cu and profiles/r2_tc_sweep.txt): a kind::tf32 M=128 MMA (K = 8) costs
// max(48, nt / 2) cycles on the tensor pipe, but an SM ingests TMA operand bytes at only ~71 B/cycle, so a 32-wide
// k-block of a 128 x nt tile costs max(4 * max(48, nt / 2), (128 + nt) * 128 / 71): every single-CTA TF32 tile is
// ingest-bound ((128 + nt) * 128 B per 2 * nt MMA cycles > 71 B/cycle for all nt <= 256).  The epilogue costs ~58
// cycles per column; ~6000 fixed per wave.  Multicast does not reduce what an SM ingests, so clusters only pay when
// L2 bandwidth is the limit; they stay available (force_cluster = 2 / AVI_TC_CLUSTER=2, AVI_TC_CA/CB) and tested.
int32_t avi_tc_plan(avi_ctx* ctx, int64_t Ma, int64_t Nb, int64_t K, bool split_k, int force_cluster, TcParams* p,
                    int allow_pair, int max_ksplit, int nt_search) {
    // split-K plans use the widest tile unless nt_search (plain store epilogue: a narrower tile means more b-chunks,
    // fewer k-splits -- fewer partial slabs for the reduction that follows -- and a shorter epilogue per CTA);
    // AVI_TC_SNT forces a width for A/B runs
    const int snt_env = env_int("AVI_TC_SNT", 0);
    if (snt_env < 0) nt_search = 0;
    const int sms = ctx->prop.multiProcessorCount;
    p->Ma = (int)Ma; p->Nb = (int)Nb;
    p->n_ablk = (int)ceil_div(Ma, BM);
    p->n_kblk = (int)ceil_div(K, BK);
    double best = 1e300;
    const int nt_hi = (int)std::min<int64_t>(256, round_up(Nb, 16));
    for (int ca = 1; ca <= 8; ca *= 2) {
        if (ca > 1 && ca > p->n_ablk) break;
        for (int cb = 1; ca * cb <= 8; cb *= 2) {
            const int csz = ca * cb;
            if (force_cluster == 0 && csz > 1) continue;
            const int maxc = max_active_clusters(ctx, csz);
            if (maxc <= 0) continue;
            // experiment overrides (read per call): AVI_TC_NT = forward tile width, AVI_TC_CA / AVI_TC_CB = cluster shape of
            // the forward (no split-K) plan, AVI_TC_BCA / AVI_TC_BCB = of the split-K plan
            const int nt_env = env_int("AVI_TC_NT", 0);
            const int ca_env = env_int(split_k ? "AVI_TC_BCA" : "AVI_TC_CA", 0), cb_env = env_int(split_k ? "AVI_TC_BCB" : "AVI_TC_CB", 0);
            if ((ca_env && ca != ca_env) || (cb_env && cb != cb_env)) continue;
            for (int nt = (split_k && !nt_search) ? nt_hi : 16; nt <= nt_hi; nt += 16) {
                if (nt % (8 * ca)) continue;
                if (nt_env && !split_k && nt != std::min(nt_env, nt_hi)) continue;
                if (snt_env > 0 && split_k && nt_search && nt != std::min(snt_env, nt_hi)) continue;
                const int n_bchunk = (int)ceil_div(Nb, nt);
                if (cb > 1 && cb > n_bchunk) continue;
                const int64_t tiles = ceil_div(p->n_ablk, ca) * ceil_div(n_bchunk, cb);
                int n_ksplit = 1, kbps = p->n_kblk;
                if (split_k) {
                    int want = (int)std::max<int64_t>(1, maxc / tiles);
                    want = std::min(want, p->n_kblk);
                    if (max_ksplit > 0) want = std::min(want, max_ksplit);
                    kbps = (int)ceil_div(p->n_kblk, want);
                    n_ksplit = (int)ceil_div(p->n_kblk, kbps);
                }
                const int64_t units = tiles * n_ksplit;
                const int64_t waves = ceil_div(units, maxc);
                const double per_kb = std::max(4.0 * std::max(48.0, nt / 2.0), (128.0 + nt) * 128.0 / 71.0) + 30.0;
                double cost = (double)waves * (per_kb * kbps + 58.0 * nt + 6000.0) + (csz > 1 ? 500.0 : 0.0);
                if (force_cluster == 2 && csz > 1) cost *= 0.25;   // testing / experiments: prefer clusters
                if (cost < best) {
                    best = cost;
                    p->ca = ca; p->cb = cb; p->nt = nt; p->n_bchunk = n_bchunk; p->n_ksplit = n_ksplit;
                    p->kb_per_split = kbps;
                }
            }
        }
    }
    // CTA pairs (cta_group::2): two a-blocks per unit, half of the b-chunk per SM
    p->pair = 0;
    // AVI_TC_PAIR: 0 never, 1 (default) when the cost model prefers it (large problems: C4 +17 %), 2 always
    const int pair_env = env_int("AVI_TC_PAIR", 1);
    if (allow_pair && pair_env && p->n_ablk >= 2) {
        const int maxp = max_active_clusters(ctx, 2);
        const int n_ag = (p->n_ablk + 1) / 2;
        for (int nt = split_k ? nt_hi : 16; nt <= nt_hi && maxp > 0; nt += 16) {
            const int n_bchunk = (int)ceil_div(Nb, nt);
            const int64_t tiles = (int64_t)n_ag * n_bchunk;
            int n_ksplit = 1, kbps = p->n_kblk;
            if (split_k) {
                int want = (int)std::max<int64_t>(1, maxp / tiles);
                want = std::min(want, p->n_kblk);
                if (max_ksplit > 0) want = std::min(want, max_ksplit);
                kbps = (int)ceil_div(p->n_kblk, want);
                n_ksplit = (int)ceil_div(p->n_kblk, kbps);
            }
            const int64_t waves = ceil_div(tiles * n_ksplit, maxp);
            // per SM: 128 rows of A and nt / 2 rows of B; measured ~25 % above the ingest model (pair handshakes)
            const double per_kb = 1.25 * std::max(4.0 * std::max(48.0, nt / 2.0), (128.0 + nt / 2.0) * 128.0 / 71.0) + 30.0;
            double cost = (double)waves * (per_kb * kbps + 58.0 * nt + 8000.0);
            if (pair_env == 2) cost *= 0.25;
            if (cost < best) {
                best = cost;
                p->pair = 1; p->ca = 1; p->cb = 1; p->nt = nt; p->n_bchunk = n_bchunk; p->n_ksplit = n_ksplit;
                p->kb_per_split = kbps;
            }
        }
    }
    if (best >= 1e300) AVI_FAIL(ctx, AVI_ERR_INVALID, "no valid tiling");
    if (getenv("AVI_TC_DEBUG"))
        fprintf(stderr, "[avi_tc_plan] Ma=%lld Nb=%lld K=%lld split_k=%d -> pair=%d ca=%d cb=%d nt=%d n_ablk=%d n_bchunk=%d n_ksplit=%d kbps=%d "
                "maxc(1,2,4,8)=%d,%d,%d,%d cost=%.0f\n", (long long)Ma, (long long)Nb, (long long)K, (int)split_k, p->pair, p->ca, p->cb, p->nt,
                p->n_ablk, p->n_bchunk, p->n_ksplit, p->kb_per_split, max_active_clusters(ctx, 1), max_active_clusters(ctx, 2),
                max_active_clusters(ctx, 4), max_active_clusters(ctx, 8), best);
    (void)sms;
    return AVI_OK;
}

int32_t avi_tc_launch(avi_ctx* ctx, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcParams& p_in) {
    TcParams p = p_in;
    const int dbg_env = env_int("AVI_TC_DBG", 0);
    p.dbg = dbg_env;
    const int prof_env = env_int("AVI_TC_PROF", 0);
    static unsigned long long* prof_buf = nullptr;
    static int prof_calls = 0;
    if (prof_env) {
        if (!prof_buf) { cudaMalloc(&prof_buf, 148 * 8 * sizeof(unsigned long long)); }
        cudaMemsetAsync(prof_buf, 0, 148 * 8 * sizeof(unsigned long long), ctx->stream);
        p.prof = prof_buf;
    }
    if (p.nt % 16 || p.nt < 16 || p.nt > 256 || p.ca < 1 || p.cb < 1 || p.ca * p.cb > 8 || p.nt % (8 * p.ca))
        AVI_FAIL(ctx, AVI_ERR_INVALID, "bad tiling");
    const int csz = p.pair ? 2 : p.ca * p.cb;
    const int64_t units = p.pair ? (int64_t)((p.n_ablk + 1) / 2) * p.n_bchunk * p.n_ksplit
                                 : ceil_div(p.n_ablk, p.ca) * ceil_div(p.n_bchunk, p.cb) * p.n_ksplit;
    if (units <= 0) return AVI_OK;
    AVI_CHECK(set_attrs(ctx));
    const int smem = smem_bytes_for(p.nt, &p.stages, p.pair != 0, epi == EPI_GLM_FWD);   // as many stages as fit
    if (smem > SMEM_LIMIT) AVI_FAIL(ctx, AVI_ERR_INVALID, "tile does not fit in shared memory");
    const int maxc = max_active_clusters(ctx, csz);
    if (maxc <= 0) AVI_FAIL(ctx, AVI_ERR_CUDA, "cluster size not launchable");
    const unsigned grid = (unsigned)(std::min<int64_t>(units, maxc) * csz);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (csz > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = csz; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    if (avi_pdl_enabled()) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    AviTimed timed(ctx, epi == EPI_GLM_FWD ? "glm_fwd" : epi == EPI_GLM_BWD ? "glm_bwd" : "gemm_store");
    cudaError_t e;
    if (p.pair) {
        if (epi == EPI_GLM_FWD && p.likelihood == AVI_GLM_BERNOULLI_LOGIT) e = cudaLaunchKernelEx(&cfg, k_gemm_tc_pair<EPI_GLM_FWD, 0>, tmA, tmB, p);
        else if (epi == EPI_GLM_FWD) e = cudaLaunchKernelEx(&cfg, k_gemm_tc_pair<EPI_GLM_FWD, 1>, tmA, tmB, p);
        else if (epi == EPI_GLM_BWD) e = cudaLaunchKernelEx(&cfg, k_gemm_tc_pair<EPI_GLM_BWD, 0>, tmA, tmB, p);
        else e = cudaLaunchKernelEx(&cfg, k_gemm_tc_pair<EPI_STORE, 0>, tmA, tmB, p);
    } else
    if (epi == EPI_GLM_FWD && p.likelihood == AVI_GLM_BERNOULLI_LOGIT) e = cudaLaunchKernelEx(&cfg, k_gemm_tc<EPI_GLM_FWD, 0>, tmA, tmB, p);
    else if (epi == EPI_GLM_FWD) e = cudaLaunchKernelEx(&cfg, k_gemm_tc<EPI_GLM_FWD, 1>, tmA, tmB, p);
    else if (epi == EPI_GLM_BWD) e = cudaLaunchKernelEx(&cfg, k_gemm_tc<EPI_GLM_BWD, 0>, tmA, tmB, p);
    else e = cudaLaunchKernelEx(&cfg, k_gemm_tc<EPI_STORE, 0>, tmA, tmB, p);
    if (e != cudaSuccess) AVI_FAIL(ctx, AVI_ERR_CUDA, std::string("tcgen05 kernel launch: ") + cudaGetErrorString(e));
    AVI_LAUNCHED(ctx);
    if (p.prof && ++prof_calls > prof_env && !ctx->capturing) {   // AVI_TC_PROF=n: report launches after the n-th
        std::vector<unsigned long long> h(148 * 8);
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h.data(), prof_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull, t6 = 0;
        for (unsigned b = 0; b < grid; ++b) { if (h[b * 8]) t0 = std::min(t0, h[b * 8]); t6 = std::max(t6, h[b * 8 + 6]); }
        double avg[8] = {0};
        int cnt[8] = {0};
        for (unsigned b = 0; b < grid; ++b)
            for (int k = 0; k < 7; ++k)
                if (h[b * 8 + k]) { avg[k] += (double)(h[b * 8 + k] - t0); cnt[k]++; }
        for (int k = 0; k < 7; ++k) avg[k] /= std::max(cnt[k], 1);
        double wait_c = 0, loop_c = 0;
        for (unsigned b = 0; b < grid; ++b) { wait_c += (double)(h[b * 8 + 7] >> 32); loop_c += (double)(h[b * 8 + 7] & 0xffffffffull); }
        fprintf(stderr, "[tc_prof] epi=%d grid=%u nt=%d total=%.2fus | avg ns since first CTA entry: entry %.0f prologue %.0f first-operands %.0f "
                "mma-issued %.0f acc-ready %.0f epi-done %.0f exit %.0f | issuer: mainloop %.0f cyc, of which stalled on operands %.0f cyc "
                "(after the first k-block)\n", epi, grid, p.nt, (t6 - t0) * 1e-3, avg[0], avg[1], avg[2], avg[3],
                avg[4], avg[5], avg[6], loop_c / grid, wait_c / grid);
    }
    return AVI_OK;
}

// C[ks][b * ldc + a] = sum_k A[a, k] B[b, k] over k-split ks (EPI_STORE).  split_k = false: one slab (C itself).
// Returns the number of slabs written; slab s starts at C + s * slab_stride.
int32_t avi_tc_gemm_store(avi_ctx* ctx, const float* A, int64_t Ma, int64_t lda, const float* B, int64_t Nb, int64_t ldb,
                          int64_t K, float* C, int64_t ldc, int64_t slab_stride, int max_slabs, int* n_slabs) {
    TcParams p{};
    AVI_CHECK(avi_tc_plan(ctx, Ma, Nb, K, max_slabs > 1, 0, &p, 1, /*the caller's slab buffer bounds the split*/ max_slabs,
                          /*nt_search=*/1));
    if (p.n_ksplit > max_slabs) AVI_FAIL(ctx, AVI_ERR_INVALID, "split-K plan exceeds the slab buffer");
    CUtensorMap tmA, tmB;
    AVI_CHECK(avi_tc_make_tmap(ctx, &tmA, A, Ma, K, lda, 128 / p.cb));
    AVI_CHECK(avi_tc_make_tmap(ctx, &tmB, B, Nb, K, ldb, p.pair ? p.nt / 2 : p.nt / p.ca));
    p.C = C; p.ldc = (int)ldc; p.slab_stride = slab_stride;
    AVI_CHECK(avi_tc_launch(ctx, EPI_STORE, tmA, tmB, p));
    if (n_slabs) *n_slabs = p.n_ksplit;
    return AVI_OK;
}
