// The fused whole-iteration kernel (interface and phase list: step_fused.cuh).
//
// One persistent CTA per SM, 640 threads, the warp roles of gemm_tc.cu inside the two contraction phases
// (warp 0 TMA producer, warp 1 tcgen05.mma issuer, warp 2 TMEM allocator, warps 4-19 epilogue); every warp works in the
// sample and tail phases.  Phases are separated by grid barriers (monotonic 64-bit counters in global memory; all CTAs
// are co-resident by construction: grid <= #SMs and the shared-memory footprint admits one CTA per SM).
//
// Memory-model notes (each one is load-bearing):
//  * data produced by generic stores in one phase and consumed by TMA (async proxy) in the next -- Zt after the sample
//    phase, R after the forward phase -- is published with fence.proxy.async.global by every writer before the barrier
//    and by the consuming producer thread after it, on top of the barrier's release / acquire at gpu scope;
//  * data produced in-kernel and consumed by ordinary loads (eps, z, the per-sample prior terms, the partial
//    log-likelihood sums) is read with ld.global.cg (`COH` in gemm_tc_dev.cuh): never through the read-only path;
//  * everything a later phase OVERWRITES at the end of the iteration (step counter, optimiser scalars, exchange
//    sequence number, lambda for log det) is read by every CTA at kernel entry, i.e. before the first barrier, so the
//    committing CTA cannot race with a CTA that is still reading;
//  * the mbarrier ring is invalidated and re-initialised between the contraction phases (their stage geometry differs),
//    after a CTA-wide barrier that follows the last accumulator hand-over of the phase.
#include <cstdio>
#include <cstdlib>

#include "step_fused.cuh"
#include "gemm_tc_dev.cuh"
#include "glm_prior.cuh"
#include "mf_finalize.cuh"

namespace {

constexpr int NW = NUM_THREADS / 32;            // 20 warps
constexpr int TAIL_MAX_PER_CTA = 120;           // coordinates per CTA in the tail (SmemCtl::tail: 2 floats each behind the staging area)
constexpr int TAIL_STAGE = 1024;                // floats of the tail's slab staging area

__device__ __forceinline__ unsigned long long ld_acquire_gpu_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// All threads of all CTAs.  Arrival: every thread publishes its generic global writes to the async proxy, the CTA
// barrier orders them before thread 0, whose gpu-scope release-add makes them visible (cumulativity) to whoever
// acquires the counter.  gbar[0] is a counter that only grows; gbar[1] holds its value at the start of the launch
// (written by CTA 0 at the very end of the previous launch, read by every CTA at entry): the k-th barrier of a launch
// of n CTAs completes at base + k * n, whatever the grid sizes of earlier launches were.
struct GridBar {
    unsigned long long* ctr;
    unsigned long long base;
    int k;
};
__device__ __forceinline__ void grid_barrier(GridBar& gb) {
    fence_proxy_async_global();
    __syncthreads();
    gb.k += 1;
    if (threadIdx.x == 0) {
        // release-add (no return value, no separate membar) / acquire-poll: the CTA barrier above makes the release
        // cumulative over the other threads' writes, the one below extends the acquire to them
        const unsigned long long target = gb.base + (unsigned long long)gb.k * gridDim.x;
        asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(gb.ctr), "l"(1ull) : "memory");
#ifdef AVI_WATCHDOG
        const unsigned long long w0 = gtime_ns();
        for (unsigned it = 0; ld_acquire_gpu_u64(gb.ctr) < target; ++it) {
            if ((it & 255u) == 255u && *(volatile unsigned int*)&tc::avi_hang_report[0]) break;
            if (gtime_ns() - w0 > 300000000ull) {
                if (atomicCAS(&tc::avi_hang_report[0], 0u, 2u) == 0u) { tc::avi_hang_report[1] = 0xBA771E5u; tc::avi_hang_report[2] = (unsigned)gb.k; tc::avi_hang_report[3] = blockIdx.x; }
                break;
            }
        }
#else
        while (ld_acquire_gpu_u64(gb.ctr) < target) { }
#endif
        fence_proxy_async_global();
    }
    __syncthreads();
}

__device__ __forceinline__ void stamp_min(unsigned long long* tl, int slot) {
    if (tl && threadIdx.x == 0) atomicMin(tl + slot, gtime_ns());
}
__device__ __forceinline__ void stamp_max(unsigned long long* tl, int slot) {
    if (tl && threadIdx.x == 0) atomicMax(tl + slot, gtime_ns());
}

// AVI_STEP_PROF: per-CTA %globaltimer stamps (StepParams.prof[blockIdx.x][32]); slot list in scripts/step_prof.py
#define PSTAMP(prof, slot) do { if (prof) (prof)[(size_t)blockIdx.x * 32 + (slot)] = gtime_ns(); } while (0)

// pipeline barriers for a phase with `stages` ring slots: one WARP, one barrier per lane (the previous phase's barriers are
// invalidated first).  which: 1 = the operand ring (full / empty), 2 = the accumulator barriers, 3 = both.
__device__ __forceinline__ void init_pipeline(SmemCtl* ctl, int stages, int stages_prev, int which = 3) {
    static_assert(2 * MAX_STAGES + 4 <= 32, "one lane per barrier");
    const int lane = threadIdx.x & 31;
    uint64_t* bar = nullptr;
    uint32_t count = 1;
    bool had = false, want = false;
    if (lane < MAX_STAGES) { if (which & 1) { bar = &ctl->full[lane]; had = lane < stages_prev; want = lane < stages; } }
    else if (lane < 2 * MAX_STAGES) { const int s = lane - MAX_STAGES; if (which & 1) { bar = &ctl->empty[s]; had = s < stages_prev; want = s < stages; } }
    else if (lane < 2 * MAX_STAGES + 2) { if (which & 2) { bar = &ctl->tmem_full[lane - 2 * MAX_STAGES]; had = stages_prev > 0; want = true; } }
    else if (lane < 2 * MAX_STAGES + 4) { if (which & 2) { bar = &ctl->tmem_empty[lane - 2 * MAX_STAGES - 2]; had = stages_prev > 0; want = true; count = EPI_WARPS; } }
    if (bar) {
        if (had) mbar_inval(bar);
        if (want) tc::mbar_init(bar, count);
    }
    tc::mbar_fence_init();
    __syncwarp();
}
// the operand ring alone, by ONE thread (the producer, while its CTA's epilogue is still running)
__device__ __forceinline__ void init_ring_thread(SmemCtl* ctl, int stages, int stages_prev) {
    for (int s = 0; s < stages_prev; ++s) { mbar_inval(&ctl->full[s]); mbar_inval(&ctl->empty[s]); }
    for (int s = 0; s < stages; ++s) { tc::mbar_init(&ctl->full[s], 1); tc::mbar_init(&ctl->empty[s], 1); }
    tc::mbar_fence_init();
}

#define FUNIT_COORDS(u) \
    const int ab = (u) % p.n_ablk, bc = ((u) / p.n_ablk) % p.n_bchunk, ks = (u) / (p.n_ablk * p.n_bchunk)

// A tile (128 a-rows x 32 k) of k-block kb into `sa`.  K-major: one box.  MN-major (TcParams.a_mn): four boxes of
// {32 a-elements, 32 k-rows}, 4 KB apart (the LBO of smem_desc_mn_sw128).
__device__ __forceinline__ void load_a_tile(uint8_t* sa, const CUtensorMap* tmA, uint64_t* bar, const TcParams& p, int kb, int ab) {
    if (!p.a_mn) { tc::tma_load_2d(sa, tmA, bar, kb * BK, ab * BM); return; }
    const int seg = p.a_seg_kb ? kb / p.a_seg_kb : 0, rb = p.a_seg_kb ? kb - seg * p.a_seg_kb : kb;
    const int a0 = ab * BM + (seg == 2 ? p.a_seg_off : 0);
    if (p.a_mn == 2) { tc::tma_load_3d(sa, tmA, bar, 0, rb * BK, a0 >> 5); return; }
#pragma unroll
    for (int fb = 0; fb < BM / 32; ++fb) tc::tma_load_2d(sa + fb * 4096, tmA, bar, a0 + fb * 32, rb * BK);
}

// B tile (NT b-rows x 32 k) of k-block kb into `sb`.  K-major: one box.  MN-major (TcParams.b_mn): ceil(NT / 32) boxes of
// {32 b-elements, 32 k-rows}, 4 KB apart.
__device__ __forceinline__ void load_b_tile(uint8_t* sb, const CUtensorMap* tmB, uint64_t* bar, const TcParams& p, int kb, int bc) {
    if (!p.b_mn) { tc::tma_load_2d(sb, tmB, bar, kb * BK, bc * p.nt); return; }
    if (p.b_mn == 2) { tc::tma_load_3d(sb, tmB, bar, 0, kb * BK, (bc * p.nt) >> 5); return; }
    const int nb = (p.nt + 31) >> 5;
    for (int j = 0; j < nb; ++j) tc::tma_load_2d(sb + j * 4096, tmB, bar, bc * p.nt + j * 32, kb * BK);
}
__device__ __forceinline__ int b_tile_bytes(const TcParams& p) { return p.b_mn ? ((p.nt + 31) >> 5) * 4096 : p.nt * BK * 4; }

// The operand that does not depend on the previous phase (X; early_op 1 = A, 2 = B) of this CTA's first unit is
// requested before the phase's dependency is satisfied: expect_tx without arrive; the producer adds the dependent
// operand (and the arrive) later.  One thread.  Returns the number of ring slots armed.
__device__ __forceinline__ int preissue_early(const CUtensorMap* tmA, const CUtensorMap* tmB, const TcParams& p,
                                              SmemCtl* ctl, uint8_t* tiles, int stages, int early_op) {
    const int units = p.n_ablk * p.n_bchunk * p.n_ksplit;
    if ((int)blockIdx.x >= units) return 0;
    const int b_bytes = b_tile_bytes(p), stage_bytes = A_TILE_BYTES + b_bytes;
    const int u = blockIdx.x;
    FUNIT_COORDS(u);
    const int kb0 = ks * p.kb_per_split, kb1 = min(p.n_kblk, kb0 + p.kb_per_split);
    int n = 0;
    for (int kb = kb0; kb < kb1 && n < stages; ++kb, ++n) {
        uint8_t* sa = tiles + n * stage_bytes;
        if (early_op == 1) {
            tc::mbar_expect_tx(&ctl->full[n], (uint32_t)A_TILE_BYTES);
            load_a_tile(sa, tmA, &ctl->full[n], p, kb, ab);
        } else {
            tc::mbar_expect_tx(&ctl->full[n], (uint32_t)b_bytes);
            load_b_tile(sa + A_TILE_BYTES, tmB, &ctl->full[n], p, kb, bc);
        }
    }
    return n;
}

// One contraction phase: the single-CTA pipeline of k_gemm_tc (gemm_tc.cu) over the units u = blockIdx.x,
// blockIdx.x + gridDim.x, ...; the ring and the accumulator stages start from their initial state.
// nextA / nextB / nextp != nullptr (forward phase): as soon as this phase's last MMA has COMPLETED -- the ring is then
// free, while the epilogue warps still work on the accumulator for microseconds -- the producer thread re-carves the ring
// for the next phase's geometry and requests that phase's static operand.  Returns (warp 0 only) the number of slots
// armed, -1 if nothing was handed over.  Doing this after the epilogue cost ~1.8 us of once-executed, instruction-fetch-
// bound code on the critical path between the phases.
template <int EPI, int LIK, int X3>
__device__ __forceinline__ int tc_phase(const CUtensorMap* tmA, const CUtensorMap* tmB, const TcParams& p, SmemCtl* ctl,
                                        uint8_t* tiles, uint32_t tmem_base, int stages, int pre_issued, int early_op,
                                        unsigned long long* prof, int ps, const CUtensorMap* nextA = nullptr,
                                        const CUtensorMap* nextB = nullptr, const TcParams* nextp = nullptr,
                                        int next_stages = 0, int next_early_op = 0) {
    int handed = -1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NT = p.nt;
    const int b_bytes = b_tile_bytes(p), stage_bytes = A_TILE_BYTES + b_bytes;
    const int units = p.n_ablk * p.n_bchunk * p.n_ksplit;
    const int first = blockIdx.x, stride = gridDim.x;
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (tc::elect_one()) {
            fence_proxy_async_global();   // operands written by generic stores of the previous phase (other CTAs)
            int stage = 0; uint32_t phase = 0;
            int kcount = 0;
            for (int u = first; u < units; u += stride) {
                FUNIT_COORDS(u);
                const int kb0 = ks * p.kb_per_split, kb1 = min(p.n_kblk, kb0 + p.kb_per_split);
                for (int kb = kb0; kb < kb1; ++kb, ++kcount) {
                    tc::mbar_wait(&ctl->empty[stage], phase ^ 1);
                    uint8_t* sa = tiles + stage * stage_bytes;
                    if (kcount < pre_issued) {
                        if (early_op == 1) {
                            tc::mbar_arrive_expect_tx(&ctl->full[stage], (uint32_t)b_bytes);
                            load_b_tile(sa + A_TILE_BYTES, tmB, &ctl->full[stage], p, kb, bc);
                        } else {
                            tc::mbar_arrive_expect_tx(&ctl->full[stage], (uint32_t)A_TILE_BYTES);
                            load_a_tile(sa, tmA, &ctl->full[stage], p, kb, ab);
                        }
                    } else {
                        tc::mbar_arrive_expect_tx(&ctl->full[stage], (uint32_t)stage_bytes);
                        load_a_tile(sa, tmA, &ctl->full[stage], p, kb, ab);
                        load_b_tile(sa + A_TILE_BYTES, tmB, &ctl->full[stage], p, kb, bc);
                    }
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
            PSTAMP(prof, ps + 0);   // last operand request issued
            if (nextp) {
                // drain: every slot's last release has landed (so no commit of this phase can hit a re-initialised
                // barrier) and the last unit's accumulator is complete (every MMA has finished reading the ring)
                const int per_cta = first < units ? (units - first + stride - 1) / stride : 0;
                if (per_cta > 0) {
                    for (int i = 0; i < stages; ++i) {
                        // what the producer would wait for before refilling slot `stage`: the release of its last use
                        // (a slot that was never filled passes at once)
                        tc::mbar_wait(&ctl->empty[stage], phase ^ 1);
                        if (++stage == stages) { stage = 0; phase ^= 1; }
                    }
                    const int last = per_cta - 1;
                    tc::mbar_wait(&ctl->tmem_full[last & 1], (uint32_t)((last >> 1) & 1));
                }
                init_ring_thread(ctl, next_stages, stages);
                handed = preissue_early(nextA, nextB, *nextp, ctl, tiles, next_stages, next_early_op);
                PSTAMP(prof, 10);   // next phase's ring carved, static operand requested
            }
        }
        handed = __reduce_max_sync(0xffffffffu, handed);   // (from whichever lane was elected; the others hold -1)
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (tc::elect_one()) {
            const uint32_t idesc = tc::idesc_tf32(BM, NT, p.a_mn, p.b_mn);
            const uint64_t a_step = p.a_mn ? 64u : 2u;   // per K = 8 MMA, in 16-byte units: 8 k-rows of 128 B | 32 B along the row
            const uint64_t b_step = p.b_mn ? 64u : 2u;
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int u = first; u < units; u += stride) {
                const int ks = u / (p.n_ablk * p.n_bchunk);
                const int kb0 = ks * p.kb_per_split, kb1 = min(p.n_kblk, kb0 + p.kb_per_split);
                tc::mbar_wait(&ctl->tmem_empty[as], aphase ^ 1);
                tc::fence_after_sync();
                const uint32_t tacc = tmem_base + (uint32_t)(as * 256);
                for (int kb = kb0; kb < kb1; ++kb) {
                    tc::mbar_wait(&ctl->full[stage], phase);
                    tc::fence_after_sync();
                    if (kb == kb0 && u == first) PSTAMP(prof, ps + 1);   // first operands landed
                    const uint32_t sa = tc::smem_u32(tiles + stage * stage_bytes);
                    const uint64_t da = p.a_mn ? tc::smem_desc_mn_sw128(sa, 4096u) : tc::smem_desc_k_sw128(sa);
                    const uint64_t db = p.b_mn ? tc::smem_desc_mn_sw128(sa + A_TILE_BYTES, 4096u) : tc::smem_desc_k_sw128(sa + A_TILE_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k)
                        tc::umma_tf32(tacc, da + a_step * (uint64_t)k, db + b_step * (uint64_t)k, idesc,
                                      (kb > kb0 || k > 0) ? 1u : 0u);
                    tc::umma_commit(&ctl->empty[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                tc::umma_commit(&ctl->tmem_full[as]);
                PSTAMP(prof, ps + 2);   // last MMA of the (last) unit issued
                if (++as == 2) { as = 0; aphase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int ew = warp - 4, quarter = warp & 3, cq = ew >> 2;
        const int et = threadIdx.x - 128;
        const int ngroups = NT / 8;
        const int c_begin = 8 * ((ngroups * cq) / 4), c_end = 8 * ((ngroups * (cq + 1)) / 4);
        int as = 0; uint32_t aphase = 0;
        for (int u = first; u < units; u += stride) {
            FUNIT_COORDS(u);
            const int a = ab * BM + quarter * 32 + lane;
            const bool a_ok = a < p.Ma;
            const uint32_t tacc = tmem_base + (uint32_t)(as * 256) + ((uint32_t)(quarter * 32) << 16);
            float s1, s2;
            epilogue_unit<EPI, LIK, true, 2, X3>(p, ctl, as, aphase, tacc, NT, a, a_ok, bc, ks, c_begin, c_end, et, s1, s2);
            if (et == 0) PSTAMP(prof, ps + 3);   // accumulator consumed (epilogue math of the unit done)
            tc::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive(&ctl->tmem_empty[as]);
            if (EPI == EPI_GLM_FWD) {
                if (p.post_on == 2) epilogue_fwd_unit_total(p, ctl, as, u, et, a_ok ? s1 : 0.0f);
                else epilogue_store_partials<EPI>(p, a_ok, a, bc, ks, cq, s1, s2);
            }
            if (EPI == EPI_GLM_BWD) epilogue_bwd_store(p, ctl, a_ok, a, bc, ks, cq, et, s1, s2);
            if (EPI == EPI_GLM_BWD && et == 0) PSTAMP(prof, ps + 4);   // slab rows stored
            if (EPI == EPI_GLM_BWD && p.post_on == 1) epilogue_bwd_combine(p, ctl, ab, et);
            if (et == 0) PSTAMP(prof, ps + 5);   // unit complete
            if (++as == 2) { as = 0; aphase ^= 1; }
        }
    }
    return handed;
}

// ---------------------------------------------------------------------------------------------------------
// Sample phase: z = mu + s .* eps for this rank's samples, eps from Philox (global sample index), plus what the
// sampling kernel's GLM hook produces (family.cu: k_sample<.., HOOK>): the TF32-rounded beta block Zt (A operand of the
// forward contraction), |eps_m|^2 and the per-sample prior terms.  Samples are dealt to CTAs in contiguous groups of
// S = ceil(Mloc / grid); inside a CTA W = 20 / S warps share one sample (fixed-order combine through shared memory).
__device__ __noinline__ void sample_phase(const StepParams& sp, SmemCtl* ctl, unsigned long long step,
                                             unsigned long long key) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = (sp.Mloc + (int)gridDim.x - 1) / (int)gridDim.x;
    const int W = S <= NW ? NW / S : 1;
    const int slots = NW / W;
    const PhiloxKeys pk(key);
    const uint32_t c2 = (uint32_t)step, c3 = eps_ctr3(step, (uint32_t)AVI_STREAM_EPS);
    const int D = sp.D, ld = sp.ld, nq = ld / 4;
    const float* mu = sp.lambda;
    const float* sc = sp.lambda + D;   // only 4-byte aligned in general
    float* red = &ctl->ys[0][0];       // [warp][3]
    const int slot = warp / W, wi = warp - slot * W;
    for (int j0 = 0; j0 < S; j0 += slots) {
        const int j = j0 + slot;
        const int m = (int)blockIdx.x * S + j;
        const bool active = slot < slots && j < S && m < sp.Mloc;
        float part = 0.f, bsq = 0.f, eta = 0.f;
        if (active) {
            float* Erow = sp.E + (size_t)m * ld;
            float* Zrow = sp.Z + (size_t)m * ld;
            for (int q = wi * 32 + lane; q < nq; q += 32 * W) {
                const float4 e = normal4((uint32_t)q, (uint32_t)(sp.m0 + m), c2, c3, pk);
                const int i = 4 * q;
                float ev[4] = {e.x, e.y, e.z, e.w}, zv[4], zt[4], mv[4] = {0.f, 0.f, 0.f, 0.f}, sv[4] = {0.f, 0.f, 0.f, 0.f};
                if (i + 3 >= D) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) ev[c] = i + c < D ? ev[c] : 0.0f;
                }
                if (i + 3 < D) {
                    const float4 m4 = *reinterpret_cast<const float4*>(mu + i);
                    mv[0] = m4.x; mv[1] = m4.y; mv[2] = m4.z; mv[3] = m4.w;
#pragma unroll
                    for (int c = 0; c < 4; ++c) sv[c] = sc[i + c];
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (i + c < D) { mv[c] = mu[i + c]; sv[c] = sc[i + c]; }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    zv[c] = fmaf(sv[c], ev[c], mv[c]);
                    part = fmaf(ev[c], ev[c], part);
                    const bool is_beta = i + c < sp.d;
                    bsq = is_beta ? fmaf(zv[c], zv[c], bsq) : bsq;
                    zt[c] = is_beta ? zv[c] : 0.0f;
                    if (i + c == sp.d) eta = zv[c];
                }
                *reinterpret_cast<float4*>(Erow + i) = make_float4(ev[0], ev[1], ev[2], ev[3]);
                *reinterpret_cast<float4*>(Zrow + i) = make_float4(zv[0], zv[1], zv[2], zv[3]);
                float hi[4], lo[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) { hi[c] = tc::round_tf32(zt[c]); lo[c] = tc::round_tf32(zt[c] - hi[c]); }
                float* row = sp.Zt + (size_t)m * sp.zt_ld + i;
                if (sp.zt_seg == 0) {
                    if (i < sp.zt_ld) *reinterpret_cast<float4*>(row) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                } else if (i < sp.zt_seg) {   // 3xTF32: [hi | hi | lo]
                    *reinterpret_cast<float4*>(row) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(row + sp.zt_seg) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(row + 2 * sp.zt_seg) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
        part = warp_sum(part); bsq = warp_sum(bsq); eta = warp_sum(eta);   // eta is non-zero in exactly one lane
        if (W > 1) {
            if (lane == 0) { red[warp * 3] = part; red[warp * 3 + 1] = bsq; red[warp * 3 + 2] = eta; }
            __syncthreads();
            if (wi == 0 && lane == 0 && slot < slots) {
                part = 0.f; bsq = 0.f; eta = 0.f;
                for (int w2 = 0; w2 < W; ++w2) {
                    part += red[(slot * W + w2) * 3]; bsq += red[(slot * W + w2) * 3 + 1]; eta += red[(slot * W + w2) * 3 + 2];
                }
            }
            __syncthreads();
        }
        if (active && wi == 0 && lane == 0) {
            sp.esq[m] = part;
            reinterpret_cast<float4*>(sp.pre)[m] = glm_prior_terms(bsq, eta, sp.d, sp.variant, sp.include_prior);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// State read by every CTA at kernel entry (see the memory-model notes at the top).
static_assert(sizeof(ObjDeviceState) == 64 && SC_ETA == 5, "the snapshot warp reads ObjDeviceState as 8 words and sc[0..5]");
struct Snap {
    unsigned long long step, key;
    long long cursor;
    int halted, tp;
    int zt_ok, zt_nparts;      // the sample buffers already hold this iteration's draws (drawn ahead by the previous tail)
    float b1t, b2t, t_avg, v_old, r_old, shift, logdet;
    unsigned int seq;
};

// Coordinate slice of a CTA in the tail phase and in the slice-major sampler: `per` coordinates, a multiple of 4 so
// that a slice is a whole number of Philox quads / 16-byte columns of the sample-major buffers.
__device__ __forceinline__ int slice_per(int D) {
    return (((D + (int)gridDim.x - 1) / (int)gridDim.x) + 3) & ~3;
}

// Slice-major sampler: this CTA draws, for ALL local samples, the coordinates of its slice for iteration `step` with the
// slice's (mu, s) taken from shared memory (lam2[2 j], lam2[2 j + 1]), writes z, eps and the TF32 copy of z, and leaves
// per-sample partial sums over the slice (|eps|^2, |beta|^2; the slice holding eta also stores eta_m).  Used by the
// tail phase to draw the NEXT iteration's samples under the lambda it has just computed, and at kernel entry when
// nothing was drawn ahead (same code, same bits either way).
__device__ __forceinline__ void draw_slice(const StepParams& sp, const float* lam2, unsigned long long step, unsigned long long key) {
    const int D = sp.D, ld = sp.ld;
    const int per = slice_per(D);
    const int c0 = (int)blockIdx.x * per;
    if (c0 >= D) return;
    const int nq = (min(ld, c0 + per) - c0) / 4;
    const PhiloxKeys pk(key);
    const uint32_t c2 = (uint32_t)step, c3 = eps_ctr3(step, (uint32_t)AVI_STREAM_EPS);
    for (int m = threadIdx.x; m < sp.Mloc; m += NUM_THREADS) {
        float e2 = 0.f, b2 = 0.f, eta = 0.f;
        for (int qq = 0; qq < nq; ++qq) {
            const int i = c0 + 4 * qq;
            const float4 e = normal4((uint32_t)(i >> 2), (uint32_t)(sp.m0 + m), c2, c3, pk);
            float ev[4] = {e.x, e.y, e.z, e.w}, zv[4], zt[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const bool in = i + c < D;
                ev[c] = in ? ev[c] : 0.0f;
                zv[c] = in ? fmaf(lam2[2 * (4 * qq + c) + 1], ev[c], lam2[2 * (4 * qq + c)]) : 0.0f;
                e2 = fmaf(ev[c], ev[c], e2);
                const bool is_beta = i + c < sp.d;
                b2 = is_beta ? fmaf(zv[c], zv[c], b2) : b2;
                zt[c] = is_beta ? tc::round_tf32(zv[c]) : 0.0f;
                if (i + c == sp.d) eta = zv[c];
            }
            *reinterpret_cast<float4*>(sp.E + (size_t)m * ld + i) = make_float4(ev[0], ev[1], ev[2], ev[3]);
            *reinterpret_cast<float4*>(sp.Z + (size_t)m * ld + i) = make_float4(zv[0], zv[1], zv[2], zv[3]);
            if (i < sp.zt_ld) *reinterpret_cast<float4*>(sp.Zt + (size_t)m * sp.zt_ld + i) = make_float4(zt[0], zt[1], zt[2], zt[3]);
        }
        sp.spart[(size_t)blockIdx.x * sp.Mloc + m] = e2;
        sp.spart[(size_t)(sp.spart_stride + blockIdx.x) * sp.Mloc + m] = b2;
        if (c0 <= sp.d && sp.d < c0 + per) sp.spart[(size_t)2 * sp.spart_stride * sp.Mloc + m] = eta;
    }
}

// Per-sample totals from the slice partials of draw_slice: |eps_m|^2 and the prior terms of the GLM (family.cu:
// k_sample<.., HOOK> produces the same pair in the sample-major sampler).  Warps 2 and 3, which have no role in the
// contraction phases, run this while the forward contraction proceeds; consumers come after the next grid barrier.
__device__ __forceinline__ void finalize_samples(const StepParams& sp, int nparts) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = (sp.Mloc + (int)gridDim.x - 1) / (int)gridDim.x;
    for (int j = warp - 2; j < S; j += 2) {
        const int m = (int)blockIdx.x * S + j;
        if (m >= sp.Mloc) break;
        float e2 = 0.f, b2 = 0.f;
        for (int c = lane; c < nparts; c += 32) {
            e2 += __ldcg(sp.spart + (size_t)c * sp.Mloc + m);
            b2 += __ldcg(sp.spart + (size_t)(sp.spart_stride + c) * sp.Mloc + m);
        }
        e2 = warp_sum(e2); b2 = warp_sum(b2);
        if (lane == 0) {
            sp.esq[m] = e2;
            reinterpret_cast<float4*>(sp.pre)[m] =
                glm_prior_terms(b2, __ldcg(sp.spart + (size_t)2 * sp.spart_stride * sp.Mloc + m), sp.d, sp.variant, sp.include_prior);
        }
    }
}

// The exchange of a multi-rank run inside the tail phase (low-latency push protocol of comm_dev.cuh): entry (class c,
// coordinate i) travels as element c * accv + i, the scalars as elements 4 * accv + {0, 1}; every rank sums the ranks'
// values in rank order => identical bits everywhere.  __noinline__ on purpose (and StepParams is a __grid_constant__
// kernel parameter, so taking its address costs nothing): code that a launch does not execute should not sit in the
// instruction stream of a kernel whose every instruction runs once.
struct Sums6 { float v0, v1, v2, v3, sl, sq; };
__device__ __noinline__ Sums6 tail_exchange(const StepParams& sp, SmemCtl* ctl, unsigned int seq, int i, bool mine, bool stl, Sums6 v) {
    const StepTail& t = sp.t;
    const int tid = threadIdx.x, accv = t.accv;
    float* sm = &ctl->ys[1][0];
    const long long sbase = 4ll * accv;
    const bool x01 = (t.xmask & STEP_X_V01) != 0, x23 = (t.xmask & STEP_X_V23) != 0 && stl;
    if (mine) {
        if (x01) { ll_push(t.comm, seq, (long long)i, v.v0); ll_push(t.comm, seq, (long long)accv + i, v.v1); }
        if (x23) { ll_push(t.comm, seq, 2ll * accv + i, v.v2); ll_push(t.comm, seq, 3ll * accv + i, v.v3); }
    }
    if (blockIdx.x == 0 && tid == NUM_THREADS - 1) {
        if (t.xmask & STEP_X_S0) ll_push(t.comm, seq, sbase, v.sl);
        if (t.xmask & STEP_X_S1) ll_push(t.comm, seq, sbase + 1, v.sq);
    }
    if (mine) {
        const long long idx[4] = {(long long)i, (long long)accv + i, 2ll * accv + i, 3ll * accv + i};
        const bool need[4] = {x01, x01, x23, x23};
        const float own[4] = {v.v0, v.v1, v.v2, v.v3};
        float out[4];
        ll_gather<4>(t.comm, seq, idx, need, own, out);
        if (x01) { v.v0 = out[0]; v.v1 = out[1]; }
        if (x23) { v.v2 = out[2]; v.v3 = out[3]; }
    }
    if (tid == NUM_THREADS - 1) {
        const long long idx[2] = {sbase, sbase + 1};
        const bool need[2] = {(t.xmask & STEP_X_S0) != 0, (t.xmask & STEP_X_S1) != 0};
        const float own[2] = {v.sl, v.sq};
        float out[2];
        ll_gather<2>(t.comm, seq, idx, need, own, out);
        sm[64] = need[0] ? out[0] : v.sl; sm[65] = need[1] ? out[1] : v.sq;
    }
    __syncthreads();
    v.sl = sm[64]; v.sq = sm[65];
    return v;
}

// The three scalars of the value slot, by the two warps without a role in the contraction phases, while the backward
// contraction runs (their inputs are final after barrier 1): scratch[1] = sum over the forward units of their
// log-likelihood totals, scratch[2] = sum_m log prior(z_m), scratch[3] = sum_m |eps_m|^2.  Lane-strided + shuffle tree:
// the same order in every CTA and on every rank.
__device__ __forceinline__ void tail_scalars(const StepParams& sp, SmemCtl* ctl) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 2) {
        float sl = 0.f;
        for (int u = lane; u < sp.t.n_units_f; u += 32) sl += __ldcg(sp.t.unit_ll + u);
        sl = warp_sum(sl);
        if (lane == 0) ctl->scratch[1] = sl;
    } else if (warp == 3) {
        float spr = 0.f, sq = 0.f;
        for (int m = lane; m < sp.Mloc; m += 32) { spr += __ldcg(sp.pre + 4 * (size_t)m); sq += __ldcg(sp.esq + m); }
        spr = warp_sum(spr); sq = warp_sum(sq);
        if (lane == 0) { ctl->scratch[2] = spr; ctl->scratch[3] = sq; }
    }
}

// Tail phase, all CTAs: CTA c owns coordinates [c * per, (c + 1) * per) of mu and of s.
//   local scalars (every CTA, same order => same bits) -> [exchange] -> value / ELBO / finiteness ->
//   gradient entries of the slice -> [DoG / DoWG: partial norms + one more grid barrier] -> update of the slice ->
//   commit (CTA 0) or, on the estimate_gradient! boundary, gradient to pinned host memory + last-CTA completion flag.
__device__ __forceinline__ void tail_phase(const StepParams& sp, SmemCtl* ctl, const Snap& sn, GridBar& gb) {
    const StepTail& t = sp.t;
    const UpdArgs a = t.a;
    const int D = sp.D, accv = t.accv, M = t.M, objective = t.objective, entropy = t.entropy;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* stage = &ctl->tail[0];      // [2][nc * nslab] slab partials of the slice (TAIL_STAGE floats)
    float* vals = &ctl->tail[TAIL_STAGE];          // [per][2]: sticking-the-landing sums of the slice
    float* sm = &ctl->ys[1][0];        // block_sum scratch (33) | [64..] broadcast slots
    const bool stl = entropy == AVI_ENT_STL || entropy == AVI_ENT_STL_ZEROGRAD;
    const bool adam = a.rule == AVI_RULE_ADAM, dog = a.rule == AVI_RULE_DOG || a.rule == AVI_RULE_DOWG;
    const bool polyavg = a.averager == AVI_AVG_POLYNOMIAL;
    const int per = slice_per(D);
    const int c0 = min(D, (int)blockIdx.x * per), c1 = min(D, c0 + per), nc = c1 - c0;
    float* lam2 = &ctl->tail[TAIL_STAGE + 2 * TAIL_MAX_PER_CTA];   // [per][2]: new (mu, s) of the slice for draw_slice
    const int NR = t.comm.nranks;

    // ---- every load this CTA needs is requested up front (one L2 round trip), the reductions follow
    //  scalars: sum_m log pi(z_m) = w * (sum of the forward units' log-likelihood totals) + sum_m log prior(z_m);
    //           sum_m |eps_m|^2
    //  (taken by warps 2 and 3 during the backward phase, tail_scalars: if every CTA fetched these few cache lines right
    //  after the barrier, 144 SMs would queue on the same L2 lines -- measured ~2 us in front of the tail)
    //  this thread's coordinate (j = tid < nc): the split-K slab partials of sum_m g and sum_m g*eps (the eta coordinate
    //  i == d comes complete from the backward epilogue), lambda and the optimiser state
    const bool mine = tid < nc;
    const int i = c0 + (mine ? tid : 0);
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
    float x0 = 0.f, x1 = 1.f, s1m0 = 0.f, s1m1 = 0.f, s2m0 = 0.f, s2m1 = 0.f, av0 = 0.f, av1 = 0.f;
    // (all threads fetch the nc x nslab partials of the slice at once -- one round trip -- and stage them in shared
    // memory; the coordinate's owner adds them up in slab order.  Slices too large for the staging area: owner loop.)
    // IMPORTANT: every global load of this block goes to a REGISTER first and the shared-memory stores come last.  A
    // store to shared memory through a generic pointer between two global loads makes the in-order issue wait for the
    // first load before it can issue the second (measured: 13 serialised L2 round trips, 2.9 us, in this block).
    const int tot = nc * t.nslab;
    const bool staged = 2 * tot <= TAIL_STAGE;
    float sg1 = 0.f, sg2 = 0.f;
    if (staged && tid < tot) {
        const int j = tid / t.nslab, q = tid - j * t.nslab, ii = c0 + j;
        if (ii < sp.d) { sg1 = __ldcg(t.part1 + (size_t)q * t.ldslab + ii); sg2 = __ldcg(t.part2 + (size_t)q * t.ldslab + ii); }
    }
    if (mine) {
        if (i >= sp.d) {
            v0 = __ldcg(t.acc + i); v1 = __ldcg(t.acc + accv + i);
        } else if (!staged) {
#pragma unroll 8
            for (int q = 0; q < t.nslab; ++q) {
                v0 += __ldcg(t.part1 + (size_t)q * t.ldslab + i);
                v1 += __ldcg(t.part2 + (size_t)q * t.ldslab + i);
            }
        }
        x0 = t.lam[i]; x1 = t.lam[D + i];
        if (t.mode == STEP_TAIL_UPDATE) {
            if (adam || dog) { s1m0 = t.m1[i]; s1m1 = t.m1[D + i]; }
            if (adam) { s2m0 = t.m2[i]; s2m1 = t.m2[D + i]; }
            if (polyavg) { av0 = t.avg[i]; av1 = t.avg[D + i]; }
        }
    }
    if (staged) {
        if (tid < tot) { stage[tid] = sg1; stage[tot + tid] = sg2; }
        for (int idx = tid + NUM_THREADS; idx < tot; idx += NUM_THREADS) {   // (slices with more than 640 partials)
            const int j = idx / t.nslab, q = idx - j * t.nslab, ii = c0 + j;
            const bool beta = ii < sp.d;
            const float a1 = beta ? __ldcg(t.part1 + (size_t)q * t.ldslab + ii) : 0.f;
            const float a2 = beta ? __ldcg(t.part2 + (size_t)q * t.ldslab + ii) : 0.f;
            stage[idx] = a1; stage[tot + idx] = a2;
        }
    }
    //  sticking the landing: sum_m eps and sum_m eps^2 of the slice (warp per coordinate, lanes over the samples)
    if (stl) {
        for (int j = warp; j < nc; j += NW) {
            float a2 = 0.f, a3 = 0.f;
            for (int m = lane; m < sp.Mloc; m += 32) {
                const float e = __ldcg(sp.E + (size_t)m * sp.ld + c0 + j);
                a2 += e; a3 = fmaf(e, e, a3);
            }
            a2 = warp_sum(a2); a3 = warp_sum(a3);
            if (lane == 0) { vals[2 * j] = a2; vals[2 * j + 1] = a3; }
        }
    }
    __syncthreads();   // stage[] / vals[] complete
    float sl = fmaf(t.w_lik, ctl->scratch[1], ctl->scratch[2]), sq = ctl->scratch[3];
    if (tid == 0) PSTAMP(sp.prof, 24);   // inputs loaded, scalars reduced
    if (stl && mine) { v2 = vals[2 * tid]; v3 = vals[2 * tid + 1]; }
    if (staged && mine && i < sp.d) {
        for (int q = 0; q < t.nslab; ++q) { v0 += stage[tid * t.nslab + q]; v1 += stage[tot + tid * t.nslab + q]; }
    }

    // ---- exchange over NVLink, multi-rank runs only (out of line: the single-rank iteration does not fetch its code)
    if (NR > 1) {
        const Sums6 x = tail_exchange(sp, ctl, sn.seq, i, mine, stl, Sums6{v0, v1, v2, v3, sl, sq});
        v0 = x.v0; v1 = x.v1; v2 = x.v2; v3 = x.v3; sl = x.sl; sq = x.sq;
    }

    MfSums S;
    S.s0 = sl; S.s1 = sq; S.s2 = 0.f; S.s3 = 0.f; S.logdet = ctl->scratch[0];
    float value, elbo, shift_next;
    mf_outputs(D, M, objective, entropy, S, sn.shift, value, elbo, shift_next);
    const bool bad = !isfinite(value);

    // ---- gradient entries of the slice: thread j < nc owns coordinate c0 + j (mu_i and s_i)
    float g0 = 0.f, g1 = 0.f;
    if (mine) {
        mf_grad_vals(v0, v1, v2, v3, x1, M, objective, entropy, S, g0, g1);
        t.grad[i] = g0; t.grad[D + i] = g1;
    }

    if (t.mode == STEP_TAIL_GRAD_OUT) {
        // estimate_gradient! boundary: the gradient goes straight to the caller-visible pinned buffer (posted writes);
        // the last CTA to finish adds the scalars, advances the step counter and release-stores the completion flag
        if (mine && t.host_out) { t.host_out[i] = g0; t.host_out[D + i] = g1; }
        __threadfence_system();
        __syncthreads();
        if (tid == 0) {
            const unsigned int tk = atomicAdd(t.done_ticket, 1u);
            if (tk == gridDim.x - 1) {
                *t.done_ticket = 0u;
                __threadfence();
                t.out[0] = value; t.out[1] = elbo; t.out[2] = S.logdet; t.out[3] = shift_next;
                const unsigned long long step_next = sn.step + 1ull;
                sp.st->step = step_next;
                sp.st->zt_kind = 0;   // the sample buffers hold this call's draws
                sp.gbar[3] = 0ull;
                if (NR > 1) *reinterpret_cast<volatile unsigned int*>(&t.comm.dev->seq) = sn.seq;
                if (t.host_out) {
                    float* tail = t.host_out + 2 * (size_t)D;
                    tail[0] = value; tail[1] = elbo; tail[2] = S.logdet; tail[3] = shift_next;
                    __threadfence_system();
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(reinterpret_cast<unsigned int*>(tail + 4)),
                                 "r"((unsigned int)step_next) : "memory");
                }
            }
        }
        return;
    }

    if (tid == 0) PSTAMP(sp.prof, 25);   // value known, gradient of the slice stored
    // ---- optimiser step (common.jl:91-94)
    float eta = a.rule == AVI_RULE_DESCENT ? a.h0 : 0.f, v_new = 0.f, r_new = 0.f;
    if (dog) {
        // two global norms (rules.jl:21-34, :52-64): per-CTA partials, one more grid barrier, fixed-order sum
        float dx2 = 0.f, g2 = 0.f;
        if (mine) {
            const float d0 = x0 - s1m0, d1 = x1 - s1m1;
            dx2 = fmaf(d0, d0, d1 * d1);
            g2 = fmaf(g0, g0, g1 * g1);
        }
        dx2 = block_sum(dx2, sm); g2 = block_sum(g2, sm);
        if (tid == 0) { t.norm_part[2 * blockIdx.x] = dx2; t.norm_part[2 * blockIdx.x + 1] = g2; }
        grid_barrier(gb);
        dx2 = 0.f; g2 = 0.f;
        for (int q = tid; q < (int)gridDim.x; q += NUM_THREADS) { dx2 += __ldcg(t.norm_part + 2 * q); g2 += __ldcg(t.norm_part + 2 * q + 1); }
        dx2 = block_sum(dx2, sm); g2 = block_sum(g2, sm);
        r_new = fmaxf(sqrtf(dx2), sn.r_old);
        if (a.rule == AVI_RULE_DOG) { v_new = sn.v_old + g2; eta = r_new / sqrtf(v_new); }
        else { const float r2 = r_new * r_new; v_new = sn.v_old + r2 * g2; eta = r2 / sqrtf(v_new); }
    }
    if (!sn.halted && !bad && mine) {
        const float w = (a.avg_param + 1.0f) / (sn.t_avg + a.avg_param);
        const float bc1 = 1.0f / (1.0f - sn.b1t), bc2 = 1.0f / (1.0f - sn.b2t);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const size_t p = (size_t)h * D + i;
            const float g = h ? g1 : g0;
            float xx = h ? x1 : x0, dx;
            if (adam) {
                const float mt = a.h1 * (h ? s1m1 : s1m0) + (1.0f - a.h1) * g;
                const float vt = a.h2 * (h ? s2m1 : s2m0) + (1.0f - a.h2) * g * g;
                t.m1[p] = mt; t.m2[p] = vt;
                dx = __fdividef(mt * bc1, sqrtf(vt * bc2) + a.h3) * a.h0;
            } else {
                dx = eta * g;
            }
            xx -= dx;
            if (h == 1 && a.op != AVI_OP_IDENTITY) {
                if (a.op == AVI_OP_CLIPSCALE) xx = fmaxf(xx, a.op_param);
                else xx = xx + (sqrtf(fmaf(xx, xx, 4.0f * eta)) - xx) * 0.5f;
            }
            t.lam[p] = xx;
            lam2[2 * tid + h] = xx;
            if (polyavg) t.avg[p] = (1.0f - w) * (h ? av1 : av0) + w * xx;
        }
    }
    if (tid == 0) PSTAMP(sp.prof, 26);   // update of the slice stored
    // ---- draw the NEXT iteration's samples for this CTA's slice under the lambda just computed
    const bool ahead = sp.draw_ahead && !sn.halted && !bad;
    if (ahead) {
        if (tid >= nc && tid < per) { lam2[2 * tid] = 0.f; lam2[2 * tid + 1] = 0.f; }   // padding coordinates of the last quad
        __syncthreads();
        draw_slice(sp, lam2, sn.step + 1ull, sn.key);
    }
    if (blockIdx.x == 0 && tid == 0) {
        if (NR > 1) *reinterpret_cast<volatile unsigned int*>(&t.comm.dev->seq) = sn.seq;
        if (!sn.halted) {
            t.out[0] = value; t.out[1] = elbo; t.out[2] = S.logdet; t.out[3] = shift_next;
            if (sn.tp < t.trace_cap) { t.trace[2 * sn.tp] = value; t.trace[2 * sn.tp + 1] = elbo; }
            sp.st->trace_pos = sn.tp + 1;
            if (bad) {
                sp.st->halted = 1;
            } else {
                t.sc[SC_T] = sn.t_avg + 1.0f;
                t.sc[SC_ETA] = eta;
                if (adam) { t.sc[SC_B1T] = sn.b1t * a.h1; t.sc[SC_B2T] = sn.b2t * a.h2; }
                if (dog) { t.sc[SC_V] = v_new; t.sc[SC_R] = r_new; }
                sp.st->step = sn.step + 1ull;
                sp.st->batch_cursor = sn.cursor + 1;
            }
        }
        // (draw_slice's stores are complete for every CTA when this launch ends; the tag is read by the next launch)
        sp.st->zt_kind = ahead ? 1 : 0;
        sp.st->zt_step = sn.step + 1ull; sp.st->zt_key = sn.key;
        sp.st->zt_nparts = (D + per - 1) / per; sp.st->zt_mloc = sp.Mloc;
        sp.gbar[3] = ahead ? (unsigned long long)(uintptr_t)sp.st : 0ull;
    }
}

// timeline slots (AVI_TIMELINE; atomicMin in 0..7, atomicMax in 8..15): first CTA 0 entered | 1 past the dependency wait |
// 2 past barrier 0 (forward starts) | 3 past barrier 1 (backward starts) | 4 past barrier 2 (tail starts);
// last CTA 8 left the sample phase | 9 left the forward phase | 10 left the backward phase | 11 done
template <int LIK, int X3, int BEPI = EPI_GLM_BWD>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k_glm_mf_step(const __grid_constant__ CUtensorMap tmZ, const __grid_constant__ CUtensorMap tmXr,
              const __grid_constant__ CUtensorMap tmXc, const __grid_constant__ CUtensorMap tmR,
              const __grid_constant__ StepParams sp) {
    extern __shared__ uint8_t smem_raw[];
    SmemCtl* ctl = reinterpret_cast<SmemCtl*>((reinterpret_cast<uintptr_t>(smem_raw) + 15) & ~(uintptr_t)15);
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ctl) + sizeof(SmemCtl) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    stamp_min(sp.tl, 0);
    unsigned long long* const prof = sp.prof;
    if (threadIdx.x == 0) PSTAMP(prof, 0);

    if (warp == 0 && lane == 0) {
        tc::tma_prefetch_desc(&tmZ); tc::tma_prefetch_desc(&tmXr); tc::tma_prefetch_desc(&tmXc); tc::tma_prefetch_desc(&tmR);
    }
    if (warp == 1) init_pipeline(ctl, sp.stages_f, 0);
    if (warp == 2) tc::tmem_alloc(&ctl->tmem_base, 512);
    pdl_trigger();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tmem_base = ctl->tmem_base;

    // forward phase: A = Zt (dependent), B = X rows.  A full-data X is static: request it before waiting for the
    // previous kernel; a minibatch copy is rewritten by the gather kernel of this iteration: request it after the wait.
    // (The dependency wait is taken per warp: warp 3 goes straight to it and fetches the state snapshot while warp 0 is
    // still issuing the requests for the static operand -- both are ~1 us of once-executed code.)
    int pre = 0;
    if (warp == 0 && sp.f.static_op == 2) {
        if (lane == 0) pre = preissue_early(&tmZ, &tmXr, sp.f, ctl, tiles, sp.stages_f, 2);
        pre = __shfl_sync(0xffffffffu, pre, 0);
    }
    if (threadIdx.x == 0) PSTAMP(prof, 1);
    pdl_wait();
    stamp_min(sp.tl, 1);
    if (threadIdx.x == 0) PSTAMP(prof, 2);
    if (warp == 0 && sp.f.static_op != 2) {   // minibatch copy of X: final once the gather kernel has completed
        if (lane == 0) pre = preissue_early(&tmZ, &tmXr, sp.f, ctl, tiles, sp.stages_f, 2);
        pre = __shfl_sync(0xffffffffu, pre, 0);
    }

    // ---- snapshot of everything the end of this iteration overwrites.  ONE warp per CTA fetches it (one lane per word,
    // a single L2 round trip) and hands it over through shared memory: when every thread of every CTA read these few
    // cache lines itself, 144 x 20 warps x ~15 loads queued on the same L2 lines for ~1.5 us.
    if (warp == 3) {
        unsigned long long v = 0ull;
        const unsigned long long* st64 = reinterpret_cast<const unsigned long long*>(sp.st);
        if (lane < 8) v = __ldcg(st64 + lane);                       // ObjDeviceState, 8 words (static_assert below)
        else if (lane == 8) v = __ldcg(sp.gbar + 1);
        else if (lane == 9) v = __ldcg(sp.gbar + 3);
        else if (lane < 16) { if (sp.t.mode == STEP_TAIL_UPDATE) v = __float_as_uint(__ldcg(sp.t.sc + (lane - 10))); }
        else if (lane == 16) { if (sp.t.mode != STEP_TAIL_NONE) v = __float_as_uint(__ldcg(sp.t.out + 3)); }
        else if (lane == 17) { if (sp.t.mode != STEP_TAIL_NONE && sp.t.comm.nranks > 1) v = *reinterpret_cast<volatile unsigned int*>(&sp.t.comm.dev->seq); }
        else if (lane >= 24 && sp.t.mode == STEP_TAIL_UPDATE) {
            // cold-L2 runs: pull this CTA's slice of lambda and of the optimiser state towards L2 now; the tail phase
            // would otherwise pay a DRAM round trip for them at the very end of the critical path
            const int per = slice_per(sp.D), c0 = min(sp.D - 1, (int)blockIdx.x * per);
            const float* base = lane < 26 ? sp.t.lam : lane < 28 ? sp.t.m1 : lane < 30 ? sp.t.m2 : sp.t.avg;
            if (base) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)(lane & 1) * sp.D + c0));
        }
        ctl->snap[lane] = v;
    }
    __syncthreads();
    GridBar gb;
    gb.ctr = sp.gbar; gb.base = ctl->snap[8]; gb.k = 0;
    Snap sn;
    sn.step = ctl->snap[0]; sn.key = ctl->snap[1]; sn.cursor = (long long)ctl->snap[2];
    sn.halted = (int)(unsigned)ctl->snap[3]; sn.tp = (int)(ctl->snap[3] >> 32);
    const int zt_kind = (int)(unsigned)ctl->snap[4], zt_mloc = (int)(unsigned)ctl->snap[5];
    sn.zt_nparts = (int)(ctl->snap[4] >> 32);
    sn.b1t = sn.b2t = sn.t_avg = sn.v_old = sn.r_old = sn.shift = sn.logdet = 0.f; sn.seq = 0;
    // (both tags: the objective's -- z, eps are its buffers -- and the target's -- Zt and the partial sums are shared by
    // every objective over that target)
    sn.zt_ok = sp.draw_ahead && zt_kind == 1 && ctl->snap[6] == sn.step && ctl->snap[7] == sn.key &&
               zt_mloc == sp.Mloc && ctl->snap[9] == (unsigned long long)(uintptr_t)sp.st;
    if (sp.t.mode != STEP_TAIL_NONE) {
        if (sp.t.mode == STEP_TAIL_UPDATE) {
            sn.t_avg = __uint_as_float((unsigned)ctl->snap[10 + SC_T]); sn.v_old = __uint_as_float((unsigned)ctl->snap[10 + SC_V]);
            sn.r_old = __uint_as_float((unsigned)ctl->snap[10 + SC_R]); sn.b1t = __uint_as_float((unsigned)ctl->snap[10 + SC_B1T]);
            sn.b2t = __uint_as_float((unsigned)ctl->snap[10 + SC_B2T]);
        }
        sn.shift = __uint_as_float((unsigned)ctl->snap[16]);
        if (sp.t.comm.nranks > 1) sn.seq = (unsigned)ctl->snap[17] + 1u;
    }

    if (threadIdx.x == 0) PSTAMP(prof, 3);
    if (sp.draw_ahead) {
        // optimiser loop: samples come slice-major, normally drawn ahead by the previous launch's tail phase
        if (!sn.zt_ok) {
            float* lam2 = &ctl->tail[TAIL_STAGE + 2 * TAIL_MAX_PER_CTA];
            const int per = slice_per(sp.D), c0 = (int)blockIdx.x * per;
            const float* src = sp.lambda_src ? sp.lambda_src : sp.lambda;
            for (int j = threadIdx.x; j < per; j += NUM_THREADS) {
                const bool in = c0 + j < sp.D;
                const float m_ = in ? src[c0 + j] : 0.f, s_ = in ? src[sp.D + c0 + j] : 0.f;
                lam2[2 * j] = m_; lam2[2 * j + 1] = s_;
                if (sp.lambda_src && in) { sp.t.lam[c0 + j] = m_; sp.t.lam[sp.D + c0 + j] = s_; }   // device copy for the tail
            }
            __syncthreads();
            if (sp.lambda_src && warp == 2) {   // this slice's share of log det(scale)
                float part = 0.f;
                for (int j = lane; j < per; j += 32) if (c0 + j < sp.D) part += __logf(lam2[2 * j + 1]);
                part = warp_sum(part);
                if (lane == 0) sp.ldpart[blockIdx.x] = part;
            }
            draw_slice(sp, lam2, sn.step, sn.key);
            sn.zt_nparts = (sp.D + per - 1) / per;
            stamp_max(sp.tl, 8);
            if (threadIdx.x == 0) PSTAMP(prof, 4);
            grid_barrier(gb);
            stamp_min(sp.tl, 2);
        }
    } else if (sp.do_sample) {
        sample_phase(sp, ctl, sn.step, sn.key);
        stamp_max(sp.tl, 8);
        if (threadIdx.x == 0) PSTAMP(prof, 4);
        grid_barrier(gb);
        stamp_min(sp.tl, 2);
    }
    if (threadIdx.x == 0) PSTAMP(prof, 5);
    if (sp.draw_ahead && warp >= 2 && warp < 4) finalize_samples(sp, sn.zt_nparts);
    // log det of the scale (sum_i log s_i) is only needed by the tail: the otherwise idle warp 2 takes it while the
    // forward contraction runs (lambda is not overwritten before the tail phase of this launch)
    if (warp == 2 && sp.t.mode != STEP_TAIL_NONE) {
        float part = 0.f;
        if (sp.lambda_src) { for (int c = lane; c < (int)gridDim.x; c += 32) part += __ldcg(sp.ldpart + c); }   // (published by barrier 0)
        else for (int i = lane; i < sp.D; i += 32) part += __logf(sp.lambda[sp.D + i]);
        part = warp_sum(part);
        if (lane == 0) ctl->scratch[0] = part;
    }

    const int handed = tc_phase<EPI_GLM_FWD, LIK, X3>(&tmZ, &tmXr, sp.f, ctl, tiles, tmem_base, sp.stages_f, pre, 2, prof, 6,
                                                      &tmXc, &tmR, &sp.b, sp.stages_b, 1);
    stamp_max(sp.tl, 9);

    // ---- drain, re-carve the ring for the backward geometry, request its static operand (X columns) while the
    //      other CTAs finish, then the barrier that publishes R
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 1) init_pipeline(ctl, sp.stages_b, sp.stages_f, /*accumulator barriers only*/ 2);
    __syncthreads();
    pre = warp == 0 ? max(handed, 0) : 0;   // (the producer re-carved the ring and requested the X columns long ago)
    if (threadIdx.x == 0) PSTAMP(prof, 12);   // arriving at barrier 1 (ring re-carved, X columns requested)
    grid_barrier(gb);
    stamp_min(sp.tl, 3);
    tc::fence_after_sync();
    if (threadIdx.x == 0) PSTAMP(prof, 13);

    if (sp.t.mode != STEP_TAIL_NONE) tail_scalars(sp, ctl);
    tc_phase<BEPI, 0, X3>(&tmXc, &tmR, sp.b, ctl, tiles, tmem_base, sp.stages_b, pre, 1, prof, 14);
    stamp_max(sp.tl, 10);

    if (sp.t.mode != STEP_TAIL_NONE) {
        if (threadIdx.x == 0) PSTAMP(prof, 20);   // arriving at barrier 2
        grid_barrier(gb);
        stamp_min(sp.tl, 4);
        if (threadIdx.x == 0) PSTAMP(prof, 21);
        tail_phase(sp, ctl, sn, gb);
        if (threadIdx.x == 0) PSTAMP(prof, 22);
    }

    // (CTA 0 has passed the last barrier of the launch: every CTA has read the old base)
    if (blockIdx.x == 0 && threadIdx.x == 0) sp.gbar[1] = gb.base + (unsigned long long)gb.k * gridDim.x;
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 2) {
        tc::fence_after_sync();
        tc::tmem_dealloc(tmem_base, 512);
    }
    stamp_max(sp.tl, 11);
    if (threadIdx.x == 0) PSTAMP(prof, 23);
}


}  // namespace

static unsigned long long* g_prof = nullptr;
static int g_prof_grid = 0;

int avi_step_fused_max_per_cta() { return TAIL_MAX_PER_CTA; }

// debug build only (-DAVI_WATCHDOG): the first wait that timed out, see tc_common.cuh; out[0] == 0: none
extern "C" int32_t avi_step_fused_hang_get(avi_ctx* ctx, uint32_t* out8) {
#ifdef AVI_WATCHDOG
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpyFromSymbol(out8, tc::avi_hang_report, 8 * sizeof(unsigned int));
    unsigned int z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(tc::avi_hang_report, z, sizeof(z));
    return 1;
#else
    (void)ctx; (void)out8;
    return 0;
#endif
}

// diagnostic (AVI_STEP_PROF=1): stamps of the most recent fused launch, [grid][32] u64; returns the grid size (0: off)
extern "C" int32_t avi_step_fused_prof_get(avi_ctx* ctx, uint64_t* out, int32_t max_ctas) {
    if (!g_prof || !out || max_ctas < g_prof_grid) return 0;
    cudaStreamSynchronize(ctx->stream);
    cudaMemcpy(out, g_prof, (size_t)g_prof_grid * 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    return g_prof_grid;
}

int32_t avi_step_fused_launch(avi_ctx* ctx, const CUtensorMap& tmZ, const CUtensorMap& tmXr, const CUtensorMap& tmXc,
                              const CUtensorMap& tmR, StepParams& sp) {
    static bool attr_done = false;
    if (!attr_done) {
        AVI_CUDA(ctx, cudaFuncSetAttribute(k_glm_mf_step<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        AVI_CUDA(ctx, cudaFuncSetAttribute(k_glm_mf_step<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        AVI_CUDA(ctx, cudaFuncSetAttribute(k_glm_mf_step<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        AVI_CUDA(ctx, cudaFuncSetAttribute(k_glm_mf_step<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        AVI_CUDA(ctx, cudaFuncSetAttribute(k_glm_mf_step<0, 0, EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        AVI_CUDA(ctx, cudaFuncSetAttribute(k_glm_mf_step<1, 0, EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT));
        attr_done = true;
    }
    const int extra = 16 + (int)sizeof(SmemCtl) + 1024;
    auto stage_bytes_of = [&](const TcParams& p) { return A_TILE_BYTES + (p.b_mn ? ((p.nt + 31) / 32) * 4096 : p.nt * BK * 4); };
    auto stages_for = [&](const TcParams& p) { return std::max(2, std::min(MAX_STAGES, (SMEM_LIMIT - extra) / stage_bytes_of(p))); };
    sp.stages_f = stages_for(sp.f); sp.stages_b = stages_for(sp.b);
    const int smem = extra + std::max(sp.stages_f * stage_bytes_of(sp.f), sp.stages_b * stage_bytes_of(sp.b));
    if (smem > SMEM_LIMIT) AVI_FAIL(ctx, AVI_ERR_INVALID, "tiles do not fit in shared memory");
    if (sp.f.pair || sp.b.pair || sp.f.ca * sp.f.cb != 1 || sp.b.ca * sp.b.cb != 1)
        AVI_FAIL(ctx, AVI_ERR_INVALID, "the fused iteration runs single-CTA tiles only");
    const int sms = ctx->prop.multiProcessorCount;
    const int64_t units_f = (int64_t)sp.f.n_ablk * sp.f.n_bchunk * sp.f.n_ksplit;
    const int64_t units_b = (int64_t)sp.b.n_ablk * sp.b.n_bchunk * sp.b.n_ksplit;
    int grid = (int)std::min<int64_t>(sms, std::max<int64_t>(units_f, units_b));
    // the tail deals D coordinates to the CTAs, at most TAIL_MAX_PER_CTA each
    if (sp.t.mode != STEP_TAIL_NONE) grid = std::max(grid, std::min(sms, (int)ceil_div(sp.D, TAIL_MAX_PER_CTA)));
    if (sp.t.mode != STEP_TAIL_NONE && ceil_div(sp.D, grid) > TAIL_MAX_PER_CTA)
        AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED, "too many coordinates for the fused tail");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(NUM_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    int na = 0;
    if (avi_pdl_enabled()) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    sp.tl = ctx->tl;
    // AVI_STEP_PROF=1: per-CTA phase stamps, copied out by avi_step_fused_prof_get
    static const bool prof_on = getenv("AVI_STEP_PROF") && atoi(getenv("AVI_STEP_PROF")) != 0;
    if (prof_on) {
        if (!g_prof) { if (cudaMalloc(&g_prof, 160 * 32 * sizeof(unsigned long long)) != cudaSuccess) g_prof = nullptr; }
        if (g_prof && !ctx->capturing) cudaMemsetAsync(g_prof, 0, 160 * 32 * sizeof(unsigned long long), ctx->stream);
        sp.prof = g_prof; g_prof_grid = grid;
    }
    AviTimed timed(ctx, sp.bwd_store ? "glm_fwd_bwd" : "glm_step");
    const bool bern = sp.f.likelihood == AVI_GLM_BERNOULLI_LOGIT, x3 = sp.f.r_seg != 0;
    if (sp.bwd_store) {
        if (x3 || sp.t.mode != STEP_TAIL_NONE || sp.do_sample || sp.draw_ahead)
            AVI_FAIL(ctx, AVI_ERR_INVALID, "the gradient-store variant runs forward + backward only, plain TF32");
        cudaError_t es = bern ? cudaLaunchKernelEx(&cfg, k_glm_mf_step<0, 0, EPI_STORE>, tmZ, tmXr, tmXc, tmR, sp)
                              : cudaLaunchKernelEx(&cfg, k_glm_mf_step<1, 0, EPI_STORE>, tmZ, tmXr, tmXc, tmR, sp);
        if (es != cudaSuccess) AVI_FAIL(ctx, AVI_ERR_CUDA, std::string("fused forward + backward launch: ") + cudaGetErrorString(es));
        AVI_LAUNCHED(ctx);
        return AVI_OK;
    }
    cudaError_t e = bern ? (x3 ? cudaLaunchKernelEx(&cfg, k_glm_mf_step<0, 1>, tmZ, tmXr, tmXc, tmR, sp)
                               : cudaLaunchKernelEx(&cfg, k_glm_mf_step<0, 0>, tmZ, tmXr, tmXc, tmR, sp))
                         : (x3 ? cudaLaunchKernelEx(&cfg, k_glm_mf_step<1, 1>, tmZ, tmXr, tmXc, tmR, sp)
                               : cudaLaunchKernelEx(&cfg, k_glm_mf_step<1, 0>, tmZ, tmXr, tmXc, tmR, sp));
    if (e != cudaSuccess) AVI_FAIL(ctx, AVI_ERR_CUDA, std::string("fused step launch: ") + cudaGetErrorString(e));
    AVI_LAUNCHED(ctx);
    return AVI_OK;
}
