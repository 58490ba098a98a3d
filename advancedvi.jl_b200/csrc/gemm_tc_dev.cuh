// Device-side pieces of the tcgen05 contraction shared by the stand-alone kernels (gemm_tc.cu) and the fused
// whole-iteration kernel (step_fused.cu): tile constants, the shared-memory control block and the three epilogues.
// COH = the epilogue's inputs may have been written EARLIER IN THE SAME KERNEL by other CTAs (fused kernel): they are
// then read with ld.global.cg (L2, the coherence point) instead of the read-only path.
#pragma once

#include "avi_internal.cuh"
#include "device_utils.cuh"
#include "gemm_tc.cuh"
#include "tc_common.cuh"

namespace {

template <bool COH>
__device__ __forceinline__ float ld_in(const float* p) { return COH ? __ldcg(p) : __ldg(p); }

constexpr int BM = 128;            // UMMA M (TMEM lanes)
constexpr int BK = 32;             // fp32 elements per 128-byte swizzle row
constexpr int MAX_STAGES = 8;
constexpr int A_TILE_BYTES = BM * BK * 4;          // 16 KB
constexpr int EPI_WARPS = 16;                      // 4 per TMEM lane quarter, each a column quarter
constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;  // 640
constexpr int SMEM_LIMIT = 232448;                 // 227 KB opt-in maximum per CTA

struct SmemCtl {
    uint64_t full[MAX_STAGES], empty[MAX_STAGES], tmem_full[2], tmem_empty[2];
    uint32_t tmem_base;
    int flag;
    float scratch[8];               // fused iteration kernel: values handed from one phase to a later one
    unsigned long long snap[32];    // fused iteration kernel: entry snapshot of the device state (one warp loads, all read)
    float red16[2][16];             // fused iteration kernel: per-warp partials of a unit's log-likelihood total
    float tail[1024 + 512];         // fused iteration kernel, tail phase: slab staging + per-coordinate sums
    alignas(16) float ys[2][512];   // FWD: y per accumulator stage (<= 256 used); BWD: reduction scratch
};

__device__ __forceinline__ unsigned long long gtime_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// phase timestamps (ns, %globaltimer) per CTA when TcParams.prof != nullptr:
// 0 kernel entry | 1 prologue done | 2 first operands landed | 3 last MMA committed (issue side)
// 4 accumulator ready (epilogue side) | 5 epilogue math done | 6 kernel exit
#define TC_STAMP(slot, cond) do { if (p.prof && (cond)) p.prof[(size_t)blockIdx.x * 8 + (slot)] = gtime_ns(); } while (0)

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// Epilogue of one work unit for one thread: waits for the accumulator stage, consumes the thread's TMEM
// row (lane) over its column range and returns the per-thread partial sums.
// POST / X3 >= 0 fix TcParams::post_on / the 3xTF32 layout at compile time (the fused iteration kernel executes every
// instruction of its epilogues once per launch, so code it cannot reach still costs instruction fetches around it);
// -1: read them from the parameters (stand-alone kernels).
template <int EPI, int LIK, bool COH = false, int POST = -1, int X3 = -1>
__device__ __forceinline__ void epilogue_unit(const TcParams& p_, SmemCtl* ctl, int as, uint32_t aphase,
                                              uint32_t tacc, int NT, int a, bool a_ok, int bc, int ks, int c_begin,
                                              int c_end, int et, float& s1, float& s2) {
    const TcParams& p = p_;
    const int post_on = POST >= 0 ? POST : p.post_on;
    const int r_seg = X3 == 0 ? 0 : p.r_seg;
    s1 = 0.0f; s2 = 0.0f;
    if (EPI == EPI_GLM_FWD) {
        int b = bc * NT + et;
        if (et < NT) ctl->ys[as][et] = b < p.Nb ? __ldg(p.y + b) : 0.0f;
        epi_bar_sync();
    }
    if (EPI == EPI_GLM_BWD) {
        // eps[b][a] for this thread's columns, fetched while the MMAs are still running
        const float* Ea = p.E + a;
        float pr1 = 0.0f, pr2 = 0.0f;
        if (post_on) {
            // Work that does not depend on this kernel's MMAs, done while they run (the epilogue warps would
            // otherwise sleep on the accumulator barrier):
            // (1) log pi(z_m) = w * sum(partial log-lik of the forward kernel) + log prior, by the a-block-0 CTAs
            const int ab0 = a / BM;
            if (ab0 == 0 && post_on == 1) {   // (post_on == 2, fused iteration: the tail phase takes sum_m log pi itself)
                float* red = &ctl->ys[0][0];   // 16 warps x 32 lanes
                const int ew = et >> 5, ln = et & 31;
                for (int m0 = (ks * p.n_bchunk + bc) * 32; m0 < p.Nb; m0 += p.n_ksplit * p.n_bchunk * 32) {
                    const int m = m0 + ln;
                    float sll = 0.0f;
                    if (m < p.Nb) {
#pragma unroll 6
                        for (int q = ew; q < p.post_nparts; q += EPI_WARPS) sll += ld_in<COH>(p.post_llpart + (size_t)q * p.post_ldll + m);
                    }
                    red[ew * 32 + ln] = sll;
                    epi_bar_sync();
                    if (ew == 0 && m < p.Nb) {
                        float t = 0.0f;
#pragma unroll
                        for (int w2 = 0; w2 < EPI_WARPS; ++w2) t += red[w2 * 32 + ln];
                        p.post_logp[m] = fmaf(p.post_w, t, ld_in<COH>(p.post_pre + 4 * (size_t)m));
                    }
                    epi_bar_sync();
                }
            }
            // (1b) eta = theta[d]: sum over ALL samples of d log pi / d eta and of its product with eps
            if (ab0 == 0 && ks == 0 && bc == 0 && (et >> 5) == EPI_WARPS - 1) {
                const int ln = et & 31;
                float t1 = 0.0f, t2 = 0.0f;
#pragma unroll 8
                for (int m = ln; m < p.Nb; m += 32) {
                    const float ge = ld_in<COH>(p.post_pre + 4 * (size_t)m + 2);
                    t1 += ge;
                    t2 = fmaf(ge, ld_in<COH>(p.E + (size_t)m * p.lde + p.Ma), t2);
                }
                t1 = warp_sum(t1); t2 = warp_sum(t2);
                if (ln == 0) { p.post_a1[p.Ma] = t1; p.post_a2[p.Ma] = t2; }
            }
            // (2) the prior part of grad_beta, -beta / sigma^2: the samples of this thread's column range are dealt
            // round-robin to the k-split CTAs of the a-block (the sum is linear: any fixed partition is exact), so
            // each thread issues only a handful of independent loads
            if (a_ok) {
                const float* Za = p.post_Z + a;
                const int first = c_begin + ((ks - c_begin) % p.n_ksplit + p.n_ksplit) % p.n_ksplit;
#pragma unroll 4
                for (int c = first; c < c_end; c += p.n_ksplit) {
                    const int b = bc * NT + c;
                    if (b < p.Nb) {
                        const float gz = -ld_in<COH>(Za + (size_t)b * p.lde) * ld_in<COH>(p.post_pre + 4 * (size_t)b + 1);
                        pr1 += gz;
                        pr2 = fmaf(gz, ld_in<COH>(Ea + (size_t)b * p.lde), pr2);
                    }
                }
            }
        }
        float e[32];
        bool waited = false;
        for (int c = c_begin; c < c_end; c += 32) {
            const int nc = min(32, c_end - c);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int b = bc * NT + c + j;
                e[j] = (j < nc && a_ok && b < p.Nb) ? ld_in<COH>(Ea + (size_t)b * p.lde) : 0.0f;
            }
            if (!waited) { tc::mbar_wait(&ctl->tmem_full[as], aphase); tc::fence_after_sync(); waited = true; TC_STAMP(4, et == 0); }
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                if (h * 8 < nc) {
                    float v[8];
                    tc::tmem_ld8(tacc + (uint32_t)(c + h * 8), v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const bool ok = bc * NT + c + h * 8 + j < p.Nb;
                        s1 += ok ? v[j] : 0.0f;
                        s2 = fmaf(v[j], e[h * 8 + j], s2);
                    }
                }
            }
        }
        if (!waited) { tc::mbar_wait(&ctl->tmem_full[as], aphase); tc::fence_after_sync(); }
        s1 += pr1; s2 += pr2;
    } else {
        if (p.dbg & 4) { while (!tc::mbar_try_wait(&ctl->tmem_full[as], aphase)) __nanosleep(500); }   // experiment: sleeping waiters
        else tc::mbar_wait(&ctl->tmem_full[as], aphase);
        tc::fence_after_sync();
        TC_STAMP(4, et == 0);
        for (int c = c_begin; c < c_end; c += 8) {
            float v[8];
            tc::tmem_ld8(tacc + (uint32_t)c, v);
            const int b0 = bc * NT + c;
            if (EPI == EPI_GLM_FWD) {
                float r[8], yv[8];
                {   // c is a multiple of 8 and ys is 16-byte aligned: two LDS.128
                    const float4 y0 = *reinterpret_cast<const float4*>(&ctl->ys[as][c]);
                    const float4 y1 = *reinterpret_cast<const float4*>(&ctl->ys[as][c + 4]);
                    yv[0] = y0.x; yv[1] = y0.y; yv[2] = y0.z; yv[3] = y0.w; yv[4] = y1.x; yv[5] = y1.y; yv[6] = y1.z; yv[7] = y1.w;
                }
                const bool full = b0 + 8 <= p.Nb;   // only the last chunk of the data rows needs per-element masks
                float lpv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float lp, rr;
                    if (p.dbg & 32) { lp = v[j]; rr = yv[j] - v[j]; }   // (dbg 32: timing experiment without the logistic terms)
                    else if (LIK == 0) {
                        const float l = v[j];
                        const float e = tc::ex2_approx(-1.4426950408889634f * fabsf(l));   // exp(-|l|) in (0, 1]
                        const float inv = tc::rcp_approx(1.0f + e);                         // in [1/2, 1)
                        const float sig = l >= 0.0f ? inv : e * inv;
                        // y l - log1pexp(l) = y l - max(l, 0) + ln(inv)
                        lp = fmaf(yv[j], l, fmaf(0.6931471805599453f, tc::lg2_approx(inv), -fmaxf(l, 0.0f)));
                        rr = yv[j] - sig;
                    } else {
                        rr = yv[j] - v[j];
                        lp = fmaf(-0.5f * rr, rr, -0.5f * AVI_LOG2PI);
                    }
                    lpv[j] = lp; r[j] = p.w * rr;
                }
                if (!full) {   // warp-uniform, rare
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (b0 + j >= p.Nb) { lpv[j] = 0.0f; r[j] = 0.0f; }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) s1 += lpv[j];
                float rlo[8];
                if (r_seg) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float hi = tc::round_tf32(r[j]);
                        rlo[j] = tc::round_tf32(r[j] - hi);
                        r[j] = hi;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) { r[j] = tc::round_tf32(r[j]); rlo[j] = 0.0f; }
                }
                // (staging this tile through shared memory for 128-byte row stores was measured slower: the
                // extra STS + barrier cost more than the 32-byte-sector stores; profiles/README.md)
                if (p.c_mn) {
                    // transposed store Rt[b][a]: one coalesced 128-byte row segment per warp and data row.  Rows past Nb:
                    // never read in the plain mode (the tensor map ends at Nb: zero fill), zeros up to the end of the
                    // segment in the 3xTF32 mode.
                    if (a_ok && !(p.dbg & 16)) {
                        float* dst = p.C + (size_t)b0 * p.ldc + a;
                        const int lim = r_seg ? r_seg : p.Nb;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (full || b0 + j < lim) {
                                dst[(size_t)j * p.ldc] = r[j];
                                if (r_seg) {
                                    dst[(size_t)(r_seg + j) * p.ldc] = rlo[j];
                                    dst[(size_t)(2 * r_seg + j) * p.ldc] = r[j];
                                }
                            }
                        }
                    }
                } else if (a_ok && b0 < (r_seg ? r_seg : p.ldc) && !(p.dbg & 16)) {   // (dbg 16: timing experiment without the R stores)
                    float4* dst = reinterpret_cast<float4*>(p.C + (size_t)a * p.ldc + b0);
                    dst[0] = make_float4(r[0], r[1], r[2], r[3]);
                    dst[1] = make_float4(r[4], r[5], r[6], r[7]);
                    if (r_seg) {   // 3xTF32: R is the B operand of the backward contraction: [hi | lo | hi]
                        float4* dl = reinterpret_cast<float4*>(p.C + (size_t)a * p.ldc + r_seg + b0);
                        dl[0] = make_float4(rlo[0], rlo[1], rlo[2], rlo[3]);
                        dl[1] = make_float4(rlo[4], rlo[5], rlo[6], rlo[7]);
                        float4* dh = reinterpret_cast<float4*>(p.C + (size_t)a * p.ldc + 2 * (size_t)r_seg + b0);
                        dh[0] = make_float4(r[0], r[1], r[2], r[3]);
                        dh[1] = make_float4(r[4], r[5], r[6], r[7]);
                    }
                }
            } else {
                float* Cs = p.C + (size_t)ks * p.slab_stride;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int b = b0 + j;
                    if (b < p.Nb && a_ok) Cs[(size_t)b * p.ldc + a] = v[j];
                }
            }
        }
    }
}

// EPI_GLM_BWD with post_on: the last CTA to finish an a-block (all k-splits and b-chunks) adds up the partial
// slabs of its 128 coordinates in a fixed order and writes the final sum_m g / sum_m g*eps; the last CTA of
// a-block 0 also produces the eta coordinate.  Replaces the separate k_glm_post_sums launch.
__device__ __forceinline__ void epilogue_bwd_combine(const TcParams& p, SmemCtl* ctl, int ab, int et) {
    // (called right after epilogue_bwd_store: slab rows written, fenced, barrier passed)
    if (et == 0) {
        const unsigned int total = (unsigned int)(p.n_bchunk * p.n_ksplit);
        const unsigned int t = atomicAdd(p.post_tickets + ab, 1u);
        ctl->flag = (t == total - 1u);
        if (t == total - 1u) p.post_tickets[ab] = 0u;   // re-arm for the next launch
    }
    epi_bar_sync();
    if (!ctl->flag) return;
    __threadfence();
    float* red1 = &ctl->ys[0][0];   // 4 groups x 128 coordinates, per array
    float* red2 = &ctl->ys[1][0];
    const int lc = et & 127, g = et >> 7, coord = ab * BM + lc;
    const int nslab = p.n_ksplit * p.n_bchunk;
    float t1 = 0.0f, t2 = 0.0f;
    if (coord < p.Ma) {
#pragma unroll 6
        for (int q = g; q < nslab; q += 4) {
            t1 += p.part1[(size_t)q * p.ldpart + coord];
            t2 += p.part2[(size_t)q * p.ldpart + coord];
        }
    }
    red1[g * 128 + lc] = t1; red2[g * 128 + lc] = t2;
    epi_bar_sync();
    if (g == 0 && coord < p.Ma) {
        p.post_a1[coord] = ((red1[lc] + red1[128 + lc]) + red1[256 + lc]) + red1[384 + lc];
        p.post_a2[coord] = ((red2[lc] + red2[128 + lc]) + red2[256 + lc]) + red2[384 + lc];
    }
    epi_bar_sync();
}

template <int EPI>
__device__ __forceinline__ void epilogue_store_partials(const TcParams& p, bool a_ok, int a, int bc, int ks, int cq,
                                                        float s1, float s2) {
    if (EPI == EPI_GLM_FWD) {
        if (a_ok) p.part1[(size_t)(bc * 4 + cq) * p.ldpart + a] = s1;
    }
}

// EPI_GLM_BWD: the four column-quarter warps of a TMEM lane combine their partial sums through shared memory
// (fixed order) so that a unit contributes ONE slab row per coordinate.  Ends with the slab rows written and
// fenced and all 512 epilogue threads past a barrier (which epilogue_bwd_combine relies on).
__device__ __forceinline__ void epilogue_bwd_store(const TcParams& p, SmemCtl* ctl, bool a_ok, int a, int bc, int ks,
                                                   int cq, int et, float s1, float s2) {
    float* red1 = &ctl->ys[0][0];           // [4 quarters][128 rows]
    float* red2 = &ctl->ys[1][0];
    const int row = a & (BM - 1);
    red1[cq * BM + row] = s1; red2[cq * BM + row] = s2;
    epi_bar_sync();
    if (cq == 0 && a_ok) {
        const size_t slab = (size_t)(ks * p.n_bchunk + bc) * p.ldpart;
        p.part1[slab + a] = ((red1[row] + red1[BM + row]) + red1[2 * BM + row]) + red1[3 * BM + row];
        p.part2[slab + a] = ((red2[row] + red2[BM + row]) + red2[2 * BM + row]) + red2[3 * BM + row];
    }
    if (p.post_on != 2) __threadfence();   // (fused iteration: the grid barrier before the tail phase publishes the slabs)
    epi_bar_sync();
}

// EPI_GLM_FWD in the fused iteration (post_on == 2): RepGradELBO only needs sum_m log pi(z_m), so a unit contributes
// ONE number -- the log-likelihood summed over its 128 samples x NT rows, fixed order -- instead of per-sample partials.
__device__ __forceinline__ void epilogue_fwd_unit_total(const TcParams& p, SmemCtl* ctl, int as, int u, int et, float s1) {
    const int ew = et >> 5, ln = et & 31;
    s1 = warp_sum(s1);
    if (ln == 0) ctl->red16[as][ew] = s1;
    epi_bar_sync();
    if (et == 0) {
        float t = 0.0f;
#pragma unroll
        for (int w2 = 0; w2 < EPI_WARPS; ++w2) t += ctl->red16[as][w2];
        p.part1[u] = t;
    }
}

}  // namespace
