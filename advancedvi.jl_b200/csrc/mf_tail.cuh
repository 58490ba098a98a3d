// The mean-field tail of one iteration -- finalize (sums -> gradient, value, elbo), finiteness check,
// rule + operator + averager, commit, and with several ranks the NVLink exchange of the partial sums -- as a device
// function with two launch shapes (opt.cu): one thread-block cluster of 8 CTAs x 256 threads (default) or one CTA of
// 1024 threads (AVI_TAIL_CLUSTER=0).
#pragma once

#include "avi_internal.cuh"
#include "device_utils.cuh"
#include "comm_dev.cuh"
#include "mf_finalize.cuh"
#include "tc_common.cuh"

// scalar state sc[]: 0 averaging t | 1 DoG v | 2 DoG r | 3 beta1^t | 4 beta2^t | 5 last step size
enum { SC_T = 0, SC_V = 1, SC_R = 2, SC_B1T = 3, SC_B2T = 4, SC_ETA = 5, SC_N = 16 };

struct UpdArgs {
    int rule, op, averager;
    float h0, h1, h2, h3;   // rule hyper-parameters
    float op_param, avg_param;
    int D, fullrank;
    long long P;
    int nparts;             // DoG/DoWG: number of partial norms
};


struct MfTailArgs {
    const float* acc; int accv; int M, objective, entropy;
    const float* logp; const float* esq; int Mloc, deferred;
    float *lam, *grad, *m1, *m2, *avg, *sc, *out;
    ObjDeviceState* st;
    float* trace; int trace_cap;
    UpdArgs a;
    unsigned long long* tl;   // step timeline (diagnostic): slots tl_s (entered), tl_s + 4 (past the wait), tl_s + 8 (done)
    int tl_s;
    CommPeers comm;    // comm.nranks > 1: the sample-sharded exchange of `acc` runs inside this kernel
    long long acc_len;
    float h0;          // entropy of the base distribution (base_dist.cuh)
};

// block sum for a 1024-thread CTA addressed by a linear thread id (any block shape); fixed tree
__device__ __forceinline__ float block_sum_1024(float v, float* sm) {
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;
    const int lane = tid & 31, w = tid >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sm[w] = v;
    __syncthreads();
    if (w == 0) {
        float t = warp_sum(sm[lane]);
        if (lane == 0) sm[32] = t;
    }
    __syncthreads();
    return sm[32];
}

// The same sums when the tail runs as ONE CLUSTER of 8 CTAs x 256 threads (tid = 256 * cta rank + thread; 2048 threads,
// so a thread owns ONE coordinate up to D = 2048): the 8 warp partials of a CTA go to its shared memory, one cluster
// barrier, then every warp reads the 64 partials of the cluster through distributed shared memory (lane l <- warps l
// and l + 32 of the cluster, in this fixed order) and finishes with a shuffle tree: every thread of every CTA gets
// the same bits.  Up to three values are reduced per barrier.  Two buffers alternate with `round`, so ONE barrier
// per call suffices: a CTA rewrites a buffer only after the next call's barrier, which every CTA reaches after its reads.
constexpr int TAIL_CLUSTER = 8, TAIL_CL_THREADS = 256, TAIL_CL_WARPS = TAIL_CL_THREADS / 32;
__device__ __forceinline__ void cluster_sum3(float& a, float& b, float& c, float (*smc)[TAIL_CL_WARPS][4], int& round) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    a = warp_sum(a); b = warp_sum(b); c = warp_sum(c);
    float (*buf)[4] = smc[round & 1];
    if (lane == 0) { buf[w][0] = a; buf[w][1] = b; buf[w][2] = c; }
    tc::cluster_sync();
    float4 p0, p1;
    {
        const int g0 = lane, g1 = lane + 32;   // global warp index = 8 * cta + warp
        const uint32_t r0 = tc::mapa_u32(tc::smem_u32(&buf[g0 % TAIL_CL_WARPS][0]), (uint32_t)(g0 / TAIL_CL_WARPS));
        const uint32_t r1 = tc::mapa_u32(tc::smem_u32(&buf[g1 % TAIL_CL_WARPS][0]), (uint32_t)(g1 / TAIL_CL_WARPS));
        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(p0.x), "=f"(p0.y), "=f"(p0.z), "=f"(p0.w) : "r"(r0) : "memory");
        asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(p1.x), "=f"(p1.y), "=f"(p1.z), "=f"(p1.w) : "r"(r1) : "memory");
    }
    ++round;
    a = warp_sum(p0.x + p1.x); b = warp_sum(p0.y + p1.y); c = warp_sum(p0.z + p1.z);
}

// Mean-field tail of one iteration in ONE launch: finalize (sums -> gradient, value, elbo)
// + finiteness check + rule + operator + averager + commit of the scalar state / trace / step counter.
// Thread t owns coordinates t, t + 1024, ... (ITEMS of them): everything it needs is fetched up front
// in one batch of independent loads, so the kernel pays the L2 latency once.
// CLUSTER = false: one CTA of 1024 threads.  CLUSTER = true: one cluster of 8 CTAs x 256 threads: the kernel is a
// chain of latencies and instruction fetches executed once, so it is spread over 8 SMs, with one coordinate per
// thread (ITEMS = ceil(D / 2048)) and all sums of a phase taken in one barrier round.
template <int ITEMS, bool CLUSTER>
__device__ __forceinline__ void mf_finalize_update_body(const MfTailArgs& t) {
    constexpr int NTH = CLUSTER ? TAIL_CLUSTER * TAIL_CL_THREADS : 1024;   // threads cooperating on the tail
    __shared__ float sm[33];
    __shared__ __align__(16) float smc[2][TAIL_CL_WARPS][4];
    int round = 0;
    // sums of up to three values over all NTH threads, result in every thread
    auto tail_sum3 = [&](float& a, float& b, float& c, int n) {
        if (CLUSTER) { cluster_sum3(a, b, c, smc, round); return; }
        a = block_sum_1024(a, sm);
        if (n > 1) b = block_sum_1024(b, sm);
        if (n > 2) c = block_sum_1024(c, sm);
    };
    auto tail_sync = [&]() { if (CLUSTER) tc::cluster_sync(); else __syncthreads(); };
    tl_min(t.tl, t.tl_s);
    pdl_trigger();
    pdl_wait();
    tl_min(t.tl, t.tl_s + 4);
    const float* __restrict__ acc = t.acc; const int accv = t.accv, M = t.M, objective = t.objective, entropy = t.entropy;
    const float* __restrict__ logp = t.logp; const float* __restrict__ esq = t.esq;
    const int Mloc = t.Mloc, deferred = t.deferred, trace_cap = t.trace_cap;
    float* __restrict__ lam = t.lam; float* __restrict__ grad = t.grad; float* __restrict__ m1 = t.m1;
    float* __restrict__ m2 = t.m2; float* __restrict__ avg = t.avg; float* __restrict__ sc = t.sc;
    float* __restrict__ out = t.out; ObjDeviceState* __restrict__ st = t.st; float* __restrict__ trace = t.trace;
    const UpdArgs a = t.a;
    // one CTA (1024 x 1 or 32 x 32 threads) or a cluster of 8 CTAs x 256
    const int tid = CLUSTER ? (int)tc::cluster_ctarank() * TAIL_CL_THREADS + (int)threadIdx.x
                            : (int)(threadIdx.x + threadIdx.y * blockDim.x);
    const int D = a.D;
    const bool stl = entropy == AVI_ENT_STL || entropy == AVI_ENT_STL_ZEROGRAD;
    const bool need23 = objective == AVI_SCOREGRAD || stl;
    const bool adam = a.rule == AVI_RULE_ADAM, dog = a.rule == AVI_RULE_DOG || a.rule == AVI_RULE_DOWG;
    const bool polyavg = a.averager == AVI_AVG_POLYNOMIAL;
    float v[ITEMS][4], x[ITEMS][2], s1m[ITEMS][2], s2m[ITEMS][2], av[ITEMS][2];
    // Fused exchange (sample sharding), low-latency protocol: every rank PUSHES its partial sums into each peer's
    // receive lane as 8-byte {value, sequence number} words (one NVLink store latency, no fence, no flag round trip),
    // then takes every entry it needs as the sum over the ranks' lanes IN RANK ORDER (own values from `acc`):
    // identical bits on all ranks, no separate all-reduce launch.  Lanes alternate with the parity of the sequence
    // number; a peer can only overwrite a lane at seq + 2 after it consumed my seq + 1 push, which I issue after
    // these reads.  Without an LL area (payload too large) the pull protocol of comm.cu runs here instead.
    const int NR = t.comm.nranks;
    const bool ll = ITEMS <= 2 && NR > 1 && t.comm.ll_cap >= t.acc_len;   // (ITEMS > 2: too many registers)
    unsigned int seq = 0;
    long long soff = 0;
    if (NR > 1) {
        seq = *reinterpret_cast<volatile unsigned int*>(&t.comm.dev->seq) + 1u;
        if (ll) {
            for (long long i = tid; i < t.acc_len; i += NTH) ll_push(t.comm, seq, i, acc[i]);
        } else {
            soff = (long long)(seq & 1u) * t.comm.slot_stride;
            float* mine = t.comm.t.data[t.comm.rank] + soff;
            for (long long i = tid; i < t.acc_len; i += NTH) mine[i] = acc[i];
            __threadfence_system();
            tail_sync();
            if (tid == 0)
                for (int r = 0; r < NR; ++r) st_release_sys(t.comm.t.flags[r] + t.comm.rank, seq);
            if (tid < NR) {
                const unsigned int* f = t.comm.t.flags[t.comm.rank] + tid;
                while ((int)(ld_acquire_sys(f) - seq) < 0) { }
            }
            tail_sync();
        }
    }
    auto acc_at = [&](size_t idx) -> float {
        if (NR <= 1) return acc[idx];
        float s = 0.f;
        for (int r = 0; r < NR; ++r) s += ld_relaxed_sys(t.comm.t.data[r] + soff + idx);
        return s;
    };
    const size_t sbase = 4 * (size_t)accv;
    float c0, c1, c2, c3;
    if constexpr (ITEMS <= 2) { if (ll) {
        // all the entries this thread needs, gathered together: ITEMS x 4 vector entries + the 4 scalars
        constexpr int NG = 4 * ITEMS + 4;
        long long gi[NG]; bool gn[NG]; float go[NG], gv[NG];
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int i = tid + k * NTH;
            const bool ok = i < D;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                gi[4 * k + c] = (long long)c * accv + i;
                gn[4 * k + c] = ok && (c < 2 || need23);
            }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) { gi[4 * ITEMS + c] = (long long)sbase + c; gn[4 * ITEMS + c] = true; }
#pragma unroll
        for (int n = 0; n < NG; ++n) go[n] = gn[n] ? acc[gi[n]] : 0.f;
        ll_gather<NG>(t.comm, seq, gi, gn, go, gv);
#pragma unroll
        for (int k = 0; k < ITEMS; ++k)
#pragma unroll
            for (int c = 0; c < 4; ++c) v[k][c] = gv[4 * k + c];
        c0 = gv[4 * ITEMS]; c1 = gv[4 * ITEMS + 1]; c2 = gv[4 * ITEMS + 2]; c3 = gv[4 * ITEMS + 3];
    } }
    if (!ll) {
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int i = tid + k * NTH;
            const bool ok = i < D;
            v[k][0] = ok ? acc_at(i) : 0.f;
            v[k][1] = ok ? acc_at((size_t)accv + i) : 0.f;
            v[k][2] = ok && need23 ? acc_at(2 * (size_t)accv + i) : 0.f;
            v[k][3] = ok && need23 ? acc_at(3 * (size_t)accv + i) : 0.f;
        }
        c0 = acc_at(sbase); c1 = acc_at(sbase + 1); c2 = acc_at(sbase + 2); c3 = acc_at(sbase + 3);
    }
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int i = tid + k * NTH;
        const bool ok = i < D;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const size_t p = (size_t)h * D + i;
            x[k][h] = ok ? lam[p] : 1.0f;
            s1m[k][h] = ok && (adam || dog) ? m1[p] : 0.f;
            s2m[k][h] = ok && adam ? m2[p] : 0.f;
            av[k][h] = ok && polyavg ? avg[p] : 0.f;
        }
    }
    float sl = 0.f, sq = 0.f;
    if (deferred)
        for (int m = tid; m < Mloc; m += NTH) { sl += logp[m]; sq += esq[m]; }
    const float shift = out[3];
    const int halted = st->halted;
    // (everything the committing thread needs later is requested now, with the rest of the batch)
    const int tp = st->trace_pos, cursor = st->batch_cursor;
    const unsigned long long step_now = st->step;
    const float b1t = sc[SC_B1T], b2t = sc[SC_B2T], t_avg = sc[SC_T], v_old = sc[SC_V], r_old = sc[SC_R];

    MfSums S;
    S.h0 = t.h0;
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) part += (tid + k * NTH < D) ? __logf(x[k][1]) : 0.f;
    tail_sum3(part, sl, sq, deferred ? 3 : 1);
    S.logdet = part;
    tl_min(t.tl, 16);   // (diagnostic) inputs loaded + reductions
    if (deferred) { S.s0 = sl; S.s1 = sq; S.s2 = 0.f; S.s3 = 0.f; }
    else { S.s0 = c0; S.s1 = c1; S.s2 = c2; S.s3 = c3; }
    float value, elbo, shift_next;
    mf_outputs(D, M, objective, entropy, S, shift, value, elbo, shift_next);
    const bool bad = !isfinite(value);
    tl_min(t.tl, 17);   // (diagnostic) all reductions done, value known

    float g[ITEMS][2];
    float dx2 = 0.f, g2 = 0.f;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int i = tid + k * NTH;
        mf_grad_vals(v[k][0], v[k][1], v[k][2], v[k][3], x[k][1], M, objective, entropy, S, g[k][0], g[k][1]);
        if (i < D) {
            grad[i] = g[k][0]; grad[D + i] = g[k][1];
            if (dog) {
                const float d0 = x[k][0] - s1m[k][0], d1 = x[k][1] - s1m[k][1];
                dx2 = fmaf(d0, d0, fmaf(d1, d1, dx2));
                g2 = fmaf(g[k][0], g[k][0], fmaf(g[k][1], g[k][1], g2));
            }
        }
    }
    float eta = a.rule == AVI_RULE_DESCENT ? a.h0 : 0.f, v_new = 0.f, r_new = 0.f;
    if (dog) {
        float unused = 0.f;
        tail_sum3(dx2, g2, unused, 2);
        r_new = fmaxf(sqrtf(dx2), r_old);
        if (a.rule == AVI_RULE_DOG) { v_new = v_old + g2; eta = r_new / sqrtf(v_new); }
        else { const float r2 = r_new * r_new; v_new = v_old + r2 * g2; eta = r2 / sqrtf(v_new); }
    }
    if (!halted && !bad) {
        const float w = (a.avg_param + 1.0f) / (t_avg + a.avg_param);
        const float bc1 = 1.0f / (1.0f - b1t), bc2 = 1.0f / (1.0f - b2t);   // Adam bias corrections, once per thread
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int i = tid + k * NTH;
            if (i >= D) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const size_t p = (size_t)h * D + i;
                float xx = x[k][h], dx;
                if (adam) {
                    const float mt = a.h1 * s1m[k][h] + (1.0f - a.h1) * g[k][h];
                    const float vt = a.h2 * s2m[k][h] + (1.0f - a.h2) * g[k][h] * g[k][h];
                    m1[p] = mt; m2[p] = vt;
                    // mt / (1 - b1^t) / (sqrt(vt / (1 - b2^t)) + eps) * eta with the bias corrections as reciprocals
                    dx = __fdividef(mt * bc1, sqrtf(vt * bc2) + a.h3) * a.h0;
                } else {
                    dx = eta * g[k][h];
                }
                xx -= dx;
                if (h == 1 && a.op != AVI_OP_IDENTITY) {
                    if (a.op == AVI_OP_CLIPSCALE) xx = fmaxf(xx, a.op_param);
                    else xx = xx + (sqrtf(fmaf(xx, xx, 4.0f * eta)) - xx) * 0.5f;
                }
                lam[p] = xx;
                if (polyavg) avg[p] = (1.0f - w) * av[k][h] + w * xx;
            }
        }
    }
    tl_min(t.tl, 18);   // (diagnostic) gradient + update stored
    if (tid == 0 && NR > 1) *reinterpret_cast<volatile unsigned int*>(&t.comm.dev->seq) = seq;
    if (tid == 0 && !halted) {
        out[0] = value; out[1] = elbo; out[2] = S.logdet; out[3] = shift_next;
        if (tp < trace_cap) { trace[2 * tp] = value; trace[2 * tp + 1] = elbo; }
        st->trace_pos = tp + 1;
        if (bad) {
            st->halted = 1;
        } else {
            sc[SC_T] = t_avg + 1.0f;
            sc[SC_ETA] = eta;
            if (adam) { sc[SC_B1T] = b1t * a.h1; sc[SC_B2T] = b2t * a.h2; }
            if (dog) { sc[SC_V] = v_new; sc[SC_R] = r_new; }
            st->step = step_now + 1ull;
            st->batch_cursor = cursor + 1;
        }
    }
    tl_max(t.tl, t.tl_s + 8);
    if (CLUSTER) tc::cluster_sync();   // no CTA may exit while a peer can still read its shared memory
}

