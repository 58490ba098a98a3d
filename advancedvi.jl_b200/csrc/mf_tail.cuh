// The mean-field tail of one iteration -- finalize (sums -> gradient, value, elbo), finiteness check,
// rule + operator + averager, commit -- as a device function, so that it can run as its own one-CTA kernel
// (opt.cu).
#pragma once

#include "avi_internal.cuh"
#include "device_utils.cuh"
#include "comm_dev.cuh"
#include "mf_finalize.cuh"

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
    CommPeers comm;    // comm.nranks > 1: the sample-sharded exchange of `acc` runs inside this kernel
    long long acc_len;
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

// Mean-field tail of one iteration in ONE launch (single CTA): finalize (sums -> gradient, value, elbo)
// + finiteness check + rule + operator + averager + commit of the scalar state / trace / step counter.
// Thread t owns coordinates t, t + 1024, ... (ITEMS of them): everything it needs is fetched up front
// in one batch of independent loads, so the kernel pays the L2 latency once.
template <int ITEMS>
__device__ __forceinline__ void mf_finalize_update_body(const MfTailArgs& t) {
    __shared__ float sm[33];
    pdl_trigger();
    pdl_wait();
    const float* __restrict__ acc = t.acc; const int accv = t.accv, M = t.M, objective = t.objective, entropy = t.entropy;
    const float* __restrict__ logp = t.logp; const float* __restrict__ esq = t.esq;
    const int Mloc = t.Mloc, deferred = t.deferred, trace_cap = t.trace_cap;
    float* __restrict__ lam = t.lam; float* __restrict__ grad = t.grad; float* __restrict__ m1 = t.m1;
    float* __restrict__ m2 = t.m2; float* __restrict__ avg = t.avg; float* __restrict__ sc = t.sc;
    float* __restrict__ out = t.out; ObjDeviceState* __restrict__ st = t.st; float* __restrict__ trace = t.trace;
    const UpdArgs a = t.a;
    const int tid = threadIdx.x + threadIdx.y * blockDim.x;   // the CTA has 1024 threads in either shape
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
            for (long long i = tid; i < t.acc_len; i += 1024) ll_push(t.comm, seq, i, acc[i]);
        } else {
            soff = (long long)(seq & 1u) * t.comm.slot_stride;
            float* mine = t.comm.t.data[t.comm.rank] + soff;
            for (long long i = tid; i < t.acc_len; i += 1024) mine[i] = acc[i];
            __threadfence_system();
            __syncthreads();
            if (tid == 0)
                for (int r = 0; r < NR; ++r) st_release_sys(t.comm.t.flags[r] + t.comm.rank, seq);
            if (tid < NR) {
                const unsigned int* f = t.comm.t.flags[t.comm.rank] + tid;
                while ((int)(ld_acquire_sys(f) - seq) < 0) { }
            }
            __syncthreads();
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
            const int i = tid + k * 1024;
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
            const int i = tid + k * 1024;
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
        const int i = tid + k * 1024;
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
        for (int m = tid; m < Mloc; m += 1024) { sl += logp[m]; sq += esq[m]; }
    const float shift = out[3];
    const int halted = st->halted;
    const float b1t = sc[SC_B1T], b2t = sc[SC_B2T], t_avg = sc[SC_T], v_old = sc[SC_V], r_old = sc[SC_R];

    MfSums S;
    float part = 0.f;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) part += (tid + k * 1024 < D) ? logf(x[k][1]) : 0.f;
    S.logdet = block_sum_1024(part, sm);
    if (deferred) { S.s0 = block_sum_1024(sl, sm); S.s1 = block_sum_1024(sq, sm); S.s2 = 0.f; S.s3 = 0.f; }
    else { S.s0 = c0; S.s1 = c1; S.s2 = c2; S.s3 = c3; }
    float value, elbo, shift_next;
    mf_outputs(D, M, objective, entropy, S, shift, value, elbo, shift_next);
    const bool bad = !isfinite(value);

    float g[ITEMS][2];
    float dx2 = 0.f, g2 = 0.f;
#pragma unroll
    for (int k = 0; k < ITEMS; ++k) {
        const int i = tid + k * 1024;
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
        dx2 = block_sum_1024(dx2, sm); g2 = block_sum_1024(g2, sm);
        r_new = fmaxf(sqrtf(dx2), r_old);
        if (a.rule == AVI_RULE_DOG) { v_new = v_old + g2; eta = r_new / sqrtf(v_new); }
        else { const float r2 = r_new * r_new; v_new = v_old + r2 * g2; eta = r2 / sqrtf(v_new); }
    }
    if (!halted && !bad) {
        const float w = (a.avg_param + 1.0f) / (t_avg + a.avg_param);
#pragma unroll
        for (int k = 0; k < ITEMS; ++k) {
            const int i = tid + k * 1024;
            if (i >= D) continue;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const size_t p = (size_t)h * D + i;
                float xx = x[k][h], dx;
                if (adam) {
                    const float mt = a.h1 * s1m[k][h] + (1.0f - a.h1) * g[k][h];
                    const float vt = a.h2 * s2m[k][h] + (1.0f - a.h2) * g[k][h] * g[k][h];
                    m1[p] = mt; m2[p] = vt;
                    dx = mt / (1.0f - b1t) / (sqrtf(vt / (1.0f - b2t)) + a.h3) * a.h0;
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
    if (tid == 0 && NR > 1) *reinterpret_cast<volatile unsigned int*>(&t.comm.dev->seq) = seq;
    if (tid == 0 && !halted) {
        out[0] = value; out[1] = elbo; out[2] = S.logdet; out[3] = shift_next;
        const int tp = st->trace_pos;
        if (tp < trace_cap) { trace[2 * tp] = value; trace[2 * tp + 1] = elbo; }
        st->trace_pos = tp + 1;
        if (bad) { st->halted = 1; return; }
        sc[SC_T] = t_avg + 1.0f;
        sc[SC_ETA] = eta;
        if (adam) { sc[SC_B1T] = b1t * a.h1; sc[SC_B2T] = b2t * a.h2; }
        if (dog) { sc[SC_V] = v_new; sc[SC_R] = r_new; }
        st->step += 1ull;
        st->batch_cursor += 1;
    }
}

