// K4: the fused SGD step -- Optimisers.update! + operator + averager of
// src/algorithms/common.jl:91-94 with the parameters resident on the device -- and the
// multi-iteration driver that replays captured iterations (graphs of 1 and of AVI_GRAPH_UNROLL = 8 iterations),
// blocking (avi_opt_steps) or without a host round trip (avi_opt_steps_begin / _enqueue / _end).
//   rules      Optimisers.Descent / Adam (third-party), DoG / DoWG  src/optimization/rules.jl:17-64
//   operators  ClipScale  src/optimization/clip_scale.jl:18-29
//              ProximalLocationScaleEntropy  src/optimization/proximal_location_scale_entropy.jl:32-61
//   averaging  PolynomialAveraging / NoAveraging  src/optimization/averaging.jl:42-53
//   non-finite value slot => the step is not applied  src/algorithms/common.jl:83-89
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "avi_internal.cuh"
#include "device_utils.cuh"
#include "fr_finalize.cuh"
#include "fr_sample.cuh"
#include "mf_finalize.cuh"
#include "mf_tail.cuh"
#include "step_fused.cuh"

namespace {

// DoG / DoWG partial norms: part[2b] = sum (x - x0)^2, part[2b+1] = sum g^2 over the CTA's slice
__global__ void __launch_bounds__(256)
k_norms(const float* __restrict__ lam, const float* __restrict__ x0, const float* __restrict__ g, long long P,
        float* __restrict__ part) {
    __shared__ float sm[33];
    float a = 0.f, b = 0.f;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long long)gridDim.x * blockDim.x) {
        float dx = lam[p] - x0[p], gg = g[p];
        a = fmaf(dx, dx, a); b = fmaf(gg, gg, b);
    }
    a = block_sum(a, sm); b = block_sum(b, sm);
    if (threadIdx.x == 0) { part[2 * blockIdx.x] = a; part[2 * blockIdx.x + 1] = b; }
}

__device__ __forceinline__ bool is_scale_diag(long long p, int D, int fullrank) {
    if (p < D) return false;
    long long q = p - D;
    // mean-field: [D, 2 D) is all there is; low-rank: the factor entries behind scale_diag are not scale diagonals
    return fullrank ? (q % (D + 1) == 0) : q < D;
}

// one parameter entry of Optimisers.update! + operator + averager (common.jl:91-94)
__device__ __forceinline__ void update_entry(long long p, float g, float* __restrict__ lam, float* __restrict__ m1,
                                             float* __restrict__ m2, float* __restrict__ avg, const UpdArgs& a,
                                             float eta, float b1t, float b2t, float w) {
    float x = lam[p], dx;
    if (a.rule == AVI_RULE_ADAM) {
        float mt = a.h1 * m1[p] + (1.0f - a.h1) * g;
        float vt = a.h2 * m2[p] + (1.0f - a.h2) * g * g;
        m1[p] = mt; m2[p] = vt;
        dx = mt / (1.0f - b1t) / (sqrtf(vt / (1.0f - b2t)) + a.h3) * a.h0;
    } else {
        dx = eta * g;
    }
    x -= dx;
    if (a.op != AVI_OP_IDENTITY && is_scale_diag(p, a.D, a.fullrank)) {
        if (a.op == AVI_OP_CLIPSCALE) x = fmaxf(x, a.op_param);
        else x = x + (sqrtf(fmaf(x, x, 4.0f * eta)) - x) * 0.5f;
    }
    lam[p] = x;
    if (a.averager == AVI_AVG_POLYNOMIAL) avg[p] = (1.0f - w) * avg[p] + w * x;
}

// COMMIT = the single CTA also commits the scalar state, the trace entry and the step counter;
// otherwise k_commit does it after every CTA has read the old scalars.
template <bool COMMIT>
__global__ void __launch_bounds__(256)
k_update(float* __restrict__ lam, const float* __restrict__ grad, float* __restrict__ m1, float* __restrict__ m2,
         float* __restrict__ avg, float* __restrict__ sc, const float* __restrict__ out,
         ObjDeviceState* __restrict__ st, float* __restrict__ trace, int trace_cap,
         const float* __restrict__ norm_part, UpdArgs a) {
    const float value = out[0];
    const bool bad = !isfinite(value);
    const int halted = st->halted;
    if (halted || bad) {
        if (COMMIT && threadIdx.x == 0 && !halted) {
            int tp = st->trace_pos;
            if (tp < trace_cap) { trace[2 * tp] = value; trace[2 * tp + 1] = out[1]; }
            st->trace_pos = tp + 1;
            st->halted = 1;
        }
        return;
    }
    float eta = 0.f, v_new = 0.f, r_new = 0.f;
    const float b1t = sc[SC_B1T], b2t = sc[SC_B2T], t_avg = sc[SC_T];
    if (a.rule == AVI_RULE_DESCENT) {
        eta = a.h0;
    } else if (a.rule == AVI_RULE_DOG || a.rule == AVI_RULE_DOWG) {
        float dx2 = 0.f, g2 = 0.f;
        for (int q = 0; q < a.nparts; ++q) { dx2 += norm_part[2 * q]; g2 += norm_part[2 * q + 1]; }
        r_new = fmaxf(sqrtf(dx2), sc[SC_R]);
        if (a.rule == AVI_RULE_DOG) { v_new = sc[SC_V] + g2; eta = r_new / sqrtf(v_new); }
        else { float r2 = r_new * r_new; v_new = sc[SC_V] + r2 * g2; eta = r2 / sqrtf(v_new); }
    }
    const float w = (a.avg_param + 1.0f) / (t_avg + a.avg_param);
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < a.P; p += (long long)gridDim.x * blockDim.x) {
        update_entry(p, grad[p], lam, m1, m2, avg, a, eta, b1t, b2t, w);
    }
    if (COMMIT) {
        __syncthreads();
        if (threadIdx.x == 0) {
            sc[SC_T] = t_avg + 1.0f;
            sc[SC_ETA] = eta;
            if (a.rule == AVI_RULE_ADAM) { sc[SC_B1T] = b1t * a.h1; sc[SC_B2T] = b2t * a.h2; }
            if (a.rule == AVI_RULE_DOG || a.rule == AVI_RULE_DOWG) { sc[SC_V] = v_new; sc[SC_R] = r_new; }
            int tp = st->trace_pos;
            if (tp < trace_cap) { trace[2 * tp] = value; trace[2 * tp + 1] = out[1]; }
            st->trace_pos = tp + 1;
            st->step += 1ull;
            st->batch_cursor += 1;
        }
    }
}

// diagnostic: file this iteration's stamps under (step - 1) % 64 (the tail has already advanced the counter)
__global__ void k_tl_commit(const unsigned long long* tl, unsigned long long* hist, const ObjDeviceState* st) {
    hist[((st->step + 63ull) % 64ull) * 32 + threadIdx.x] = tl[threadIdx.x];
}

template <int ITEMS>
__global__ void __launch_bounds__(1024)
k_mf_finalize_update(MfTailArgs t) {
    mf_finalize_update_body<ITEMS, false>(t);
}
// the same tail spread over one thread-block cluster of 8 CTAs (mf_tail.cuh)
template <int ITEMS>
__global__ void __cluster_dims__(TAIL_CLUSTER, 1, 1) __launch_bounds__(TAIL_CL_THREADS)
k_mf_finalize_update_cl(MfTailArgs t) {
    mf_finalize_update_body<ITEMS, true>(t);
}

// Full-rank: gradient of the scale block computed on the fly from the reduced sums (no D x D gradient pass of its
// own) + rule + operator + averager.  Only the lower triangle is touched: the strict upper triangle of L, of its
// gradient and of the Adam moments is identically zero.  Scalars / trace / step are committed by k_commit.
__global__ void __launch_bounds__(256)
k_fr_update(float* __restrict__ lam, float* __restrict__ grad, float* __restrict__ m1, float* __restrict__ m2,
            float* __restrict__ avg, const float* __restrict__ sc, const float* __restrict__ out,
            const ObjDeviceState* __restrict__ st, const float* __restrict__ C1, const float* __restrict__ C2,
            const float* __restrict__ scal, int M, int objective, int entropy, UpdArgs a) {
    // CTA 0: the location block; CTA 1 + j: column j of L, rows i >= j only (contiguous in the column-major layout:
    // coalesced, no index divisions, nothing above the diagonal is touched)
    if (st->halted || !isfinite(out[0])) return;
    const int D = a.D;
    const float eta = a.rule == AVI_RULE_DESCENT ? a.h0 : 0.f;
    const float b1t = sc[SC_B1T], b2t = sc[SC_B2T], w = (a.avg_param + 1.0f) / (sc[SC_T] + a.avg_param);
    const float bc1 = 1.0f / (1.0f - b1t), bc2 = 1.0f / (1.0f - b2t);   // Adam bias corrections as reciprocals
    const int j = (int)blockIdx.x - 1;
    const int i0 = j < 0 ? 0 : j;
    for (int i = (i0 & ~31) + (int)threadIdx.x; i < D; i += 256) {
        if (i < i0) continue;
        size_t p;
        float g;
        if (j < 0) {
            p = (size_t)i;
            g = grad[p];   // location block: written by k_finalize_fr_vec
        } else {
            const size_t idx = (size_t)j * D + i;
            p = (size_t)D + idx;
            g = fr_grad_entry(C1, C2, scal, lam[p], idx, i, j, M, objective, entropy);
            grad[p] = g;
        }
        float x = lam[p], dx;
        if (a.rule == AVI_RULE_ADAM) {
            const float mt = a.h1 * m1[p] + (1.0f - a.h1) * g;
            const float vt = a.h2 * m2[p] + (1.0f - a.h2) * g * g;
            m1[p] = mt; m2[p] = vt;
            dx = __fdividef(mt * bc1, sqrtf(vt * bc2) + a.h3) * a.h0;
        } else {
            dx = eta * g;
        }
        x -= dx;
        if (a.op != AVI_OP_IDENTITY && i == j) {   // diagonal of the scale
            if (a.op == AVI_OP_CLIPSCALE) x = fmaxf(x, a.op_param);
            else x = x + (sqrtf(fmaf(x, x, 4.0f * eta)) - x) * 0.5f;
        }
        lam[p] = x;
        if (a.averager == AVI_AVG_POLYNOMIAL) avg[p] = (1.0f - w) * avg[p] + w * x;
    }
}

// scalar state, trace entry, step counter: what remains of an iteration once every parameter entry is updated
__device__ __forceinline__ void commit_body(float* __restrict__ sc, const float* __restrict__ out,
                                            ObjDeviceState* __restrict__ st, float* __restrict__ trace, int trace_cap,
                                            const float* __restrict__ norm_part, const UpdArgs& a) {
    const float value = out[0];
    if (st->halted) return;
    int tp = st->trace_pos;
    if (tp < trace_cap) { trace[2 * tp] = value; trace[2 * tp + 1] = out[1]; }
    st->trace_pos = tp + 1;
    if (!isfinite(value)) { st->halted = 1; return; }
    float eta = a.rule == AVI_RULE_DESCENT ? a.h0 : 0.f;
    if (a.rule == AVI_RULE_DOG || a.rule == AVI_RULE_DOWG) {
        float dx2 = 0.f, g2 = 0.f;
        for (int q = 0; q < a.nparts; ++q) { dx2 += norm_part[2 * q]; g2 += norm_part[2 * q + 1]; }
        float r_new = fmaxf(sqrtf(dx2), sc[SC_R]), v_new;
        if (a.rule == AVI_RULE_DOG) { v_new = sc[SC_V] + g2; eta = r_new / sqrtf(v_new); }
        else { float r2 = r_new * r_new; v_new = sc[SC_V] + r2 * g2; eta = r2 / sqrtf(v_new); }
        sc[SC_V] = v_new; sc[SC_R] = r_new;
    }
    if (a.rule == AVI_RULE_ADAM) { sc[SC_B1T] *= a.h1; sc[SC_B2T] *= a.h2; }
    sc[SC_T] += 1.0f;
    sc[SC_ETA] = eta;
    st->step += 1ull;
    st->batch_cursor += 1;
}

__global__ void k_commit(float* __restrict__ sc, const float* __restrict__ out, ObjDeviceState* __restrict__ st,
                         float* __restrict__ trace, int trace_cap, const float* __restrict__ norm_part, UpdArgs a) {
    if (threadIdx.x != 0) return;
    commit_body(sc, out, st, trace, trace_cap, norm_part, a);
}

// Full-rank update, tiled: CTA = FRU_ROWS rows x 32 columns of L (entries on or below the diagonal only), a warp-wide
// access runs along a column (contiguous in the column-major layout).  On top of k_fr_update's arithmetic it
//   * rewrites the tile of Lr3 -- the transposed 3xTF32 split of L that the NEXT iteration's z = L eps contraction reads
//     as its A operand (family_fr.cu) -- through a shared-memory transpose, so the sampling stage needs no
//     transpose-and-split pass over the D x D matrix of its own;
//   * commits the scalar state in the last CTA to finish (ticket), instead of a k_commit launch: every CTA has read the
//     old scalars before it takes its ticket.
// The trailing CTAs (blockIdx.x >= ntc * ntr) update the location block.
template <int FRU_ROWS>
__global__ void __launch_bounds__(256)
k_fr_update_t(float* __restrict__ lam, float* __restrict__ grad, float* __restrict__ m1, float* __restrict__ m2,
              float* __restrict__ avg, float* __restrict__ sc, const float* __restrict__ out,
              ObjDeviceState* __restrict__ st, const float* __restrict__ C1, const float* __restrict__ C2,
              const float* __restrict__ scal, int M, int objective, int entropy, UpdArgs a, int ntc, int ntr,
              float* __restrict__ Lr3, int seg, unsigned int* __restrict__ ticket, float* __restrict__ trace,
              int trace_cap, FrDrawAhead da) {
    __shared__ float tile[FRU_ROWS / 32][32][33];
    const int D = a.D;
    const bool live = !(st->halted || !isfinite(out[0]));
    if ((int)blockIdx.x < da.nctas) {
        // ---- the NEXT iteration's eps (two samples per CTA, 128 threads each): eps depends on (key, step, sample,
        // coordinate) only, so it is drawn here, under the shadow of the optimiser-state stream, instead of by a launch
        // of its own at the top of the next iteration.  A step that is not applied keeps its counter: nothing to draw.
        __shared__ float ssum[2][4];
        const int half = threadIdx.x >> 7, t = threadIdx.x & 127;
        const int m = 2 * (int)blockIdx.x + half;
        const unsigned long long step_next = st->step + 1ull;
        float tot = 0.0f;
        if (live && m < da.Mloc) {
            const PhiloxKeys pk(st->key);
            tot = fr_sample_row_part(t, m, da.m0 + m, D, da.ld, seg, (uint32_t)step_next, eps_ctr3(step_next, AVI_STREAM_EPS), pk,
                                     da.E, da.Er3);
        }
        if ((t & 31) == 0) ssum[half][t >> 5] = tot;
        __syncthreads();
        if (live && m < da.Mloc && t == 0) da.esq[m] = ((ssum[half][0] + ssum[half][1]) + ssum[half][2]) + ssum[half][3];
    }
    const float eta = a.rule == AVI_RULE_DESCENT ? a.h0 : 0.f;
    const float b1t = sc[SC_B1T], b2t = sc[SC_B2T], w = (a.avg_param + 1.0f) / (sc[SC_T] + a.avg_param);
    const float bc1 = 1.0f / (1.0f - b1t), bc2 = 1.0f / (1.0f - b2t);   // Adam bias corrections as reciprocals
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int tile_id = (int)blockIdx.x - da.nctas;   // (negative: a sampling CTA, handled above)
    const bool adam = a.rule == AVI_RULE_ADAM, poly = a.averager == AVI_AVG_POLYNOMIAL;
    // one parameter entry of Optimisers.update! + operator + averager on register values (k_fr_update's arithmetic)
    auto entry = [&](float x, float g, float& mt, float& vt, float& av, bool diag) -> float {
        float dx;
        if (adam) {
            mt = a.h1 * mt + (1.0f - a.h1) * g;
            vt = a.h2 * vt + (1.0f - a.h2) * g * g;
            dx = __fdividef(mt * bc1, sqrtf(vt * bc2) + a.h3) * a.h0;
        } else {
            dx = eta * g;
        }
        x -= dx;
        if (a.op != AVI_OP_IDENTITY && diag) {   // diagonal of the scale
            if (a.op == AVI_OP_CLIPSCALE) x = fmaxf(x, a.op_param);
            else x = x + (sqrtf(fmaf(x, x, 4.0f * eta)) - x) * 0.5f;
        }
        if (poly) av = (1.0f - w) * av + w * x;
        return x;
    };
    if (live && tile_id >= 0) {
        if (tile_id >= ntc * ntr) {   // location block (gradient written by the finalize stage)
            const int i = (tile_id - ntc * ntr) * 256 + (int)threadIdx.x;
            if (i < D) {
                float mt = adam ? m1[i] : 0.f, vt = adam ? m2[i] : 0.f, av = poly ? avg[i] : 0.f;
                const float x = entry(lam[i], grad[i], mt, vt, av, false);
                lam[i] = x;
                if (adam) { m1[i] = mt; m2[i] = vt; }
                if (poly) avg[i] = av;
            }
        } else {
            const int j0 = (tile_id % ntc) * 32, i0 = (tile_id / ntc) * FRU_ROWS;
            if (i0 + FRU_ROWS > j0) {   // (tiles entirely above the diagonal hold nothing)
                // every load of the thread's 8 entries is issued before the first dependent instruction: the kernel
                // is a stream over 36 bytes per entry and its duration is the number of exposed memory round trips
                constexpr int NE = (FRU_ROWS / 32) * 4;
                float x[NE], mt[NE], vt[NE], av[NE], c1[NE], c2[NE];
                bool ok[NE];
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const int i = i0 + (e >> 2) * 32 + tx, j = j0 + ty * 4 + (e & 3);
                    ok[e] = i < D && j < D && i >= j;
                    const size_t idx = ok[e] ? (size_t)j * D + i : 0;
                    const size_t p = (size_t)D + idx;
                    x[e] = ok[e] ? lam[p] : 0.f;
                    c1[e] = ok[e] ? C1[idx] : 0.f;
                    c2[e] = ok[e] && objective != AVI_REPGRAD ? C2[idx] : 0.f;
                    mt[e] = ok[e] && adam ? m1[p] : 0.f;
                    vt[e] = ok[e] && adam ? m2[p] : 0.f;
                    av[e] = ok[e] && poly ? avg[p] : 0.f;
                }
                float g[NE];
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const int i = i0 + (e >> 2) * 32 + tx, j = j0 + ty * 4 + (e & 3);
                    g[e] = fr_grad_value(c1[e], c2[e], scal, x[e], i == j, M, objective, entropy);
                    if (ok[e]) x[e] = entry(x[e], g[e], mt[e], vt[e], av[e], i == j);
                }
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    const int i = i0 + (e >> 2) * 32 + tx, j = j0 + ty * 4 + (e & 3);
                    if (ok[e]) {
                        const size_t p = (size_t)D + (size_t)j * D + i;
                        grad[p] = g[e];
                        lam[p] = x[e];
                        if (adam) { m1[p] = mt[e]; m2[p] = vt[e]; }
                        if (poly) avg[p] = av[e];
                    }
                }
                if (Lr3) {
#pragma unroll
                    for (int e = 0; e < NE; ++e) tile[e >> 2][ty * 4 + (e & 3)][tx] = x[e];
                    __syncthreads();
                    // row i of Lr3 holds [hi | hi | lo] of L(i, .) in segments of seg: 32 consecutive columns per warp store
#pragma unroll
                    for (int r = ty; r < FRU_ROWS; r += 8) {
                        const int i = i0 + r, j = j0 + tx;
                        if (i < D && j < D && i >= j) {
                            const float xv = tile[r >> 5][tx][r & 31];
                            const float hi = tc::round_tf32(xv), lo = tc::round_tf32(xv - hi);
                            float* dst = Lr3 + (size_t)i * 3 * seg + j;
                            dst[0] = hi; dst[seg] = hi; dst[2 * (size_t)seg] = lo;
                        }
                    }
                }
            }
        }
    }
    // last CTA out commits (its ticket is taken after this CTA's reads of sc / st above)
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int t = atomicAdd(ticket, 1u);
        if (t == gridDim.x - 1) {
            *ticket = 0u;
            __threadfence();
            commit_body(sc, out, st, trace, trace_cap, nullptr, a);
        }
    }
}

__global__ void k_begin_call(ObjDeviceState* st) { st->trace_pos = 0; st->batch_cursor = 0; st->halted = 0; }

UpdArgs make_args(const avi_opt* op) {
    UpdArgs a{};
    a.rule = op->rule; a.op = op->op; a.averager = op->averager;
    a.h0 = op->hyper[0]; a.h1 = op->hyper[1]; a.h2 = op->hyper[2]; a.h3 = op->hyper[3];
    a.op_param = op->op_param; a.avg_param = op->avg_param;
    a.D = op->obj->D; a.fullrank = op->obj->family == AVI_FULLRANK; a.P = op->P;
    return a;
}

constexpr int NORM_BLOCKS_MAX = 256;

// ... and does it also draw the next iteration's eps?  (Normal(0, 1) base, the one-CTA-per-sample sampler shape)
bool fr_draws_ahead(const avi_opt* op) {
    static const bool on = !(getenv("AVI_FR_DRAW_AHEAD") && atoi(getenv("AVI_FR_DRAW_AHEAD")) == 0);
    const avi_obj* o = op->obj;
    return on && o->base.kind == AVI_BASE_NORMAL && o->Mloc <= 8192;
}

// does the iteration's update kernel maintain the split of L for the next iteration's sampling stage?
bool fr_maintains_split(const avi_opt* op) {
    static const bool on = !(getenv("AVI_FR_TILED_UPDATE") && atoi(getenv("AVI_FR_TILED_UPDATE")) == 0) &&
                           !(getenv("AVI_FR_KEEP_SPLIT") && atoi(getenv("AVI_FR_KEEP_SPLIT")) == 0);
    const avi_obj* o = op->obj;
    return on && o->family == AVI_FULLRANK && op->rule != AVI_RULE_DOG && op->rule != AVI_RULE_DOWG && o->Mloc > 0 &&
           avi_fr_tc_ok(o, o->Mloc);
}

// enqueue ONE iteration of `step` (common.jl:75-104 minus the callback)
int32_t enqueue_iteration(avi_opt* op, bool subsampled, int64_t batch) {
    avi_obj* o = op->obj;
    avi_ctx* ctx = op->ctx;
    if (subsampled) AVI_CHECK(o->model->subsample_dev(op->idx_dev, batch, o->d_state));
    if (ctx->tl) {   // diagnostic: reset the min slots (all ones) and max slots (zero) of this iteration's stamps
        for (int half = 0; half < 2; ++half) {
            AVI_CUDA(ctx, cudaMemsetAsync(ctx->tl + 16 * half, 0xFF, 8 * sizeof(unsigned long long), ctx->stream));
            AVI_CUDA(ctx, cudaMemsetAsync(ctx->tl + 16 * half + 8, 0, 8 * sizeof(unsigned long long), ctx->stream));
        }
    }
    UpdArgs a = make_args(op);
    {   // one launch for the whole iteration when the objective / target / sharding allow it (step_fused.cu)
        StepTail ft{};
        ft.mode = STEP_TAIL_UPDATE;
        ft.lam = op->lam; ft.m1 = op->m1; ft.m2 = op->m2; ft.avg = op->avg; ft.sc = op->sc;
        ft.trace = op->trace; ft.trace_cap = op->trace_cap; ft.a = a; ft.norm_part = op->norm_part;
        bool taken = false;
        AVI_CHECK(avi_objective_fused(o, op->lam, ft, &taken));
        if (taken) {
            if (ctx->tl) k_tl_commit<<<1, 32, 0, ctx->stream>>>(ctx->tl, ctx->tl_hist, o->d_state);
            return AVI_OK;
        }
    }
    MfTailArgs tail{};
    tail.tl = ctx->tl; tail.tl_s = 3;
    tail.acc = o->acc; tail.accv = o->accv; tail.M = o->M; tail.objective = o->objective; tail.entropy = o->entropy;
    tail.logp = o->logp; tail.esq = o->esq; tail.Mloc = o->Mloc; tail.deferred = avi_obj_defers_scalars(o) ? 1 : 0;
    tail.lam = op->lam; tail.grad = o->grad; tail.m1 = op->m1; tail.m2 = op->m2; tail.avg = op->avg; tail.sc = op->sc;
    tail.out = o->out; tail.st = o->d_state; tail.trace = op->trace; tail.trace_cap = op->trace_cap; tail.a = a;
    tail.h0 = o->base.h0;
    // sample sharding over the native NVLink exchange: the tail kernel performs the all-reduce itself
    tail.comm.nranks = 1; tail.acc_len = o->acc_len;
    const bool mf_tail = o->family == AVI_MEANFIELD && o->D <= 8 * 1024;
    o->fused_exchange = mf_tail && o->shard_axis == AVI_SHARD_SAMPLES && ctx->nranks > 1 &&
                        avi_comm_peers(ctx, o->acc_len, &tail.comm);
    if (!o->fused_exchange) tail.comm.nranks = 1;
    // full-rank: the update kernel below keeps the transposed split of L (the sampling stage's A operand) in step with
    // lambda, so the sampling stage skips its own pass over the D x D matrix (steps_launch refreshes it before the first
    // iteration of a call when anything else touched lambda or the buffer)
    const bool dog_rule = op->rule == AVI_RULE_DOG || op->rule == AVI_RULE_DOWG;
    const bool maintain = fr_maintains_split(op);
    o->fr.Lr3_maintained = maintain;
    o->fr.eps_ahead = maintain && fr_draws_ahead(op);   // (... and draws the next iteration's eps: k_fr_update_t)
    int32_t rc_local = avi_objective_local(o, op->lam);
    o->fr.Lr3_maintained = false; o->fr.eps_ahead = false;
    o->fused_exchange = false;
    AVI_CHECK(rc_local);
    // (running this tail inside the last CTA of the target's final kernel was measured SLOWER than its own
    // launch: profiles/README.md)
    if (o->family == AVI_MEANFIELD && o->D <= 8 * 1024) {
        static const bool tail_cluster = !(getenv("AVI_TAIL_CLUSTER") && atoi(getenv("AVI_TAIL_CLUSTER")) == 0);
        const int items = (int)ceil_div(o->D, tail_cluster ? TAIL_CLUSTER * TAIL_CL_THREADS : 1024);
#define LAUNCH_TAIL(IT) (tail_cluster ? avi_launch_pdl(ctx, k_mf_finalize_update_cl<IT>, dim3(TAIL_CLUSTER), dim3(TAIL_CL_THREADS), 0, tail) \
                                      : avi_launch_pdl(ctx, k_mf_finalize_update<IT>, dim3(1), dim3(1024), 0, tail))
        if (items <= 1) LAUNCH_TAIL(1);
        else if (items <= 2) LAUNCH_TAIL(2);
        else if (items <= 4) LAUNCH_TAIL(4);
        else LAUNCH_TAIL(8);
#undef LAUNCH_TAIL
        AVI_LAUNCHED(ctx);
        if (ctx->tl) k_tl_commit<<<1, 32, 0, ctx->stream>>>(ctx->tl, ctx->tl_hist, o->d_state);
        return AVI_OK;
    }
    if (o->family == AVI_FULLRANK && !dog_rule) {
        AVI_CHECK(avi_objective_finalize(o, op->lam, o->grad, o->out, /*skip_fr_matrix=*/true));
        const float* scal = o->acc + 4 * (size_t)o->accv;
        const float* C1 = scal + ACC_NSCAL;
        const float* C2 = C1 + (size_t)o->D * o->D;
        static const bool tiled = !(getenv("AVI_FR_TILED_UPDATE") && atoi(getenv("AVI_FR_TILED_UPDATE")) == 0);
        if (tiled) {
            // gradient of the scale block + rule + operator + averager + next iteration's split of L + commit: one launch
            static const int fru_rows = getenv("AVI_FRU_ROWS") ? atoi(getenv("AVI_FRU_ROWS")) : 32;   // rows per CTA: 32 | 64 | 128 (measured on C3: 88.6 | 90.2 | 93.7 us per step)
            const int rows = fru_rows == 64 || fru_rows == 128 ? fru_rows : 32;
            const int ntc = (int)ceil_div(o->D, 32), ntr = (int)ceil_div(o->D, rows), nloc = (int)ceil_div(o->D, 256);
            auto kern = rows == 32 ? k_fr_update_t<32> : rows == 128 ? k_fr_update_t<128> : k_fr_update_t<64>;
            FrDrawAhead da{};
            if (maintain && fr_draws_ahead(op)) {
                da.nctas = (int)ceil_div(o->Mloc, 2); da.Mloc = o->Mloc; da.m0 = o->m0; da.ld = o->ld;
                da.E = o->E; da.Er3 = o->fr.Er3; da.esq = o->esq;
            }
            kern<<<(unsigned)(da.nctas + ntc * ntr + nloc), 256, 0, ctx->stream>>>(
                op->lam, o->grad, op->m1, op->m2, op->avg, op->sc, o->out, o->d_state, C1, C2, scal, o->M, o->objective,
                o->entropy, a, ntc, ntr, maintain ? o->fr.Lr3 : (float*)nullptr, (int)round_up(o->D, 32), op->ticket,
                op->trace, op->trace_cap, da);
            AVI_LAUNCHED(ctx);
            return AVI_OK;
        }
        k_fr_update<<<(unsigned)o->D + 1u, 256, 0, ctx->stream>>>(op->lam, o->grad, op->m1, op->m2, op->avg, op->sc, o->out, o->d_state, C1,
                                                 C2, scal, o->M, o->objective, o->entropy, a);
        AVI_LAUNCHED(ctx);
        k_commit<<<1, 32, 0, ctx->stream>>>(op->sc, o->out, o->d_state, op->trace, op->trace_cap, op->norm_part, a);
        AVI_LAUNCHED(ctx);
        return AVI_OK;
    }
    AVI_CHECK(avi_objective_finalize(o, op->lam, o->grad, o->out));
    const bool dog = op->rule == AVI_RULE_DOG || op->rule == AVI_RULE_DOWG;
    const int nb_upd = (int)std::min<int64_t>(ceil_div(op->P, 256 * 4), 4 * ctx->prop.multiProcessorCount);
    if (dog) {
        a.nparts = (int)std::min<int64_t>(ceil_div(op->P, 1024), NORM_BLOCKS_MAX);
        k_norms<<<a.nparts, 256, 0, ctx->stream>>>(op->lam, op->m1, o->grad, op->P, op->norm_part);
        AVI_LAUNCHED(ctx);
    }
    if (nb_upd <= 1) {
        k_update<true><<<1, 256, 0, ctx->stream>>>(op->lam, o->grad, op->m1, op->m2, op->avg, op->sc, o->out,
                                                   o->d_state, op->trace, op->trace_cap, op->norm_part, a);
        AVI_LAUNCHED(ctx);
    } else {
        k_update<false><<<nb_upd, 256, 0, ctx->stream>>>(op->lam, o->grad, op->m1, op->m2, op->avg, op->sc, o->out,
                                                         o->d_state, op->trace, op->trace_cap, op->norm_part, a);
        AVI_LAUNCHED(ctx);
        k_commit<<<1, 32, 0, ctx->stream>>>(op->sc, o->out, o->d_state, op->trace, op->trace_cap, op->norm_part, a);
        AVI_LAUNCHED(ctx);
    }
    return AVI_OK;
}

void drop_graph(avi_opt* op) {
    if (op->graph_exec) cudaGraphExecDestroy(op->graph_exec);
    if (op->graph) cudaGraphDestroy(op->graph);
    if (op->graph_u_exec) cudaGraphExecDestroy(op->graph_u_exec);
    if (op->graph_u) cudaGraphDestroy(op->graph_u);
    op->graph_exec = nullptr; op->graph = nullptr; op->graph_u_exec = nullptr; op->graph_u = nullptr;
}

// One avi_opt_steps call = prepare (size the trace, upload minibatch indices, reset the per-call device counters,
// capture the iteration graph if needed) + launch (n graph replays, no host synchronisation) + finish (copy the
// trace back, synchronise once, bookkeeping).  The three phases are also exported separately
// (avi_opt_steps_begin / _enqueue / _end) so that a caller can keep the device queue full.
int32_t steps_prepare(avi_opt* op, int32_t n, const int32_t* idx_host, int64_t batch) {
    avi_obj* o = op->obj;
    avi_ctx* ctx = op->ctx;
    cudaSetDevice(ctx->device);
    const bool subsampled = idx_host != nullptr;
    if (n > op->trace_cap) {
        AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        avi_free(op->trace);
        if (op->h_trace) cudaFreeHost(op->h_trace);
        op->h_trace = nullptr;
        op->trace_cap = std::max(n, 1024);
        AVI_CHECK(avi_alloc(ctx, &op->trace, 2 * (size_t)op->trace_cap));
        AVI_CUDA(ctx, cudaMallocHost(&op->h_trace, 2 * (size_t)op->trace_cap * sizeof(float)));
        drop_graph(op);
    }
    if (subsampled) {
        const int64_t need = (int64_t)n * batch;
        // the gather kernel trusts these indices: reject anything outside [0, rows) here (0-based; a Julia caller's
        // 1:n must be shifted by the glue)
        const int64_t rows = o->model->rows_full();
        if (rows >= 0) {
            if (batch > rows) AVI_FAIL(ctx, AVI_ERR_INVALID, "minibatch larger than the data set");
            int64_t bad = -1;   // (branch-free min / max scan: the per-element early-exit loop cost 4 us per iteration on C5)
            if (avi_check_indices(idx_host, need, rows, &bad) != AVI_OK)
                AVI_FAIL(ctx, AVI_ERR_INVALID, "minibatch index " + std::to_string(idx_host[bad]) + " outside [0, " +
                                                   std::to_string(rows) + ")");
        }
        if (need > op->idx_cap) {
            AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            avi_free(op->idx_dev);
            // (the captured iteration holds this pointer: room for calls of up to 1024 iterations, so that a longer call
            // after a short one does not pay for a re-capture)
            op->idx_cap = std::max<int64_t>(need, 1024 * batch);
            AVI_CHECK(avi_alloc(ctx, &op->idx_dev, (size_t)op->idx_cap));
            drop_graph(op);
        }
        AVI_CUDA(ctx, cudaMemcpyAsync(op->idx_dev, idx_host, need * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    }
    k_begin_call<<<1, 1, 0, ctx->stream>>>(o->d_state);
    AVI_LAUNCHED(ctx);

    static const bool no_graph = getenv("AVI_NO_GRAPH") && atoi(getenv("AVI_NO_GRAPH")) != 0;
    op->use_graph = !no_graph && !ctx->timing && !o->model->needs_sync_eval() && !(ctx->nranks > 1 && !ctx->comm_capturable);
    if (op->use_graph) {
        const int64_t gen = avi_graph_key(o, !subsampled);
        if (op->graph_exec && (op->graph_gen != gen || op->graph_subsampled != subsampled || op->graph_batch != batch))
            drop_graph(op);
        if (!op->graph_exec) {
            // Allocation inside a capture is illegal: run the objective once eagerly so that every
            // lazily sized buffer exists.  It only writes scratch (no optimiser state, no step counter).
            if (subsampled) AVI_CHECK(o->model->subsample_dev(op->idx_dev, batch, o->d_state));
            {
                bool fused = false;   // the single-kernel path sizes its own buffers without launching anything
                StepTail ft{};
                ft.mode = STEP_TAIL_UPDATE;
                AVI_CHECK(avi_objective_fused(o, op->lam, ft, &fused, /*dry_run=*/true));
                if (!fused) AVI_CHECK(avi_objective_local(o, op->lam));
            }
            auto capture = [&](int iters, cudaGraph_t* graph, cudaGraphExec_t* exec) -> int32_t {
                const int64_t launches0 = ctx->launches;
                AVI_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
                ctx->capturing = true;
                int32_t rc = AVI_OK;
                for (int it = 0; it < iters && rc == AVI_OK; ++it) rc = enqueue_iteration(op, subsampled, batch);
                ctx->capturing = false;
                cudaGraph_t g = nullptr;
                cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
                op->graph_launches = (ctx->launches - launches0) / iters;
                ctx->launches = launches0;
                if (rc != AVI_OK) { if (g) cudaGraphDestroy(g); return rc; }
                if (e != cudaSuccess) AVI_FAIL(ctx, AVI_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
                *graph = g;
                AVI_CUDA(ctx, cudaGraphInstantiate(exec, g, 0));
                return AVI_OK;
            };
            AVI_CHECK(capture(1, &op->graph, &op->graph_exec));
            // several iterations per graph launch (AVI_GRAPH_UNROLL, default 8): inside one graph the tail -> sampling
            // edge of consecutive iterations is a programmatic dependent launch too, which saves the ~2.5 us that a
            // graph boundary costs (measured +6 % steps/s on C2).  Full-batch objectives only (a minibatch iteration
            // reads its indices at a per-iteration offset that the device state already tracks, but keep the simple
            // path there).
            static const int unroll = getenv("AVI_GRAPH_UNROLL") ? std::max(1, atoi(getenv("AVI_GRAPH_UNROLL"))) : 8;
            op->graph_unroll = subsampled ? 1 : unroll;
            if (op->graph_unroll > 1) AVI_CHECK(capture(op->graph_unroll, &op->graph_u, &op->graph_u_exec));
            op->graph_gen = avi_graph_key(o, !subsampled);
            op->graph_subsampled = subsampled; op->graph_batch = batch;
        }
    }
    op->call_cap = n; op->call_enqueued = 0; op->call_subsampled = subsampled; op->call_batch = batch;
    return AVI_OK;
}

int32_t steps_launch(avi_opt* op, int32_t n) {
    avi_ctx* ctx = op->ctx;
    if (op->call_enqueued + n > op->call_cap) AVI_FAIL(ctx, AVI_ERR_INVALID, "more iterations enqueued than avi_opt_steps_begin reserved");
    if (n > 0 && fr_maintains_split(op)) {
        // the iterations assume Lr3 == split(L of op->lam): true if the last writer of both was this optimiser's own
        // update kernel; otherwise (first call, warm start, the objective evaluated at another lambda meanwhile,
        // buffers reallocated) rebuild it once, outside the captured iteration
        FrWork& w = op->obj->fr;
        if (w.Lr3_owner != op || w.Lr3_version != op->lam_version || w.owner_generation != op->obj->generation) {
            AVI_CHECK(avi_fr_refresh_split(op->obj, op->lam));
            if (fr_draws_ahead(op)) AVI_CHECK(avi_fr_draw_current(op->obj, op->lam));   // eps of the step about to run
        }
        op->lam_version += n;
        w.Lr3_owner = op; w.Lr3_version = op->lam_version; w.owner_generation = op->obj->generation;
    } else {
        op->lam_version += n;
    }
    if (op->use_graph) {
        int32_t left = n;
        if (op->graph_u_exec)
            for (; left >= op->graph_unroll; left -= op->graph_unroll) AVI_CUDA(ctx, cudaGraphLaunch(op->graph_u_exec, ctx->stream));
        for (; left > 0; --left) AVI_CUDA(ctx, cudaGraphLaunch(op->graph_exec, ctx->stream));
        ctx->launches += (int64_t)n * op->graph_launches;
    } else {
        for (int32_t it = 0; it < n; ++it) AVI_CHECK(enqueue_iteration(op, op->call_subsampled, op->call_batch));
    }
    op->call_enqueued += n;
    return AVI_OK;
}

int32_t steps_finish(avi_opt* op, float* value_host, float* elbo_host, int32_t* n_done) {
    avi_obj* o = op->obj;
    avi_ctx* ctx = op->ctx;
    const int n = op->call_enqueued;
    const bool was_subsampled = op->call_subsampled;
    op->call_cap = 0; op->call_enqueued = 0; op->call_subsampled = false;
    if (n_done) *n_done = 0;
    // AdvancedVI.subsample(prob, batch) returns a NEW problem and never alters `prob` (src/AdvancedVI.jl:303-313):
    // after a subsampled run the target is back on its full data (host-side view switch only; the enqueued work
    // carries its own pointers)
    if (was_subsampled) AVI_CHECK(o->model->subsample(nullptr, 0));
    if (n <= 0) return AVI_OK;
    ObjDeviceState hs{};
    AVI_CUDA(ctx, cudaMemcpyAsync(op->h_trace, op->trace, 2 * (size_t)n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AVI_CUDA(ctx, cudaMemcpyAsync(&hs, o->d_state, sizeof(hs), cudaMemcpyDeviceToHost, ctx->stream));
    AVI_CUDA(ctx, avi_stream_wait(ctx));
    const int recorded = std::min(hs.trace_pos, n);
    const int done = hs.halted ? recorded - 1 : recorded;
    for (int i = 0; i < recorded; ++i) {
        if (value_host) value_host[i] = op->h_trace[2 * i];
        if (elbo_host) elbo_host[i] = op->h_trace[2 * i + 1];
    }
    o->step = hs.step;
    op->iteration += done;
    if (n_done) *n_done = done;
    return AVI_OK;
}

int32_t run_steps(avi_opt* op, int32_t n, const int32_t* idx_host, int64_t batch, float* value_host, float* elbo_host,
                  int32_t* n_done) {
    if (n_done) *n_done = 0;
    if (n <= 0) return AVI_OK;
    if (op->call_cap) AVI_FAIL(op->ctx, AVI_ERR_STATE, "avi_opt_steps_begin is open: call avi_opt_steps_end first");
    AVI_CHECK(steps_prepare(op, n, idx_host, batch));
    static const bool dbg_launch = getenv("AVI_DEBUG_LAUNCH") && atoi(getenv("AVI_DEBUG_LAUNCH")) != 0;
    const auto t0 = std::chrono::steady_clock::now();
    int32_t rc = steps_launch(op, n);
    const auto t1 = std::chrono::steady_clock::now();
    if (rc != AVI_OK) { op->call_cap = 0; op->call_enqueued = 0; return rc; }
    rc = steps_finish(op, value_host, elbo_host, n_done);
    if (dbg_launch) {   // is the host's launch loop or the device the limit?
        const auto t2 = std::chrono::steady_clock::now();
        fprintf(stderr, "[avi_opt_steps] n=%d: host launch loop %.1f us/iteration, launch + wait %.1f us/iteration\n", n,
                std::chrono::duration<double, std::micro>(t1 - t0).count() / n,
                std::chrono::duration<double, std::micro>(t2 - t0).count() / n);
    }
    return rc;
}

}  // namespace

// The element loop of avi_host_update, written so that the host compiler vectorises it (restrict-qualified streams,
// branch-free body per rule) and cloned for AVX2 / AVX-512 hosts: at P = 2050 the scalar loop cost 25 us per call,
// a quarter of the whole estimate_gradient! + update round trip.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
#define AVI_HOST_CLONES __attribute__((target_clones("avx512f", "avx2", "default")))
#else
#define AVI_HOST_CLONES
#endif
AVI_HOST_CLONES
static void host_update_loop(int rule, const float* h, int op_kind, float op_param, int averager, float w, float b1t, float b2t,
                             int64_t P, int64_t scale_offset, float* __restrict__ lambda, const float* __restrict__ grad,
                             float* __restrict__ m1, float* __restrict__ m2, float* __restrict__ lambda_avg) {
    const int64_t clip_from = (op_kind == AVI_OP_CLIPSCALE && scale_offset >= 0) ? scale_offset : P;
    if (rule == AVI_RULE_ADAM) {
        const float c1 = 1.0f / (1.0f - b1t), c2 = 1.0f / (1.0f - b2t);
        const float be1 = h[1], be2 = h[2], eps = h[3], eta = h[0];
        for (int64_t p = 0; p < P; ++p) {
            const float g = grad[p];
            const float mt = be1 * m1[p] + (1.0f - be1) * g;
            const float vt = be2 * m2[p] + (1.0f - be2) * g * g;
            m1[p] = mt; m2[p] = vt;
            lambda[p] -= (mt * c1) / (std::sqrt(vt * c2) + eps) * eta;
        }
    } else {
        const float eta = h[0];
        for (int64_t p = 0; p < P; ++p) lambda[p] -= eta * grad[p];
    }
    for (int64_t p = clip_from; p < P; ++p) lambda[p] = lambda[p] > op_param ? lambda[p] : op_param;
    if (averager == AVI_AVG_POLYNOMIAL)
        for (int64_t p = 0; p < P; ++p) lambda_avg[p] = (1.0f - w) * lambda_avg[p] + w * lambda[p];
}

extern "C" {

// AVI_OK iff every idx[j], j < n, lies in [0, rows); otherwise AVI_ERR_INVALID and *first_bad = the first offending
// position.  Blocks of 4096 entries are scanned with a branch-free min / max (the host compiler vectorises it); only a
// block that fails is walked element by element.
int32_t avi_check_indices(const int32_t* idx, int64_t n, int64_t rows, int64_t* first_bad) {
    if (n < 0 || (n > 0 && !idx)) return AVI_ERR_INVALID;
    if (first_bad) *first_bad = -1;
    const int64_t B = 4096;
    for (int64_t j0 = 0; j0 < n; j0 += B) {
        const int64_t j1 = std::min(n, j0 + B);
        int32_t lo = INT32_MAX, hi = INT32_MIN;
        for (int64_t j = j0; j < j1; ++j) {
            const int32_t v = idx[j];
            lo = v < lo ? v : lo;
            hi = v > hi ? v : hi;
        }
        if (lo < 0 || (int64_t)hi >= rows) {
            for (int64_t j = j0; j < j1; ++j)
                if (idx[j] < 0 || (int64_t)idx[j] >= rows) {
                    if (first_bad) *first_bad = j;
                    return AVI_ERR_INVALID;
                }
        }
    }
    return AVI_OK;
}

int32_t avi_opt_create(avi_obj* obj, int32_t rule, const float* hyper, int32_t n_hyper, int32_t op_kind,
                       float op_param, int32_t averager, float avg_param, const float* lambda0_host, int64_t P,
                       avi_opt** out) {
    if (!obj || !out) return AVI_ERR_INVALID;
    avi_ctx* ctx = obj->ctx;
    *out = nullptr;
    if (!lambda0_host || P != obj->P) AVI_FAIL(ctx, AVI_ERR_INVALID, "lambda0 must have num_params entries");
    if (rule < AVI_RULE_DESCENT || rule > AVI_RULE_DOWG) AVI_FAIL(ctx, AVI_ERR_INVALID, "rule");
    if (op_kind < AVI_OP_IDENTITY || op_kind > AVI_OP_PROXENTROPY) AVI_FAIL(ctx, AVI_ERR_INVALID, "operator");
    if (averager != AVI_AVG_NONE && averager != AVI_AVG_POLYNOMIAL) AVI_FAIL(ctx, AVI_ERR_INVALID, "averager");
    if (op_kind == AVI_OP_PROXENTROPY && obj->family == AVI_LOWRANK)
        AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED,
                 "ProximalLocationScaleEntropy is defined for MvLocationScale only "
                 "(src/optimization/proximal_location_scale_entropy.jl:46-61)");
    if (op_kind == AVI_OP_PROXENTROPY && rule == AVI_RULE_ADAM)
        AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED,
                 "ProximalLocationScaleEntropy only supports Descent, DoG and DoWG "
                 "(src/optimization/proximal_location_scale_entropy.jl:29-44)");
    cudaSetDevice(ctx->device);
    avi_opt* op = new avi_opt();
    op->ctx = ctx; op->obj = obj; op->rule = rule; op->op = op_kind; op->averager = averager;
    op->op_param = op_param; op->avg_param = avg_param; op->P = P;
    // defaults: Descent(0.1); Adam(1e-3, (0.9, 0.999), 1e-8); DoG/DoWG(alpha = 1e-6)
    const float dflt[4][4] = {{0.1f, 0, 0, 0}, {1e-3f, 0.9f, 0.999f, 1e-8f}, {1e-6f, 0, 0, 0}, {1e-6f, 0, 0, 0}};
    for (int i = 0; i < 4; ++i) op->hyper[i] = (hyper && i < n_hyper) ? hyper[i] : dflt[rule][i];
    int32_t rc = avi_alloc(ctx, &op->lam, (size_t)P);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &op->m1, (size_t)P);
    if (rc == AVI_OK && rule == AVI_RULE_ADAM) rc = avi_alloc(ctx, &op->m2, (size_t)P);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &op->avg, (size_t)P);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &op->sc, SC_N);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &op->norm_part, 2 * NORM_BLOCKS_MAX);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &op->ticket, 4);
    if (rc != AVI_OK) { avi_opt_destroy(op); return rc; }
    float sc[SC_N] = {0};
    sc[SC_T] = 1.0f;   // PolynomialAveraging state starts at (x0, 1)  (averaging.jl:42)
    sc[SC_B1T] = op->hyper[1]; sc[SC_B2T] = op->hyper[2];
    if (rule == AVI_RULE_DOG || rule == AVI_RULE_DOWG) {
        double nrm = 0.0;
        for (int64_t p = 0; p < P; ++p) nrm += (double)lambda0_host[p] * lambda0_host[p];
        sc[SC_R] = op->hyper[0] * (1.0f + (float)std::sqrt(nrm));   // rules.jl:21-23, :52-54
    }
    cudaError_t ce = cudaSuccess;
    if (rule == AVI_RULE_DOG || rule == AVI_RULE_DOWG)
        ce = avi_copy(ctx, op->m1, lambda0_host, (size_t)P * sizeof(float), cudaMemcpyHostToDevice);   // x0
    if (ce == cudaSuccess) ce = avi_copy(ctx, op->lam, lambda0_host, (size_t)P * sizeof(float), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = avi_copy(ctx, op->avg, lambda0_host, (size_t)P * sizeof(float), cudaMemcpyHostToDevice);
    if (ce == cudaSuccess) ce = avi_copy(ctx, op->sc, sc, sizeof(sc), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) {
        avi_set_error(ctx, std::string("avi_opt_create: initial state upload: ") + cudaGetErrorString(ce));
        avi_opt_destroy(op);
        return AVI_ERR_CUDA;
    }
    *out = op;
    return AVI_OK;
}

int32_t avi_opt_destroy(avi_opt* op) {
    if (!op) return AVI_OK;
    cudaStreamSynchronize(op->ctx->stream);
    drop_graph(op);
    avi_free(op->lam); avi_free(op->m1); avi_free(op->m2); avi_free(op->avg); avi_free(op->sc); avi_free(op->ticket);
    avi_free(op->norm_part); avi_free(op->trace); avi_free(op->idx_dev);
    if (op->h_trace) cudaFreeHost(op->h_trace);
    delete op;
    return AVI_OK;
}

int32_t avi_opt_steps(avi_opt* op, int32_t n, float* value_host, float* elbo_host, int32_t* n_done) {
    if (!op) return AVI_ERR_INVALID;
    return run_steps(op, n, nullptr, 0, value_host, elbo_host, n_done);
}

int32_t avi_opt_steps_begin(avi_opt* op, int32_t capacity) {
    if (!op) return AVI_ERR_INVALID;
    if (capacity <= 0) AVI_FAIL(op->ctx, AVI_ERR_INVALID, "capacity must be positive");
    if (op->call_cap) AVI_FAIL(op->ctx, AVI_ERR_STATE, "avi_opt_steps_begin is already open");
    return steps_prepare(op, capacity, nullptr, 0);
}

int32_t avi_opt_steps_enqueue(avi_opt* op, int32_t n) {
    if (!op) return AVI_ERR_INVALID;
    if (!op->call_cap) AVI_FAIL(op->ctx, AVI_ERR_STATE, "avi_opt_steps_enqueue without avi_opt_steps_begin");
    if (n <= 0) return AVI_OK;
    return steps_launch(op, n);
}

int32_t avi_opt_steps_end(avi_opt* op, float* value_host, float* elbo_host, int32_t* n_done) {
    if (!op) return AVI_ERR_INVALID;
    if (!op->call_cap) AVI_FAIL(op->ctx, AVI_ERR_STATE, "avi_opt_steps_end without avi_opt_steps_begin");
    return steps_finish(op, value_host, elbo_host, n_done);
}

int32_t avi_opt_steps_subsampled(avi_opt* op, int32_t n, const int32_t* idx_host, int64_t batch, float* value_host,
                                 float* elbo_host, int32_t* n_done) {
    if (!op) return AVI_ERR_INVALID;
    if (!idx_host || batch <= 0) AVI_FAIL(op->ctx, AVI_ERR_INVALID, "minibatch indices missing");
    return run_steps(op, n, idx_host, batch, value_host, elbo_host, n_done);
}

int32_t avi_host_update(int32_t rule, const float* hyper, int32_t n_hyper, int32_t op_kind, float op_param,
                        int32_t averager, float avg_param, int64_t P, int64_t scale_offset, float* lambda,
                        const float* grad, float* m1, float* m2, float* lambda_avg, float* st) {
    if (!lambda || !grad || !st || P <= 0) return AVI_ERR_INVALID;
    if (rule != AVI_RULE_DESCENT && rule != AVI_RULE_ADAM) return AVI_ERR_UNSUPPORTED;
    if (op_kind != AVI_OP_IDENTITY && op_kind != AVI_OP_CLIPSCALE) return AVI_ERR_UNSUPPORTED;
    if (rule == AVI_RULE_ADAM && (!m1 || !m2)) return AVI_ERR_INVALID;
    if (averager == AVI_AVG_POLYNOMIAL && !lambda_avg) return AVI_ERR_INVALID;
    const float dflt[2][4] = {{0.1f, 0, 0, 0}, {1e-3f, 0.9f, 0.999f, 1e-8f}};
    float h[4];
    for (int i = 0; i < 4; ++i) h[i] = (hyper && i < n_hyper) ? hyper[i] : dflt[rule][i];
    if (st[SC_T] == 0.0f) { st[SC_T] = 1.0f; st[SC_B1T] = h[1]; st[SC_B2T] = h[2]; }   // first call (averaging.jl:42)
    const float b1t = st[SC_B1T], b2t = st[SC_B2T];
    const float w = (avg_param + 1.0f) / (st[SC_T] + avg_param);
    host_update_loop(rule, h, op_kind, op_param, averager, w, b1t, b2t, P, scale_offset, lambda, grad, m1, m2, lambda_avg);
    st[SC_T] += 1.0f;
    if (rule == AVI_RULE_ADAM) { st[SC_B1T] = b1t * h[1]; st[SC_B2T] = b2t * h[2]; }
    return AVI_OK;
}

// ---- `step` for callers whose parameters live in HOST memory (common.jl:75-104): estimate_gradient! through the
// zero-copy boundary + Optimisers.update! + operator + averager on the host, one call per iteration.  The handle binds the
// caller-owned arrays once, so the per-iteration call carries two pointers instead of fifteen arguments.
struct avi_hoststep {
    avi_obj* obj;
    int32_t rule, op_kind, averager, n_hyper;
    float hyper[4], op_param, avg_param;
    int64_t P, scale_offset;
    float *lambda, *grad, *lambda_avg;     // caller-owned, P entries each (lambda_avg may be null without averaging)
    std::vector<float> m1, m2;
    float st[SC_N];
    double last_estimate_us, last_update_us;
};

int32_t avi_hoststep_create(avi_obj* obj, int32_t rule, const float* hyper, int32_t n_hyper, int32_t op_kind, float op_param,
                            int32_t averager, float avg_param, int64_t scale_offset, float* lambda, float* grad,
                            float* lambda_avg, avi_hoststep** out) {
    if (!obj || !out) return AVI_ERR_INVALID;
    avi_ctx* ctx = obj->ctx;
    *out = nullptr;
    if (!lambda || !grad) AVI_FAIL(ctx, AVI_ERR_INVALID, "hoststep: lambda and grad buffers are required");
    if (rule != AVI_RULE_DESCENT && rule != AVI_RULE_ADAM) AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED, "hoststep: Descent and Adam only");
    if (op_kind != AVI_OP_IDENTITY && op_kind != AVI_OP_CLIPSCALE) AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED, "hoststep: IdentityOperator and ClipScale only");
    if (averager != AVI_AVG_NONE && averager != AVI_AVG_POLYNOMIAL) AVI_FAIL(ctx, AVI_ERR_INVALID, "averager");
    if (averager == AVI_AVG_POLYNOMIAL && !lambda_avg) AVI_FAIL(ctx, AVI_ERR_INVALID, "hoststep: PolynomialAveraging needs lambda_avg");
    if (n_hyper < 0 || n_hyper > 4) AVI_FAIL(ctx, AVI_ERR_INVALID, "hoststep: at most 4 hyperparameters");
    avi_hoststep* hs = new avi_hoststep();
    hs->obj = obj; hs->rule = rule; hs->op_kind = op_kind; hs->averager = averager; hs->n_hyper = n_hyper;
    for (int i = 0; i < 4; ++i) hs->hyper[i] = (hyper && i < n_hyper) ? hyper[i] : 0.0f;
    hs->op_param = op_param; hs->avg_param = avg_param; hs->P = obj->P; hs->scale_offset = scale_offset;
    hs->lambda = lambda; hs->grad = grad; hs->lambda_avg = lambda_avg;
    hs->m1.assign((size_t)obj->P, 0.0f); hs->m2.assign((size_t)obj->P, 0.0f);
    for (int i = 0; i < SC_N; ++i) hs->st[i] = 0.0f;
    hs->last_estimate_us = hs->last_update_us = 0.0;
    *out = hs;
    return AVI_OK;
}

int32_t avi_hoststep_step(avi_hoststep* hs, float* value, float* elbo) {
    if (!hs) return AVI_ERR_INVALID;
    const auto t0 = std::chrono::steady_clock::now();
    float v = 0.0f, e = 0.0f;
    int32_t rc = avi_obj_estimate_gradient(hs->obj, hs->lambda, hs->P, hs->grad, &v, &e);
    if (rc != AVI_OK) return rc;
    const auto t1 = std::chrono::steady_clock::now();
    if (value) *value = v;
    if (elbo) *elbo = e;
    if (!std::isfinite(v)) return AVI_OK;   // the caller raises (common.jl:83-89); parameters stay untouched
    rc = avi_host_update(hs->rule, hs->hyper, hs->n_hyper, hs->op_kind, hs->op_param, hs->averager, hs->avg_param, hs->P,
                         hs->scale_offset, hs->lambda, hs->grad, hs->m1.data(), hs->m2.data(), hs->lambda_avg, hs->st);
    const auto t2 = std::chrono::steady_clock::now();
    hs->last_estimate_us = std::chrono::duration<double, std::micro>(t1 - t0).count();
    hs->last_update_us = std::chrono::duration<double, std::micro>(t2 - t1).count();
    return rc;
}

int32_t avi_hoststep_timing(const avi_hoststep* hs, double* estimate_us, double* update_us, double* enqueue_us, double* wait_us) {
    if (!hs) return AVI_ERR_INVALID;
    if (estimate_us) *estimate_us = hs->last_estimate_us;
    if (update_us) *update_us = hs->last_update_us;
    if (enqueue_us) *enqueue_us = hs->obj->last_launch_us;
    if (wait_us) *wait_us = hs->obj->last_wait_us;
    return AVI_OK;
}

int32_t avi_hoststep_destroy(avi_hoststep* hs) {
    delete hs;
    return AVI_OK;
}

int32_t avi_opt_get(avi_opt* op, float* lambda_host, float* lambda_avg_host, float* grad_host) {
    if (!op) return AVI_ERR_INVALID;
    avi_ctx* ctx = op->ctx;
    const size_t nb = (size_t)op->P * sizeof(float);
    if (lambda_host) AVI_CUDA(ctx, cudaMemcpyAsync(lambda_host, op->lam, nb, cudaMemcpyDeviceToHost, ctx->stream));
    if (lambda_avg_host)   // NoAveraging: value(avg_st) is the current iterate (averaging.jl:9-13)
        AVI_CUDA(ctx, cudaMemcpyAsync(lambda_avg_host, op->averager == AVI_AVG_POLYNOMIAL ? op->avg : op->lam, nb,
                                      cudaMemcpyDeviceToHost, ctx->stream));
    if (grad_host) AVI_CUDA(ctx, cudaMemcpyAsync(grad_host, op->obj->grad, nb, cudaMemcpyDeviceToHost, ctx->stream));
    AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AVI_OK;
}

int64_t avi_opt_iteration(const avi_opt* op) { return op ? op->iteration : -1; }

// ---- warm start (src/optimize.jl:50, :58-62): [header | lam | m1 | m2 | avg | sc | out] ----------
struct StateHeader {
    uint64_t magic, P;
    int32_t rule, op, averager, family;
    int64_t iteration;
    uint64_t key, step;
};
static const uint64_t STATE_MAGIC = 0x4156493130305354ull;   // "AVI100ST"

int64_t avi_opt_state_nbytes(const avi_opt* op) {
    if (!op) return -1;
    return (int64_t)sizeof(StateHeader) + (int64_t)sizeof(float) * (4 * op->P + SC_N + 4);
}

int32_t avi_opt_state_export(avi_opt* op, void* buf_host, int64_t nbytes) {
    if (!op || !buf_host) return AVI_ERR_INVALID;
    avi_ctx* ctx = op->ctx;
    if (nbytes < avi_opt_state_nbytes(op)) AVI_FAIL(ctx, AVI_ERR_INVALID, "buffer too small");
    AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    StateHeader h{};
    h.magic = STATE_MAGIC; h.P = (uint64_t)op->P; h.rule = op->rule; h.op = op->op; h.averager = op->averager;
    h.family = op->obj->family; h.iteration = op->iteration; h.key = op->obj->key; h.step = op->obj->step;
    char* p = static_cast<char*>(buf_host);
    std::memcpy(p, &h, sizeof(h)); p += sizeof(h);
    const size_t nb = (size_t)op->P * sizeof(float);
    const float* parts[4] = {op->lam, op->m1, op->m2, op->avg};
    for (int i = 0; i < 4; ++i) {
        if (parts[i]) AVI_CUDA(ctx, avi_copy(ctx, p, parts[i], nb, cudaMemcpyDeviceToHost));
        else std::memset(p, 0, nb);
        p += nb;
    }
    AVI_CUDA(ctx, avi_copy(ctx, p, op->sc, SC_N * sizeof(float), cudaMemcpyDeviceToHost)); p += SC_N * sizeof(float);
    AVI_CUDA(ctx, avi_copy(ctx, p, op->obj->out, 4 * sizeof(float), cudaMemcpyDeviceToHost));
    return AVI_OK;
}

int32_t avi_opt_state_import(avi_opt* op, const void* buf_host, int64_t nbytes) {
    if (!op || !buf_host) return AVI_ERR_INVALID;
    avi_ctx* ctx = op->ctx;
    if (nbytes < avi_opt_state_nbytes(op)) AVI_FAIL(ctx, AVI_ERR_INVALID, "buffer too small");
    StateHeader h{};
    const char* p = static_cast<const char*>(buf_host);
    std::memcpy(&h, p, sizeof(h)); p += sizeof(h);
    if (h.magic != STATE_MAGIC || h.P != (uint64_t)op->P || h.rule != op->rule || h.op != op->op ||
        h.averager != op->averager || h.family != op->obj->family)
        AVI_FAIL(ctx, AVI_ERR_STATE, "state does not belong to an optimiser of this configuration");
    AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const size_t nb = (size_t)op->P * sizeof(float);
    float* parts[4] = {op->lam, op->m1, op->m2, op->avg};
    for (int i = 0; i < 4; ++i) {
        if (parts[i]) AVI_CUDA(ctx, avi_copy(ctx, parts[i], p, nb, cudaMemcpyHostToDevice));
        p += nb;
    }
    AVI_CUDA(ctx, avi_copy(ctx, op->sc, p, SC_N * sizeof(float), cudaMemcpyHostToDevice)); p += SC_N * sizeof(float);
    AVI_CUDA(ctx, avi_copy(ctx, op->obj->out, p, 4 * sizeof(float), cudaMemcpyHostToDevice));
    op->iteration = h.iteration;
    op->lam_version += 1;   // (lambda replaced: a maintained split of L is stale)
    return avi_obj_seed(op->obj, h.key, h.step);
}

}  // extern "C"
