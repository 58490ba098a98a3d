// Internal declarations shared by the translation units of libavi_b200.so.
// Not part of the ABI (include/avi.h is).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/avi.h"
#include "base_dist.cuh"

#define AVI_VERSION 100

// ---------------------------------------------------------------------------------------------
// error plumbing
void avi_set_error(const avi_ctx* ctx, const std::string& msg);
#define AVI_FAIL(ctx, code, msg)                                   \
    do {                                                           \
        avi_set_error((ctx), std::string(__func__) + ": " + (msg)); \
        return (code);                                             \
    } while (0)
#define AVI_CUDA(ctx, expr)                                                              \
    do {                                                                                 \
        cudaError_t e__ = (expr);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            avi_set_error((ctx), std::string(__func__) + ": " #expr ": " +               \
                                     cudaGetErrorString(e__));                           \
            return AVI_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)
#define AVI_CHECK(expr)                       \
    do {                                      \
        int32_t s__ = (expr);                 \
        if (s__ != AVI_OK) return s__;        \
    } while (0)
// after a kernel launch: count it and surface launch-configuration errors
#define AVI_LAUNCHED(ctx)                                                        \
    do {                                                                         \
        (ctx)->launches++;                                                       \
        cudaError_t e__ = cudaGetLastError();                                    \
        if (e__ != cudaSuccess) {                                                \
            avi_set_error((ctx), std::string(__func__) + ": kernel launch: " +   \
                                     cudaGetErrorString(e__));                   \
            return AVI_ERR_CUDA;                                                 \
        }                                                                        \
    } while (0)

static inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
static inline int64_t ceil_div(int64_t x, int64_t m) { return (x + m - 1) / m; }

// ---------------------------------------------------------------------------------------------
struct avi_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaDeviceProp prop{};
    mutable std::string err;
    int64_t launches = 0;
    bool capturing = false;   // a CUDA graph capture is open on `stream`
    // exchange step (multi-rank)
    int rank = 0, nranks = 1;
    avi_allreduce_fn ar_fn = nullptr;
    void* ar_user = nullptr;
    // per-kernel device timing for bench.py's roofline (CUDA events around the named hot kernels;
    // forces eager launches while enabled)
    bool timing = false;
    struct KTimer {
        std::string name;
        std::vector<cudaEvent_t> ev;   // start/stop pairs not yet folded in
        double total_ms = 0.0;
        int64_t count = 0;
    };
    std::vector<KTimer> timers;
    // AVI_TIMELINE=1: per-kernel %globaltimer stamps of the fused iteration (device_utils.cuh), history of 64 steps
    unsigned long long* tl = nullptr;
    unsigned long long* tl_hist = nullptr;
    bool comm_capturable = false;   // the exchange is a kernel of ours (comm.cu), safe inside a graph
    void* comm = nullptr;           // struct CommState* (comm.cu)
};

void avi_ktime_mark(avi_ctx* ctx, const char* name);   // records one event; calls come in start/stop pairs
struct AviTimed {
    avi_ctx* c; const char* n;
    AviTimed(avi_ctx* ctx, const char* name) : c(ctx), n(name) { if (c->timing) avi_ktime_mark(c, n); }
    ~AviTimed() { if (c->timing) avi_ktime_mark(c, n); }
};

// Launch with programmatic dependent launch (PDL): the grid may start while its predecessor on the stream is
// still draining; the kernel must execute pdl_wait() (device_utils.cuh) before touching anything the
// predecessor wrote, and pdl_trigger() lets ITS successor start early.  Inside a captured graph this becomes
// a programmatic edge.  AVI_PDL=0 turns it off (plain stream order).
bool avi_pdl_enabled();
// Wait for the ctx stream on the hot blocking calls (estimate_gradient!, avi_opt_steps): polls cudaStreamQuery instead of
// sleeping in cudaStreamSynchronize, which returns ~10 us after the work is done.  AVI_SPIN_SYNC=0: plain synchronise.
cudaError_t avi_stream_wait(avi_ctx* ctx);
template <typename... KArgs, typename... Args>
static inline cudaError_t avi_launch_pdl(avi_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                         Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = avi_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// device allocation with error plumbing (zero-filled)
int32_t avi_dev_alloc(avi_ctx* ctx, void** p, size_t bytes);
template <typename T>
static inline int32_t avi_alloc(avi_ctx* ctx, T** p, size_t count) {
    return avi_dev_alloc(ctx, reinterpret_cast<void**>(p), count * sizeof(T));
}
// blocking copies ordered on the ctx stream (never the legacy stream: see avi_dev_alloc)
static inline cudaError_t avi_copy(avi_ctx* ctx, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind) {
    cudaError_t e = cudaMemcpyAsync(dst, src, bytes, kind, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    return e;
}
template <typename T>
static inline void avi_free(T*& p) {
    if (p) cudaFree(p);
    p = nullptr;
}

// Device-resident step state read by the kernels (so a captured CUDA graph can be replayed
// with a moving step counter / minibatch cursor).
struct ObjDeviceState {
    unsigned long long step;   // optimisation step whose eps is drawn next
    unsigned long long key;    // Philox key
    long long batch_cursor;    // index of the minibatch used by the next subsampled step
    int halted;                // set when the value slot was not finite: later steps are no-ops
    int trace_pos;             // next slot of the per-call (value, elbo) trace
    // Samples drawn AHEAD by the fused iteration kernel (step_fused.cu): the tail phase of iteration t writes z, eps and
    // the tensor-core copy of z for iteration t + 1 (it holds the new lambda and eps does not depend on lambda), so the
    // next launch starts its forward contraction at once.  zt_kind == 1: the sample buffers hold exactly the draws of
    // (zt_step, zt_key) under the current lambda, as zt_nparts per-sample partial sums; any other sampler resets it.
    int zt_kind, zt_nparts, zt_mloc, zt_pad;
    unsigned long long zt_step, zt_key;
};

// A target may ask the mean-field sampling kernel to produce its per-sample preprocessing in the same
// pass over z (kind 1: hierarchical GLM -> TF32-rounded copy of beta for the tensor-core contraction and
// the prior terms of glm_prior.cuh), saving one launch per step.
struct SampleHook {
    int kind = 0;
    int d = 0, variant = 0, include_prior = 1;
    float* Zt = nullptr;
    int zt_ld = 0, zt_seg = 0;   // row pitch of Zt; zt_seg > 0: 3xTF32 split [hi | hi | lo] in segments of zt_seg
    float4* pre = nullptr;
    unsigned long long* tl = nullptr;     // step timeline (diagnostic)
    unsigned long long* zt_owner = nullptr;   // cleared by whoever rewrites the target's sample-derived buffers (step_fused.cu)
    const void* pf_ptr = nullptr;         // static operand of the next kernel to pull into L2 meanwhile (may be null)
    unsigned long long pf_bytes = 0;
    // full-rank sampler: also write the 3xTF32 split of eps, rows [hi | lo | hi] in segments of er_seg (family_fr.cu)
    float* Er3 = nullptr;
    int er_seg = 0;
};

struct FusedStepArgs;   // step_fused.cuh


// Layout shared by every sample-major buffer: row m (one Monte-Carlo sample) holds `ld` floats,
// coordinate i at [m * ld + i]  ==  a D x M column-major matrix with leading dimension ld.
struct avi_model {
    avi_ctx* ctx = nullptr;
    int D = 0;
    int capability = 1;
    int64_t generation = 0;   // bumped whenever device buffers are reallocated (captured graphs go stale)
    virtual ~avi_model() {}
    // identifies the active data view (full data or a minibatch of some size) for captured graphs: switching between
    // views of the same shape keeps pointers and tensor maps, so graphs are keyed on (generation, view_key)
    virtual int64_t view_key() const { return -1; }
    virtual int64_t rows_full() const { return -1; }   // number of data rows minibatch indices may address (-1: n/a)
    // logp[m] = log pi(z_m); G (nullable) = grad log pi(z_m), same layout as Z.
    virtual int32_t eval(const float* Z, int ld, int M, float* logp, float* G) = 0;
    // Fused mean-field path: a1[i] = sum_m G[m][i], a2[i] = sum_m G[m][i] * E[m][i] without
    // materialising G (a1, a2 hold D floats and are overwritten).  Optional.
    virtual bool has_gradsums() const { return false; }
    virtual int32_t eval_gradsums(const float* Z, const float* E, int ld, int M, float* logp, float* a1, float* a2) {
        return AVI_ERR_UNSUPPORTED;
    }
    virtual int32_t subsample(const int32_t* idx_host, int64_t batch) {
        return AVI_OK;   // AdvancedVI.subsample default: identity (src/AdvancedVI.jl:313)
    }
    virtual int32_t set_gemm_mode(int mode) { return AVI_OK; }
    // fills *h and returns true when the next eval / eval_gradsums on these samples may skip its own pass
    // (the promise is bound to these samples: Z pointer and count; any other eval ignores and clears it)
    virtual bool sample_hook(const float* Z, int ld, int M, SampleHook* h) { return false; }
    virtual void clear_hook() {}
    // Device-side minibatch selection for the fused multi-step loop: the rows of iteration k are
    // idx_dev[k * batch .. (k+1) * batch) with k = st->batch_cursor read ON THE DEVICE.
    virtual int32_t subsample_dev(const int32_t* idx_dev, int64_t batch, const ObjDeviceState* st) {
        return AVI_ERR_UNSUPPORTED;
    }
    // restrict the target to the data rows [r0, r0 + nr) (multi-rank row sharding)
    virtual int32_t set_row_shard(int64_t r0, int64_t nr) { return AVI_ERR_UNSUPPORTED; }
    virtual bool needs_sync_eval() const { return false; }   // host callback: not graph-capturable
    // The whole mean-field RepGradELBO iteration (sample -> log-density + gradient sums -> [exchange] -> finalize +
    // update) as ONE kernel launch (step_fused.cuh).  Optional.
    virtual bool fused_step_ok(int Mloc) const { return false; }
    virtual bool fused_host_lambda_ok() const { return false; }   // fused_step can take lambda from mapped pinned host memory
    virtual int32_t set_fused_step(int mode) { return AVI_ERR_UNSUPPORTED; }
    virtual int32_t fused_step(const FusedStepArgs& a) { return AVI_ERR_UNSUPPORTED; }
};

// accumulator layout (floats): 4 vectors of `accv` entries + ACC_NSCAL scalars
//   RepGrad : v0 = sum_m g, v1 = sum_m g*eps, v2 = sum_m eps, v3 = sum_m eps^2
//             s0 = sum_m logp, s1 = sum_m |eps_m|^2
//   ScoreGrad: v0 = sum_m f*eps, v1 = sum_m f*eps^2, v2 = sum_m eps, v3 = sum_m eps^2
//             s0 = sum_m logp, s1 = sum |eps|^2, s2 = sum f, s3 = sum f^2
//   full-rank: v0 = sum_m w (RepGrad) / sum_m f*u (ScoreGrad, u = L^-T eps), v2 = sum_m u,
//              followed by one (RepGrad) or two (ScoreGrad) D x D column-major matrix blocks.
// This vector is what the exchange step all-reduces when the samples are sharded.
enum { ACC_NSCAL = 8 };

// scratch of the tensor-core full-rank contractions (family_fr.cu)
struct FrWork {
    float *Lr3 = nullptr, *Er3 = nullptr, *Et3 = nullptr, *Wt3 = nullptr, *Ut3 = nullptr, *zslab = nullptr;
    size_t Lr3_cap = 0, Er3_cap = 0, Et3_cap = 0, Wt3_cap = 0, Ut3_cap = 0, zslab_cap = 0;
    // true while the optimiser loop's update kernel rewrites Lr3 together with lambda (opt.cu: k_fr_update_t), so the
    // sampling stage does not transpose + split L itself
    bool Lr3_maintained = false;
    const void* Lr3_owner = nullptr;   // the optimiser whose update kernel last wrote Lr3 (nullptr: anyone else did)
    int64_t Lr3_version = -1;          // ... and the avi_opt::lam_version it corresponds to
    int64_t owner_generation = -1;     // ... and the objective's generation then (shard / target / base changes)
    // true while the optimiser loop's update kernel also draws the NEXT iteration's eps (E, Er3, |eps|^2): the sampling
    // stage then starts at the contraction (same validity rules as Lr3)
    bool eps_ahead = false;
};

// what k_fr_outer_prep needs to finish the location block and the value slot in the same launch (family_fr.cu)
struct FrPrepFinalize {
    int on = 0;
    float h0 = AVI_H0;   // entropy of the base distribution
    const float* lambda = nullptr;
    int M = 0, objective = 0, entropy = 0, Mloc = 0, accv = 0;
    float *grad = nullptr, *out = nullptr;
    const float *logp = nullptr, *esq = nullptr;
};

// arguments of the draw-ahead CTAs of the tiled full-rank update kernel (opt.cu); nctas == 0: none
struct FrDrawAhead {
    int nctas = 0, Mloc = 0, m0 = 0, ld = 0;
    float *E = nullptr, *Er3 = nullptr, *esq = nullptr;
};

struct avi_obj {
    double last_launch_us = 0.0, last_wait_us = 0.0;   // estimate_gradient!: host time to enqueue / to see the completion flag
    avi_ctx* ctx = nullptr;
    avi_model* model = nullptr;
    int family = 0, objective = 0, entropy = 0;
    int D = 0, M = 0;         // M = global number of Monte-Carlo samples
    int m0 = 0, Mloc = 0;     // this rank's shard
    int shard_axis = 0;       // AVI_SHARD_*
    bool fused_exchange = false;   // the caller's next kernel performs the sample-shard exchange of acc itself
    int64_t generation = 0;   // bumped when buffers / shard / target change (captured graphs go stale)
    int ld = 0, accv = 0;     // leading dimension of sample-major buffers; padded vector length
    int cap_M = 0;            // sample capacity of the buffers below
    int64_t P = 0;
    int64_t acc_len = 0;
    unsigned long long key = 0, step = 0;   // host mirror of d_state
    // device buffers
    ObjDeviceState* d_state = nullptr;
    float* d_lambda = nullptr;   // P (used by the host-buffer entry points)
    float* Z = nullptr;          // cap_M x ld
    float* E = nullptr;          // cap_M x ld
    float* G = nullptr;          // cap_M x ld
    float* U = nullptr;          // cap_M x ld (full-rank: L^{-T} eps)
    // low-rank family (family_lr.cu): rank, pitch of the factor draws, u_fact draws, [H | dH/dD | dH/dU]
    int rank = 0, ldr = 0;
    float* E2 = nullptr;         // cap_M x ldr
    float* lr_ent = nullptr;     // [H | dH/dD (D) | dH/dU (D * rank) | B^-1 (32 x 32) | log det B | sum log D]
    float* V = nullptr;          // cap_M x ldr : U' w per sample (logq-based estimators; w itself lives in U)
    float* logp = nullptr;       // cap_M
    float* esq = nullptr;        // cap_M : |eps_m|^2
    float* fbuf = nullptr;       // cap_M : ScoreGrad f_m
    float* acc = nullptr;        // acc_len
    float* grad = nullptr;       // P
    float* out = nullptr;        // 4 : value, elbo, logdet, ScoreGrad centring shift
    FrWork fr;
    BaseDist base;               // base distribution of MvLocationScale (base_dist.cuh); Normal(0, 1) by default
    bool fr_vec_done = false;    // the local phase already produced grad[0 .. D) and out[0 .. 3) (full-rank, see family.cu)
    // pinned host staging
    float* h_lambda = nullptr;   // P
    float* h_grad = nullptr;     // P + 8: gradient | value, elbo, logdet, shift | completion flag (u32)
    // captured estimate_gradient! (H2D lambda -> kernels -> D2H gradient): one graph launch per call
    cudaGraph_t eg_graph = nullptr;
    cudaGraphExec_t eg_exec = nullptr;
    int64_t eg_gen = -1;
    int64_t eg_launches = 0;
    int eg_calls = 0;
};

struct avi_opt {
    avi_ctx* ctx = nullptr;
    avi_obj* obj = nullptr;
    int rule = 0, op = 0, averager = 0;
    float hyper[4] = {0, 0, 0, 0};
    float op_param = 0, avg_param = 0;
    int64_t P = 0;
    int64_t iteration = 0;
    // device state
    float* lam = nullptr;    // P current iterate
    float* m1 = nullptr;     // P Adam first moment  | DoG/DoWG x0
    float* m2 = nullptr;     // P Adam second moment
    float* avg = nullptr;    // P averaged iterate
    float* sc = nullptr;     // 16 scalars: see opt.cu
    float* norm_part = nullptr;   // per-CTA partial norms (DoG/DoWG)
    unsigned int* ticket = nullptr;   // last-CTA election of the tiled full-rank update (always left at 0)
    int64_t lam_version = 0;      // host-side count of everything that rewrote lam (iterations enqueued, state imports)
    float* trace = nullptr;       // 2 * trace_cap (value, elbo) per iteration of one call
    int trace_cap = 0;
    float* h_trace = nullptr;     // pinned
    int32_t* idx_dev = nullptr;   // minibatch indices of one avi_opt_steps_subsampled call
    int64_t idx_cap = 0;
    // captured iteration
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    cudaGraph_t graph_u = nullptr;          // graph_unroll iterations in one graph
    cudaGraphExec_t graph_u_exec = nullptr;
    int graph_unroll = 1;
    bool graph_subsampled = false;
    int64_t graph_batch = 0;
    int64_t graph_gen = -1;
    int64_t graph_launches = 0;   // kernels per captured iteration
    // open avi_opt_steps call (begin / enqueue / end)
    int call_cap = 0, call_enqueued = 0;
    bool call_subsampled = false, use_graph = false;
    int64_t call_batch = 0;
};

// validity key of a captured graph: buffers of the objective, buffers of the target and (unless the captured work
// selects its own minibatch view) the active data view
static inline int64_t avi_graph_key(const avi_obj* o, bool with_view) {
    return (o->generation * 1000003 + o->model->generation) * 1000003 + (with_view ? o->model->view_key() : -7);
}

// ---------------------------------------------------------------------------------------------
// entry points implemented across the .cu files (all enqueue on ctx->stream)
int32_t avi_obj_ensure_capacity(avi_obj* o, int M);
// rand(rng, q, M): Z = mu + scale * eps for samples [m0, m0 + Mloc) of the step/key held in *st
// (or in *ov when ov != nullptr, a host value).  E, esq always written.
int32_t avi_family_sample(avi_obj* o, const float* lambda, float* Z, float* E, float* esq, int Mloc,
                          int m0, const ObjDeviceState* st, const ObjDeviceState* ov, const SampleHook* hook = nullptr);
int avi_lr_max_rank();                                                                                 // family_lr.cu
int32_t avi_lr_affine(avi_obj* o, const float* lambda, const float* E1, const float* E2, float* Z, int Mloc);
int32_t avi_lr_entropy(avi_obj* o, const float* lambda);      // -> o->lr_ent
int32_t avi_lr_finalize(avi_obj* o, const float* lambda, float* grad, float* out); // acc, lr_ent -> gradient, value, elbo
bool avi_lr_needs_logq(const avi_obj* o);
int32_t avi_lr_logq(avi_obj* o, const float* lambda, int Mloc, bool forward_only = false);   // w, U'w, log q per sample (+ G += w for RepGrad)
int32_t avi_lr_logq_sums(avi_obj* o, int Mloc);
int32_t avi_axpy(avi_ctx* ctx, const float* x, float* y, int64_t n);                                   // y += x
int32_t avi_rowsq_sum(avi_ctx* ctx, const float* A, const float* B, int ld, int D, int M, float* out);   // out[m] = |A_m + B_m|^2
int32_t avi_colsum_add(avi_ctx* ctx, const float* W, int ld, int Mloc, int D, float* tmp, float* dst);  // dst += column sums
int32_t avi_obj_stage_lambda(avi_obj* o);   // o->h_lambda (pinned, mapped) -> o->d_lambda by a kernel
int32_t avi_objective_local(avi_obj* o, const float* lambda);          // sample + model + reduce -> acc
// The whole iteration as one launch when objective, target and sharding allow it (step_fused.cuh): `tail` carries the
// mode and the optimiser / host-output pointers, the rest is filled here.  *taken = false: nothing was enqueued, use
// the multi-kernel path.
struct StepTail;
// lambda_src != nullptr: lambda has not been staged to the device; the kernel reads it from this mapped pinned host buffer
int32_t avi_objective_fused(avi_obj* o, const float* lambda, const StepTail& tail, bool* taken, bool dry_run = false,
                            const float* lambda_src = nullptr);
// acc -> grad (skip_fr_matrix: leave the D x D block of a full-rank gradient to the caller's fused update)
int32_t avi_objective_finalize(avi_obj* o, const float* lambda, float* grad, float* out, bool skip_fr_matrix = false,
                               bool fuse_advance = false);
// forward-only chunk for estimate_objective: sums_dev = {sum logp, sum |eps|^2, logdet}
int32_t avi_objective_forward_chunk(avi_obj* o, const float* lambda, int m0, int Mc, const ObjDeviceState* ov,
                                    float* sums_dev, bool lowrank_logq = false);
int32_t avi_exchange(avi_ctx* ctx, float* buf, int64_t count);
int64_t avi_comm_capacity(avi_ctx* ctx);   // comm.cu: floats per native exchange, -1 = not connected (callback exchange)
struct CommPeers;
bool avi_comm_peers(avi_ctx* ctx, int64_t count, CommPeers* out);   // comm.cu          // all-reduce (no-op single rank)
int32_t avi_obj_advance(avi_obj* o);
bool avi_obj_defers_scalars(const avi_obj* o);                                    // step += 1 on the device

// generic SIMT fp32 GEMM:  C[a*sc_r + b*sc_c] = alpha * sum_k A[a*sa_r + k*sa_k] * B[b*sb_r + k*sb_k],
// a < Ma, b < Nb, k < K; bounds-checked.
int32_t avi_gemm_simt(avi_ctx* ctx, const float* A, long long sa_r, long long sa_k, const float* B,
                      long long sb_r, long long sb_k, float* C, long long sc_r, long long sc_c, int Ma,
                      int Nb, int K, float alpha, int tri_b = 0);
// U = L^{-T} E for a column-major lower-triangular L (D x D): row m of U solves L' u = e_m.
// (base: the right-hand side is -score(e) of the base distribution, == e for Normal(0, 1))
int32_t avi_trsm_lt(avi_ctx* ctx, const float* L, int D, const float* E, float* U, int ld, int M, const BaseDist& base = BaseDist{});

// full-rank family on the tensor cores (family_fr.cu)
bool avi_fr_tc_ok(const avi_obj* o, int Mloc);
int32_t avi_fr_affine_prepare(avi_obj* o, int Mloc, float** Er3, int* seg);
int32_t avi_fr_refresh_split(avi_obj* o, const float* lambda);   // Lr3 = transposed 3xTF32 split of L(lambda)
int32_t avi_fr_draw_current(avi_obj* o, const float* lambda);    // family.cu: eps (E, Er3, |eps|^2) of the objective's current step
int32_t avi_fr_affine_tc(avi_obj* o, const float* lambda, const float* E, float* Z, int Mloc, bool er3_done = false,
                         const SampleHook* hook = nullptr);
int32_t avi_fr_outer_tc(avi_obj* o, const float* W, const float* E, float* C, int Mloc, int which, bool reuse_E,
                        float* colsum = nullptr, const FrPrepFinalize* pf = nullptr);
void avi_fr_free(avi_obj* o);

// models
int32_t avi_model_mvnormal_diag_make(avi_ctx* ctx, const float* mu, const float* sigma, int D, avi_model** out);
int32_t avi_model_glm_make(avi_ctx* ctx, const float* X, const float* y, int64_t n, int d, int64_t n_data,
                           int likelihood, int variant, int gemm_mode, avi_model** out);
int32_t avi_model_hostcallback_make(avi_ctx* ctx, int D, int capability, avi_logdensity_fn cb, void* user,
                                    avi_model** out);
