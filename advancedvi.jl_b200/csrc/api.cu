// extern "C" entry points of libavi_b200.so (include/avi.h): lifecycle, targets, objective.
// The fused step lives in opt.cu.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <atomic>
#include <cstring>
#include <mutex>
#include <vector>

#include "avi_internal.cuh"
#include "device_utils.cuh"
#include "step_fused.cuh"

static thread_local std::string g_create_error;

void avi_set_error(const avi_ctx* ctx, const std::string& msg) {
    if (ctx) ctx->err = msg;
    else g_create_error = msg;
}

int32_t avi_dev_alloc(avi_ctx* ctx, void** p, size_t bytes) {
    *p = nullptr;
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    // zero-fill ON THE CTX STREAM: a legacy-stream cudaMemset is not ordered against our
    // non-blocking stream and could land after kernels that already wrote the buffer
    if (e == cudaSuccess) e = ctx ? cudaMemsetAsync(*p, 0, bytes, ctx->stream) : cudaMemset(*p, 0, bytes);
    if (e != cudaSuccess) {
        avi_set_error(ctx, std::string("device allocation of ") + std::to_string(bytes) + " bytes failed: " +
                               cudaGetErrorString(e));
        if (*p) cudaFree(*p);
        *p = nullptr;
        return AVI_ERR_CUDA;
    }
    return AVI_OK;
}

void avi_ktime_mark(avi_ctx* ctx, const char* name) {
    avi_ctx::KTimer* t = nullptr;
    for (auto& k : ctx->timers)
        if (k.name == name) t = &k;
    if (!t) { ctx->timers.push_back(avi_ctx::KTimer{}); t = &ctx->timers.back(); t->name = name; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, ctx->stream);
    t->ev.push_back(e);
}

cudaError_t avi_stream_wait(avi_ctx* ctx) {
    static const bool spin = !(getenv("AVI_SPIN_SYNC") && atoi(getenv("AVI_SPIN_SYNC")) == 0);
    if (!spin) return cudaStreamSynchronize(ctx->stream);
    cudaError_t e;
    while ((e = cudaStreamQuery(ctx->stream)) == cudaErrorNotReady) {
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    return e;
}

bool avi_pdl_enabled() {
    static const bool on = !(getenv("AVI_PDL") && atoi(getenv("AVI_PDL")) == 0);
    return on;
}

static void ktime_fold(avi_ctx* ctx) {
    cudaStreamSynchronize(ctx->stream);
    for (auto& k : ctx->timers) {
        for (size_t i = 0; i + 1 < k.ev.size(); i += 2) {
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, k.ev[i], k.ev[i + 1]) == cudaSuccess) { k.total_ms += ms; k.count++; }
        }
        for (auto e : k.ev) cudaEventDestroy(e);
        k.ev.clear();
    }
}

int32_t avi_glm_set_data_shard(avi_model* model, int32_t nshards, int64_t rows_global, int32_t include_prior);
int32_t avi_comm_exchange(avi_ctx* ctx, float* buf, int64_t count);   // comm.cu
void avi_comm_destroy(avi_ctx* ctx);

extern "C" {

int32_t avi_version(void) { return AVI_VERSION; }

const char* avi_last_error(const avi_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int32_t avi_ctx_create(int32_t device, avi_ctx** out) {
    if (!out) return AVI_ERR_INVALID;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        avi_set_error(nullptr, std::string("avi_ctx_create: no CUDA device (this library has no CPU fallback): ") +
                                   cudaGetErrorString(e));
        return AVI_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        avi_set_error(nullptr, "avi_ctx_create: device index out of range");
        return AVI_ERR_INVALID;
    }
    avi_ctx* ctx = new avi_ctx();
    ctx->device = device;
    e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&ctx->prop, device);
    if (e == cudaSuccess && ctx->prop.major != 10) {
        avi_set_error(nullptr, "avi_ctx_create: built for sm_100a (B200); found compute capability " +
                                   std::to_string(ctx->prop.major) + "." + std::to_string(ctx->prop.minor));
        delete ctx;
        return AVI_ERR_UNSUPPORTED;
    }
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        avi_set_error(nullptr, std::string("avi_ctx_create: ") + cudaGetErrorString(e));
        delete ctx;
        return AVI_ERR_CUDA;
    }
    if (getenv("AVI_TIMELINE") && atoi(getenv("AVI_TIMELINE")) != 0) {
        avi_dev_alloc(ctx, reinterpret_cast<void**>(&ctx->tl), 32 * sizeof(unsigned long long));
        avi_dev_alloc(ctx, reinterpret_cast<void**>(&ctx->tl_hist), 64 * 32 * sizeof(unsigned long long));
    }
    *out = ctx;
    return AVI_OK;
}

int32_t avi_ctx_timeline_get(avi_ctx* ctx, uint64_t* hist_host) {
    if (!ctx || !hist_host) return AVI_ERR_INVALID;
    if (!ctx->tl_hist) AVI_FAIL(ctx, AVI_ERR_STATE, "timeline disabled (set AVI_TIMELINE=1 before avi_ctx_create)");
    AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    AVI_CUDA(ctx, cudaMemcpy(hist_host, ctx->tl_hist, 64 * 32 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return AVI_OK;
}

int32_t avi_ctx_destroy(avi_ctx* ctx) {
    if (!ctx) return AVI_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    avi_comm_destroy(ctx);
    if (ctx->tl) cudaFree(ctx->tl);
    if (ctx->tl_hist) cudaFree(ctx->tl_hist);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
    return AVI_OK;
}

int32_t avi_ctx_synchronize(avi_ctx* ctx) {
    if (!ctx) return AVI_ERR_INVALID;
    AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AVI_OK;
}

int32_t avi_ctx_info(avi_ctx* ctx, int32_t* sm_count, int64_t* hbm_bytes, int32_t* cc_major, int32_t* cc_minor) {
    if (!ctx) return AVI_ERR_INVALID;
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (hbm_bytes) *hbm_bytes = (int64_t)ctx->prop.totalGlobalMem;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    return AVI_OK;
}

int64_t avi_ctx_launch_count(const avi_ctx* ctx) { return ctx ? ctx->launches : 0; }

void* avi_ctx_stream(avi_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int32_t avi_ctx_timing(avi_ctx* ctx, int32_t enable) {
    if (!ctx) return AVI_ERR_INVALID;
    if (!enable && ctx->timing) ktime_fold(ctx);
    if (enable && !ctx->timing) ctx->timers.clear();
    ctx->timing = enable != 0;
    return AVI_OK;
}

int32_t avi_ctx_timing_get(avi_ctx* ctx, const char* name, double* total_ms, int64_t* count) {
    if (!ctx || !name) return AVI_ERR_INVALID;
    if (ctx->timing) ktime_fold(ctx);
    for (auto& k : ctx->timers)
        if (k.name == name) {
            if (total_ms) *total_ms = k.total_ms;
            if (count) *count = k.count;
            return AVI_OK;
        }
    if (total_ms) *total_ms = 0.0;
    if (count) *count = 0;
    return AVI_OK;
}

int32_t avi_ctx_set_allreduce(avi_ctx* ctx, avi_allreduce_fn fn, void* user, int32_t rank, int32_t nranks) {
    if (!ctx || nranks < 1 || rank < 0 || rank >= nranks) return AVI_ERR_INVALID;
    ctx->ar_fn = fn; ctx->ar_user = user; ctx->rank = rank; ctx->nranks = nranks;
    return AVI_OK;
}

// ---- targets ----------------------------------------------------------------------------------
int32_t avi_model_mvnormal_diag_create(avi_ctx* ctx, const float* mu_host, const float* sigma_host, int32_t D,
                                       avi_model** out) {
    if (!ctx || !out) return AVI_ERR_INVALID;
    cudaSetDevice(ctx->device);
    return avi_model_mvnormal_diag_make(ctx, mu_host, sigma_host, D, out);
}

int32_t avi_model_glm_create(avi_ctx* ctx, const float* X_host, const float* y_host, int64_t n, int32_t d,
                             int64_t n_data, int32_t likelihood, int32_t variant, int32_t gemm_mode,
                             avi_model** out) {
    if (!ctx || !out) return AVI_ERR_INVALID;
    cudaSetDevice(ctx->device);
    return avi_model_glm_make(ctx, X_host, y_host, n, d, n_data, likelihood, variant, gemm_mode, out);
}

int32_t avi_model_hostcallback_create(avi_ctx* ctx, int32_t D, int32_t capability, avi_logdensity_fn cb, void* user,
                                      avi_model** out) {
    if (!ctx || !out) return AVI_ERR_INVALID;
    return avi_model_hostcallback_make(ctx, D, capability, cb, user, out);
}

int32_t avi_model_subsample(avi_model* model, const int32_t* idx_host, int64_t batch) {
    if (!model) return AVI_ERR_INVALID;
    return model->subsample(idx_host, batch);
}

int32_t avi_model_set_data_shard(avi_model* model, int32_t nshards, int64_t rows_global, int32_t include_prior) {
    if (!model) return AVI_ERR_INVALID;
    int32_t rc = avi_glm_set_data_shard(model, nshards, rows_global, include_prior);
    if (rc != AVI_OK) avi_set_error(model->ctx, "avi_model_set_data_shard: unsupported target or bad arguments");
    return rc;
}

int32_t avi_model_dimension(const avi_model* model) { return model ? model->D : -1; }
int32_t avi_model_capability(const avi_model* model) { return model ? model->capability : -1; }
int32_t avi_model_set_gemm_mode(avi_model* model, int32_t gemm_mode) {
    return model ? model->set_gemm_mode(gemm_mode) : AVI_ERR_INVALID;
}
int32_t avi_model_set_fused_step(avi_model* model, int32_t mode) {
    if (!model) return AVI_ERR_INVALID;
    const int32_t rc = model->set_fused_step(mode);
    if (rc != AVI_OK) avi_set_error(model->ctx, "avi_model_set_fused_step: unsupported target or mode");
    return rc;
}

int32_t avi_model_logdensity(avi_model* model, const float* Z_dev, int32_t ldz, int32_t M, float* logp_dev) {
    if (!model || !Z_dev || !logp_dev || ldz < model->D || (ldz % 4)) return AVI_ERR_INVALID;
    model->clear_hook();
    return model->eval(Z_dev, ldz, M, logp_dev, nullptr);
}

int32_t avi_model_logdensity_and_gradient(avi_model* model, const float* Z_dev, int32_t ldz, int32_t M,
                                          float* logp_dev, float* G_dev) {
    if (!model || !Z_dev || !logp_dev || ldz < model->D || (ldz % 4)) return AVI_ERR_INVALID;
    model->clear_hook();
    return model->eval(Z_dev, ldz, M, logp_dev, G_dev);
}

int32_t avi_model_logdensity_and_gradient_host(avi_model* model, const float* Z_host, int32_t M, float* logp_host,
                                               float* G_host) {
    if (!model || !Z_host || !logp_host || M <= 0) return AVI_ERR_INVALID;
    avi_ctx* ctx = model->ctx;
    const int D = model->D, ld = (int)round_up(D, 4);
    float *Z = nullptr, *G = nullptr, *lp = nullptr;
    int32_t rc = avi_alloc(ctx, &Z, (size_t)M * ld);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &lp, (size_t)M);
    if (rc == AVI_OK && G_host) rc = avi_alloc(ctx, &G, (size_t)M * ld);
    if (rc == AVI_OK) {
        cudaError_t e = cudaMemcpy2DAsync(Z, ld * sizeof(float), Z_host, D * sizeof(float), D * sizeof(float), M,
                                          cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { avi_set_error(ctx, cudaGetErrorString(e)); rc = AVI_ERR_CUDA; }
    }
    model->clear_hook();
    if (rc == AVI_OK) rc = model->eval(Z, ld, M, lp, G);
    if (rc == AVI_OK) {
        cudaError_t e = cudaMemcpyAsync(logp_host, lp, M * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess && G_host)
            e = cudaMemcpy2DAsync(G_host, D * sizeof(float), G, ld * sizeof(float), D * sizeof(float), M,
                                  cudaMemcpyDeviceToHost, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { avi_set_error(ctx, cudaGetErrorString(e)); rc = AVI_ERR_CUDA; }
    }
    avi_free(Z); avi_free(G); avi_free(lp);
    return rc;
}

int32_t avi_model_destroy(avi_model* model) {
    if (!model) return AVI_OK;
    cudaStreamSynchronize(model->ctx->stream);
    delete model;
    return AVI_OK;
}

// ---- objective ---------------------------------------------------------------------------------
static void obj_free_buffers(avi_obj* o) {
    avi_free(o->Z); avi_free(o->E); avi_free(o->G); avi_free(o->U); avi_free(o->E2); avi_free(o->V);
    o->logp = nullptr;   // (lives behind G)
    avi_free(o->esq); avi_free(o->fbuf);
}

}  // extern "C"

int32_t avi_obj_ensure_capacity(avi_obj* o, int M) {
    if (M <= o->cap_M) return AVI_OK;
    avi_ctx* ctx = o->ctx;
    AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    obj_free_buffers(o);
    o->generation++;
    o->cap_M = M;
    const size_t n = (size_t)M * o->ld;
    AVI_CHECK(avi_alloc(ctx, &o->Z, n));
    AVI_CHECK(avi_alloc(ctx, &o->E, n));
    // log pi(z_m) sits right behind the gradient block: under row sharding both are partial sums over the data rows and
    // travel in ONE exchange (family.cu)
    AVI_CHECK(avi_alloc(ctx, &o->G, n + (size_t)round_up(M, 4)));
    o->logp = o->G + n;
    if (o->family == AVI_FULLRANK) AVI_CHECK(avi_alloc(ctx, &o->U, n));
    if (o->family == AVI_LOWRANK) {
        AVI_CHECK(avi_alloc(ctx, &o->E2, (size_t)M * o->ldr));
        // w = Sigma^-1 (z - mu) and U'w per sample: the log q based estimators, and estimate_objective with any entropy
        AVI_CHECK(avi_alloc(ctx, &o->U, n));
        AVI_CHECK(avi_alloc(ctx, &o->V, (size_t)M * o->ldr));
    }
    AVI_CHECK(avi_alloc(ctx, &o->esq, (size_t)M));
    AVI_CHECK(avi_alloc(ctx, &o->fbuf, (size_t)M));
    return AVI_OK;
}

extern "C" {

static int32_t obj_push_state(avi_obj* o) {
    ObjDeviceState s{};
    s.step = o->step; s.key = o->key; s.batch_cursor = 0; s.halted = 0; s.trace_pos = 0;
    AVI_CUDA(o->ctx, cudaMemcpyAsync(o->d_state, &s, sizeof(s), cudaMemcpyHostToDevice, o->ctx->stream));
    AVI_CUDA(o->ctx, cudaStreamSynchronize(o->ctx->stream));
    return AVI_OK;
}

static int32_t obj_create(avi_ctx* ctx, avi_model* model, int32_t family, int32_t rank, int32_t objective,
                          int32_t entropy, int32_t M, avi_obj** out);

int32_t avi_obj_create(avi_ctx* ctx, avi_model* model, int32_t family, int32_t objective, int32_t entropy,
                       int32_t M, avi_obj** out) {
    if (!ctx || !model || !out) return AVI_ERR_INVALID;
    *out = nullptr;
    if (family != AVI_MEANFIELD && family != AVI_FULLRANK)
        AVI_FAIL(ctx, AVI_ERR_INVALID, "family (the low-rank family is created with avi_obj_create_lowrank)");
    return obj_create(ctx, model, family, 0, objective, entropy, M, out);
}

int32_t avi_obj_create_lowrank(avi_ctx* ctx, avi_model* model, int32_t rank, int32_t objective, int32_t entropy,
                               int32_t M, avi_obj** out) {
    if (!ctx || !model || !out) return AVI_ERR_INVALID;
    *out = nullptr;
    if (rank < 1 || rank > avi_lr_max_rank()) AVI_FAIL(ctx, AVI_ERR_INVALID, "rank must be in 1..32");
    if (ctx->nranks > 1 && !(objective == AVI_REPGRAD && (entropy == AVI_ENT_CLOSEDFORM || entropy == AVI_ENT_CLOSEDFORM_ZEROGRAD)))
        AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED, "the low-rank family runs its log q based estimators on one rank only");
    return obj_create(ctx, model, AVI_LOWRANK, rank, objective, entropy, M, out);
}

static int32_t obj_create(avi_ctx* ctx, avi_model* model, int32_t family, int32_t rank, int32_t objective,
                          int32_t entropy, int32_t M, avi_obj** out) {
    if (objective != AVI_REPGRAD && objective != AVI_SCOREGRAD) AVI_FAIL(ctx, AVI_ERR_INVALID, "objective");
    if (entropy < AVI_ENT_CLOSEDFORM || entropy > AVI_ENT_STL_ZEROGRAD) AVI_FAIL(ctx, AVI_ERR_INVALID, "entropy");
    if (M < 1) AVI_FAIL(ctx, AVI_ERR_INVALID, "n_samples must be >= 1");
    if (objective == AVI_REPGRAD && model->capability < 1)
        AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED,
                 "RepGradELBO needs a target with first-order capability (logdensity_and_gradient): the native path "
                 "has no AD backend to differentiate through logdensity");
    cudaSetDevice(ctx->device);
    avi_obj* o = new avi_obj();
    o->ctx = ctx; o->model = model; o->family = family; o->objective = objective; o->entropy = entropy;
    o->D = model->D; o->M = M; o->m0 = 0; o->Mloc = M;
    o->ld = (int)round_up(o->D, 4);
    o->accv = (int)round_up(o->D, 32);
    o->rank = rank; o->ldr = (int)round_up(std::max(rank, 1), 4);
    o->P = family == AVI_MEANFIELD ? 2LL * o->D
           : family == AVI_FULLRANK ? (int64_t)o->D + (int64_t)o->D * o->D
                                    : 2LL * o->D + (int64_t)o->D * rank;
    // payload of the exchange: vector sums | scalars | full-rank: two D x D blocks / low-rank: sum_m g u_fact' (D x r)
    o->acc_len = 4LL * o->accv + ACC_NSCAL +
                 (family == AVI_FULLRANK ? 2LL * o->D * o->D : family == AVI_LOWRANK ? 2LL * o->D * rank : 0);
    int32_t rc = avi_alloc(ctx, &o->d_state, 1);
    if (rc == AVI_OK && family == AVI_LOWRANK) rc = avi_alloc(ctx, &o->lr_ent, 1 + (size_t)o->D + (size_t)o->D * rank + 32 * 32 + 2);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &o->d_lambda, (size_t)o->P);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &o->acc, (size_t)o->acc_len);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &o->grad, (size_t)o->P + 4);   // + {value, elbo, logdet, shift} (estimate_gradient!)
    if (rc == AVI_OK) rc = avi_alloc(ctx, &o->out, 4);
    if (rc == AVI_OK) rc = avi_obj_ensure_capacity(o, M);
    if (rc == AVI_OK) {
        cudaError_t e = cudaMallocHost(&o->h_lambda, (size_t)o->P * sizeof(float));
        if (e == cudaSuccess) e = cudaMallocHost(&o->h_grad, ((size_t)o->P + 8) * sizeof(float));
        if (e == cudaSuccess) std::memset(o->h_grad, 0, ((size_t)o->P + 8) * sizeof(float));
        if (e != cudaSuccess) { avi_set_error(ctx, std::string("pinned allocation: ") + cudaGetErrorString(e)); rc = AVI_ERR_CUDA; }
    }
    if (rc == AVI_OK) rc = obj_push_state(o);
    if (rc != AVI_OK) { avi_obj_destroy(o); return rc; }
    *out = o;
    return AVI_OK;
}

int32_t avi_obj_destroy(avi_obj* o) {
    if (!o) return AVI_OK;
    cudaStreamSynchronize(o->ctx->stream);
    if (o->eg_exec) cudaGraphExecDestroy(o->eg_exec);
    if (o->eg_graph) cudaGraphDestroy(o->eg_graph);
    obj_free_buffers(o);
    avi_fr_free(o);
    avi_free(o->d_state); avi_free(o->d_lambda); avi_free(o->acc); avi_free(o->grad); avi_free(o->out);
    avi_free(o->lr_ent);
    if (o->h_lambda) cudaFreeHost(o->h_lambda);
    if (o->h_grad) cudaFreeHost(o->h_grad);
    delete o;
    return AVI_OK;
}

int32_t avi_obj_set_base(avi_obj* obj, int32_t base, float param) {
    if (!obj) return AVI_ERR_INVALID;
    avi_ctx* ctx = obj->ctx;
    if (obj->family == AVI_LOWRANK && base != AVI_BASE_NORMAL)
        AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED, "the low-rank family is Gaussian (src/families/location_scale_low_rank.jl:119-135)");
    BaseDist b;
    if (!avi_base_make(base, param, &b))
        AVI_FAIL(ctx, AVI_ERR_INVALID, "base distribution: AVI_BASE_NORMAL, AVI_BASE_LAPLACE or AVI_BASE_STUDENT_T with 0 < nu < 1e6");
    if (b.kind != obj->base.kind || b.nu != obj->base.nu) obj->generation++;   // captured iterations hold the old draws
    obj->base = b;
    return AVI_OK;
}

int32_t avi_base_constants(int32_t base, float param, float* entropy, float* log_normaliser) {
    BaseDist b;
    if (!entropy || !log_normaliser || !avi_base_make(base, param, &b)) return AVI_ERR_INVALID;
    *entropy = b.h0;
    // base_nl2(u) = -2 log phi(u) - log(2 pi) has the additive constant nl2_c = -2 log phi(0) - log(2 pi)
    *log_normaliser = (float)(-0.5 * ((double)b.nl2_c + 1.8378770664093453));
    return AVI_OK;
}

int32_t avi_obj_set_model(avi_obj* obj, avi_model* model) {
    if (!obj || !model) return AVI_ERR_INVALID;
    if (model->D != obj->D) AVI_FAIL(obj->ctx, AVI_ERR_INVALID, "dimension of the new target differs");
    if (obj->model != model) obj->generation++;
    obj->model = model;
    return AVI_OK;
}

int32_t avi_obj_seed(avi_obj* obj, uint64_t key, uint64_t step) {
    if (!obj) return AVI_ERR_INVALID;
    obj->key = key; obj->step = step;
    obj->fr.Lr3_owner = nullptr;   // (anything an optimiser loop drew ahead belongs to the old key / step)
    return obj_push_state(obj);
}

int32_t avi_obj_get_step(const avi_obj* obj, uint64_t* step) {
    if (!obj || !step) return AVI_ERR_INVALID;
    *step = obj->step;
    return AVI_OK;
}

int32_t avi_obj_set_sample_shard(avi_obj* obj, int32_t m0, int32_t M_local) {
    if (!obj) return AVI_ERR_INVALID;
    if (m0 < 0 || M_local < 0 || m0 + M_local > obj->M) AVI_FAIL(obj->ctx, AVI_ERR_INVALID, "shard outside [0, M)");
    obj->m0 = m0; obj->Mloc = M_local;
    obj->shard_axis = (M_local == obj->M) ? obj->shard_axis : AVI_SHARD_SAMPLES;
    obj->generation++;
    return AVI_OK;
}

int32_t avi_obj_set_shard_axis(avi_obj* obj, int32_t axis) {
    if (!obj || axis < AVI_SHARD_NONE || axis > AVI_SHARD_ROWS) return AVI_ERR_INVALID;
    if (axis == AVI_SHARD_ROWS && (obj->m0 != 0 || obj->Mloc != obj->M))
        AVI_FAIL(obj->ctx, AVI_ERR_INVALID, "row sharding needs every rank to hold all samples");
    obj->shard_axis = axis;
    obj->generation++;
    return AVI_OK;
}

int64_t avi_obj_num_params(const avi_obj* obj) { return obj ? obj->P : -1; }

static int32_t check_lambda(avi_obj* o, const float* lambda_host, int64_t P) {
    if (!lambda_host || P != o->P) AVI_FAIL(o->ctx, AVI_ERR_INVALID, "lambda must have num_params entries");
    return AVI_OK;
}

int32_t avi_obj_estimate_gradient(avi_obj* o, const float* lambda_host, int64_t P, float* grad_host, float* value,
                                  float* elbo) {
    if (!o) return AVI_ERR_INVALID;
    avi_ctx* ctx = o->ctx;
    AVI_CHECK(check_lambda(o, lambda_host, P));
    const auto tm0 = std::chrono::steady_clock::now();
    cudaSetDevice(ctx->device);
    std::memcpy(o->h_lambda, lambda_host, (size_t)P * sizeof(float));
    // Mean-field: no copy-engine node in the chain.  lambda is staged in by a kernel reading the pinned buffer, and the
    // finalize kernel writes gradient + scalars + a completion flag (the new step counter) straight into pinned host
    // memory; the host spins on that flag.  (AVI_ZERO_COPY=0: copy nodes + stream wait, as for the full-rank family.)
    static const bool zc_env = !(getenv("AVI_ZERO_COPY") && atoi(getenv("AVI_ZERO_COPY")) == 0);
    const bool zero_copy = zc_env && o->family == AVI_MEANFIELD && P <= (1 << 16);
    auto enqueue = [&]() -> int32_t {
        if (zero_copy) {
            StepTail ft{};
            ft.mode = STEP_TAIL_GRAD_OUT;
            ft.lam = o->d_lambda; ft.host_out = o->h_grad;
            bool taken = false;
            // ONE launch: lambda read from the pinned buffer slice by slice, sample -> contractions -> gradient + completion
            // flag written to pinned host memory (step_fused.cu)
            AVI_CHECK(avi_objective_fused(o, o->d_lambda, ft, &taken, false, o->h_lambda));
            if (taken) { o->step += 1; return AVI_OK; }
            AVI_CHECK(avi_obj_stage_lambda(o));
            {   // targets / modes that need all of lambda on the device first: stage-in kernel + the one launch
                AVI_CHECK(avi_objective_fused(o, o->d_lambda, ft, &taken));
                if (taken) { o->step += 1; return AVI_OK; }
            }
            AVI_CHECK(avi_objective_local(o, o->d_lambda));
            AVI_CHECK(avi_objective_finalize(o, o->d_lambda, o->grad, o->out, false, /*fuse_advance=*/true));
            o->step += 1;
            return AVI_OK;
        }
        AVI_CUDA(ctx, cudaMemcpyAsync(o->d_lambda, o->h_lambda, (size_t)P * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        AVI_CHECK(avi_objective_local(o, o->d_lambda));
        AVI_CHECK(avi_objective_finalize(o, o->d_lambda, o->grad, o->out));
        AVI_CHECK(avi_obj_advance(o));
        AVI_CUDA(ctx, cudaMemcpyAsync(o->h_grad, o->grad, (size_t)P * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        AVI_CUDA(ctx, cudaMemcpyAsync(o->h_grad + P, o->out, 4 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        return AVI_OK;
    };
    // The whole call (copies included: the staging buffers are pinned and fixed) is captured as a CUDA graph
    // on its second use and replayed afterwards; anything that invalidates device pointers bumps `generation`.
    static const bool no_graph = getenv("AVI_NO_GRAPH") && atoi(getenv("AVI_NO_GRAPH")) != 0;
    const bool capturable = !no_graph && !ctx->timing && !o->model->needs_sync_eval() &&
                            !(ctx->nranks > 1 && !ctx->comm_capturable);
    const int64_t gen = avi_graph_key(o, true);
    if (o->eg_exec && o->eg_gen != gen) {
        cudaGraphExecDestroy(o->eg_exec); cudaGraphDestroy(o->eg_graph);
        o->eg_exec = nullptr; o->eg_graph = nullptr; o->eg_calls = 0;
    }
    if (zero_copy)   // arm the completion flag: anything but the value this call will publish (o->step + 1)
        *reinterpret_cast<volatile unsigned int*>(o->h_grad + P + 4) = ~(unsigned int)(o->step + 1);
    if (capturable && o->eg_exec) {
        AVI_CUDA(ctx, cudaGraphLaunch(o->eg_exec, ctx->stream));
        ctx->launches += o->eg_launches;
        o->step += 1;
    } else if (capturable && o->eg_calls >= 1 && o->eg_gen == gen) {
        const int64_t l0 = ctx->launches;
        const unsigned long long step0 = o->step;
        AVI_CUDA(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        ctx->capturing = true;
        int32_t rc = enqueue();
        ctx->capturing = false;
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
        o->eg_launches = ctx->launches - l0;
        ctx->launches = l0;
        o->step = step0;
        if (rc != AVI_OK) { if (g) cudaGraphDestroy(g); return rc; }
        if (e != cudaSuccess) AVI_FAIL(ctx, AVI_ERR_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
        o->eg_graph = g;
        AVI_CUDA(ctx, cudaGraphInstantiate(&o->eg_exec, o->eg_graph, 0));
        o->eg_gen = avi_graph_key(o, true);
        AVI_CUDA(ctx, cudaGraphLaunch(o->eg_exec, ctx->stream));
        ctx->launches += o->eg_launches;
        o->step += 1;
    } else {
        AVI_CHECK(enqueue());   // first call: also sizes every lazily allocated buffer
        o->eg_calls++;
        o->eg_gen = avi_graph_key(o, true);
    }
    const auto tm1 = std::chrono::steady_clock::now();
    if (zero_copy) {
        // completion flag = low 32 bits of the device step counter after this call (== the host mirror o->step)
        volatile unsigned int* flag = reinterpret_cast<volatile unsigned int*>(o->h_grad + P + 4);
        const unsigned int want = (unsigned int)o->step;
        for (unsigned long long spins = 0; *flag != want; ++spins) {
#if defined(__x86_64__)
            __builtin_ia32_pause();
#endif
            if ((spins & 0xFFFFull) == 0xFFFFull) {   // every ~64k polls: has the stream failed or finished without the flag?
                cudaError_t qe = cudaStreamQuery(ctx->stream);
                if (qe != cudaErrorNotReady) {
                    AVI_CUDA(ctx, qe);
                    if (*flag != want) AVI_FAIL(ctx, AVI_ERR_STATE, "estimate_gradient: stream finished without the completion flag");
                }
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    } else {
        AVI_CUDA(ctx, avi_stream_wait(ctx));
    }
    const auto tm2 = std::chrono::steady_clock::now();
    if (grad_host) std::memcpy(grad_host, o->h_grad, (size_t)P * sizeof(float));
    if (value) *value = o->h_grad[P];
    if (elbo) *elbo = o->h_grad[P + 1];
    o->last_launch_us = std::chrono::duration<double, std::micro>(tm1 - tm0).count();   // diagnostics (avi_hoststep_timing)
    o->last_wait_us = std::chrono::duration<double, std::micro>(tm2 - tm1).count();
    return AVI_OK;
}

int32_t avi_obj_estimate_objective(avi_obj* o, const float* lambda_host, int64_t P, int32_t n_samples,
                                   int32_t objective, int32_t entropy, uint64_t key, float* neg_elbo) {
    if (!o || !neg_elbo) return AVI_ERR_INVALID;
    avi_ctx* ctx = o->ctx;
    AVI_CHECK(check_lambda(o, lambda_host, P));
    if (n_samples < 1) AVI_FAIL(ctx, AVI_ERR_INVALID, "n_samples must be >= 1");
    const bool lowrank = o->family == AVI_LOWRANK;
    const bool closed = objective == AVI_REPGRAD && (entropy == AVI_ENT_CLOSEDFORM || entropy == AVI_ENT_CLOSEDFORM_ZEROGRAD);
    cudaSetDevice(ctx->device);
    std::memcpy(o->h_lambda, lambda_host, (size_t)P * sizeof(float));
    AVI_CUDA(ctx, cudaMemcpyAsync(o->d_lambda, o->h_lambda, (size_t)P * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    const int chunk = std::max(o->cap_M, std::min(n_samples, 32768));
    AVI_CHECK(avi_obj_ensure_capacity(o, std::min(chunk, n_samples)));
    ObjDeviceState ov{};
    ov.key = key; ov.step = 0;
    double s_logp = 0.0, s_esq = 0.0;
    float logdet = 0.0f, lr_H = 0.0f;
    if (lowrank) {   // entropy and the capacitance inverse B^-1 once; log q per sample (if needed) inside the chunks
        AVI_CHECK(avi_lr_entropy(o, o->d_lambda));
        AVI_CUDA(ctx, cudaMemcpyAsync(&lr_H, o->lr_ent, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    }
    for (int m0 = 0; m0 < n_samples; m0 += chunk) {
        const int Mc = std::min(chunk, n_samples - m0);
        AVI_CHECK(avi_objective_forward_chunk(o, o->d_lambda, m0, Mc, &ov, o->out, lowrank && !closed));
        float h[4];
        AVI_CUDA(ctx, cudaMemcpyAsync(h, o->out, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
        AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        s_logp += h[0]; s_esq += h[1]; logdet = h[2];
    }
    // restore the ScoreGrad centring slot clobbered above
    AVI_CUDA(ctx, cudaMemsetAsync(o->out, 0, 4 * sizeof(float), ctx->stream));
    const double D = o->D, LOG2PI = 1.8378770664093453, H0 = o->base.h0;   // (1.4189385... for Normal(0, 1))
    const double energy = s_logp / n_samples;
    double ent;
    if (lowrank)   // H from k_lr_entropy, or -mean log q(z) (the chunk sums then carry sum log q in the second slot)
        ent = closed ? (double)lr_H : -s_esq / n_samples;
    else if (closed)
        ent = D * H0 + logdet;
    else   // -mean log q(z) with scale \ (z - mu) == eps
        ent = 0.5 * s_esq / n_samples + 0.5 * D * LOG2PI + logdet;
    *neg_elbo = (float)(-(energy + ent));
    return AVI_OK;
}

int32_t avi_obj_rand(avi_obj* o, const float* lambda_host, int64_t P, float* Z_host, float* eps_host) {
    if (!o) return AVI_ERR_INVALID;
    avi_ctx* ctx = o->ctx;
    AVI_CHECK(check_lambda(o, lambda_host, P));
    cudaSetDevice(ctx->device);
    std::memcpy(o->h_lambda, lambda_host, (size_t)P * sizeof(float));
    AVI_CUDA(ctx, cudaMemcpyAsync(o->d_lambda, o->h_lambda, (size_t)P * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    if (o->Mloc <= 0) return AVI_OK;
    AVI_CHECK(avi_family_sample(o, o->d_lambda, o->Z, o->E, o->esq, o->Mloc, o->m0, o->d_state, nullptr));
    const size_t w = (size_t)o->D * sizeof(float), pitch = (size_t)o->ld * sizeof(float);
    if (Z_host) AVI_CUDA(ctx, cudaMemcpy2DAsync(Z_host, w, o->Z, pitch, w, o->Mloc, cudaMemcpyDeviceToHost, ctx->stream));
    if (eps_host) AVI_CUDA(ctx, cudaMemcpy2DAsync(eps_host, w, o->E, pitch, w, o->Mloc, cudaMemcpyDeviceToHost, ctx->stream));
    AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return AVI_OK;
}

// gaussian_expectation_gradient_and_hessian! (src/algorithms/gauss_expected_grad_hess.jl:20-58), first-order branch
// (Stein / Price identity, :33-58): u ~ N(0, I), z = C u + m, then
//   log pi_avg = mean log pi(z_b),  grad = mean grad log pi(z_b),  hess = C' \ mean(u_b grad log pi(z_b)').
// Device path: the full-rank sampling kernels, the target's batched log-density + gradient, column sums, the
// sample contraction U' G on the tensor cores (3xTF32, as the full-rank gradient) accumulated over chunks of
// <= 4096 samples, and one triangular solve with D right-hand sides.  The 1 / n_samples is applied on the host.
int32_t avi_obj_gauss_expected_grad_hess(avi_obj* o, const float* lambda_host, int64_t P, int32_t n_samples,
                                         float* logpi_avg, float* grad_host, float* hess_host) {
    if (!o || !logpi_avg || !grad_host || !hess_host) return AVI_ERR_INVALID;
    avi_ctx* ctx = o->ctx;
    if (o->family != AVI_FULLRANK)
        AVI_FAIL(ctx, AVI_ERR_INVALID, "gaussian_expectation_gradient_and_hessian needs a full-rank (triangular scale) Gaussian");
    if (o->model->capability < 1)
        AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED, "the target must provide logdensity_and_gradient (capability >= 1)");
    AVI_CHECK(check_lambda(o, lambda_host, P));
    if (n_samples < 1) AVI_FAIL(ctx, AVI_ERR_INVALID, "n_samples must be >= 1");
    if (ctx->nranks > 1 && o->shard_axis == AVI_SHARD_ROWS) AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED, "row-sharded targets are not supported here");
    cudaSetDevice(ctx->device);
    const int D = o->D, ld = o->ld, accv = o->accv;
    const int chunk = std::min(n_samples, 4096);
    AVI_CHECK(avi_obj_ensure_capacity(o, chunk));
    std::memcpy(o->h_lambda, lambda_host, (size_t)P * sizeof(float));
    AVI_CUDA(ctx, cudaMemcpyAsync(o->d_lambda, o->h_lambda, (size_t)P * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    float* scal = o->acc + 4 * (size_t)accv;
    float* C1 = scal + ACC_NSCAL;              // per-chunk contraction, then the solved Hessian
    float* C2 = C1 + (size_t)D * D;            // running sum of u g' over the chunks
    float* gsum = o->acc + accv;               // running sum of g
    AVI_CUDA(ctx, cudaMemsetAsync(C2, 0, (size_t)D * D * sizeof(float), ctx->stream));
    AVI_CUDA(ctx, cudaMemsetAsync(gsum, 0, (size_t)accv * sizeof(float), ctx->stream));
    std::vector<float> lp((size_t)chunk);
    double s_logp = 0.0;
    for (int m0 = 0; m0 < n_samples; m0 += chunk) {
        const int Mc = std::min(chunk, n_samples - m0);
        AVI_CHECK(avi_family_sample(o, o->d_lambda, o->Z, o->E, o->esq, Mc, m0, o->d_state, nullptr));
        AVI_CHECK(o->model->eval(o->Z, ld, Mc, o->logp, o->G));
        AVI_CHECK(avi_colsum_add(ctx, o->G, ld, Mc, D, o->acc, gsum));
        // C1[j * D + i] = sum_b U[b][i] G[b][j]  (column-major D x D)
        if (avi_fr_tc_ok(o, Mc)) AVI_CHECK(avi_fr_outer_tc(o, o->E, o->G, C1, Mc, 0, false));
        else AVI_CHECK(avi_gemm_simt(ctx, o->G, 1, ld, o->E, 1, ld, C1, D, 1, D, D, Mc, 1.0f));
        AVI_CHECK(avi_axpy(ctx, C1, C2, (int64_t)D * D));
        AVI_CUDA(ctx, cudaMemcpyAsync(lp.data(), o->logp, (size_t)Mc * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int b = 0; b < Mc; ++b) s_logp += lp[(size_t)b];
    }
    // hess = C' \ A: column j of A is a right-hand side (A is column-major: "row" j of a [D][D] sample-major buffer)
    AVI_CHECK(avi_trsm_lt(ctx, o->d_lambda + D, D, C2, C1, D, D));
    AVI_CHECK(avi_obj_advance(o));
    AVI_CUDA(ctx, cudaMemcpyAsync(grad_host, gsum, (size_t)D * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AVI_CUDA(ctx, cudaMemcpyAsync(hess_host, C1, (size_t)D * D * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const float inv = 1.0f / (float)n_samples;
    for (int i = 0; i < D; ++i) grad_host[i] *= inv;
    for (size_t e = 0; e < (size_t)D * D; ++e) hess_host[e] *= inv;
    *logpi_avg = (float)(s_logp / n_samples);
    return AVI_OK;
}

// rand_batch_match_samples_with_objective! (src/algorithms/fisherminbatchmatch.jl:81-111): the full-rank sampling
// kernels, the target's batched log-density + gradient, T = G C (row b: C' grad_b; exact-fp32 SIMT contraction over
// the lower triangle) and per-sample |u_b + T_b|^2; sums finish on the host in double.  Chunks of <= 4096 samples.
int32_t avi_obj_batch_match_samples(avi_obj* o, const float* lambda_host, int64_t P, int32_t n_samples, float* u_host,
                                    float* z_host, float* grad_host, float* fisher, float* logpi_avg) {
    if (!o || !fisher || !logpi_avg) return AVI_ERR_INVALID;
    avi_ctx* ctx = o->ctx;
    if (o->family != AVI_FULLRANK || o->base.kind != AVI_BASE_NORMAL)
        AVI_FAIL(ctx, AVI_ERR_INVALID, "rand_batch_match_samples_with_objective! needs a full-rank (triangular scale) Gaussian");
    if (o->model->capability < 1)
        AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED, "`FisherMinBatchMatch` requires at least first-order differentiation capability");
    AVI_CHECK(check_lambda(o, lambda_host, P));
    if (n_samples < 1) AVI_FAIL(ctx, AVI_ERR_INVALID, "n_samples must be >= 1");
    if (ctx->nranks > 1 && o->shard_axis == AVI_SHARD_ROWS) AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED, "row-sharded targets are not supported here");
    cudaSetDevice(ctx->device);
    const int D = o->D, ld = o->ld;
    const int chunk = std::min(n_samples, 4096);
    AVI_CHECK(avi_obj_ensure_capacity(o, chunk));
    std::memcpy(o->h_lambda, lambda_host, (size_t)P * sizeof(float));
    AVI_CUDA(ctx, cudaMemcpyAsync(o->d_lambda, o->h_lambda, (size_t)P * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<float> lp((size_t)chunk), fr((size_t)chunk);
    double s_logp = 0.0, s_fisher = 0.0;
    const size_t row = (size_t)D * sizeof(float), pitch = (size_t)ld * sizeof(float);
    for (int m0 = 0; m0 < n_samples; m0 += chunk) {
        const int Mc = std::min(chunk, n_samples - m0);
        AVI_CHECK(avi_family_sample(o, o->d_lambda, o->Z, o->E, o->esq, Mc, m0, o->d_state, nullptr));
        AVI_CHECK(o->model->eval(o->Z, ld, Mc, o->logp, o->G));
        // T[b][j] = sum_{i >= j} G[b][i] C[i + D j]: a = sample b (rows of G), b = column j of C, k = i
        AVI_CHECK(avi_gemm_simt(ctx, o->G, ld, 1, o->d_lambda + D, D, 1, o->U, ld, 1, Mc, D, D, 1.0f));
        AVI_CHECK(avi_rowsq_sum(ctx, o->E, o->U, ld, D, Mc, o->fbuf));   // fbuf[b] = |u_b + T_b|^2
        AVI_CUDA(ctx, cudaMemcpyAsync(lp.data(), o->logp, (size_t)Mc * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        AVI_CUDA(ctx, cudaMemcpyAsync(fr.data(), o->fbuf, (size_t)Mc * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        if (u_host) AVI_CUDA(ctx, cudaMemcpy2DAsync(u_host + (size_t)m0 * D, row, o->E, pitch, row, (size_t)Mc, cudaMemcpyDeviceToHost, ctx->stream));
        if (z_host) AVI_CUDA(ctx, cudaMemcpy2DAsync(z_host + (size_t)m0 * D, row, o->Z, pitch, row, (size_t)Mc, cudaMemcpyDeviceToHost, ctx->stream));
        if (grad_host) AVI_CUDA(ctx, cudaMemcpy2DAsync(grad_host + (size_t)m0 * D, row, o->G, pitch, row, (size_t)Mc, cudaMemcpyDeviceToHost, ctx->stream));
        AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int b = 0; b < Mc; ++b) { s_logp += lp[(size_t)b]; s_fisher += fr[(size_t)b]; }
    }
    AVI_CHECK(avi_obj_advance(o));
    AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *fisher = (float)(s_fisher / n_samples);
    *logpi_avg = (float)(s_logp / n_samples);
    return AVI_OK;
}

// Random.shuffle(rng, dataset) of src/reshuffling.jl:29 as a pure function of (key, shuffle_index):
// Fisher-Yates, descending i, j = (word_i * (i + 1)) >> 32 with word_i the i-th Philox4x32-10 output
// word of counter (i / 4, shuffle_index, 0, STREAM_SHUFFLE).  Host-side integer arithmetic.
int32_t avi_shuffle(uint64_t key, uint64_t shuffle_index, int64_t n, int32_t* perm_inout) {
    if (n < 0 || (n > 0 && !perm_inout)) return AVI_ERR_INVALID;
    uint32_t w[4] = {0, 0, 0, 0};
    int64_t have = -1;
    for (int64_t i = n - 1; i >= 1; --i) {
        const int64_t blk = i / 4;
        if (blk != have) {
            philox4x32_10((uint32_t)blk, (uint32_t)shuffle_index, 0u, (uint32_t)AVI_STREAM_SHUFFLE, (uint32_t)key,
                          (uint32_t)(key >> 32), w);
            have = blk;
        }
        const uint64_t j = ((uint64_t)w[i & 3] * (uint64_t)(i + 1)) >> 32;
        int32_t t = perm_inout[i]; perm_inout[i] = perm_inout[j]; perm_inout[j] = t;
    }
    return AVI_OK;
}

}  // extern "C"

int32_t avi_exchange(avi_ctx* ctx, float* buf, int64_t count) {
    if (ctx->nranks <= 1 || count <= 0) return AVI_OK;
    int32_t rc = avi_comm_exchange(ctx, buf, count);   // peer-memory one-shot all-reduce when connected
    if (rc != AVI_ERR_UNSUPPORTED) return rc;
    if (!ctx->ar_fn) AVI_FAIL(ctx, AVI_ERR_COMM, "multi-rank context without an exchange (avi_ctx_set_allreduce / avi_comm_connect)");
    if (ctx->capturing) AVI_FAIL(ctx, AVI_ERR_STATE, "callback exchange inside a captured step");
    int32_t r = ctx->ar_fn(ctx->ar_user, buf, count, (void*)ctx->stream);
    if (r != 0) AVI_FAIL(ctx, AVI_ERR_COMM, "all-reduce callback returned " + std::to_string(r));
    return AVI_OK;
}
