// Simple targets: the diagonal-Gaussian test target and the host-callback target.
//   logpdf(MvNormal(mu, Diagonal(sigma.^2)), z)   test/models/normal.jl:8-11, :56-75
//   any LogDensityProblem through a per-sample host callback: the per-sample loop of
//   src/algorithms/repgradelbo.jl:84-86 and the logdensity_and_gradient pullback of
//   src/mixedad_logdensity.jl:23-34, kept for targets without a native kernel.
#include "avi_internal.cuh"
#include "device_utils.cuh"

namespace {

__global__ void __launch_bounds__(256)
k_mvnormal_diag(const float* __restrict__ mu, const float* __restrict__ sigma, int D, const float* __restrict__ Z,
                int ld, float* __restrict__ logp, float* __restrict__ G) {
    __shared__ float sm[33];
    const int m = blockIdx.x;
    float q = 0.f, ls = 0.f;
    for (int i = threadIdx.x; i < ld; i += blockDim.x) {
        float g = 0.0f;
        if (i < D) {
            float s = __ldg(sigma + i);
            float r = (Z[(size_t)m * ld + i] - __ldg(mu + i)) / s;
            q = fmaf(r, r, q);
            ls += logf(s);
            g = -r / s;
        }
        if (G) G[(size_t)m * ld + i] = g;
    }
    q = block_sum(q, sm);
    ls = block_sum(ls, sm);
    if (threadIdx.x == 0) logp[m] = -0.5f * (float)D * AVI_LOG2PI - ls - 0.5f * q;
}

struct MvNormalDiag : avi_model {
    float *mu = nullptr, *sigma = nullptr;
    ~MvNormalDiag() override { avi_free(mu); avi_free(sigma); }
    int32_t eval(const float* Z, int ld, int M, float* logp, float* G) override {
        if (M <= 0) return AVI_OK;
        k_mvnormal_diag<<<M, 256, 0, ctx->stream>>>(mu, sigma, D, Z, ld, logp, G);
        AVI_LAUNCHED(ctx);
        return AVI_OK;
    }
};

struct HostCallback : avi_model {
    avi_logdensity_fn cb = nullptr;
    void* user = nullptr;
    std::vector<float> hz, hg, hl;
    bool needs_sync_eval() const override { return true; }
    int32_t eval(const float* Z, int ld, int M, float* logp, float* G) override {
        if (M <= 0) return AVI_OK;
        if (ctx->capturing) AVI_FAIL(ctx, AVI_ERR_STATE, "host-callback target inside a captured step");
        if (G && capability < 1)
            AVI_FAIL(ctx, AVI_ERR_UNSUPPORTED,
                     "the target has no gradient (capability 0): RepGradELBO needs logdensity_and_gradient; "
                     "use ScoreGradELBO or supply the gradient");
        hz.resize((size_t)M * ld); hl.resize(M);
        if (G) hg.assign((size_t)M * ld, 0.0f);
        AVI_CUDA(ctx, cudaMemcpyAsync(hz.data(), Z, hz.size() * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
        AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (int m = 0; m < M; ++m) {
            int32_t rc = cb(user, hz.data() + (size_t)m * ld, D, &hl[m], G ? hg.data() + (size_t)m * ld : nullptr);
            if (rc != 0) AVI_FAIL(ctx, AVI_ERR_CALLBACK, "logdensity callback returned " + std::to_string(rc));
        }
        AVI_CUDA(ctx, cudaMemcpyAsync(logp, hl.data(), M * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        if (G) AVI_CUDA(ctx, cudaMemcpyAsync(G, hg.data(), hg.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        AVI_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return AVI_OK;
    }
};

}  // namespace

int32_t avi_model_mvnormal_diag_make(avi_ctx* ctx, const float* mu, const float* sigma, int D, avi_model** out) {
    if (D <= 0 || !mu || !sigma) AVI_FAIL(ctx, AVI_ERR_INVALID, "bad arguments");
    for (int i = 0; i < D; ++i)
        if (!(sigma[i] > 0.0f)) AVI_FAIL(ctx, AVI_ERR_INVALID, "sigma must be positive");
    MvNormalDiag* mdl = new MvNormalDiag();
    mdl->ctx = ctx; mdl->D = D; mdl->capability = 1;
    int32_t rc = avi_alloc(ctx, &mdl->mu, D);
    if (rc == AVI_OK) rc = avi_alloc(ctx, &mdl->sigma, D);
    if (rc != AVI_OK) { delete mdl; return rc; }
    avi_copy(ctx, mdl->mu, mu, D * sizeof(float), cudaMemcpyHostToDevice);
    avi_copy(ctx, mdl->sigma, sigma, D * sizeof(float), cudaMemcpyHostToDevice);
    *out = mdl;
    return AVI_OK;
}

int32_t avi_model_hostcallback_make(avi_ctx* ctx, int D, int capability, avi_logdensity_fn cb, void* user,
                                    avi_model** out) {
    if (D <= 0 || !cb) AVI_FAIL(ctx, AVI_ERR_INVALID, "bad arguments");
    HostCallback* mdl = new HostCallback();
    mdl->ctx = ctx; mdl->D = D; mdl->capability = capability > 0 ? 1 : 0;
    mdl->cb = cb; mdl->user = user;
    *out = mdl;
    return AVI_OK;
}
