/* avi.h -- C ABI of the B200-native ELBO-gradient engine (libavi_b200.so).
 *
 * Drop-in boundary for ONE path of TuringLang/AdvancedVI.jl v0.7.0: the Monte-Carlo ELBO
 * gradient estimator over the location-scale Gaussian family and the SGD step around it.
 * Host code (Julia via @ccall, or the Python mirror in advancedvi.jl_b200/) binds exactly
 * these symbols; no torch / CUDA types appear in any signature.  File:line citations are
 * relative to the reference checkout (/root/reference).
 *
 * Conventions
 *   - every function returns int32 status (AVI_OK == 0); avi_last_error(ctx) gives the
 *     message of the last failure on that ctx (ctx == NULL: last failure of a create call);
 *   - element type is float32 only (the Julia glue raises ArgumentError otherwise);
 *   - matrices are column-major, Monte-Carlo samples are COLUMNS (src/utils.jl:6);
 *   - flat parameter vector lambda: mean-field [mu(D); diag(scale)(D)]
 *     (src/families/location_scale.jl:39-43), full-rank [mu(D); vec(L)(D*D)] with the
 *     zero strict upper triangle present (Optimisers.destructure through Functors);
 *   - `*_host` pointers are host memory borrowed for the call, `*_dev` are device memory;
 *   - a ctx and its handles are used by one host thread at a time; one CUDA stream per ctx;
 *   - a non-finite objective value is NOT an error here: the caller's `step` raises
 *     (src/algorithms/common.jl:83-89);
 *   - there is no CPU fallback: without a CUDA device avi_ctx_create fails.
 */
#ifndef AVI_H
#define AVI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct avi_ctx avi_ctx;     /* device + stream + (optional) communicator          */
typedef struct avi_model avi_model; /* target log-density: LogDensityProblems plugin side */
typedef struct avi_obj avi_obj;     /* variational objective + family (L3 + L2 of SURVEY) */
typedef struct avi_opt avi_opt;     /* fused SGD step: rule + operator + averager         */

enum { AVI_OK = 0, AVI_ERR_INVALID = 1, AVI_ERR_CUDA = 2, AVI_ERR_UNSUPPORTED = 3,
       AVI_ERR_COMM = 4, AVI_ERR_STATE = 5, AVI_ERR_CALLBACK = 6 };

/* MeanFieldGaussian / FullRankGaussian (src/families/location_scale.jl:124-141) */
enum { AVI_MEANFIELD = 0, AVI_FULLRANK = 1, AVI_LOWRANK = 2 };
/* RepGradELBO (src/algorithms/repgradelbo.jl:21-24) / ScoreGradELBO (scoregradelbo.jl:15-17) */
enum { AVI_REPGRAD = 0, AVI_SCOREGRAD = 1 };
/* entropy estimators, src/algorithms/entropy.jl:11-15, 25-29, 40-46, 57-65, 78-90 */
enum { AVI_ENT_CLOSEDFORM = 0, AVI_ENT_MONTECARLO = 1, AVI_ENT_STL = 2,
       AVI_ENT_CLOSEDFORM_ZEROGRAD = 3, AVI_ENT_STL_ZEROGRAD = 4 };
/* Optimisers.Descent / Adam, src/optimization/rules.jl:48-64 (DoG), :17-34 (DoWG) */
enum { AVI_RULE_DESCENT = 0, AVI_RULE_ADAM = 1, AVI_RULE_DOG = 2, AVI_RULE_DOWG = 3 };
/* IdentityOperator (src/AdvancedVI.jl:199), ClipScale (clip_scale.jl:8-29),
 * ProximalLocationScaleEntropy (proximal_location_scale_entropy.jl:20-61) */
enum { AVI_OP_IDENTITY = 0, AVI_OP_CLIPSCALE = 1, AVI_OP_PROXENTROPY = 2 };
/* NoAveraging / PolynomialAveraging (src/optimization/averaging.jl:7-53) */
enum { AVI_AVG_NONE = 0, AVI_AVG_POLYNOMIAL = 1 };
/* GLM targets: likelihood and which docs model the prior follows
 * (docs/src/tutorials/subsampling.md:26-38 | README.md:47-58 + :91-106) */
enum { AVI_GLM_BERNOULLI_LOGIT = 0, AVI_GLM_GAUSSIAN = 1 };
enum { AVI_GLM_SUBSAMPLING = 0, AVI_GLM_BASIC = 1 };
/* arithmetic of the X*beta / X'*r contractions */
enum { AVI_GEMM_SIMT_FP32 = 0, AVI_GEMM_TF32 = 1, AVI_GEMM_TF32X3 = 2 };
/* which axis a multi-rank run partitions */
enum { AVI_SHARD_NONE = 0, AVI_SHARD_SAMPLES = 1, AVI_SHARD_ROWS = 2 };
/* base distribution `dist` of MvLocationScale(location, scale, dist) (src/families/location_scale.jl:15-19) */
enum { AVI_BASE_NORMAL = 0, AVI_BASE_LAPLACE = 1, AVI_BASE_STUDENT_T = 2 };

/* ---- lifecycle ------------------------------------------------------------------------- */
int32_t avi_version(void);
const char* avi_last_error(const avi_ctx* ctx);
int32_t avi_ctx_create(int32_t device, avi_ctx** out);
int32_t avi_ctx_destroy(avi_ctx* ctx);
int32_t avi_ctx_synchronize(avi_ctx* ctx);
/* device info for callers that size work (SM count, bytes of HBM) */
int32_t avi_ctx_info(avi_ctx* ctx, int32_t* sm_count, int64_t* hbm_bytes, int32_t* cc_major, int32_t* cc_minor);
/* number of kernels launched on this ctx since creation (bench.py "gpu_launches") */
int64_t avi_ctx_launch_count(const avi_ctx* ctx);

/* Device timing of the named hot kernels ("sample", "glm_step", "glm_fwd_bwd", "glm_fwd", "glm_bwd", "gemm_store") with CUDA events
 * on the ctx stream, for roofline reporting.  While enabled, avi_opt_steps launches eagerly (no graph). */
int32_t avi_ctx_timing(avi_ctx* ctx, int32_t enable);
/* Diagnostic step timeline (AVI_TIMELINE=1 in the environment at avi_ctx_create): %globaltimer stamps (ns) of the
 * last 64 fused iterations, hist[(step % 64) * 32 + slot]; slot id / 4 + id / 8 + id = first CTA entered / first CTA
 * past its dependency wait / last CTA done, id 0 sample, 1 forward, 2 backward, 3 tail (slots 16.. : experiments).  hist_host: 2048 uint64. */
int32_t avi_ctx_timeline_get(avi_ctx* ctx, uint64_t* hist_host);
int32_t avi_ctx_timing_get(avi_ctx* ctx, const char* name, double* total_ms, int64_t* count);

/* Multi-rank plumbing.  The exchange step of the path is ONE sum-all-reduce of the partial
 * accumulator [grad sums ; scalars] per step.  The library does not own a communicator:
 * the host supplies the exchange as a callback that runs between the local phase and the
 * replicated finalize (NCCL through torch.distributed in the Python mirror, NCCL.jl /
 * MPI.jl from Julia).  buf_dev is device memory of `count` floats, reduced in place on the
 * ctx stream whose handle is `stream` (a cudaStream_t). */
typedef int32_t (*avi_allreduce_fn)(void* user, float* buf_dev, int64_t count, void* stream);
int32_t avi_ctx_set_allreduce(avi_ctx* ctx, avi_allreduce_fn fn, void* user, int32_t rank, int32_t nranks);
/* Native exchange over NVLink peer memory (one-shot all-reduce kernel of ours, graph-capturable; takes
 * precedence over the callback).  Each rank: avi_comm_buffer (allocates its symmetric buffer for payloads
 * of up to max_floats floats, writes its 64-byte CUDA IPC handle), the host all-gathers the handles,
 * then avi_comm_connect with the nranks x 64-byte table. */
int32_t avi_comm_buffer(avi_ctx* ctx, int64_t max_floats, char* handle_out_64);
int32_t avi_comm_connect(avi_ctx* ctx, int32_t rank, int32_t nranks, const char* handles);
/* Unmap the peers' buffers (collective by convention: every rank calls it, then the ranks synchronise, before any
 * rank destroys its context or calls avi_comm_buffer again). */
int32_t avi_comm_disconnect(avi_ctx* ctx);
/* Enqueue a device-side rendezvous of the connected ranks on the ctx stream (no-op for a single rank). */
int32_t avi_comm_barrier(avi_ctx* ctx);
/* the cudaStream_t every call on this ctx enqueues on (for hosts that order their own work after it) */
void* avi_ctx_stream(avi_ctx* ctx);

/* ---- targets (replace the user's per-sample logdensity called at
 *      src/algorithms/repgradelbo.jl:84-86 and the MixedADLogDensityProblem pullback,
 *      src/mixedad_logdensity.jl:23-34) ---------------------------------------------------- */
/* logpdf(MvNormal(mu, Diagonal(sigma.^2)), z): test/models/normal.jl:8-11, :56-75 */
int32_t avi_model_mvnormal_diag_create(avi_ctx* ctx, const float* mu_host, const float* sigma_host,
                                       int32_t D, avi_model** out);
/* Hierarchical GLM, theta = [beta(d); log sigma].  X_host is n x d column-major and y_host has
 * n entries (0/1 for Bernoulli-logit); both are copied to the device once.  n_data is the
 * full-data size used for the likelihood adjustment n_data / n_batch. */
int32_t avi_model_glm_create(avi_ctx* ctx, const float* X_host, const float* y_host, int64_t n,
                             int32_t d, int64_t n_data, int32_t likelihood, int32_t variant,
                             int32_t gemm_mode, avi_model** out);
/* Any other LogDensityProblem: cb is called once per sample with z (D floats, host) and must
 * write *logp and, when grad != NULL, grad (D floats).  Non-zero return aborts the call. */
typedef int32_t (*avi_logdensity_fn)(void* user, const float* z, int32_t D, float* logp, float* grad);
int32_t avi_model_hostcallback_create(avi_ctx* ctx, int32_t D, int32_t capability,
                                      avi_logdensity_fn cb, void* user, avi_model** out);
/* AdvancedVI.subsample(prob, batch) (src/AdvancedVI.jl:303-313;
 * docs/src/tutorials/subsampling.md:99-102): restrict a GLM target to the rows idx_host[0..batch)
 * (0-based) with likelihood adjustment n_data / batch.  idx_host == NULL restores all rows. */
int32_t avi_model_subsample(avi_model* model, const int32_t* idx_host, int64_t batch);
/* Multi-rank data sharding (n-axis): this target instance was created from one of
 * `nshards` disjoint row slices holding rows_global rows in total.  The likelihood adjustment
 * becomes n_data / rows_global (n_data / (batch * nshards) after avi_model_subsample) and the
 * prior terms are included only when include_prior != 0 (exactly one rank), so that the sum of
 * the per-rank log-densities and gradients is the full-data one. */
int32_t avi_model_set_data_shard(avi_model* model, int32_t nshards, int64_t rows_global, int32_t include_prior);
int32_t avi_model_dimension(const avi_model* model);      /* LogDensityProblems.dimension    */
int32_t avi_model_capability(const avi_model* model);     /* LogDensityProblems.capabilities */
int32_t avi_model_set_gemm_mode(avi_model* model, int32_t gemm_mode);
/* How iterations over this target are launched: 0 = one kernel per stage (sample / forward / backward / tail),
 * 1 = (default; env AVI_FUSED_STEP) the whole iteration as ONE persistent kernel where the fixed per-launch costs
 * matter, 2 = always the single kernel.  Results agree up to summation order.  AVI_ERR_UNSUPPORTED for targets
 * without a fused path. */
int32_t avi_model_set_fused_step(avi_model* model, int32_t mode);
/* Batched LogDensityProblems.logdensity / logdensity_and_gradient on device data:
 * Z_dev is D x M column-major with leading dimension ldz; logp_dev has M entries; G_dev is
 * D x M with leading dimension ldz. */
int32_t avi_model_logdensity(avi_model* model, const float* Z_dev, int32_t ldz, int32_t M, float* logp_dev);
int32_t avi_model_logdensity_and_gradient(avi_model* model, const float* Z_dev, int32_t ldz, int32_t M,
                                          float* logp_dev, float* G_dev);
/* Same with host buffers (D x M column-major, ld = D); copies in and out. */
int32_t avi_model_logdensity_and_gradient_host(avi_model* model, const float* Z_host, int32_t M,
                                               float* logp_host, float* G_host /* may be NULL */);
int32_t avi_model_destroy(avi_model* model);

/* ---- objective: init / estimate_gradient! / estimate_objective
 *      (src/algorithms/abstractobjective.jl:25-86) ----------------------------------------- */
int32_t avi_obj_create(avi_ctx* ctx, avi_model* model, int32_t family, int32_t objective,
                       int32_t entropy, int32_t M, avi_obj** out);
/* MvLocationScaleLowRank / LowRankGaussian (src/families/location_scale_low_rank.jl:16-24, :119-135): covariance
 * diag(scale_diag^2) + scale_factors scale_factors', lambda = [location (D); scale_diag (D); vec(scale_factors)
 * (D x rank, column-major)] (Functors order, :26), P = 2 D + D rank.  rank <= 32.  Supported: RepGradELBO with
 * ClosedFormEntropy (what KLMinRepGradDescent uses by default, docs/src/families.md:185-190),
 * ClosedFormEntropyZeroGradient, MonteCarloEntropy and StickingTheLandingEntropy, ScoreGradELBO (VarGrad) -- log q(z)
 * and its gradients go through the rank x rank capacitance matrix (Woodbury) --, estimate_gradient!, estimate_objective,
 * rand and the fused step with Descent / Adam / DoG / DoWG, IdentityOperator / ClipScale (on scale_diag,
 * clip_scale.jl:31-41) and both averagers; every entropy estimator of entropy.jl.  AVI_ERR_UNSUPPORTED: the proximal
 * operator (defined for MvLocationScale only, proximal_location_scale_entropy.jl:46-61) and the log q based estimators
 * on more than one rank. */
int32_t avi_obj_create_lowrank(avi_ctx* ctx, avi_model* model, int32_t rank, int32_t objective, int32_t entropy,
                               int32_t M, avi_obj** out);
/* Base distribution of MvLocationScale(location, scale, dist) for the mean-field and full-rank families: Normal(0, 1)
 * (default; MeanFieldGaussian / FullRankGaussian), Laplace(0, 1) or TDist(param) (docs/src/families.md:72-101).  The base
 * supplies the draws of rand (location_scale.jl:71-87: z = scale * u + location, u iid from dist), logpdf
 * (:59-63: sum_i logpdf(dist, u_i) - logdet(scale)) and entropy (:52-57: D * entropy(dist) + logdet(scale)); every
 * objective, entropy estimator, estimate_objective and the optimiser loop work with any base.  Draws are counter-based
 * (Philox) like the normals: Laplace by inversion, Student-t by Bailey's polar transform (no rejection).  The
 * single-launch iteration is a Normal(0, 1) kernel: other bases run the multi-kernel path.  AVI_ERR_UNSUPPORTED for the
 * low-rank family (LowRankGaussian is Gaussian by definition) and for ProximalLocationScaleEntropy-style zero-gradient
 * closed forms nothing changes (the entropy constant does not enter a gradient). */
int32_t avi_obj_set_base(avi_obj* obj, int32_t base, float param);
/* The two constants the kernels take from a base distribution, computed on the host (no device needed): *entropy =
 * entropy(dist) (location_scale.jl:52-57 multiplies it by D) and *log_normaliser = log phi(0), the additive constant of
 * logpdf(dist, u) (Normal: -log(2 pi)/2; Laplace: -log 2; TDist(nu): lgamma((nu+1)/2) - lgamma(nu/2) - log(nu pi)/2).
 * AVI_ERR_INVALID for an unknown base or nu outside (0, 1e6). */
int32_t avi_base_constants(int32_t base, float param, float* entropy, float* log_normaliser);
/* set_objective_state_problem (repgradelbo.jl:31-39, scoregradelbo.jl:24-32) */
int32_t avi_obj_set_model(avi_obj* obj, avi_model* model);
/* eps[i, m] at step t is a pure function of (key, t, m, i) (Philox4x32-10 + Box-Muller);
 * the host draws `key` from its rng once, so "same seed => identical run"
 * (test/algorithms/klminrepgraddescent.jl:40-57) holds. */
int32_t avi_obj_seed(avi_obj* obj, uint64_t key, uint64_t step);
int32_t avi_obj_get_step(const avi_obj* obj, uint64_t* step);
/* partition the Monte-Carlo samples: this rank evaluates samples [m0, m0 + M_local) of M */
int32_t avi_obj_set_sample_shard(avi_obj* obj, int32_t m0, int32_t M_local);
/* which axis the ranks of this ctx partition: AVI_SHARD_SAMPLES (set by avi_obj_set_sample_shard)
 * exchanges the partial gradient sums; AVI_SHARD_ROWS (targets sharded with
 * avi_model_set_data_shard, every rank holds all M samples) exchanges log pi and its gradient sums */
int32_t avi_obj_set_shard_axis(avi_obj* obj, int32_t axis);
int64_t avi_obj_num_params(const avi_obj* obj);
/* estimate_gradient! (repgradelbo.jl:151-177 / scoregradelbo.jl:96-117): lambda in, gradient
 * of the value slot out; *value = -ELBO (RepGrad) or VarGrad (ScoreGrad), *elbo = info.elbo.
 * Uses the eps of the current step and then advances the step counter.  Blocking. */
int32_t avi_obj_estimate_gradient(avi_obj* obj, const float* lambda_host, int64_t P,
                                  float* grad_host, float* value, float* elbo);
/* estimate_objective (repgradelbo.jl:112-118 / scoregradelbo.jl:58-65 / common.jl:29-38):
 * forward only with n_samples draws and the given entropy estimator; returns -ELBO.
 * objective == AVI_SCOREGRAD ignores `entropy`. */
int32_t avi_obj_estimate_objective(avi_obj* obj, const float* lambda_host, int64_t P, int32_t n_samples,
                                   int32_t objective, int32_t entropy, uint64_t key, float* neg_elbo);
/* rand(rng, q, M) (src/families/location_scale.jl:71-87): Z_host and eps_host (either may be
 * NULL) receive D x M column-major draws of the current step WITHOUT advancing it. */
int32_t avi_obj_rand(avi_obj* obj, const float* lambda_host, int64_t P, float* Z_host, float* eps_host);
/* gaussian_expectation_gradient_and_hessian!(rng, q, n_samples, grad_buf, hess_buf, prob), first-order (Stein / Price)
 * branch -- src/algorithms/gauss_expected_grad_hess.jl:20-58: with u ~ N(0, I), z = C u + m,
 * *logpi_avg = mean log pi(z), grad_host (D) = mean grad log pi(z), hess_host (D x D, column-major, not symmetrised)
 * = C' \ mean(u grad log pi(z)').  obj must be a full-rank objective (lambda = [m; vec(C)]); draws come from the
 * objective's Philox stream at its current step, which then advances.  The sampling stage of KLMinWassFwdBwd,
 * KLMinNaturalGradDescent and KLMinSqrtNaturalGradDescent (klminwassfwdbwd.jl:101, klminnaturalgraddescent.jl:120). */
/* rand_batch_match_samples_with_objective!(rng, q, n_samples, prob, u_buf, grad_buf) -- the sampling stage of
 * FisherMinBatchMatch, src/algorithms/fisherminbatchmatch.jl:81-111: u ~ N(0, I) (D x n), z = C u + mu, grad[:, b] =
 * grad log pi(z_b), *logpi_avg = mean log pi(z_b) and the Fisher-divergence estimate *fisher = sum |-u - C' grad|^2 / n.
 * u_host, z_host, grad_host: D x n_samples column-major (any may be NULL).  obj must be a full-rank Gaussian objective
 * (lambda = [mu; vec(C)]) over a target with capability >= 1; the draws are the objective's Philox stream at its
 * current step, which then advances.  The d x d batch-and-match update itself stays on the host (:140-190). */
int32_t avi_obj_batch_match_samples(avi_obj* obj, const float* lambda_host, int64_t P, int32_t n_samples, float* u_host,
                                    float* z_host, float* grad_host, float* fisher, float* logpi_avg);
int32_t avi_obj_gauss_expected_grad_hess(avi_obj* obj, const float* lambda_host, int64_t P, int32_t n_samples,
                                         float* logpi_avg, float* grad_host, float* hess_host);
int32_t avi_obj_destroy(avi_obj* obj);

/* ---- minibatch order: ReshufflingBatchSubsampling (src/reshuffling.jl:27-32) --------------------
 * In-place Fisher-Yates shuffle of perm_inout[0..n) driven by Philox4x32-10 (key, shuffle_index);
 * the k-th reshuffle of a run uses shuffle_index = k, so a run is a pure function of its key. */
int32_t avi_shuffle(uint64_t key, uint64_t shuffle_index, int64_t n, int32_t* perm_inout);

/* ---- fused step: Optimisers.update! + operator + averager
 *      (src/algorithms/common.jl:91-94) with parameters resident on the device ------------- */
/* hyper: Descent {eta}; Adam {eta, beta1, beta2, epsilon}; DoG/DoWG {alpha}.
 * op_param: ClipScale epsilon.  avg_param: PolynomialAveraging eta. */
int32_t avi_opt_create(avi_obj* obj, int32_t rule, const float* hyper, int32_t n_hyper, int32_t op,
                       float op_param, int32_t averager, float avg_param, const float* lambda0_host,
                       int64_t P, avi_opt** out);
/* n iterations of `step` (common.jl:69-104) without the callback; per-iteration value slot and
 * elbo go to the host arrays (n entries each, either may be NULL).  Stops early and returns
 * AVI_OK with *n_done < n when the value slot is not finite (the caller raises).  One stream
 * synchronisation at the end. */
int32_t avi_opt_steps(avi_opt* opt, int32_t n, float* value_host, float* elbo_host, int32_t* n_done);
/* The same call in three phases, for callers that keep the device queue full (no host round trip between
 * iterations): _begin reserves trace space for `capacity` iterations and does every piece of setup (graph capture
 * included); _enqueue launches n more iterations and returns WITHOUT synchronising (sum of n <= capacity);
 * _end copies the trace back, synchronises once and reports like avi_opt_steps.  Full-batch objectives only. */
int32_t avi_opt_steps_begin(avi_opt* opt, int32_t capacity);
int32_t avi_opt_steps_enqueue(avi_opt* opt, int32_t n);
int32_t avi_opt_steps_end(avi_opt* opt, float* value_host, float* elbo_host, int32_t* n_done);
/* Range check of minibatch indices (0-based rows of the target), what avi_opt_steps_subsampled applies to idx_host before
 * anything reaches the device: AVI_OK iff every idx[j] lies in [0, rows); otherwise AVI_ERR_INVALID with *first_bad (may
 * be NULL) = the first offending position.  Host only. */
int32_t avi_check_indices(const int32_t* idx, int64_t n, int64_t rows, int64_t* first_bad);
/* same, with the minibatch of every iteration given up front: idx_host holds n * batch row
 * indices (SubsampledObjective, src/algorithms/subsampledobjective.jl:64-90) */
int32_t avi_opt_steps_subsampled(avi_opt* opt, int32_t n, const int32_t* idx_host, int64_t batch,
                                 float* value_host, float* elbo_host, int32_t* n_done);
/* Host-side counterpart of the fused update, for callers that stay on the estimate_gradient! boundary and keep the
 * parameters on the host (what Optimisers.update! + the operator + the averager do in `step`, common.jl:91-94, on
 * Julia arrays): Descent / Adam on a mean-field lambda = [mu (D); diag scale (D)] of P = 2 D entries (or any flat
 * vector with scale_offset = -1: no operator), ClipScale on the scale entries, PolynomialAveraging.  state16 holds the
 * scalar state between calls (zero-initialise it; entries: 0 t of the averager, 3 beta1^t, 4 beta2^t).  m1 / m2 /
 * lambda_avg may be NULL when the rule / averager does not use them.  No device work. */
int32_t avi_host_update(int32_t rule, const float* hyper, int32_t n_hyper, int32_t op, float op_param, int32_t averager,
                        float avg_param, int64_t P, int64_t scale_offset, float* lambda, const float* grad, float* m1,
                        float* m2, float* lambda_avg, float* state16);
/* `step` (src/algorithms/common.jl:75-104) for callers whose parameters stay in HOST memory: one call = estimate_gradient!
 * (host lambda in, host gradient out through the zero-copy boundary) + Optimisers.update! + apply(operator) +
 * apply(averager) on the host (common.jl:91-94).  The handle binds the caller-owned arrays once: lambda (P, updated in
 * place), grad (P, receives every gradient), lambda_avg (P or NULL without averaging).  Descent / Adam, IdentityOperator /
 * ClipScale (entries from scale_offset on), NoAveraging / PolynomialAveraging.  A non-finite value slot leaves the
 * parameters untouched and is reported through `value` (the caller raises, common.jl:83-89).  _timing: wall-clock
 * microseconds of the last call: estimate_gradient! as a whole, the host update, and inside the former the time to
 * enqueue the launch and the time until the completion flag was seen (any pointer may be NULL). */
typedef struct avi_hoststep avi_hoststep;
int32_t avi_hoststep_create(avi_obj* obj, int32_t rule, const float* hyper, int32_t n_hyper, int32_t op, float op_param,
                            int32_t averager, float avg_param, int64_t scale_offset, float* lambda, float* grad,
                            float* lambda_avg, avi_hoststep** out);
int32_t avi_hoststep_step(avi_hoststep* hs, float* value, float* elbo);
int32_t avi_hoststep_timing(const avi_hoststep* hs, double* estimate_us, double* update_us, double* enqueue_us, double* wait_us);
int32_t avi_hoststep_destroy(avi_hoststep* hs);
/* current iterate, averaged iterate (output(), common.jl:63-67) and last gradient */
int32_t avi_opt_get(avi_opt* opt, float* lambda_host, float* lambda_avg_host, float* grad_host);
int64_t avi_opt_iteration(const avi_opt* opt);
/* warm start (src/optimize.jl:50, :58-62): serialise / restore the complete device state */
int64_t avi_opt_state_nbytes(const avi_opt* opt);
int32_t avi_opt_state_export(avi_opt* opt, void* buf_host, int64_t nbytes);
int32_t avi_opt_state_import(avi_opt* opt, const void* buf_host, int64_t nbytes);
int32_t avi_opt_destroy(avi_opt* opt);

#ifdef __cplusplus
}
#endif
#endif /* AVI_H */
