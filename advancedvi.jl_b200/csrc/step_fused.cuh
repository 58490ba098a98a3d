// Host-side interface of the fused whole-iteration kernel (step_fused.cu): ONE launch per `step`
// (src/algorithms/common.jl:75-104 minus the callback) for the mean-field family over a hierarchical GLM target:
//   sample phase   rand(rng, q, M)                               src/families/location_scale.jl:80-87
//   -- grid barrier --
//   forward phase  X * beta + log-likelihood + weighted residual  docs/src/tutorials/subsampling.md:35-36 for all M samples
//   -- grid barrier --
//   backward phase X' R reduced against eps (mean-field pullback) src/families/location_scale.jl:86, repgradelbo.jl:142-149
//   -- grid barrier --
//   tail phase     [NVLink exchange of the partial sums] + closed-form gradient + value / ELBO + finiteness check +
//                  Optimisers rule + operator + averager + commit src/algorithms/common.jl:83-94
// The contraction phases are the tcgen05 pipelines of gemm_tc.cu; what the fusion removes is three kernel boundaries
// (prologue, first-operand latency, teardown: ~8 us per contraction launch at the benchmark shape) and it lets the
// static operand of the NEXT phase stream in while the current one drains.
#pragma once

#include <cuda.h>
#include <stdint.h>

#include "avi_internal.cuh"
#include "comm_dev.cuh"
#include "gemm_tc.cuh"
#include "mf_tail.cuh"

enum { STEP_TAIL_NONE = 0, STEP_TAIL_UPDATE = 1, STEP_TAIL_GRAD_OUT = 2 };
// which classes of partial sums are summed over the ranks by the fused exchange
enum { STEP_X_V01 = 1, STEP_X_V23 = 2, STEP_X_S0 = 4, STEP_X_S1 = 8 };

struct StepTail {
    int mode;                      // STEP_TAIL_*
    float* acc; int accv;          // [v0 | v1 | v2 | v3] as in avi_internal.cuh (v0, v1 written by the backward phase)
    int M, objective, entropy;
    float* logp;                   // [Mloc] log pi(z_m) (unused by the kernel: the tail takes sum_m log pi from unit_ll and pre)
    const float* unit_ll; int n_units_f; float w_lik;   // per forward unit: log-likelihood total; likelihood adjustment
    const float *part1, *part2; int nslab, ldslab;      // split-K slabs of sum_m g, sum_m g*eps (backward phase)
    float *lam, *grad, *m1, *m2, *avg, *sc, *out;
    float* trace; int trace_cap;
    UpdArgs a;
    float* norm_part;              // DoG / DoWG: [2 * grid] partial norms
    float* host_out;               // STEP_TAIL_GRAD_OUT: mapped pinned [grad (2 D) | value, elbo, logdet, shift | flag]
    unsigned int* done_ticket;     // STEP_TAIL_GRAD_OUT: last-CTA election
    CommPeers comm; int xmask; long long acc_len;
};

struct StepParams {
    TcParams f, b;                 // forward / backward contraction plans (no clusters, no CTA pairs)
    int stages_f, stages_b;
    // bwd_store != 0 (requires t.mode == STEP_TAIL_NONE, do_sample == 0, draw_ahead == 0): the backward phase writes the
    // FULL gradient block as split-K slabs b.C[ks][sample][coordinate] (plain store epilogue) instead of the mean-field
    // sums, and the forward phase leaves per-sample partial log-likelihoods (f.post_on = 0): the batched
    // logdensity_and_gradient of the target for the families that need every g_m (full-rank, low-rank, Stein / BaM stages)
    int bwd_store;
    // sample phase
    int do_sample;                 // 0: Z, Zt, E, esq, pre come from the stand-alone sampling kernel
    const float* lambda; int D, ld, m0, Mloc;
    ObjDeviceState* st;
    float *Z, *E, *esq;
    int d, variant, include_prior;
    float* Zt; int zt_ld, zt_seg;
    float* pre;                    // float4 per sample
    // slice-major sampling, one iteration ahead (optimiser loop, plain TF32 mode): the tail phase draws the next
    // iteration's samples; spart = [2][grid][Mloc] per-sample partial sums (|eps|^2, |beta|^2) + [Mloc] eta
    int draw_ahead; float* spart; int spart_stride;
    // estimate_gradient! boundary: lambda arrives in MAPPED PINNED HOST memory; each CTA fetches only its own slice over
    // PCIe (one round trip), keeps a device copy in t.lam and contributes a partial of log det(scale) to ldpart[grid]
    const float* lambda_src; float* ldpart;
    StepTail t;
    unsigned long long* gbar;      // grid barrier: [0] monotonic arrival counter | [1] its value at the start of the launch;
                                   // [2] completion ticket (estimate_gradient! boundary) | [3] who drew the target's samples ahead
    unsigned long long* tl;        // AVI_TIMELINE: %globaltimer stamps per phase (diagnostic)
    unsigned long long* prof;      // AVI_STEP_PROF: per-CTA stamps [grid][32] (diagnostic)
};

// what the objective layer hands to a target that can run the fused iteration (avi_model::fused_step)
struct FusedStepArgs {
    bool dry_run;                  // size every buffer the launch needs, launch nothing (before a graph capture)
    const float* lambda; int D, ld, m0, Mloc;
    const float* lambda_src;       // != nullptr: lambda is still in mapped pinned host memory (see StepParams)
    ObjDeviceState* st;
    float *Z, *E, *esq, *logp;
    StepTail t;
};

int avi_step_fused_max_per_cta();   // coordinates of the tail one CTA can take
int32_t avi_step_fused_launch(avi_ctx* ctx, const CUtensorMap& tmZ, const CUtensorMap& tmXr, const CUtensorMap& tmXc,
                              const CUtensorMap& tmR, StepParams& sp);
