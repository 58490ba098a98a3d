// K1 (eps-sample + affine transform + |eps|^2) and K3 (reduction over the M samples to the
// gradient of the variational parameters) for the location-scale Gaussian family, plus the
// objective orchestration:  local phase -> [exchange] -> finalize.
//
// Replaces (reference file:line, paths under /root/reference):
//   rand(rng, q, M)                      src/families/location_scale.jl:71-87
//   entropy(q) / logpdf(q, z)            src/families/location_scale.jl:52-63
//   estimate_entropy (5 estimators)      src/algorithms/entropy.jl:13-15, 27-29, 42-46, 59-65, 80-90
//   estimate_repgradelbo_ad_forward + AD src/algorithms/repgradelbo.jl:142-177
//   estimate_scoregradelbo_ad_forward    src/algorithms/scoregradelbo.jl:87-117
// Gradients are the closed forms of SURVEY.md Appendix A (checked against finite differences of
// the restated forward in tests/test_oracle_gradients.py).
#include <cstdlib>

#include "avi_internal.cuh"
#include "device_utils.cuh"
#include "glm_prior.cuh"
#include "fr_finalize.cuh"
#include "mf_finalize.cuh"
#include "tc_common.cuh"
#include "step_fused.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// K1: one WARP per Monte-Carlo sample; lane l owns the coordinate quads l, l + 32, ... (one Philox block
// each; a warp instruction moves 512 contiguous bytes).  Writes Z = mu + s .* eps (mean-field), E = eps
// (both zero in the padding columns i >= D) and |eps_m|^2 (fixed shuffle tree: no block barrier, no
// atomics).  FULLRANK writes only E (Z = L * eps + mu comes from k_fr_affine).  HOOK: SampleHook.
// SPLIT = warps cooperating on one sample: 1 for large M, SAMPLE_WARPS for small M (latency: the row is
// spread over the whole CTA, partial sums combined in a fixed order through smem).
// SPLIT == 1 is the bandwidth shape: CTAs stride over groups of SAMPLE_WARPS samples and (mean-field) keep
// mu and s zero-padded in shared memory, so the per-quad work is Philox + Box-Muller + 2 LDS.128 + 2 STG.128.
constexpr int SAMPLE_WARPS = 4;
// GBASE: base distribution other than Normal(0, 1) (base_dist.cuh); the Gaussian instantiations carry none of that code
template <bool FULLRANK, bool HOOK, int SPLIT, bool GBASE = false>
__global__ void __launch_bounds__(32 * SAMPLE_WARPS)
k_sample(const float* __restrict__ lambda, int D, int ld, int m0, int Mloc, const ObjDeviceState* __restrict__ st,
         ObjDeviceState st_val, int use_val, uint32_t stream_id, float* __restrict__ Z,
         float* __restrict__ E, float* __restrict__ esq, SampleHook hk, BaseDist bd) {
    extern __shared__ __align__(16) float s_ms[];   // STAGE: [ld] mu, [ld] s
    constexpr bool STAGE = !FULLRANK && SPLIT == 1;
    if (HOOK) tl_min(hk.tl, 0);
    if (HOOK && hk.zt_owner && blockIdx.x == 0 && threadIdx.x == 0) *hk.zt_owner = 0ull;
    pdl_trigger();
    if (HOOK && hk.pf_bytes)   // stream the forward kernel's X into L2 while this kernel runs (static data: before the wait)
        l2_prefetch_span(hk.pf_ptr, hk.pf_bytes, blockIdx.x * SAMPLE_WARPS + (threadIdx.x >> 5), gridDim.x * SAMPLE_WARPS, 8192);
    pdl_wait();   // lambda and the step counter come from the previous iteration's tail
    if (HOOK) tl_min(hk.tl, 4);
    const unsigned long long step = use_val ? st_val.step : st->step;
    const PhiloxKeys pk(use_val ? st_val.key : st->key);
    // these draws overwrite whatever the fused iteration kernel may have drawn ahead (ObjDeviceState::zt_kind)
    if (blockIdx.x == 0 && threadIdx.x == 0 && st) const_cast<ObjDeviceState*>(st)->zt_kind = 0;
    const uint32_t c2 = (uint32_t)step, c3 = eps_ctr3(step, stream_id);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* mu = lambda;
    const float* sc = lambda + D;   // only 4-byte aligned in general
    if (STAGE) {
        for (int i = threadIdx.x; i < ld; i += 32 * SAMPLE_WARPS) {
            s_ms[i] = i < D ? __ldg(mu + i) : 0.0f;
            s_ms[ld + i] = i < D ? __ldg(sc + i) : 0.0f;
        }
        __syncthreads();
    }
    const int m_step = SPLIT == 1 ? (int)gridDim.x * SAMPLE_WARPS : Mloc;
    for (int m = SPLIT == 1 ? blockIdx.x * SAMPLE_WARPS + warp : blockIdx.x; m < Mloc; m += m_step) {
        float part = 0.0f, bsq = 0.0f, eta = 0.0f;
        float* Erow = E ? E + (size_t)m * ld : nullptr;   // (mean-field forward-only callers do not need eps: E == nullptr)
        float* Zrow = FULLRANK ? nullptr : Z + (size_t)m * ld;
        float* Er3row = FULLRANK && hk.Er3 ? hk.Er3 + (size_t)m * 3 * hk.er_seg : nullptr;
        const int qend = (Er3row && hk.er_seg > ld ? hk.er_seg : ld) / 4;
        for (int q = SPLIT == 1 ? lane : threadIdx.x; q < qend; q += 32 * SPLIT) {
            const int i = 4 * q;
            if (FULLRANK && i >= ld) {   // zero tail of the split rows beyond the sample buffers' pitch
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                *reinterpret_cast<float4*>(Er3row + i) = z4;
                *reinterpret_cast<float4*>(Er3row + hk.er_seg + i) = z4;
                *reinterpret_cast<float4*>(Er3row + 2 * hk.er_seg + i) = z4;
                continue;
            }
            const float4 e = GBASE ? base_draw4(bd, (uint32_t)q, (uint32_t)(m0 + m), c2, c3, pk)   // u ~ dist
                                   : normal4((uint32_t)q, (uint32_t)(m0 + m), c2, c3, pk);
            float ev[4] = {e.x, e.y, e.z, e.w}, zv[4] = {0.f, 0.f, 0.f, 0.f}, zt[4];
            if (i + 3 >= D) {   // the row's last quad: zero the padding columns
#pragma unroll
                for (int c = 0; c < 4; ++c) ev[c] = i + c < D ? ev[c] : 0.0f;
            }
            if (!FULLRANK) {
                float mv[4] = {0.f, 0.f, 0.f, 0.f}, sv[4] = {0.f, 0.f, 0.f, 0.f};
                if (STAGE) {
                    const float4 m4 = *reinterpret_cast<const float4*>(s_ms + i);
                    const float4 s4 = *reinterpret_cast<const float4*>(s_ms + ld + i);
                    mv[0] = m4.x; mv[1] = m4.y; mv[2] = m4.z; mv[3] = m4.w;
                    sv[0] = s4.x; sv[1] = s4.y; sv[2] = s4.z; sv[3] = s4.w;
                } else if (i + 3 < D) {
                    const float4 m4 = *reinterpret_cast<const float4*>(mu + i);
                    mv[0] = m4.x; mv[1] = m4.y; mv[2] = m4.z; mv[3] = m4.w;
#pragma unroll
                    for (int c = 0; c < 4; ++c) sv[c] = __ldg(sc + i + c);
                } else {
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        if (i + c < D) { mv[c] = __ldg(mu + i + c); sv[c] = __ldg(sc + i + c); }
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) zv[c] = fmaf(sv[c], ev[c], mv[c]);   // padding: 0 * 0 + 0
            }
            if (!GBASE) {
#pragma unroll
                for (int c = 0; c < 4; ++c) part = fmaf(ev[c], ev[c], part);
            } else {   // sum_i -2 log phi(u_i) - log 2 pi: |eps|^2's role in log q(z) for any base (base_dist.cuh)
#pragma unroll
                for (int c = 0; c < 4; ++c) part += i + c < D ? base_nl2(bd, ev[c]) : 0.0f;
            }
            if (HOOK) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const bool is_beta = i + c < hk.d;
                    bsq = is_beta ? fmaf(zv[c], zv[c], bsq) : bsq;
                    zt[c] = is_beta ? zv[c] : 0.0f;
                    if (i + c == hk.d) eta = zv[c];
                }
            }
            if (FULLRANK || Erow) *reinterpret_cast<float4*>(Erow + i) = make_float4(ev[0], ev[1], ev[2], ev[3]);
            if (FULLRANK && Er3row) {   // B-operand pattern of the 3xTF32 contraction Z = L * eps: [hi | lo | hi]
                float hi[4], lo[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) { hi[c] = tc::round_tf32(ev[c]); lo[c] = tc::round_tf32(ev[c] - hi[c]); }
                *reinterpret_cast<float4*>(Er3row + i) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<float4*>(Er3row + hk.er_seg + i) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                *reinterpret_cast<float4*>(Er3row + 2 * hk.er_seg + i) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            }
            if (!FULLRANK) *reinterpret_cast<float4*>(Zrow + i) = make_float4(zv[0], zv[1], zv[2], zv[3]);
            if (HOOK && hk.Zt) {
                float hi[4], lo[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) { hi[c] = tc::round_tf32(zt[c]); lo[c] = tc::round_tf32(zt[c] - hi[c]); }
                float* row = hk.Zt + (size_t)m * hk.zt_ld + i;
                if (hk.zt_seg == 0) {
                    *reinterpret_cast<float4*>(row) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                } else if (i < hk.zt_seg) {   // 3xTF32: [hi | hi | lo]
                    *reinterpret_cast<float4*>(row) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(row + hk.zt_seg) = make_float4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<float4*>(row + 2 * hk.zt_seg) = make_float4(lo[0], lo[1], lo[2], lo[3]);
                }
            }
        }
        float tot = warp_sum(part);
        if (HOOK) { bsq = warp_sum(bsq); eta = warp_sum(eta); }   // eta is non-zero in exactly one lane
        if (SPLIT > 1) {
            __shared__ float sm[3][SAMPLE_WARPS];
            if (lane == 0) { sm[0][warp] = tot; sm[1][warp] = bsq; sm[2][warp] = eta; }
            __syncthreads();
            tot = 0.f; bsq = 0.f; eta = 0.f;
#pragma unroll
            for (int w = 0; w < SAMPLE_WARPS; ++w) { tot += sm[0][w]; bsq += sm[1][w]; eta += sm[2][w]; }
        }
        if (lane == 0 && (SPLIT == 1 || warp == 0)) {
            esq[m] = tot;
            if (HOOK) hk.pre[m] = glm_prior_terms(bsq, eta, hk.d, hk.variant, hk.include_prior);
        }
    }
    if (HOOK) tl_max(hk.tl, 8);
}

// full-rank: Z[m][i] += mu[i] after the L * eps contraction; zero the padding columns
__global__ void k_fr_add_mu(const float* __restrict__ lambda, int D, int ld, float* __restrict__ Z) {
    const int m = blockIdx.x;
    for (int i = threadIdx.x; i < ld; i += blockDim.x) {
        float* p = Z + (size_t)m * ld + i;
        *p = i < D ? *p + __ldg(lambda + i) : 0.0f;
    }
}

// ------------------------------------------------------------------------------------------
// scalars: s0 = sum logp, s1 = sum |eps|^2; ScoreGrad additionally f_m = log q(z_m) - log pi(z_m)
// (shifted by f_shift = out[3], see k_finalize_*) and s2 = sum f, s3 = sum f^2.
// Single CTA => fixed summation order.
__global__ void __launch_bounds__(1024)
k_scalars(const float* __restrict__ lambda, int D, int fullrank, int objective, const float* __restrict__ logp,
          const float* __restrict__ esq, int Mloc, const float* __restrict__ out, float* __restrict__ fbuf,
          float* __restrict__ scal) {
    __shared__ float sm[33];
    float ld_part = 0.0f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        size_t idx = fullrank ? (size_t)D + (size_t)i * (D + 1) : (size_t)D + i;
        ld_part += logf(__ldg(lambda + idx));
    }
    const float logdet = block_sum(ld_part, sm);
    const float shift = out[3];
    float a = 0.f, b = 0.f, c = 0.f, d = 0.f;
    for (int m = threadIdx.x; m < Mloc; m += blockDim.x) {
        float lp = logp[m], es = esq[m];
        a += lp; b += es;
        if (objective == AVI_SCOREGRAD) {
            // log q(z_m) with scale \ (z - mu) == eps:  -|eps|^2/2 - D log(2 pi)/2 - logdet
            float f = (-0.5f * es - 0.5f * (float)D * AVI_LOG2PI - logdet) - lp - shift;
            fbuf[m] = f;
            c += f; d = fmaf(f, f, d);
        }
    }
    a = block_sum(a, sm); b = block_sum(b, sm); c = block_sum(c, sm); d = block_sum(d, sm);
    if (threadIdx.x == 0) {
        scal[0] = a; scal[1] = b; scal[2] = c; scal[3] = d;
        scal[4] = 0.f; scal[5] = 0.f; scal[6] = 0.f; scal[7] = 0.f;
    }
}

// ------------------------------------------------------------------------------------------
// K3, mean-field: per-coordinate sums over the local samples.  CTA = 32 coordinates x 32 sample
// groups; coalesced 128 B rows; fixed-order combine across the 32 groups.
//   RepGrad  : v0 = sum g, v1 = sum g*eps        (skipped when the target produced them itself)
//   ScoreGrad: v0 = sum f*eps, v1 = sum f*eps^2
//   both     : v2 = sum eps, v3 = sum eps^2
__global__ void __launch_bounds__(1024)
k_reduce_mf(const float* __restrict__ G, const float* __restrict__ E, const float* __restrict__ fbuf,
            int ld, int Mloc, int D, int accv, int objective, int skip_g, float* __restrict__ acc, BaseDist bd) {
    __shared__ float sm[4][32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * 32 + tx;
    float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
    if (i < D) {
        for (int m = ty; m < Mloc; m += 32) {
            float e = E[(size_t)m * ld + i];
            const float sc = base_negscore(bd, e);   // -d log phi / du: e itself for Normal(0, 1)
            v2 += sc; v3 = fmaf(sc, e, v3);
            if (objective == AVI_SCOREGRAD) {
                float f = fbuf[m];
                v0 = fmaf(f, sc, v0); v1 = fmaf(f * sc, e, v1);
            } else if (!skip_g) {
                float g = G[(size_t)m * ld + i];
                v0 += g; v1 = fmaf(g, e, v1);
            }
        }
    }
    sm[0][ty][tx] = v0; sm[1][ty][tx] = v1; sm[2][ty][tx] = v2; sm[3][ty][tx] = v3;
    __syncthreads();
    if (ty < 4 && i < D) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) s += sm[ty][r][tx];
        if (!(skip_g && ty < 2)) acc[(size_t)ty * accv + i] = s;
    }
}

// ------------------------------------------------------------------------------------------
// finalize, mean-field: global sums -> gradient of the value slot, value, elbo.
// Gradient entries are independent; the scalars are recomputed by every CTA in the same fixed
// order (CTA 0 writes them), so the grid size does not change any result.
__global__ void __launch_bounds__(1024)
k_finalize_mf(const float* __restrict__ acc, int accv, const float* __restrict__ lambda, int D, int M,
              int objective, int entropy, const float* __restrict__ logp, const float* __restrict__ esq, int Mloc,
              int deferred, float* __restrict__ grad, float* __restrict__ out, float* __restrict__ host_out,
              ObjDeviceState* __restrict__ advance_st, float h0) {
    __shared__ float sm[33];
    pdl_trigger();
    pdl_wait();
    const float* s = lambda + D;
    MfSums S = mf_collect_sums(lambda, D, acc + 4 * (size_t)accv, logp, esq, Mloc, deferred, sm);
    S.h0 = h0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < D; i += gridDim.x * blockDim.x) {
        float gm, gs;
        mf_grad_entry(acc, accv, __ldg(s + i), i, M, objective, entropy, S, gm, gs);
        grad[i] = gm; grad[D + i] = gs;
        if (host_out) { host_out[i] = gm; host_out[D + i] = gs; }   // posted writes into mapped pinned host memory
    }
    if (host_out) {
        // estimate_gradient! boundary (launched as ONE CTA): the gradient went straight to the caller-visible pinned
        // buffer [grad (2 D) | value, elbo, logdet, shift | done flag]; once every thread's stores are fenced, thread 0
        // adds the scalars, advances the step counter and release-stores the new counter value as the completion flag
        // the host spins on -- no device-to-host copy node, no stream query round trip
        __threadfence_system();
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float value, elbo, shift_next;
        mf_outputs(D, M, objective, entropy, S, out[3], value, elbo, shift_next);
        out[0] = value; out[1] = elbo; out[2] = S.logdet; out[3] = shift_next;
        // estimate_gradient! boundary: the scalars also go right behind the gradient (one device-to-host copy) and
        // the step counter advances here instead of in a launch of its own
        if (advance_st) {
            const unsigned long long step_next = advance_st->step + 1ull;
            advance_st->step = step_next;
            if (host_out) {
                float* tail = host_out + 2 * (size_t)D;
                tail[0] = value; tail[1] = elbo; tail[2] = S.logdet; tail[3] = shift_next;
                __threadfence_system();
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(reinterpret_cast<unsigned int*>(tail + 4)),
                             "r"((unsigned int)step_next) : "memory");
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// full-rank helpers
// W[m][i] = G[m][i] + U[m][i]   (STL: w = g + L^-T eps), or W[m][i] = f_m * U[m][i] (ScoreGrad)
__global__ void k_fr_make_w(const float* G, const float* __restrict__ U, const float* __restrict__ fbuf, int ld,
                            int Mloc, int mode, float* W) {   // W may alias G
    const int m = blockIdx.x;
    for (int i = threadIdx.x; i < ld; i += blockDim.x) {
        size_t p = (size_t)m * ld + i;
        W[p] = mode == 0 ? G[p] + U[p] : fbuf[m] * U[p];
    }
}

// column sums of a sample-major buffer: v[i] = sum_m W[m][i]
__global__ void __launch_bounds__(1024)
k_colsum(const float* __restrict__ W, int ld, int Mloc, int D, float* __restrict__ v) {
    __shared__ float sm[32][33];
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i = blockIdx.x * 32 + tx;
    float a = 0.f;
    if (i < D)
        for (int m = ty; m < Mloc; m += 32) a += W[(size_t)m * ld + i];
    sm[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && i < D) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) s += sm[r][tx];
        v[i] = s;
    }
}

// grad = [ . ; vec(tril(...)) ]: RepGrad  -C1/M -+ diag(1/L_ii);  ScoreGrad  (C1 - fbar C2)/M.
// C*[j*D + i] = sum_m W[m][i] E[m][j] (column-major L layout).
__global__ void k_finalize_fr_mat(const float* __restrict__ C1, const float* __restrict__ C2,
                                  const float* __restrict__ scal, const float* __restrict__ lambda, int D,
                                  int M, int objective, int entropy, float* __restrict__ grad) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)D * D) return;
    const int j = (int)(idx / D), i = (int)(idx % D);   // column-major: entry (i, j)
    float g = 0.0f;
    if (i >= j) g = fr_grad_entry(C1, C2, scal, __ldg(lambda + D + idx), idx, i, j, M, objective, entropy);
    grad[D + idx] = g;
}

__global__ void __launch_bounds__(1024)
k_finalize_fr_vec(const float* __restrict__ acc, int accv, const float* __restrict__ lambda, int D, int M,
                  int objective, int entropy, float* __restrict__ grad, float* __restrict__ out,
                  const float* __restrict__ logp, const float* __restrict__ esq, int Mloc, int deferred, float h0) {
    __shared__ float sm[33];
    fr_vec_finalize(acc, accv, lambda, D, M, objective, entropy, grad, out, logp, esq, Mloc, deferred != 0, true, sm, h0);
}

__global__ void k_advance(ObjDeviceState* st) { st->step += 1ull; }

// estimate_gradient! boundary: lambda from the caller's pinned (mapped) host buffer into device memory by a kernel
// (a few KB: one PCIe read round trip) instead of a copy-engine node in front of the compute chain
__global__ void __launch_bounds__(256)
k_stage_in(const float* __restrict__ host_src, float* __restrict__ dst, long long n) {
    pdl_trigger();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = host_src[i];
}

// forward-only chunk sums for estimate_objective: out = {sum logp, sum |eps|^2, logdet}
__global__ void __launch_bounds__(1024)
k_forward_sums(const float* __restrict__ lambda, int D, int fullrank, const float* __restrict__ logp,
               const float* __restrict__ esq, int Mc, float* __restrict__ out) {
    __shared__ float sm[33];
    float ld_part = 0.0f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        size_t idx = fullrank ? (size_t)D + (size_t)i * (D + 1) : (size_t)D + i;
        ld_part += logf(__ldg(lambda + idx));
    }
    const float logdet = block_sum(ld_part, sm);
    float a = 0.f, b = 0.f;
    for (int m = threadIdx.x; m < Mc; m += blockDim.x) { a += logp[m]; b += esq[m]; }
    a = block_sum(a, sm); b = block_sum(b, sm);
    if (threadIdx.x == 0) { out[0] = a; out[1] = b; out[2] = logdet; }
}

}  // namespace

// ==============================================================================================
static int32_t family_sample_impl(avi_obj* o, const float* lambda, float* Z, float* E, float* esq, int Mloc, int m0,
                                  const ObjDeviceState* st, const ObjDeviceState* ov, const SampleHook* hook, bool sample_only);
int32_t avi_family_sample(avi_obj* o, const float* lambda, float* Z, float* E, float* esq, int Mloc, int m0,
                          const ObjDeviceState* st, const ObjDeviceState* ov, const SampleHook* hook) {
    return family_sample_impl(o, lambda, Z, E, esq, Mloc, m0, st, ov, hook, false);
}
// full-rank: the eps-draw of the objective's current step alone (E, the split of eps, |eps|^2), no contraction
int32_t avi_fr_draw_current(avi_obj* o, const float* lambda) {
    if (o->family != AVI_FULLRANK || o->Mloc <= 0) return AVI_OK;
    return family_sample_impl(o, lambda, o->Z, o->E, o->esq, o->Mloc, o->m0, o->d_state, nullptr, nullptr, true);
}
static int32_t family_sample_impl(avi_obj* o, const float* lambda, float* Z, float* E, float* esq, int Mloc, int m0,
                                  const ObjDeviceState* st, const ObjDeviceState* ov, const SampleHook* hook, bool sample_only) {
    avi_ctx* ctx = o->ctx;
    if (Mloc <= 0) return AVI_OK;
    ObjDeviceState sv{};
    int use_val = 0;
    if (ov) { sv = *ov; use_val = 1; }
    AviTimed timed(ctx, "sample");
    // small batches: spread each sample over the whole CTA.  Large batches: warp per sample, CTAs stride over
    // the samples with mu / s staged in shared memory (needs 8 * ld bytes; wider families keep the CTA shape)
    const size_t stage_bytes = o->family == AVI_MEANFIELD ? 2 * (size_t)o->ld * sizeof(float) : 0;
    const bool split = Mloc <= 8192 || stage_bytes > 48 * 1024;
    const unsigned sgrid = split ? (unsigned)Mloc
                                 : (unsigned)std::min<int64_t>(ceil_div(Mloc, SAMPLE_WARPS), (int64_t)ctx->prop.multiProcessorCount * 12);
#define LAUNCH_SAMPLE_B(FR, HK, HOOKV, GB)                                                                           \
    do {                                                                                                             \
        if (split) avi_launch_pdl(ctx, k_sample<FR, HK, SAMPLE_WARPS, GB>, dim3(sgrid), dim3(32 * SAMPLE_WARPS), 0,   \
            lambda, o->D, o->ld, m0, Mloc, st, sv, use_val, (uint32_t)AVI_STREAM_EPS, Z, E, esq, HOOKV, o->base);     \
        else avi_launch_pdl(ctx, k_sample<FR, HK, 1, GB>, dim3(sgrid), dim3(32 * SAMPLE_WARPS), stage_bytes,          \
            lambda, o->D, o->ld, m0, Mloc, st, sv, use_val, (uint32_t)AVI_STREAM_EPS, Z, E, esq, HOOKV, o->base);     \
    } while (0)
#define LAUNCH_SAMPLE(FR, HK, HOOKV)                                                                                 \
    do {                                                                                                             \
        if (o->base.kind == AVI_BASE_NORMAL) LAUNCH_SAMPLE_B(FR, HK, HOOKV, false);                                   \
        else LAUNCH_SAMPLE_B(FR, HK, HOOKV, true);                                                                    \
    } while (0)
    if (o->family == AVI_MEANFIELD) {
        if (hook && hook->kind == 1) LAUNCH_SAMPLE(false, true, *hook);
        else LAUNCH_SAMPLE(false, false, SampleHook{});
        AVI_LAUNCHED(ctx);
    } else if (o->family == AVI_LOWRANK) {
        // u_diag -> E (D per sample, eps stream), u_fact -> E2 (rank per sample, its own Philox stream), then
        // z = scale_diag .* u_diag + scale_factors * u_fact + location (location_scale_low_rank.jl:79-86)
        LAUNCH_SAMPLE(true, false, SampleHook{});
        AVI_LAUNCHED(ctx);
        avi_launch_pdl(ctx, k_sample<true, false, SAMPLE_WARPS>, dim3((unsigned)Mloc), dim3(32 * SAMPLE_WARPS), 0, lambda,
                       o->rank, o->ldr, m0, Mloc, st, sv, use_val, (uint32_t)AVI_STREAM_EPS_FACTORS, (float*)nullptr, o->E2,
                       o->fbuf, SampleHook{}, BaseDist{});
        AVI_LAUNCHED(ctx);
        AVI_CHECK(avi_lr_affine(o, lambda, E, o->E2, Z, Mloc));
    } else {
        // Z[m][i] = mu[i] + sum_{j <= i} E[m][j] * L[i + D*j]  (scale * eps, location_scale.jl:76)
        const bool tc = avi_fr_tc_ok(o, Mloc);
        SampleHook fh{};   // (kind 0: only the split-eps fields are read by the full-rank sampler)
        if (tc) AVI_CHECK(avi_fr_affine_prepare(o, Mloc, &fh.Er3, &fh.er_seg));
        if (!(o->fr.eps_ahead && tc && E == o->E && esq == o->esq)) {   // (else: drawn by the previous iteration's update kernel)
            LAUNCH_SAMPLE(true, false, fh);
            AVI_LAUNCHED(ctx);
        }
        if (sample_only) return AVI_OK;
        if (tc) {
            AVI_CHECK(avi_fr_affine_tc(o, lambda, E, Z, Mloc, /*er3_done=*/true, hook));
        } else {   // very large forward-only batches (estimate_objective): exact-fp32 SIMT contraction
            AVI_CHECK(avi_gemm_simt(ctx, E, o->ld, 1, lambda + o->D, 1, o->D, Z, o->ld, 1, Mloc, o->D, o->D, 1.0f, 1));
            k_fr_add_mu<<<Mloc, 256, 0, ctx->stream>>>(lambda, o->D, o->ld, Z);
            AVI_LAUNCHED(ctx);
        }
    }
#undef LAUNCH_SAMPLE
#undef LAUNCH_SAMPLE_B
    return AVI_OK;
}

// Mean-field RepGrad without a sample-sharded exchange: sum logp / sum |eps|^2 feed only the value slot,
// so the finalize kernel takes them from the per-sample vectors itself (one launch less per step).
bool avi_obj_defers_scalars(const avi_obj* o) {
    return (o->family == AVI_MEANFIELD || o->family == AVI_FULLRANK) && o->objective == AVI_REPGRAD && o->Mloc > 0 &&
           !(o->shard_axis == AVI_SHARD_SAMPLES && o->ctx->nranks > 1);
}

// dst[i] += sum_m W[m][i] (tmp receives the plain column sums) and y += x: small helpers of the Stein estimator
__global__ void k_axpy(const float* __restrict__ x, float* __restrict__ y, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) y[i] += x[i];
}
int32_t avi_axpy(avi_ctx* ctx, const float* x, float* y, int64_t n) {
    if (n <= 0) return AVI_OK;
    k_axpy<<<(unsigned)std::min<int64_t>(ceil_div(n, 256), 1184), 256, 0, ctx->stream>>>(x, y, (long long)n);
    AVI_LAUNCHED(ctx);
    return AVI_OK;
}
// out[m] = sum_i (A[m][i] + B[m][i])^2: one CTA per sample, fixed summation order
__global__ void __launch_bounds__(256)
k_rowsq_sum(const float* __restrict__ A, const float* __restrict__ B, int ld, int D, float* __restrict__ out) {
    __shared__ float sm[33];
    const int m = blockIdx.x;
    float part = 0.f;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const float v = A[(size_t)m * ld + i] + B[(size_t)m * ld + i];
        part = fmaf(v, v, part);
    }
    part = block_sum(part, sm);
    if (threadIdx.x == 0) out[m] = part;
}
int32_t avi_rowsq_sum(avi_ctx* ctx, const float* A, const float* B, int ld, int D, int M, float* out) {
    if (M <= 0) return AVI_OK;
    k_rowsq_sum<<<(unsigned)M, 256, 0, ctx->stream>>>(A, B, ld, D, out);
    AVI_LAUNCHED(ctx);
    return AVI_OK;
}
int32_t avi_colsum_add(avi_ctx* ctx, const float* W, int ld, int Mloc, int D, float* tmp, float* dst) {
    k_colsum<<<(unsigned)ceil_div(D, 32), dim3(32, 32), 0, ctx->stream>>>(W, ld, Mloc, D, tmp);
    AVI_LAUNCHED(ctx);
    return avi_axpy(ctx, tmp, dst, D);
}

int32_t avi_obj_stage_lambda(avi_obj* o) {
    avi_ctx* ctx = o->ctx;
    cudaError_t e = avi_launch_pdl(ctx, k_stage_in, dim3((unsigned)ceil_div(o->P, 256)), dim3(256), 0, (const float*)o->h_lambda,
                                   o->d_lambda, (long long)o->P);
    if (e != cudaSuccess) AVI_FAIL(ctx, AVI_ERR_CUDA, std::string("stage-in launch: ") + cudaGetErrorString(e));
    AVI_LAUNCHED(ctx);
    return AVI_OK;
}

int32_t avi_obj_advance(avi_obj* o) {
    k_advance<<<1, 1, 0, o->ctx->stream>>>(o->d_state);
    AVI_LAUNCHED(o->ctx);
    o->step += 1;
    return AVI_OK;
}

// local phase: everything that only needs this rank's samples.  Leaves the partial sums in o->acc.
int32_t avi_objective_local(avi_obj* o, const float* lambda) {
    avi_ctx* ctx = o->ctx;
    const int D = o->D, ld = o->ld, Mloc = o->Mloc, accv = o->accv;
    float* scal = o->acc + 4 * (size_t)accv;
    if (Mloc <= 0) {   // a rank without samples contributes zeros
        AVI_CUDA(ctx, cudaMemsetAsync(o->acc, 0, o->acc_len * sizeof(float), ctx->stream));
        if (o->shard_axis == AVI_SHARD_SAMPLES && ctx->nranks > 1 && !o->fused_exchange)
            AVI_CHECK(avi_exchange(ctx, o->acc, o->acc_len));
        return AVI_OK;
    }
    SampleHook hook;
    // (the full-rank family runs the hook in the kernel that finishes z = L eps + mu: family_fr.cu)
    const bool hooked = (o->family == AVI_MEANFIELD || (o->family == AVI_FULLRANK && avi_fr_tc_ok(o, Mloc))) &&
                        o->model->sample_hook(o->Z, ld, Mloc, &hook);
    {
        const int32_t rc_s = avi_family_sample(o, lambda, o->Z, o->E, o->esq, Mloc, o->m0, o->d_state, nullptr, hooked ? &hook : nullptr);
        if (rc_s != AVI_OK) { o->model->clear_hook(); return rc_s; }
    }
    const bool rep = o->objective == AVI_REPGRAD;
    const bool stl = o->entropy == AVI_ENT_STL || o->entropy == AVI_ENT_STL_ZEROGRAD;
    const bool rows = o->shard_axis == AVI_SHARD_ROWS && ctx->nranks > 1;
    int skip_g = 0;
    bool logp_sent = false;
    if (rep) {
        if (o->family == AVI_MEANFIELD && o->model->has_gradsums()) {
            AVI_CHECK(o->model->eval_gradsums(o->Z, o->E, ld, Mloc, o->logp, o->acc, o->acc + accv));
            skip_g = 1;
            if (rows) AVI_CHECK(avi_exchange(ctx, o->acc, 2LL * accv));
        } else {
            AVI_CHECK(o->model->eval(o->Z, ld, Mloc, o->logp, o->G));
            if (rows) {   // G and (when the buffers are full: logp starts where G ends) log pi in one exchange
                const int64_t cap = avi_comm_capacity(ctx);
                logp_sent = Mloc == o->cap_M && (cap < 0 || (int64_t)Mloc * ld + Mloc <= cap);
                AVI_CHECK(avi_exchange(ctx, o->G, (int64_t)Mloc * ld + (logp_sent ? Mloc : 0)));
            }
        }
    } else {
        AVI_CHECK(o->model->eval(o->Z, ld, Mloc, o->logp, nullptr));
    }
    // row sharding: every rank holds all samples and a slice of the data rows; log pi and its
    // gradient are sums over rows, everything after this point is replicated arithmetic
    if (rows && !logp_sent) AVI_CHECK(avi_exchange(ctx, o->logp, Mloc));
    if (!avi_obj_defers_scalars(o)) {
        k_scalars<<<1, 1024, 0, ctx->stream>>>(lambda, D, o->family == AVI_FULLRANK, o->objective, o->logp, o->esq,
                                               Mloc, o->out, o->fbuf, scal);
        AVI_LAUNCHED(ctx);
    }
    if (o->family == AVI_LOWRANK) {
        const bool logq = avi_lr_needs_logq(o);
        if (logq) {
            // w = Sigma^-1 (z - mu), U'w and log q(z) per sample through the r x r capacitance inverse (family_lr.cu);
            // RepGrad + STL / MonteCarlo: G += w.  Replaces the scalar sums of k_scalars (whose log q is the mean-field one).
            AVI_CHECK(avi_lr_entropy(o, lambda));
            AVI_CHECK(avi_lr_logq(o, lambda, Mloc));
        }
        if (rep) {
            // v0 = sum_m g, v1 = sum_m g .* u_diag (the mean-field reduction), CU[k * D + i] = sum_m g[m][i] u_fact[m][k]
            k_reduce_mf<<<(unsigned)ceil_div(D, 32), dim3(32, 32), 0, ctx->stream>>>(o->G, o->E, o->fbuf, ld, Mloc, D, accv,
                                                                                   AVI_REPGRAD, 0, o->acc, BaseDist{});
            AVI_LAUNCHED(ctx);
            AVI_CHECK(avi_gemm_simt(ctx, o->E2, 1, o->ldr, o->G, 1, ld, scal + ACC_NSCAL, D, 1, o->rank, D, Mloc, 1.0f));
        }
        if (logq) AVI_CHECK(avi_lr_logq_sums(o, Mloc));
    } else if (o->family == AVI_MEANFIELD) {
        // with a fused target and a closed-form entropy nothing else is needed (v2, v3 unused)
        if (!(skip_g && !stl)) {
            k_reduce_mf<<<(unsigned)ceil_div(D, 32), dim3(32, 32), 0, ctx->stream>>>(
                o->G, o->E, o->fbuf, ld, Mloc, D, accv, o->objective, skip_g, o->acc, o->base);
            AVI_LAUNCHED(ctx);
        }
    } else {
        float* C1 = scal + ACC_NSCAL;
        float* C2 = C1 + (size_t)D * D;
        const float* W = o->G;
        if (!rep || stl) {
            AVI_CHECK(avi_trsm_lt(ctx, lambda + D, D, o->E, o->U, ld, Mloc, o->base));
            k_fr_make_w<<<Mloc, 256, 0, ctx->stream>>>(o->G, o->U, o->fbuf, ld, Mloc, rep ? 0 : 1, o->G);
            AVI_LAUNCHED(ctx);
        }
        // C1[j*D + i] = sum_m W[m][i] * E[m][j]: contraction over the samples.  On the tensor-core path the launch that
        // transposes and splits W and eps also takes the column sums of W and -- RepGrad without a sample-shard
        // exchange in between -- finishes the location block and the value slot (o->fr_vec_done)
        const bool tc_ok = avi_fr_tc_ok(o, Mloc);
        FrPrepFinalize pf{};
        if (tc_ok && avi_obj_defers_scalars(o)) {
            pf.on = 1; pf.lambda = lambda; pf.M = o->M; pf.objective = o->objective; pf.entropy = o->entropy;
            pf.grad = o->grad; pf.out = o->out; pf.logp = o->logp; pf.esq = o->esq; pf.Mloc = Mloc; pf.accv = accv;
            pf.h0 = o->base.h0;
        }
        if (!tc_ok) {
            k_colsum<<<(unsigned)ceil_div(D, 32), dim3(32, 32), 0, ctx->stream>>>(W, ld, Mloc, D, o->acc);
            AVI_LAUNCHED(ctx);
            AVI_CHECK(avi_gemm_simt(ctx, o->E, 1, ld, W, 1, ld, C1, D, 1, D, D, Mloc, 1.0f));
        } else {
            AVI_CHECK(avi_fr_outer_tc(o, W, o->E, C1, Mloc, 0, false, o->acc, &pf));
            o->fr_vec_done = pf.on != 0;
        }
        if (!rep) {
            k_colsum<<<(unsigned)ceil_div(D, 32), dim3(32, 32), 0, ctx->stream>>>(o->U, ld, Mloc, D,
                                                                                  o->acc + 2 * (size_t)accv);
            AVI_LAUNCHED(ctx);
            if (tc_ok) AVI_CHECK(avi_fr_outer_tc(o, o->U, o->E, C2, Mloc, 1, true));
            else AVI_CHECK(avi_gemm_simt(ctx, o->E, 1, ld, o->U, 1, ld, C2, D, 1, D, D, Mloc, 1.0f));
        }
    }
    // sample sharding: the partial sums are the exchange payload (full-rank RepGrad: the second D x D block is unused)
    if (o->shard_axis == AVI_SHARD_SAMPLES && ctx->nranks > 1 && !o->fused_exchange)
        AVI_CHECK(avi_exchange(ctx, o->acc, o->family == AVI_FULLRANK && rep ? o->acc_len - (int64_t)D * D : o->acc_len));
    return AVI_OK;
}

int32_t avi_objective_fused(avi_obj* o, const float* lambda, const StepTail& tail, bool* taken, bool dry_run, const float* lambda_src) {
    avi_ctx* ctx = o->ctx;
    *taken = false;
    if (o->family != AVI_MEANFIELD || o->objective != AVI_REPGRAD || o->Mloc <= 0) return AVI_OK;
    if (o->base.kind != AVI_BASE_NORMAL) return AVI_OK;   // (the single-launch iteration draws Normal(0, 1) itself)
    if (!o->model->fused_step_ok(o->Mloc)) return AVI_OK;
    if (lambda_src && !o->model->fused_host_lambda_ok()) return AVI_OK;
    StepTail t = tail;
    t.acc = o->acc; t.accv = o->accv; t.M = o->M; t.objective = o->objective; t.entropy = o->entropy;
    t.logp = o->logp; t.grad = o->grad; t.out = o->out; t.acc_len = o->acc_len;
    t.a.D = o->D;
    t.comm.nranks = 1; t.xmask = 0;
    if (ctx->nranks > 1 && o->shard_axis != AVI_SHARD_NONE) {
        // the exchange runs inside the kernel's tail phase: needs the peer-mapped low-latency lanes of comm.cu
        if (!avi_comm_peers(ctx, o->acc_len, &t.comm) || t.comm.ll_cap < o->acc_len) return AVI_OK;
        t.xmask = o->shard_axis == AVI_SHARD_ROWS ? (STEP_X_V01 | STEP_X_S0)
                                                  : (STEP_X_V01 | STEP_X_V23 | STEP_X_S0 | STEP_X_S1);
    }
    if (ceil_div(o->D, ctx->prop.multiProcessorCount) > avi_step_fused_max_per_cta()) return AVI_OK;
    FusedStepArgs fa{};
    fa.dry_run = dry_run;
    fa.lambda_src = lambda_src;
    fa.lambda = lambda; fa.D = o->D; fa.ld = o->ld; fa.m0 = o->m0; fa.Mloc = o->Mloc; fa.st = o->d_state;
    fa.Z = o->Z; fa.E = o->E; fa.esq = o->esq; fa.logp = o->logp; fa.t = t;
    AVI_CHECK(o->model->fused_step(fa));
    *taken = true;
    return AVI_OK;
}

int32_t avi_objective_forward_chunk(avi_obj* o, const float* lambda, int m0, int Mc, const ObjDeviceState* ov,
                                    float* sums_dev, bool lowrank_logq) {
    avi_ctx* ctx = o->ctx;
    // forward only: eps itself is not needed downstream (|eps_m|^2 is), so the mean-field sampler writes z alone --
    // the algorithmic 4 (2 D + D M) bytes of SURVEY.md 8(d) K1
    AVI_CHECK(avi_family_sample(o, lambda, o->Z, o->family == AVI_MEANFIELD ? nullptr : o->E, o->esq, Mc, m0, o->d_state, ov));
    AVI_CHECK(o->model->eval(o->Z, o->ld, Mc, o->logp, nullptr));
    // low-rank family: the second sum is sum_m log q(z_m) (through the capacitance inverse prepared by avi_lr_entropy)
    if (lowrank_logq) AVI_CHECK(avi_lr_logq(o, lambda, Mc, /*forward_only=*/true));
    k_forward_sums<<<1, 1024, 0, ctx->stream>>>(lambda, o->D, o->family == AVI_FULLRANK, o->logp, o->esq, Mc, sums_dev);
    AVI_LAUNCHED(ctx);
    return AVI_OK;
}

int32_t avi_objective_finalize(avi_obj* o, const float* lambda, float* grad, float* out, bool skip_fr_matrix,
                               bool fuse_advance) {
    avi_ctx* ctx = o->ctx;
    const int D = o->D, accv = o->accv;
    if (o->family == AVI_LOWRANK) {
        if (!avi_lr_needs_logq(o)) AVI_CHECK(avi_lr_entropy(o, lambda));   // (otherwise computed before the log q kernel)
        return avi_lr_finalize(o, lambda, grad, out);
    }
    if (o->family == AVI_MEANFIELD) {
        unsigned nb = (unsigned)std::min<int64_t>(ceil_div(D, 256), 64);
        // fuse_advance (estimate_gradient!): one CTA; gradient, scalars and completion flag written straight into the
        // pinned host buffer o->h_grad, step counter advanced by this kernel
        cudaError_t e = avi_launch_pdl(ctx, k_finalize_mf, dim3(fuse_advance ? 1u : nb), dim3(fuse_advance ? 1024u : 256u), 0,
                                       (const float*)o->acc, accv, lambda, D, o->M,
                                       o->objective, o->entropy, (const float*)o->logp, (const float*)o->esq, o->Mloc,
                                       avi_obj_defers_scalars(o) ? 1 : 0, grad, out,
                                       fuse_advance ? o->h_grad : (float*)nullptr,
                                       fuse_advance ? o->d_state : (ObjDeviceState*)nullptr, o->base.h0);
        if (e != cudaSuccess) AVI_FAIL(ctx, AVI_ERR_CUDA, std::string("finalize launch: ") + cudaGetErrorString(e));
        AVI_LAUNCHED(ctx);
    } else {
        const float* scal = o->acc + 4 * (size_t)accv;
        const float* C1 = scal + ACC_NSCAL;
        const float* C2 = C1 + (size_t)D * D;
        size_t n = (size_t)D * D;
        // the matrix part reads fbar from scal before k_finalize_fr_vec rewrites out[3]
        if (!skip_fr_matrix) {   // (the fused update kernel computes these entries on the fly)
            k_finalize_fr_mat<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(C1, C2, scal, lambda, D, o->M,
                                                                                  o->objective, o->entropy, grad);
            AVI_LAUNCHED(ctx);
        }
        // (already done by the pullback's preparation launch when the local phase could: avi_objective_local)
        const bool done = o->fr_vec_done && grad == o->grad && out == o->out;
        o->fr_vec_done = false;
        if (!done) {
            k_finalize_fr_vec<<<1, 1024, 0, ctx->stream>>>(o->acc, accv, lambda, D, o->M, o->objective, o->entropy, grad, out,
                                                           o->logp, o->esq, o->Mloc, avi_obj_defers_scalars(o) ? 1 : 0, o->base.h0);
            AVI_LAUNCHED(ctx);
        }
    }
    return AVI_OK;
}
