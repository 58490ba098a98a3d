// Device helpers: Philox4x32-10, Box-Muller, deterministic block reductions.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define AVI_LOG2PI 1.8378770664093453f
#define AVI_H0 1.4189385332046727f   // entropy(Normal(0,1)) = (log 2pi + 1) / 2

enum { AVI_STREAM_EPS = 0, AVI_STREAM_SHUFFLE = 1, AVI_STREAM_DATA = 2 };

// Philox4x32-10 (Salmon et al., SC'11).  Same counter/key convention as oracle/philox.py.
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                       uint32_t k0, uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// uint32 -> (0,1), exactly representable in fp32 (oracle/philox.py: uniform23)
__device__ __forceinline__ float uniform23(uint32_t x) {
    return ((float)(x >> 9) + 0.5f) * 1.1920928955078125e-07f;   // 2^-23
}

// Box-Muller pair from two uniforms in (0,1), on the SFU pipes (MUFU lg2 / rsq / sin / cos) so that the
// sampling kernel stays HBM-bound.  ln(u) loses absolute accuracy through cancellation as u -> 1
// (lg2.approx has a fixed absolute error), so that range uses the series of ln(1 + t), t = u - 1.
// Accuracy vs the fp64 oracle: |d eps| < 3e-6 (tests/test_gpu_parity.py::test_rand_matches_oracle).
__device__ __forceinline__ void box_muller_fast(float u0, float u1, float& n0, float& n1) {
    float ln_u;
    if (u0 > 0.875f) {
        const float t = u0 - 1.0f;   // exact
        ln_u = t * (1.0f + t * (-0.5f + t * (0.33333334f + t * (-0.25f + t * (0.2f + t * (-0.16666667f +
               t * (0.14285715f + t * (-0.125f))))))));
    } else {
        ln_u = 0.6931471805599453f * __log2f(u0);
    }
    const float r = sqrtf(-2.0f * ln_u);
    // angle 2 pi u1 folded into (-pi, pi] where sin.approx / cos.approx are most accurate
    const float th = 6.283185307179586f * (u1 - (u1 > 0.5f ? 1.0f : 0.0f));
    n0 = r * __cosf(th);
    n1 = r * __sinf(th);
}

// four standard normals for coordinates 4q .. 4q+3 of Monte-Carlo sample m at step `step`
__device__ __forceinline__ float4 normal4(uint32_t q, uint32_t m, unsigned long long step, uint32_t stream,
                                          unsigned long long key) {
    uint32_t x[4];
    philox4x32_10(q, m, (uint32_t)step, (stream & 0xFFu) | ((uint32_t)((step >> 32) & 0xFFFFFFu) << 8),
                  (uint32_t)key, (uint32_t)(key >> 32), x);
    float4 e;
    box_muller_fast(uniform23(x[0]), uniform23(x[1]), e.x, e.y);
    box_muller_fast(uniform23(x[2]), uniform23(x[3]), e.z, e.w);
    return e;
}

// programmatic dependent launch (see avi_launch_pdl): no-ops when the grid was launched without the attribute
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Deterministic block sum (fixed tree); result valid in every thread.  blockDim.x multiple of 32,
// at most 1024 threads.  `sm` must hold 33 floats.
__device__ __forceinline__ float block_sum(float v, float* sm) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();   // protect sm reuse across consecutive calls
    if (lane == 0) sm[w] = v;
    __syncthreads();
    if (w == 0) {
        float t = lane < nw ? sm[lane] : 0.0f;
        t = warp_sum(t);
        if (lane == 0) sm[32] = t;
    }
    __syncthreads();
    return sm[32];
}

__device__ __forceinline__ float softplus_f(float x) {   // log1pexp
    return fmaxf(x, 0.0f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float sigmoid_f(float x) {
    float e = expf(-fabsf(x));
    float r = 1.0f / (1.0f + e);
    return x >= 0.0f ? r : e * r;
}
