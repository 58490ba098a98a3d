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

// four standard normals for coordinates 4q .. 4q+3 of Monte-Carlo sample m at step `step`
__device__ __forceinline__ float4 normal4(uint32_t q, uint32_t m, unsigned long long step, uint32_t stream,
                                          unsigned long long key) {
    uint32_t x[4];
    philox4x32_10(q, m, (uint32_t)step, (stream & 0xFFu) | ((uint32_t)((step >> 32) & 0xFFFFFFu) << 8),
                  (uint32_t)key, (uint32_t)(key >> 32), x);
    float4 e;
    float r0 = sqrtf(-2.0f * logf(uniform23(x[0])));
    float s0, c0;
    sincospif(2.0f * uniform23(x[1]), &s0, &c0);
    e.x = r0 * c0; e.y = r0 * s0;
    float r1 = sqrtf(-2.0f * logf(uniform23(x[2])));
    float s1, c1;
    sincospif(2.0f * uniform23(x[3]), &s1, &c1);
    e.z = r1 * c1; e.w = r1 * s1;
    return e;
}

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
