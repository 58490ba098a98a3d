// Device helpers: Philox4x32-10, Box-Muller, deterministic block reductions.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define AVI_LOG2PI 1.8378770664093453f
#define AVI_H0 1.4189385332046727f   // entropy(Normal(0,1)) = (log 2pi + 1) / 2

enum { AVI_STREAM_EPS = 0, AVI_STREAM_SHUFFLE = 1, AVI_STREAM_DATA = 2, AVI_STREAM_EPS_FACTORS = 3 };

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

// uint32 -> (0,1), exactly representable in fp32 (oracle/philox.py: uniform23): (v + 0.5) * 2^-23 has 24
// significant bits, so the single fma is exact
// Computed WITHOUT an int -> float conversion: I2FP runs on the XU pipe next to the MUFU ops, and that pipe is what
// bounds the sampler (ncu: 87 % busy, 12 XU ops per Philox quad of which 4 were conversions).  1 + v 2^-23 is the float
// with mantissa bits v; subtracting 1 - 2^-24 (the largest float below 1) leaves v 2^-23 + 2^-24 exactly.
__device__ __forceinline__ float uniform23(uint32_t x) {
    return __uint_as_float(0x3f800000u | (x >> 9)) - 0.99999994f;
}

// Box-Muller pair from two Philox words, on the SFU pipes (MUFU lg2 / sqrt / sin / cos) and branch-free so
// that the sampling kernel stays HBM-bound (the first version, with a divergent series branch, the IEEE sqrtf
// and a compare/select angle fold, was issue-bound at 49 % of HBM peak: profiles/README.md).
//   radius: r = sqrt(-2 ln u0).  lg2.approx has a fixed ABSOLUTE error, so ln(u) loses relative accuracy
//           through cancellation as u -> 1; above 1 - 2^-5 the series of ln(1 + t), t = u0 - 1 (exact), is
//           selected instead (5 terms: relative error < 5e-9).
//   angle : 2 pi u1 folded into (-pi, pi), where sin.approx / cos.approx are most accurate.  u1 > 1/2 is the
//           top bit of the 23-bit integer, so the fold u1 - 1 is the SIGNED reading of the same bits and
//           th = fl(2 pi) * (vs + 0.5) * 2^-23 is one exact-product fma (bit-identical to 2 pi * (u1 - [u1 > 1/2])).
// Accuracy vs the fp64 oracle: |d eps| < 3e-6 (tests/test_gpu_parity.py::test_rand_matches_oracle).
__device__ __forceinline__ void box_muller_fast(uint32_t x0, uint32_t x1, float& n0, float& n1) {
    const float u0 = uniform23(x0);
    const float t = u0 - 1.0f;   // exact
    const float a_series = t * (-2.0f + t * (1.0f + t * (-0.66666669f + t * (0.5f + t * (-0.4f)))));
    float l2;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(u0));   // u0 >= 2^-24: never subnormal
    const float a_lg2 = -1.3862943611198906f * l2;
    const float a = u0 > 0.96875f ? a_series : a_lg2;   // -2 ln u0 > 0
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
    // (float)((int32_t)x1 >> 9) without I2FP (see uniform23): the signed 23-bit value vs plus 2^22 is the unsigned field
    // with its top bit flipped; 2^23 + that is a float by construction, and 2^23 + 2^22 subtracts exactly.
    // (sin / cos as quadrant + Taylor polynomials on the FMA pipe instead of MUFU -- 1.3e-7 max error -- made the sampler
    // SLOWER, 69 vs 56 us at M = 32768: it is bound by instruction issue along its dependency chains, not by the XU pipe.)
    const float vsf = __uint_as_float(0x4b000000u | ((x1 >> 9) ^ 0x400000u)) - 12582912.0f;
    const float th = fmaf(vsf, 6.283185307179586f * 1.1920928955078125e-07f, 6.283185307179586f * 5.9604644775390625e-08f);
    n0 = r * __cosf(th);
    n1 = r * __sinf(th);
}

// Philox round keys (k + r * Weyl constant): warp-uniform, hoisted out of the per-quad loop
struct PhiloxKeys {
    uint32_t k0[10], k1[10];
    __device__ __forceinline__ explicit PhiloxKeys(unsigned long long key) {
        uint32_t a = (uint32_t)key, b = (uint32_t)(key >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) { k0[r] = a; k1[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
    }
};

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                              const PhiloxKeys& pk, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t lo0, hi0, lo1, hi1;   // one IMAD.WIDE each
        asm("{\n\t.reg .b64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}"
            : "=r"(lo0), "=r"(hi0) : "r"(c0), "r"(0xD2511F53u));
        asm("{\n\t.reg .b64 p;\n\tmul.wide.u32 p, %2, %3;\n\tmov.b64 {%0, %1}, p;\n\t}"
            : "=r"(lo1), "=r"(hi1) : "r"(c2), "r"(0xCD9E8D57u));
        c0 = hi1 ^ c1 ^ pk.k0[r];
        c2 = hi0 ^ c3 ^ pk.k1[r];
        c1 = lo1;
        c3 = lo0;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// counter words 2 and 3 of the eps stream (oracle/philox.py): step low word, stream id | step high bits
__device__ __forceinline__ uint32_t eps_ctr3(unsigned long long step, uint32_t stream) {
    return (stream & 0xFFu) | ((uint32_t)((step >> 32) & 0xFFFFFFu) << 8);
}

// four standard normals for coordinates 4q .. 4q+3 of Monte-Carlo sample m at step `step`
__device__ __forceinline__ float4 normal4(uint32_t q, uint32_t m, uint32_t c2, uint32_t c3, const PhiloxKeys& pk) {
    uint32_t x[4];
    philox4x32_10(q, m, c2, c3, pk, x);
    float4 e;
    box_muller_fast(x[0], x[1], e.x, e.y);
    box_muller_fast(x[2], x[3], e.z, e.w);
    return e;
}
__device__ __forceinline__ float4 normal4(uint32_t q, uint32_t m, unsigned long long step, uint32_t stream,
                                          unsigned long long key) {
    const PhiloxKeys pk(key);
    return normal4(q, m, (uint32_t)step, eps_ctr3(step, stream), pk);
}

// programmatic dependent launch (see avi_launch_pdl): no-ops when the grid was launched without the attribute
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Step timeline (AVI_TIMELINE=1, diagnostic): every kernel of the fused iteration stamps %globaltimer into
// tl[16]: [id] first CTA entered, [4 + id] first CTA past its dependency wait, [8 + id] last CTA done
// (id: 0 sample, 1 forward, 2 backward, 3 tail).  Null pointer: nothing happens.
__device__ __forceinline__ unsigned long long tl_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void tl_min(unsigned long long* tl, int slot) {
    if (tl && threadIdx.x == 0 && threadIdx.y == 0) atomicMin(tl + slot, tl_now());
}
__device__ __forceinline__ void tl_max(unsigned long long* tl, int slot) {
    if (tl && threadIdx.x == 0 && threadIdx.y == 0) atomicMax(tl + slot, tl_now());
}

// Ask the memory system to pull [base, base + bytes) into L2 (cp.async.bulk.prefetch.L2, a per-warp instruction with
// a uniform address): warp `widx` of `nwarps` issues every nwarps-th chunk of `chunk` bytes.  Used to stream the NEXT
// kernel's static operand from HBM while the current kernel computes (an L2 hit costs a tag lookup).  base must be
// 16-byte aligned, chunk a multiple of 16.  Call with the whole warp.
__device__ __forceinline__ void l2_prefetch_span(const void* base, unsigned long long bytes, unsigned widx,
                                                 unsigned nwarps, unsigned chunk, unsigned pace_ns = 0) {
    const char* p = static_cast<const char*>(base);
    if ((threadIdx.x & 31) == 0) {
        for (unsigned long long off = (unsigned long long)widx * chunk; off < bytes; off += (unsigned long long)nwarps * chunk) {
            const unsigned long long left = bytes - off;
            const unsigned sz = (unsigned)(left < chunk ? left : chunk) & ~15u;
            if (sz) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p + off), "r"(sz) : "memory");
            if (pace_ns) __nanosleep(pace_ns);   // spread the requests so they do not queue ahead of demand loads
        }
    }
}
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
