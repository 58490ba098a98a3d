// Device-side pieces of the one-shot all-reduce over NVLink peer memory (see comm.cu), shared with the
// mean-field tail kernel, which fuses the exchange into the finalize + update launch (mf_tail.cuh).
#pragma once

#include <stdint.h>

constexpr int AVI_MAX_RANKS = 16;
constexpr int AVI_FLAG_WORDS = 64;   // flag area: AVI_MAX_RANKS words used, padded to 256 B

struct CommDev {
    unsigned int seq, arrive, depart, pad;
};

// One 8-byte word of the low-latency ("LL") receive area: a payload float and the sequence number of the exchange
// it belongs to, written by ONE 8-byte store, so a reader that sees the expected sequence number has the value
// (no fence, no separate flag: the exchange costs one NVLink store latency).
struct LLWord { float v; unsigned int seq; };

struct PeerTable {
    float* data[AVI_MAX_RANKS];          // slot 0 of rank r (slot 1 follows at +slot_stride floats)
    unsigned int* flags[AVI_MAX_RANKS];  // flags[r][q] : rank q published sequence number ...
    LLWord* ll[AVI_MAX_RANKS];           // LL receive area of rank r: [parity 2][source rank AVI_MAX_RANKS][ll_cap]
};

// what a fused kernel needs to run the exchange itself
struct CommPeers {
    int nranks, rank;
    long long slot_stride;
    long long ll_cap;                    // floats per (parity, source) LL lane; 0: LL not available
    PeerTable t;
    CommDev* dev;
};

__device__ __forceinline__ LLWord* ll_lane(const CommPeers& c, int owner, unsigned int seq, int src) {
    return c.t.ll[owner] + ((size_t)(seq & 1u) * AVI_MAX_RANKS + src) * (size_t)c.ll_cap;
}
__device__ __forceinline__ void st_ll(LLWord* p, float v, unsigned int seq) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(seq) : "memory");
}
__device__ __forceinline__ LLWord ld_ll(const LLWord* p) {
    unsigned int a, b;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "l"(p) : "memory");
    LLWord w; w.v = __uint_as_float(a); w.seq = b;
    return w;
}
// push my value of element i to every peer's receive lane for me (fire and forget)
__device__ __forceinline__ void ll_push(const CommPeers& c, unsigned int seq, long long i, float v) {
    for (int r = 0; r < c.nranks; ++r)
        if (r != c.rank) st_ll(ll_lane(c, r, seq, c.rank) + i, v, seq);
}
// out[k] = sum over ranks IN RANK ORDER of element idx[k] (own[k] is this rank's value): identical bits on every
// rank.  Per peer the N words are requested together (independent loads) and re-polled until all have arrived.
template <int N>
__device__ __forceinline__ void ll_gather(const CommPeers& c, unsigned int seq, const long long (&idx)[N],
                                          const bool (&need)[N], const float (&own)[N], float (&out)[N]) {
#pragma unroll
    for (int k = 0; k < N; ++k) out[k] = 0.f;
    for (int r = 0; r < c.nranks; ++r) {
        if (r == c.rank) {
#pragma unroll
            for (int k = 0; k < N; ++k) out[k] += need[k] ? own[k] : 0.f;
            continue;
        }
        const LLWord* lane = ll_lane(c, c.rank, seq, r);
        LLWord w[N];
        bool all;
        do {
            all = true;
#pragma unroll
            for (int k = 0; k < N; ++k)
                if (need[k]) w[k] = ld_ll(lane + idx[k]);
#pragma unroll
            for (int k = 0; k < N; ++k) all = all && (!need[k] || w[k].seq == seq);
        } while (!all);
#pragma unroll
        for (int k = 0; k < N; ++k) out[k] += need[k] ? w[k].v : 0.f;
    }
}

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_v4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
