// Device-side pieces of the one-shot all-reduce over NVLink peer memory (see comm.cu), shared with the
// mean-field tail kernel, which fuses the exchange into the finalize + update launch (mf_tail.cuh).
#pragma once

#include <stdint.h>

constexpr int AVI_MAX_RANKS = 16;
constexpr int AVI_FLAG_WORDS = 64;   // flag area: AVI_MAX_RANKS words used, padded to 256 B

struct CommDev {
    unsigned int seq, arrive, depart, pad;
};

struct PeerTable {
    float* data[AVI_MAX_RANKS];          // slot 0 of rank r (slot 1 follows at +slot_stride floats)
    unsigned int* flags[AVI_MAX_RANKS];  // flags[r][q] : rank q published sequence number ...
};

// what a fused kernel needs to run the exchange itself
struct CommPeers {
    int nranks, rank;
    long long slot_stride;
    PeerTable t;
    CommDev* dev;
};

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
