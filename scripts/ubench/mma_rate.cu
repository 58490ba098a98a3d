// Micro-benchmark: issue rate of tcgen05.mma.kind::tf32 (SS mode, M = 128) as a function of N, with and
// without a tcgen05.commit per 4 MMAs.  One CTA per SM, operands = whatever is in shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../advancedvi.jl_b200/csrc mma_rate.cu -o mma_rate -lcuda
#include <cstdio>
#include <cstdlib>
#include "tc_common.cuh"

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db),
                 "r"(idesc), "r"(acc) : "memory");
}

template <int MODE>   // 0: tf32 back-to-back, 1: tf32 commit every 4, 2: bf16 back-to-back (kind::f16)
__global__ void __launch_bounds__(128, 1) k(int N, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* tiles = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar[8];
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 48 * 1024 * 4 / 4; i += 128) reinterpret_cast<float*>(tiles)[i] = 0.f;
    if (threadIdx.x == 0) { for (int s = 0; s < 8; ++s) tc::mbar_init(&bar[s], 1); tc::mbar_fence_init(); }
    if (warp == 1) tc::tmem_alloc(&tbase, 512);
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const uint32_t tm = tbase;
    if (warp == 0 && tc::elect_one()) {
        // bf16 idesc: D=F32 (bit4), A=B=BF16 (1 at bits 7, 10)
        const uint32_t idesc = MODE == 2 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24))
                                         : tc::idesc_tf32(128, N);
        const int stage_bytes = 48 * 1024;
        long long t0 = clock64();
        uint32_t ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < iters; ++i) {
            const int st = i & 3;
            const uint32_t sa = tc::smem_u32(tiles + st * stage_bytes);
            const uint64_t da = tc::smem_desc_k_sw128(sa), db = tc::smem_desc_k_sw128(sa + 16384);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                if (MODE == 2) umma_bf16(tm, da + 2 * kk, db + 2 * kk, idesc, 1u);
                else tc::umma_tf32(tm, da + 2 * kk, db + 2 * kk, idesc, 1u);
            }
            if (MODE == 1) tc::umma_commit(&bar[1 + st]);
        }
        tc::umma_commit(&bar[0]);
        tc::mbar_wait(&bar[0], 0);
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
        (void)ph;
    }
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tm, 512);
}

int main() {
    long long* d; cudaMalloc(&d, 148 * 8);
    long long h[148];
    const int smem = 4 * 48 * 1024 + 2048;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 256;   // x4 MMAs
    for (int grid : {1, 148})
    for (int mode = 0; mode < 3; ++mode)
        for (int N : {64, 128, 144, 192, 256}) {
            for (int rep = 0; rep < 2; ++rep) {
                if (mode == 0) k<0><<<grid, 128, smem>>>(N, iters, d);
                if (mode == 1) k<1><<<grid, 128, smem>>>(N, iters, d);
                if (mode == 2) k<2><<<grid, 128, smem>>>(N, iters, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("err %s\n", cudaGetErrorString(e)); return 1; }
            }
            cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("grid %3d mode %d (%s) N %3d: %.1f cycles per MMA (M=128, K=%d)\n", grid, mode,
                   mode == 0 ? "tf32" : mode == 1 ? "tf32+commit/4" : "bf16", N, (double)mx / (iters * 4), mode == 2 ? 16 : 8);
        }
    return 0;
}
