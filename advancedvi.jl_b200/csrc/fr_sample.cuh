// eps-draw of the full-rank family for ONE Monte-Carlo sample by 128 threads (4 warps), bit-for-bit what
// k_sample<FULLRANK, no hook, SPLIT = SAMPLE_WARPS> (family.cu) does for its sample: thread t owns the coordinate quads
// t, t + 128, ...; E row (zero padding columns), the 3xTF32 split of the row [hi | lo | hi] in segments of `seg` (the B
// operand of z = L eps, family_fr.cu) and |eps|^2 by a fixed tree (warp shuffles, then the four warp totals in order).
// Used by the tiled full-rank update kernel (opt.cu) to draw the NEXT iteration's eps while it streams the optimiser
// state: eps does not depend on lambda, only on (key, step, sample, coordinate).
#pragma once

#include "device_utils.cuh"
#include "tc_common.cuh"

// t = thread index within the group of 128 (0..127); sm4 = 4 floats of shared memory owned by the group.
// The caller executes __syncthreads() between fr_sample_row_part and fr_sample_row_finish (all threads of the CTA).
__device__ __forceinline__ float fr_sample_row_part(int t, int m, int m_global, int D, int ld, int seg, uint32_t c2, uint32_t c3,
                                                    const PhiloxKeys& pk, float* __restrict__ E, float* __restrict__ Er3) {
    float part = 0.0f;
    float* Erow = E + (size_t)m * ld;
    float* Er3row = Er3 ? Er3 + (size_t)m * 3 * seg : nullptr;
    const int qend = (Er3row && seg > ld ? seg : ld) / 4;
    for (int q = t; q < qend; q += 128) {
        const int i = 4 * q;
        if (i >= ld) {   // zero tail of the split rows beyond the sample buffers' pitch
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(Er3row + i) = z4;
            *reinterpret_cast<float4*>(Er3row + seg + i) = z4;
            *reinterpret_cast<float4*>(Er3row + 2 * seg + i) = z4;
            continue;
        }
        const float4 e = normal4((uint32_t)q, (uint32_t)m_global, c2, c3, pk);
        float ev[4] = {e.x, e.y, e.z, e.w};
        if (i + 3 >= D) {
#pragma unroll
            for (int c = 0; c < 4; ++c) ev[c] = i + c < D ? ev[c] : 0.0f;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) part = fmaf(ev[c], ev[c], part);
        *reinterpret_cast<float4*>(Erow + i) = make_float4(ev[0], ev[1], ev[2], ev[3]);
        if (Er3row) {
            float hi[4], lo[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) { hi[c] = tc::round_tf32(ev[c]); lo[c] = tc::round_tf32(ev[c] - hi[c]); }
            *reinterpret_cast<float4*>(Er3row + i) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(Er3row + seg + i) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            *reinterpret_cast<float4*>(Er3row + 2 * seg + i) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        }
    }
    return warp_sum(part);
}
