// fp32 SIMT building block used by the fp32-tier kernels: one thread owns ONE output column of a
// 128-wide layer for ROWS rows of a shared-memory tile.  Weights are read transposed ([k][128],
// coalesced across the warp, reused across ROWS rows from a register); activations are read as
// warp-wide broadcast float4 from shared memory.
#pragma once
#include "common.cuh"

namespace cb2 {

// acc[r] += sum_k sIn[r*LD + k] * Wt[k*ldw]   for k in [0, KD); KD % 4 == 0, LD % 4 == 0.
// `Wt` already points at this thread's column.
template <int ROWS, int KD, int LD>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ sIn, const float* __restrict__ Wt, int ldw,
                                          float (&acc)[ROWS]) {
#pragma unroll 2
    for (int k = 0; k < KD; k += 4) {
        const float w0 = __ldg(Wt + (size_t)(k + 0) * ldw);
        const float w1 = __ldg(Wt + (size_t)(k + 1) * ldw);
        const float w2 = __ldg(Wt + (size_t)(k + 2) * ldw);
        const float w3 = __ldg(Wt + (size_t)(k + 3) * ldw);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
            const float4 a = *reinterpret_cast<const float4*>(sIn + r * LD + k);
            acc[r] = fmaf(a.x, w0, acc[r]);
            acc[r] = fmaf(a.y, w1, acc[r]);
            acc[r] = fmaf(a.z, w2, acc[r]);
            acc[r] = fmaf(a.w, w3, acc[r]);
        }
    }
}

template <int ROWS>
__device__ __forceinline__ void zero_acc(float (&acc)[ROWS]) {
#pragma unroll
    for (int r = 0; r < ROWS; ++r) acc[r] = 0.0f;
}

}  // namespace cb2
