// Shared device/host helpers for the codlad_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define CB2_H 128            // hidden width of the denoiser
#define CB2_MOD_ENC 1152     // 9*H adaLN outputs per encoder layer
#define CB2_MOD_DEC 768      // 6*H per decoder layer
#define CB2_MOD_FIN 256      // 2*H final layer
#define CB2_MOD_TOTAL (3 * CB2_MOD_ENC + 3 * CB2_MOD_DEC + CB2_MOD_FIN)   // 6016
#define CB2_MOD_ENC_OFF(l) ((l) * CB2_MOD_ENC)
#define CB2_MOD_DEC_OFF(l) (3 * CB2_MOD_ENC + (l) * CB2_MOD_DEC)
#define CB2_MOD_FIN_OFF (3 * CB2_MOD_ENC + 3 * CB2_MOD_DEC)

namespace cb2 {

void set_error(const char* fmt, ...);

#define CB2_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (call);                                                                \
        if (_e != cudaSuccess) {                                                                \
            cb2::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
            return (int)_e;                                                                     \
        }                                                                                       \
    } while (0)

#define CB2_LAUNCH_CHECK() CB2_CUDA(cudaGetLastError())

__device__ __forceinline__ float gelu_erf(float x) {
    // torch.nn.GELU() default (approximate='none'): 0.5 x (1 + erf(x / sqrt(2)))
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

__device__ __forceinline__ float silu(float x) { return x / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// LayerNorm statistics of one 128-wide row held as 4 values per lane of one warp.
// Matches torch.layer_norm: biased variance, rstd = 1/sqrt(var + eps).
__device__ __forceinline__ void warp_ln_stats(const float v[4], float eps, float& mean, float& rstd) {
    float s = warp_sum((v[0] + v[1]) + (v[2] + v[3]));
    mean = s * (1.0f / 128.0f);
    float d0 = v[0] - mean, d1 = v[1] - mean, d2 = v[2] - mean, d3 = v[3] - mean;
    float q = warp_sum((d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3));
    rstd = rsqrtf(q * (1.0f / 128.0f) + eps);
}

}  // namespace cb2
