// Internal: launchers of the training-step kernels (train_gemm.cu, train_ops.cu); the C ABI is include/codlad_b200_train.h.
#pragma once
#include <cuda_runtime.h>

namespace cb2 {
namespace train {

int gemm(const float* A, const float* B, float* C, int M, int N, int K, long long lda, long long ldb, long long ldc, int a_kc, int b_kc,
         int accumulate, cudaStream_t s);

}  // namespace train
}  // namespace cb2
