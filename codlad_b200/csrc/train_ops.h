// Internal: launchers of the training-step kernels (train_gemm.cu, train_ops.cu); the C ABI is include/codlad_b200_train.h.
#pragma once
#include <cuda_runtime.h>

namespace cb2 {
namespace train {

int gemm(const float* A, const float* B, float* C, int M, int N, int K, long long lda, long long ldb, long long ldc, int a_kc, int b_kc,
         int accumulate, cudaStream_t s);

// TF32 tensor-core paths (train_gemm_tc.cu)
bool tc_shape_ok(const float* A, const float* B, const float* C, int M, int N, int K, long long lda, long long ldb, long long ldc, int a_kc, int b_kc);
int gemm_tc_nt(const float* A, const float* B, float* C, int M, int N, int K, long long lda, long long ldb, long long ldc, int accumulate, cudaStream_t s);
int gemm_tc_nt_bias_gelu(const float* A, const float* B, const float* bias, float* Z, float* Y, int M, int N, int K, long long lda, long long ldb,
                         long long ldz, cudaStream_t s);
int gemm_tc_tn(const float* A, const float* B, float* C, int P, int Q, int R, long long lda, long long ldb, long long ldc, int accumulate, cudaStream_t s);
extern int g_gemm_mode;      // 0 = fp32 SIMT everywhere, 1 = TF32 tensor cores where the shape allows

}  // namespace train
}  // namespace cb2
