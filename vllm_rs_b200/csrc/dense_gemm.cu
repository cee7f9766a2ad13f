// Dense fp16 GEMM for the unquantised linears on the path (lm_head; fp16 checkpoints):
//   C[m, n] = A[m, k] x W[n, k]^T, fp32 accumulate, fp16 or fp32 output.
// Reference: Linear::forward -> candle -> cuBLAS (/root/reference/src/models/layers/linear.rs:75-123)
// and the f32 logits cast in /root/reference/src/models/llama.rs:311-320.
// This is a plain library GEMM (cuBLAS), exactly what the reference uses for it.
#include <cublas_v2.h>

#include <mutex>

#include "../../include/b200_vllm.h"
#include "common.cuh"

namespace b200 {
static std::mutex g_cublas_mu;
static cublasHandle_t g_cublas[16] = {nullptr};
static cublasHandle_t cublas_for_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 16) return nullptr;
  std::lock_guard<std::mutex> lk(g_cublas_mu);
  if (!g_cublas[dev]) {
    if (cublasCreate(&g_cublas[dev]) != CUBLAS_STATUS_SUCCESS) {
      g_cublas[dev] = nullptr;
      return nullptr;
    }
    cublasSetMathMode(g_cublas[dev], CUBLAS_DEFAULT_MATH);
  }
  return g_cublas[dev];
}
}  // namespace b200

extern "C" void b200_dense_gemm_f16(const void* a, const void* w, void* c, int32_t m, int32_t n,
                                    int32_t k, int32_t out_f32, int64_t stream) {
  using namespace b200;
  if (m <= 0) return;
  B200_CHECK_ARG(a && w && c && n > 0 && k > 0, "dense_gemm: bad arguments");
  cublasHandle_t h = cublas_for_device();
  if (!h) {
    set_error(2, "dense_gemm: cublasCreate failed");
    return;
  }
  cublasSetStream(h, reinterpret_cast<cudaStream_t>(stream));
  const float alpha = 1.f, beta = 0.f;
  cublasStatus_t s = cublasGemmEx(h, CUBLAS_OP_T, CUBLAS_OP_N, n, m, k, &alpha, w, CUDA_R_16F, k, a,
                                  CUDA_R_16F, k, &beta, c, out_f32 ? CUDA_R_32F : CUDA_R_16F, n,
                                  CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
  count_launch(1);
  if (s != CUBLAS_STATUS_SUCCESS) set_error(2, "dense_gemm: cublasGemmEx failed");
}
