// Dense fp16 GEMM for the unquantised linears on the path (lm_head; fp16 checkpoints):
//   C[m, n] = A[m, k] x W[n, k]^T, fp32 accumulate, fp16 or fp32 output.
// Reference: Linear::forward -> candle -> cuBLAS (/root/reference/src/models/layers/linear.rs:75-123)
// and the f32 logits cast in /root/reference/src/models/llama.rs:311-320.
// Default: the hand-written tcgen05 kernel of dense_gemm_tc.cu.  B200_DENSE=cublas (or mode 0) selects the
// library GEMM the reference itself calls -- ONLY as an explicit A/B arm: a shape the tcgen05 kernel does not
// take (K % 64 != 0, N % 8 != 0, unaligned pointers) latches an error, there is no silent library fallback.
// B200_DENSE=tc1 forces single-CTA tiles for M > 128.
#include <cublas_v2.h>

#include <cstdlib>
#include <cstring>
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
bool launch_dense_gemm_tc(const void* x, const void* w, void* c, const void* bias, int M, int N, int K,
                          int out_f32, int mode, cudaStream_t st);
// 0 = cuBLAS, 1 = tcgen05 single-CTA tiles, 2 = tcgen05 CTA pairs
static int dense_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("B200_DENSE");
    mode = 2;
    if (e && !strcmp(e, "cublas")) mode = 0;
    if (e && !strcmp(e, "tc1")) mode = 1;
  }
  return mode;
}
}  // namespace b200

extern "C" void b200_dense_gemm_ex(const void* a, const void* w, const void* bias, void* c, int32_t m,
                                   int32_t n, int32_t k, int32_t out_f32, int32_t mode, int64_t stream) {
  using namespace b200;
  if (m <= 0) return;
  B200_CHECK_ARG(a && w && c && n > 0 && k > 0, "dense_gemm: bad arguments");
  if (mode < 0) mode = dense_mode();
  if (mode != 0) {
    if (!launch_dense_gemm_tc(a, w, c, bias, m, n, k, out_f32, mode, reinterpret_cast<cudaStream_t>(stream)))
      set_error(1, "dense_gemm: shape not taken by the tcgen05 kernel (needs K % 64 == 0, N % 8 == 0, 16-byte "
                   "aligned pointers); the cuBLAS arm must be asked for explicitly (mode 0 / B200_DENSE=cublas)");
    return;
  }
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
  if (bias) {
    if (out_f32)
      set_error(1, "dense_gemm: bias with fp32 output is only fused in the tcgen05 kernel");
    else
      b200_add_bias_f16(c, bias, m, n, stream);
  }
}

extern "C" void b200_dense_gemm_f16(const void* a, const void* w, void* c, int32_t m, int32_t n,
                                    int32_t k, int32_t out_f32, int64_t stream) {
  b200_dense_gemm_ex(a, w, nullptr, c, m, n, k, out_f32, -1, stream);
}
