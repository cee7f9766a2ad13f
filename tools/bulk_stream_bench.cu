// Micro-benchmark: how fast can ONE producer thread per SM stream HBM into a shared-memory ring
// with cp.async.bulk (the weight path of wna16_tc_dec.cu)?  Sweeps copy size and ring depth.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_stream_bench bulk_stream_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n));
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0,1,0,P;\n\t}\n"
                 : "=r"(ok) : "r"(s32(b)), "r"(ph) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory");
}
__device__ __forceinline__ void expect_tx(uint64_t* b, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk(void* dst, const void* src, uint32_t bytes, uint64_t* b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(s32(dst)), "l"(src), "r"(bytes), "r"(s32(b)) : "memory");
}

template <int BYTES, int STAGES>
__global__ void __launch_bounds__(64, 1) stream_kernel(const uint8_t* src, int n, int consumers, unsigned* sink) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[STAGES], empty[STAGES];
  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const uint8_t* base = src + (size_t)blockIdx.x * n * BYTES;
  if (threadIdx.x == 0) {
    for (int i = 0; i < n; ++i) {
      const int s = i % STAGES;
      mbar_wait(&empty[s], ((i / STAGES) & 1) ^ 1);
      expect_tx(&full[s], BYTES);
      bulk(smem + s * BYTES, base + (size_t)i * BYTES, BYTES, &full[s]);
    }
  } else if (threadIdx.x == 32) {
    unsigned acc = 0;
    for (int i = 0; i < n; ++i) {
      const int s = i % STAGES;
      mbar_wait(&full[s], (i / STAGES) & 1);
      acc += *reinterpret_cast<volatile unsigned*>(smem + s * BYTES);
      mbar_arrive(&empty[s]);
    }
    if (acc == 0x12345678u) *sink = acc;
  }
}

template <int BYTES, int STAGES>
void run(const uint8_t* buf, size_t total, unsigned* sink, int ctas) {
  const int n = (int)(total / ctas / BYTES);
  const size_t smem = (size_t)BYTES * STAGES;
  cudaFuncSetAttribute(stream_kernel<BYTES, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  stream_kernel<BYTES, STAGES><<<ctas, 64, smem>>>(buf, n, 1, sink);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 3; ++r) stream_kernel<BYTES, STAGES><<<ctas, 64, smem>>>(buf, n, 1, sink);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double bytes = 3.0 * ctas * (double)n * BYTES;
  printf("ctas=%d bytes/copy=%5d stages=%2d in-flight=%3zu KB/SM : %7.1f GB/s  (%.2f GB/s per SM)  err=%s\n", ctas, BYTES, STAGES,
         smem / 1024, bytes / ms / 1e6, bytes / ms / 1e6 / ctas, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const size_t total = (size_t)1536 << 20;
  uint8_t* buf; unsigned* sink;
  cudaMalloc(&buf, total); cudaMalloc(&sink, 4);
  cudaMemset(buf, 1, total);
  run<4096, 16>(buf, total, sink, 148);
  run<4096, 32>(buf, total, sink, 148);
  run<4096, 48>(buf, total, sink, 148);
  run<8192, 8>(buf, total, sink, 148);
  run<8192, 16>(buf, total, sink, 148);
  run<8192, 24>(buf, total, sink, 148);
  run<16384, 4>(buf, total, sink, 148);
  run<16384, 8>(buf, total, sink, 148);
  run<16384, 12>(buf, total, sink, 148);
  run<32768, 6>(buf, total, sink, 148);
  run<2048, 32>(buf, total, sink, 148);
  run<4096, 16>(buf, total, sink, 296);
  return 0;
}
