// Issue-rate micro-benchmark of the ALU instructions the dequant / e4m3 paths are made of (sm_100a).
// One CTA per SM, W warps per CTA, each thread runs ILP independent chains of N instructions; reports
// warp-instructions per cycle per SM sub-partition.  nvcc -arch=sm_100a -O3 -o build/alu_rate_bench tools/alu_rate_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>

constexpr int kIters = 4096, kIlp = 8;

#define CHAIN(NAME, ...)                                                                           \
  __global__ void NAME(uint32_t* out, long long* cycles, uint32_t seed) {                          \
    uint32_t r[kIlp];                                                                              \
    for (int i = 0; i < kIlp; ++i) r[i] = seed + threadIdx.x * 7 + i;                              \
    const uint32_t b = seed | 0x3c003c00u, c = 0x64006400u;                                        \
    __syncthreads();                                                                               \
    const long long t0 = clock64();                                                                \
    for (int it = 0; it < kIters; ++it) {                                                          \
      _Pragma("unroll") for (int i = 0; i < kIlp; ++i) { __VA_ARGS__ }                                \
    }                                                                                              \
    const long long t1 = clock64();                                                                \
    uint32_t acc = 0;                                                                              \
    for (int i = 0; i < kIlp; ++i) acc ^= r[i];                                                    \
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;                                              \
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;                                            \
  }

CHAIN(k_lop3, asm volatile("lop3.b32 %0, %0, %1, %2, 0xEA;" : "+r"(r[i]) : "r"(b), "r"(c));)
CHAIN(k_hfma2, asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));)
CHAIN(k_hmul2, asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(b));)
CHAIN(k_hadd2, asm volatile("sub.rn.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(c));)
CHAIN(k_ffma, asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*reinterpret_cast<float*>(&r[i])) : "f"(1.0001f), "f"(0.5f));)
CHAIN(k_imad, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(b), "r"(c));)
CHAIN(k_prmt, asm volatile("prmt.b32 %0, %0, %1, 0x6420;" : "+r"(r[i]) : "r"(b));)
CHAIN(k_shf, asm volatile("shr.u32 %0, %0, 1; add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));)
CHAIN(k_cvt_e4m3, asm volatile("{.reg .b16 l, h; mov.b32 {l, h}, %0; cvt.rn.f16x2.e4m3x2 %0, l;}" : "+r"(r[i]));)
// the dequant mix of one packed word: shift, 4 lop3, 2 sub, 2 fma, 4 mul (13 instructions per 8 weights)
CHAIN(k_mix, {
  uint32_t y, v0, v1, v2, v3;
  asm volatile("shr.u32 %0, %1, 8;" : "=r"(y) : "r"(r[i]));
  asm volatile("lop3.b32 %0, %1, 0x000F000F, %2, 0xEA;" : "=r"(v0) : "r"(r[i]), "r"(c));
  asm volatile("lop3.b32 %0, %1, 0x00F000F0, %2, 0xEA;" : "=r"(v1) : "r"(r[i]), "r"(c));
  asm volatile("lop3.b32 %0, %1, 0x000F000F, %2, 0xEA;" : "=r"(v2) : "r"(y), "r"(c));
  asm volatile("lop3.b32 %0, %1, 0x00F000F0, %2, 0xEA;" : "=r"(v3) : "r"(y), "r"(c));
  asm volatile("sub.rn.f16x2 %0, %0, %1;" : "+r"(v0) : "r"(c));
  asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(v1) : "r"(b), "r"(c));
  asm volatile("sub.rn.f16x2 %0, %0, %1;" : "+r"(v2) : "r"(c));
  asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(v3) : "r"(b), "r"(c));
  asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(v0) : "r"(b));
  asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(v1) : "r"(b));
  asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(v2) : "r"(b));
  asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(v3) : "r"(b));
  r[i] = v0 ^ v1 ^ v2 ^ v3;
})

template <typename K>
static void run(const char* name, K kern, int per_iter) {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&cyc, 148 * 8);
  for (int warps : {4, 8, 16, 32}) {
    kern<<<148, warps * 32>>>(out, cyc, 12345);
    kern<<<148, warps * 32>>>(out, cyc, 12345);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double instr = double(kIters) * kIlp * per_iter * warps / 4.0;   // warp-instructions per SMSP
    printf("%-10s %2d warps/SM (%d/SMSP): %.3f warp-instr/cycle/SMSP\n", name, warps, warps / 4, instr / avg);
  }
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  run("lop3", k_lop3, 1);
  run("hfma2", k_hfma2, 1);
  run("hmul2", k_hmul2, 1);
  run("hsub2", k_hadd2, 1);
  run("ffma", k_ffma, 1);
  run("imad", k_imad, 1);
  run("prmt", k_prmt, 1);
  run("shr+add", k_shf, 2);
  run("cvt.e4m3", k_cvt_e4m3, 1);
  run("w4 mix/13", k_mix, 14);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  return 0;
}
