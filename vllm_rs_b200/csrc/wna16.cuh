// Shared pieces of the W4A16 / W8A16 kernels: the B200 tile-major weight layout, zero-point
// decoding for AWQ / GPTQ checkpoints and register dequantisation to fp16.
//
// Quantised formats follow what the reference loader implies
// (/root/reference/src/models/layers/wna16.rs:56-83,127-152):
//   GPTQ  qweight i32 [K/pack, N], element (k,n) = (w[k/pack][n] >> bits*(k%pack)) & mask
//         qzeros  i32 [K/G, N/pack], zero = field(n%pack) + 1, scales f16 [K/G, N]
//   AWQ   qweight i32 [K, N/8], nibble j of a word holds output 8c + {0,2,4,6,1,3,5,7}[j]
//         qzeros  i32 [K/G, N/8] same nibble order, no +1
//   w = (q - z) * s, rounded once to fp16.
#pragma once
#include "../../include/b200_vllm.h"
#include "common.cuh"

namespace b200 {

constexpr int kTileN = 128;    // out-channels per weight tile
constexpr int kChunkK = 64;    // K values per pipeline stage

struct WnaParams {
  const __half* a;        // [M, K]
  const uint32_t* bp;     // prepacked weight
  const __half* scales;   // [K/G, N]
  const uint32_t* qzeros; // checkpoint packing, may be NULL (GPTQ symmetric)
  const __half* bias;     // [N] or NULL
  __half* c;              // [M, N] (or [M, N/2] with SILU_MUL)
  int M, K, N, bits, group_size, qfmt, epilogue;
  int no_split;           // force one K split (direct fp16 store, no workspace traffic)
  void* workspace;
};

// position of output channel r (0..7) inside an AWQ-packed word
__device__ __forceinline__ int awq_nibble_of(int r) { return ((r & 1) << 2) | (r >> 1); }

// integer zero point of (group g, channel n)
template <int BITS>
__device__ __forceinline__ int load_zero(const uint32_t* __restrict__ qzeros, int qfmt, int g,
                                         int n, int N) {
  if (qzeros == nullptr) return 1 << (BITS - 1);
  constexpr int PACK = 32 / BITS;
  const uint32_t w = qzeros[static_cast<size_t>(g) * (N / PACK) + n / PACK];
  if (BITS == 4 && qfmt == B200_QFMT_AWQ) return (w >> (4 * awq_nibble_of(n & 7))) & 0xF;
  return static_cast<int>((w >> (BITS * (n % PACK))) & ((1u << BITS) - 1)) + 1;
}

__device__ __forceinline__ uint32_t lop3_or_and(uint32_t x, uint32_t mask, uint32_t orv) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(r) : "r"(x), "r"(mask), "r"(orv));  // (x&mask)|orv
  return r;
}
__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ __half2 u32_as_h2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

// Per-group dequant constants for one channel.
struct DqConst {
  __half2 s2;      // scale
  __half2 zlo;     // 1024 + z   (low-nibble path)
  __half2 zhi;     // 64 + z     (high-nibble path, value arrives as 1024 + 16 q)
};
__device__ __forceinline__ DqConst make_dq(__half s, int z) {
  DqConst d;
  d.s2 = __half2half2(s);
  d.zlo = __half2half2(__int2half_rn(1024 + z));
  d.zhi = __half2half2(__int2half_rn(64 + z));
  return d;
}

// One packed 4-bit word (tile-major nibble order: nibble p<4 -> j=2p, nibble p+4 -> j=2p+1)
// -> four half2 holding k offsets (0,1) (2,3) (4,5) (6,7), each fp16((q - z) * s).
__device__ __forceinline__ void dequant4_word(uint32_t x, const DqConst& d, uint32_t (&out)[4]) {
  const __half2 sixteenth = __float2half2_rn(0.0625f);
  const uint32_t y = x >> 8;
  __half2 v0 = u32_as_h2(lop3_or_and(x, 0x000F000Fu, 0x64006400u));
  __half2 v1 = u32_as_h2(lop3_or_and(x, 0x00F000F0u, 0x64006400u));
  __half2 v2 = u32_as_h2(lop3_or_and(y, 0x000F000Fu, 0x64006400u));
  __half2 v3 = u32_as_h2(lop3_or_and(y, 0x00F000F0u, 0x64006400u));
  v0 = __hsub2(v0, d.zlo);
  v1 = __hfma2(v1, sixteenth, __hneg2(d.zhi));
  v2 = __hsub2(v2, d.zlo);
  v3 = __hfma2(v3, sixteenth, __hneg2(d.zhi));
  out[0] = h2_as_u32(__hmul2(v0, d.s2));
  out[1] = h2_as_u32(__hmul2(v1, d.s2));
  out[2] = h2_as_u32(__hmul2(v2, d.s2));
  out[3] = h2_as_u32(__hmul2(v3, d.s2));
}

// One packed 8-bit word (natural byte order, k offsets 0..3) -> two half2.
__device__ __forceinline__ void dequant8_word(uint32_t x, const DqConst& d, uint32_t (&out)[2]) {
  uint32_t lo, hi;
  asm("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(lo) : "r"(x), "r"(0x64646464u));
  asm("prmt.b32 %0, %1, %2, 0x4342;" : "=r"(hi) : "r"(x), "r"(0x64646464u));
  out[0] = h2_as_u32(__hmul2(__hsub2(u32_as_h2(lo), d.zlo), d.s2));
  out[1] = h2_as_u32(__hmul2(__hsub2(u32_as_h2(hi), d.zlo), d.s2));
}

// ---------------------------------------------------------------- split-K meeting point
// Workspace layout (b200_wna16_workspace_bytes): [0, 16 KB) tile arrival counters, zero at rest
// and self-resetting; then fp32 partials [splits][M][N] with no invariant (nothing to zero).
constexpr int kCounterBytes = 16384;
constexpr int kMaxSplitCols = 148 * 128 * 2;     // launchers keep splits * N below this

struct SplitOut {
  float* partials;
  uint32_t* counters;
  const __half* bias;
  __half* c;
  int M, N, splits, epilogue;
};

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Called by `nthreads` (128 or 512) threads of a CTA AFTER the partials of rows [m_lo, m_hi) of
// tile `nt` were stored and fenced by their writers.  tid % 128 = channel, tid / 128 = row phase.
// The last CTA to arrive for the tile (or, with SILU_MUL, for the gate/up tile pair) sums the
// partials in split order -- deterministic -- applies bias / SiLU*mul and writes fp16.  Loads of
// up to 8 rows are issued together so the L2 latency is paid `splits` times, not M * splits.
// `s_flag` is a shared int; `counter_base` separates token tiles; `bar_id` a free named barrier.
__device__ __forceinline__ void split_arrive_and_reduce(const SplitOut& o, int nt, int tid,
                                                        int nthreads, int m_lo, int m_hi,
                                                        int counter_base, int* s_flag, int bar_id) {
  const int N = o.N, M = o.M;
  const int tiles = N / kTileN;
  const bool silu = o.epilogue == B200_EPI_SILU_MUL;
  const int half_tiles = tiles >> 1;
  const int cidx = counter_base + (silu ? nt % half_tiles : nt);
  const unsigned target = silu ? 2u * o.splits : static_cast<unsigned>(o.splits);
  // Release / acquire through ONE thread (the cooperative-groups grid.sync pattern): the
  // barrier orders every writer's stores before thread 0's fence + atomic (cumulativity), and
  // thread 0's acquire fence before the second barrier orders the readers' loads (ld.cg, no L1).
  named_bar_sync(bar_id, nthreads);
  if (tid == 0) {
    __threadfence();
    const unsigned old = atomicAdd(&o.counters[cidx], 1u);
    const int last = (old == target - 1) ? 1 : 0;
    if (last) {
      o.counters[cidx] = 0;                          // self-reset (graph replay safe)
      __threadfence();
    }
    *s_flag = last;
  }
  named_bar_sync(bar_id, nthreads);
  if (!*s_flag) return;
  const int ch = tid & 127, phase = tid >> 7, nphase = nthreads >> 7;
  constexpr int R = 8;                               // rows in flight per thread
  const size_t row_stride = static_cast<size_t>(nphase) * N;
  if (silu) {
    const int ng = (nt % half_tiles) * kTileN + ch, nu = ng + (N >> 1);
    const float bg = o.bias ? __half2float(o.bias[ng]) : 0.f;
    const float bu = o.bias ? __half2float(o.bias[nu]) : 0.f;
    for (int mb = m_lo + phase; mb < m_hi; mb += R * nphase) {
      float g[R], u[R];
#pragma unroll
      for (int j = 0; j < R; ++j) { g[j] = bg; u[j] = bu; }
      for (int s = 0; s < o.splits; ++s) {
        const float* p = o.partials + (static_cast<size_t>(s) * M + mb) * N;
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (mb + j * nphase < m_hi) {
            g[j] += __ldcg(p + j * row_stride + ng);
            u[j] += __ldcg(p + j * row_stride + nu);
          }
      }
#pragma unroll
      for (int j = 0; j < R; ++j)
        if (mb + j * nphase < m_hi) {
          // round the two GEMM outputs to fp16 first: the values an unfused pipeline would see
          const float gg = __half2float(__float2half_rn(g[j]));
          const float uu = __half2float(__float2half_rn(u[j]));
          o.c[static_cast<size_t>(mb + j * nphase) * (N >> 1) + ng] =
              __float2half_rn(gg / (1.f + __expf(-gg)) * uu);
        }
    }
  } else {
    const int n = nt * kTileN + ch;
    const float bb = o.bias ? __half2float(o.bias[n]) : 0.f;
    for (int mb = m_lo + phase; mb < m_hi; mb += R * nphase) {
      float acc[R];
#pragma unroll
      for (int j = 0; j < R; ++j) acc[j] = bb;
      for (int s = 0; s < o.splits; ++s) {
        const float* p = o.partials + (static_cast<size_t>(s) * M + mb) * N + n;
#pragma unroll
        for (int j = 0; j < R; ++j)
          if (mb + j * nphase < m_hi) acc[j] += __ldcg(p + j * row_stride);
      }
#pragma unroll
      for (int j = 0; j < R; ++j)
        if (mb + j * nphase < m_hi)
          o.c[static_cast<size_t>(mb + j * nphase) * N + n] = __float2half_rn(acc[j]);
    }
  }
}

// launchers implemented in wna16_gemv.cu / wna16_tc.cu
void launch_wna16_gemv(const WnaParams& p, cudaStream_t st);
bool launch_wna16_tc(const WnaParams& p, cudaStream_t st);   // false: shape unsupported
bool launch_wna16_tc_dec(const WnaParams& p, cudaStream_t st);   // decode shapes (M <= 64)

}  // namespace b200
