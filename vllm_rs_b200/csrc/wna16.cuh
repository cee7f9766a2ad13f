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
#include <cuda_fp8.h>

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
  const struct FusedEpi* fe;   // optional fused epilogue description (host pointer, copied by value)
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
#if defined(B200_DQ_EXPERIMENT) && B200_DQ_EXPERIMENT == 2
  // TIMING EXPERIMENT ONLY (side builds, wrong values): the lop3-only core of a bias-folded dequant
  out[0] = h2_as_u32(v0); out[1] = h2_as_u32(v1); out[2] = h2_as_u32(v2); out[3] = h2_as_u32(v3);
  return;
#endif
  v0 = __hsub2(v0, d.zlo);
  v1 = __hfma2(v1, sixteenth, __hneg2(d.zhi));
  v2 = __hsub2(v2, d.zlo);
  v3 = __hfma2(v3, sixteenth, __hneg2(d.zhi));
#if defined(B200_DQ_EXPERIMENT) && B200_DQ_EXPERIMENT == 1
  // TIMING EXPERIMENT ONLY (side builds, unscaled values): the exact (q - z) core of a post-MMA-scaled dequant
  out[0] = h2_as_u32(v0); out[1] = h2_as_u32(v1); out[2] = h2_as_u32(v2); out[3] = h2_as_u32(v3);
  return;
#endif
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
constexpr int kSkSlotCols = 148 * 128 * 4;       // stream-K kernel (wna16_sk.cu): per-CTA partial slots, [M][128] fp32 each

// internal epilogue kinds beyond the public B200_EPI_* flags (used by the fused decode plan)
constexpr int kEpiResidNorm = 2;   // x += y (fp16), out = x * w_next, ssq_out[m][tile] = sum x^2 over the tile
constexpr int kSsqStride = 64;    // floats per row of sums of squares: one slot per 128-channel tile (hidden <=
                                  // 8192).  Tile finalisers STORE their slot (no atomics, nothing to zero) and the
                                  // consumer adds the slots in a fixed butterfly order: run-to-run deterministic
constexpr int kEpiRopeKv = 3;      // fused q|k|v: RoPE(q, k), K/V rows written to the paged cache
constexpr int kEpiPushPeers = 4;   // row-parallel GEMM of a TP rank: the fp16 partial tile is stored into
                                   // EVERY rank's receive buffer over NVLink and a per-source tile counter is
                                   // bumped there -- the all-reduce's transfer, tile by tile, under the GEMM
constexpr int kMaxTpRanks = 8;

// Everything the tile finaliser may fuse around a GEMM of the decoder layer
// (llama.rs:107-131, attention.rs:648-839, others.rs:11-66 restated as GEMM epilogues).
struct FusedEpi {
  // consumer side of a folded RMSNorm: row m of the accumulators is scaled by
  // rsqrt(ssq_in[m] * inv_dim + eps) (the per-token factor commutes with the GEMM)
  const float* ssq_in;
  int ssq_tiles;       // slots written per row of ssq_in (= producer N / 128)
  float inv_dim, eps;
  // kEpiResidNorm: residual stream (in/out), weight of the NEXT norm, its sum-of-squares target
  __half* resid;
  const __half* norm_w;
  float* ssq_out;
  // kEpiRopeKv
  const int64_t* positions;
  const int64_t* slots;
  const __half* cos;
  const __half* sin;
  __half* k_cache;
  __half* v_cache;
  int H, Hkv, D;
  int kv_fp8;      // caches hold e4m3 bytes (--fp8-kvcache): K/V rows are rounded to fp16, then to e4m3
  // kEpiPushPeers: rank r's receive slot for THIS rank's partial [M, N] and its tile counter
  __half* push[kMaxTpRanks];
  uint32_t* push_cnt[kMaxTpRanks];
  int push_world;
  int out_bf16;    // plain epilogue only: C (and the bias) are bf16 (b200_marlin_*_bf16, wna16_sk.cu)
};

struct SplitOut {
  float* partials;
  uint32_t* counters;
  const __half* bias;
  __half* c;
  int M, N, splits, epilogue;
  FusedEpi fe;
};

__device__ __forceinline__ bool epi_is_plain(const SplitOut& o) {
  return o.epilogue == B200_EPI_NONE && o.fe.ssq_in == nullptr;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---- tile finaliser.  A thread owns FOUR consecutive output channels of the 128-wide tile
// (one 16-byte load per split and row) and a warp owns whole rows: the partials of a row come in
// as 32 x float4, the sum of squares of a row is one warp reduction and one atomic.
struct EpiCol {
  int n0, n1;          // first of the 4 columns this thread sums, partner columns (up-projection /
                       // RoPE pair) or -1
  bool pair, kv_col, silu;
  float b0[4], b1[4], nw[4];
};
struct EpiRow {        // per-row operands that do not depend on the sums (requested early)
  float rs;
  float e0[4], e1[4];
  int sl;
  int64_t pos;
};

__device__ __forceinline__ void half4_to_float(const void* p, float (&f)[4]) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 lo = __half22float2(u32_as_h2(u.x)), hi = __half22float2(u32_as_h2(u.y));
  f[0] = lo.x; f[1] = lo.y; f[2] = hi.x; f[3] = hi.y;
}
__device__ __forceinline__ uint32_t bf16x2_bits(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void store_bf16x4(void* p, const float (&f)[4]) {
  uint2 u;
  u.x = bf16x2_bits(f[0], f[1]);
  u.y = bf16x2_bits(f[2], f[3]);
  *reinterpret_cast<uint2*>(p) = u;
}
__device__ __forceinline__ void store_half4(void* p, const float (&f)[4]) {
  uint2 u;
  u.x = h2_as_u32(__floats2half2_rn(f[0], f[1]));
  u.y = h2_as_u32(__floats2half2_rn(f[2], f[3]));
  *reinterpret_cast<uint2*>(p) = u;
}

// lane's share of a row's sum of squares (slots lane, lane + 32); warp_sum() of it is the row's total, added in
// the same order wherever it is consumed (GEMM finalisers, rowscale_kernel)
__device__ __forceinline__ float ssq_lane_partial(const float* __restrict__ row, int tiles, int lane) {
  float v = lane < tiles ? row[lane] : 0.f;
  if (lane + 32 < tiles) v += row[lane + 32];
  return v;
}

template <int EPI>
__device__ __forceinline__ EpiCol epi_col(const SplitOut& o, int nt, int ch) {
  const FusedEpi& fe = o.fe;
  EpiCol c;
  c.silu = EPI == B200_EPI_SILU_MUL;
  c.n1 = -1;
  if (c.silu) {
    const int half_tiles = (o.N / kTileN) >> 1;
    c.n0 = (nt % half_tiles) * kTileN + ch;
    c.n1 = c.n0 + (o.N >> 1);
  } else {
    c.n0 = nt * kTileN + ch;
    if (EPI == kEpiRopeKv && c.n0 < (fe.H + fe.Hkv) * fe.D) {
      const int d = c.n0 % fe.D, hd = fe.D >> 1;
      c.n1 = d < hd ? c.n0 + hd : c.n0 - hd;
    }
  }
  c.pair = c.n1 >= 0;
  c.kv_col = EPI == kEpiRopeKv && c.n0 >= fe.H * fe.D;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    c.b0[e] = 0.f;
    c.b1[e] = 0.f;
    c.nw[e] = 0.f;
  }
  if (o.bias) {
    half4_to_float(o.bias + c.n0, c.b0);
    if (c.pair) half4_to_float(o.bias + c.n1, c.b1);
  }
  if (EPI == kEpiResidNorm) half4_to_float(fe.norm_w + c.n0, c.nw);
  return c;
}

template <int EPI>
__device__ __forceinline__ void epi_row_prefetch(const SplitOut& o, const EpiCol& c, int m, bool valid,
                                                 EpiRow& r) {
  const FusedEpi& fe = o.fe;
  r.rs = 0.f;
  r.sl = -1;
  r.pos = 0;
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    r.e0[e] = 0.f;
    r.e1[e] = 0.f;
  }
  if (!valid) return;
  if (fe.ssq_in) r.rs = ssq_lane_partial(fe.ssq_in + m * kSsqStride, fe.ssq_tiles, (c.n0 >> 2) & 31);
  if (EPI == kEpiResidNorm) {
    half4_to_float(fe.resid + static_cast<size_t>(m) * o.N + c.n0, r.e0);
  } else if (EPI == kEpiRopeKv) {
    if (c.pair) r.pos = fe.positions[m];
    if (c.kv_col) r.sl = static_cast<int>(fe.slots[m]);
  }
}
// second wave of operand loads: those whose address depends on a first-wave value (cos / sin rows
// of the token's position).  Issued after the partial loads so the two latencies overlap.
template <int EPI>
__device__ __forceinline__ void epi_row_prefetch2(const SplitOut& o, const EpiCol& c, bool valid, EpiRow& r) {
  if (EPI != kEpiRopeKv) return;
  const FusedEpi& fe = o.fe;
  if (valid && c.pair) {
    const int hd = fe.D >> 1;
    const int d = c.n0 % fe.D;
    const size_t ci = static_cast<size_t>(r.pos) * hd + (d < hd ? d : d - hd);
    half4_to_float(fe.cos + ci, r.e0);
    half4_to_float(fe.sin + ci, r.e1);
  }
}

// a0 / a1: fp32 sums of the thread's 4 columns / partner columns for row m.  Called by whole
// warps with a warp-uniform m (the sum-of-squares path reduces across the warp = the tile row).
template <int EPI>
__device__ __forceinline__ void epi_row_apply(const SplitOut& o, const EpiCol& c, int m, bool valid,
                                              const float (&a0)[4], const float (&a1)[4],
                                              const EpiRow& r, int lane) {
  const FusedEpi& fe = o.fe;
  const int N = o.N;
  constexpr int epi = EPI;
  const float scale = fe.ssq_in ? rsqrtf(warp_sum(r.rs) * fe.inv_dim + fe.eps) : 1.f;
  float v0[4], v1[4], out[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v0[e] = a0[e] * scale + c.b0[e];
    v1[e] = a1[e] * scale + c.b1[e];
  }
  if (c.silu) {
    if (valid) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // round the two GEMM outputs to fp16 first: the values an unfused pipeline would see
        const float gg = __half2float(__float2half_rn(v0[e]));
        const float uu = __half2float(__float2half_rn(v1[e]));
        out[e] = gg / (1.f + __expf(-gg)) * uu;
      }
      store_half4(o.c + static_cast<size_t>(m) * (N >> 1) + c.n0, out);
    }
  } else if (epi == kEpiResidNorm) {
    float sq = 0.f;
    if (valid) {
      const size_t idx = static_cast<size_t>(m) * N + c.n0;
      float xn[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // llama.rs:126-130: the GEMM output is fp16, the residual add is done in fp16
        xn[e] = __half2float(__float2half_rn(r.e0[e] + __half2float(__float2half_rn(v0[e]))));
        out[e] = xn[e] * c.nw[e];
        sq += xn[e] * xn[e];
      }
      store_half4(fe.resid + idx, xn);
      store_half4(o.c + idx, out);
    }
    sq = warp_sum(sq);            // the warp holds the whole 128-channel row of this tile
    if (lane == 0 && valid) fe.ssq_out[m * kSsqStride + (c.n0 >> 7)] = sq;
  } else if (epi == kEpiRopeKv) {
    if (valid) {
      const int D = fe.D, hd = D >> 1;
      const int hh = c.n0 / D, d = c.n0 - hh * D;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        out[e] = v0[e];
        if (c.pair) {
          // the reference rotates the fp16 GEMM outputs (rotary_emb.rs:88-104)
          const float x0 = __half2float(__float2half_rn(v0[e])), x1 = __half2float(__float2half_rn(v1[e]));
          out[e] = d < hd ? x0 * r.e0[e] - x1 * r.e1[e] : x0 * r.e0[e] + x1 * r.e1[e];
        }
      }
      if (!c.kv_col) {
        store_half4(o.c + static_cast<size_t>(m) * N + c.n0, out);       // q stays in the fused buffer
      } else if (r.sl >= 0) {
        const bool is_k = hh < fe.H + fe.Hkv;
        __half* cache = is_k ? fe.k_cache : fe.v_cache;
        const int kvh = hh - fe.H - (is_k ? 0 : fe.Hkv);
        const size_t ci = (static_cast<size_t>(r.sl) * fe.Hkv + kvh) * D + d;
        if (fe.kv_fp8) {
          __half2_raw a = __floats2half2_rn(out[0], out[1]), b = __floats2half2_rn(out[2], out[3]);
          *reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(cache) + ci) =
              static_cast<uint32_t>(__nv_cvt_halfraw2_to_fp8x2(a, __NV_SATFINITE, __NV_E4M3)) |
              (static_cast<uint32_t>(__nv_cvt_halfraw2_to_fp8x2(b, __NV_SATFINITE, __NV_E4M3)) << 16);
        } else {
          store_half4(cache + ci, out);
        }
      }
    }
  } else if (epi == kEpiPushPeers) {
    if (valid) {
      const size_t idx = static_cast<size_t>(m) * N + c.n0;
      for (int p = 0; p < fe.push_world; ++p) store_half4(fe.push[p] + idx, v0);   // peer-mapped (NVLink) stores
    }
  } else if (valid) {
    if (fe.out_bf16)
      store_bf16x4(o.c + static_cast<size_t>(m) * N + c.n0, v0);
    else
      store_half4(o.c + static_cast<size_t>(m) * N + c.n0, v0);
  }
}

// After the finaliser of a kEpiPushPeers tile: make the tile's remote stores visible system-wide,
// then bump this rank's tile counter on every peer (release, system scope).
__device__ __forceinline__ void push_signal(const FusedEpi& fe, int tid, int nthreads, int bar_id) {
  __threadfence_system();
  named_bar_sync(bar_id, nthreads);
  if (tid < fe.push_world)
    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(fe.push_cnt[tid]) : "memory");
}

// Rows [m_lo, m_hi) of tile nt: sum the partials in split order (deterministic) and apply the
// epilogue.  lane = 4 channels, warp = row phase.  ALL loads of a batch (2 rows x up to 4 splits
// x float4) are issued before the first add: an in-order warp that meets a dependent add between
// two loads pays an L2 round trip per add -- the previous scalar version took 3 us for 32 loads,
// the longest piece of every GEMM's serial tail (tools/dec_trace.py).
template <int EPI>
__device__ __forceinline__ void split_reduce_rows(const SplitOut& o, int nt, int tid, int nthreads,
                                                  int m_lo, int m_hi) {
  const int N = o.N, M = o.M;
  const int lane = tid & 31, phase = tid >> 5, nphase = nthreads >> 5;
  constexpr int R = 2;                                 // rows in flight per thread
  constexpr bool kMayPair = EPI == B200_EPI_SILU_MUL || EPI == kEpiRopeKv;
  constexpr int U = kMayPair ? 2 : 4;                  // splits per load batch
  const EpiCol c = epi_col<EPI>(o, nt, lane * 4);
  const size_t split_stride = static_cast<size_t>(M) * N;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int mb = m_lo + phase; mb < m_hi; mb += R * nphase) {
    float a0[R][4], a1[R][4];
    EpiRow er[R];
    // (1) operands of the epilogue that do not depend on the sums are requested first
#pragma unroll
    for (int j = 0; j < R; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        a0[j][e] = 0.f;
        a1[j][e] = 0.f;
      }
      epi_row_prefetch<EPI>(o, c, mb + j * nphase, mb + j * nphase < m_hi, er[j]);
    }
    // (2) the partials: load a whole batch, then add in split order
    for (int s0 = 0; s0 < o.splits; s0 += U) {
      float4 v0[U][R], v1[U][R];
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int m = mb + j * nphase;
          const bool ok = (s0 + u < o.splits) && (m < m_hi);
          const float* row = o.partials + (s0 + u) * split_stride + static_cast<size_t>(m) * N;
          v0[u][j] = ok ? *reinterpret_cast<const float4*>(row + c.n0) : zero4;
          if (kMayPair) v1[u][j] = (ok && c.pair) ? *reinterpret_cast<const float4*>(row + c.n1) : zero4;
        }
      if (s0 == 0) {
#pragma unroll
        for (int j = 0; j < R; ++j) epi_row_prefetch2<EPI>(o, c, mb + j * nphase < m_hi, er[j]);
      }
#pragma unroll
      for (int u = 0; u < U; ++u)
#pragma unroll
        for (int j = 0; j < R; ++j) {
          a0[j][0] += v0[u][j].x; a0[j][1] += v0[u][j].y; a0[j][2] += v0[u][j].z; a0[j][3] += v0[u][j].w;
          if (kMayPair) {
            a1[j][0] += v1[u][j].x; a1[j][1] += v1[u][j].y; a1[j][2] += v1[u][j].z; a1[j][3] += v1[u][j].w;
          }
        }
    }
    // (3) epilogue (m is warp-uniform: a warp shares its row phase)
#pragma unroll
    for (int j = 0; j < R; ++j)
      epi_row_apply<EPI>(o, c, mb + j * nphase, mb + j * nphase < m_hi, a0[j], a1[j], er[j], lane);
  }
}

// Called by `nthreads` (128 or 512) threads of a CTA AFTER the partials of rows [m_lo, m_hi) of
// tile `nt` were stored.  The last CTA to arrive for
// the tile (or, with SILU_MUL, for the gate/up tile pair) sums the partials in split order --
// deterministic -- and applies the epilogue.  Loads of 8 rows are issued together so the L2
// latency is paid `splits` times, not M * splits.  `s_flag` is a shared int; `counter_base`
// separates token tiles; `bar_id` a free named barrier.
__device__ __forceinline__ bool split_arrive(const SplitOut& o, int nt, int tid, int nthreads,
                                             int counter_base, int* s_flag, int bar_id) {
  const int tiles = o.N / kTileN;
  const bool silu = o.epilogue == B200_EPI_SILU_MUL;
  const int half_tiles = tiles >> 1;
  const int cidx = counter_base + (silu ? nt % half_tiles : nt);
  const unsigned target = silu ? 2u * o.splits : static_cast<unsigned>(o.splits);
  // Release / acquire through ONE thread (the cooperative-groups grid.sync pattern): the
  // barrier orders every writer's stores before thread 0's fence + atomic (cumulativity), and
  // thread 0's acquire fence before the second barrier orders the readers' loads (ld.cg, no L1).
  named_bar_sync(bar_id, nthreads);
  if (tid == 0) {
    // acq_rel RMW at gpu scope: releases this CTA's partials, acquires the others' (no
    // sequentially-consistent fence needed: one location, one reader).  The self-reset only
    // has to be visible to the NEXT launch, which the kernel boundary orders.
    unsigned old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(o.counters + cidx) : "memory");
    const int last = (old == target - 1) ? 1 : 0;
    if (last) o.counters[cidx] = 0;                  // self-reset (graph replay safe)
    *s_flag = last;
  }
  named_bar_sync(bar_id, nthreads);
  return *s_flag != 0;
}

__device__ __forceinline__ void split_arrive_and_reduce(const SplitOut& o, int nt, int tid,
                                                        int nthreads, int m_lo, int m_hi,
                                                        int counter_base, int* s_flag, int bar_id) {
  if (!split_arrive(o, nt, tid, nthreads, counter_base, s_flag, bar_id)) return;
  switch (o.epilogue) {
    case B200_EPI_SILU_MUL: split_reduce_rows<B200_EPI_SILU_MUL>(o, nt, tid, nthreads, m_lo, m_hi); break;
    case kEpiResidNorm: split_reduce_rows<kEpiResidNorm>(o, nt, tid, nthreads, m_lo, m_hi); break;
    case kEpiRopeKv: split_reduce_rows<kEpiRopeKv>(o, nt, tid, nthreads, m_lo, m_hi); break;
    case kEpiPushPeers:
      split_reduce_rows<kEpiPushPeers>(o, nt, tid, nthreads, m_lo, m_hi);
      push_signal(o.fe, tid, nthreads, bar_id);
      break;
    default: split_reduce_rows<B200_EPI_NONE>(o, nt, tid, nthreads, m_lo, m_hi); break;
  }
}

// launchers implemented in wna16_gemv.cu / wna16_tc.cu
void launch_wna16_gemv(const WnaParams& p, cudaStream_t st);
bool launch_wna16_tc(const WnaParams& p, cudaStream_t st);   // false: shape unsupported
bool launch_wna16_tc_dec(const WnaParams& p, cudaStream_t st);   // decode shapes (M <= 64)
// stream-K decode kernel (M <= 32); bf16: A, scales and C are bf16 (plain epilogue, 4-bit only)
bool launch_wna16_sk(const WnaParams& p, cudaStream_t st, bool force = false, bool bf16 = false, int flavour = -1);
// Largest M the automatic dispatch still gives to the flat GEMV (env B200_GEMV_MAX_M, default 0: the tcgen05
// decode kernel streams the weights faster than the GEMV even for one token, profiles/r1_v7_batch_sweep.txt;
// the GEMV stays as `variant 1` and as the fallback for shapes the tensor-core kernels do not take).
int gemv_max_m();

}  // namespace b200
