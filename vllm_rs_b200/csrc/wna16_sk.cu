// W4A16 / W8A16 GEMM on tcgen05 for DECODE batches (M <= 32): persistent stream-K variant.
//
// Replaces marlin_*_4bit_f16 / gemm_half_q_half_alt for decode shapes (/root/reference/src/utils/gptq.rs:118-195,
// wna16.rs:263-306).  Same data layout, epilogues and split meeting point as wna16_tc_dec.cu; what changes is how
// the work is cut and who does what.  Motivation (round-2 measurements of the round-1 kernel,
// profiles/r2_b_dec_gemm_diet_experiments.txt, r2_c_dec_gemm_sass_histogram.txt): two thirds of its issued
// instructions are pipeline bookkeeping, gate|up runs 1.51 waves, ~4 us of serial epilogue per launch.
// Outcome (profiles/r2_j_sk_vs_dec.txt, ten measured iterations): bit-reproducible, the SAME weight arithmetic as
// the reference and the round-1 kernel (w = fp16((q - z) * s), one rounding, before the MMA), parity green on every
// shape / epilogue at kernel level and through the whole runner with every decode GEMM forced onto it
// (tests/test_runner_gpu.py::test_runner_parity_with_the_streamk_gemm_forced).  FASTER than the round-1 kernel on
// big-N projections in isolation (gate|up: 25.6 vs 30.0 us at one token, 31.2 vs 32.4 us at 32), slower on the
// small-N ones (whole layer at 32 tokens 68.4 vs 65.7 us); in the decode step it wins up to 16 tokens -- the
// automatic dispatch uses it there (sk_mode below).
//
//   * WORK = units of 128 K (two 64-K chunks) of one 128-channel tile.  The tile-major weight layout makes
//     the flattened (tile, unit) list ONE contiguous byte range, so CTA b of a one-CTA-per-SM grid simply
//     streams bytes [b*share, (b+1)*share) * unit_bytes: no wave tail, every SM busy on every shape.
//     A tile cut by a CTA boundary is finished by the split meeting point (per-CTA partial slots, arrival
//     counter, last arriver reduces in contributor order: deterministic).  Small-N projections are cut into the
//     same <= 4 aligned pieces per tile as the round-1 kernel, so a finaliser sums at most four partial tiles.
//   * three dequant groups (4 warps = 128 channels each) take the units round robin; a group dequantises a unit
//     in registers (zero point and scale of the unit read from global memory three units ahead: no staging ring),
//     tcgen05.st's the fp16 weights to TMEM, and its own issuer warp runs the eight tcgen05.mma of the unit into
//     the group's accumulator of the current tile segment (two per group, alternating between segments).
//   * at the end of a tile segment each group reads its accumulator out (tcgen05.ld, one unit later so the MMAs
//     have finished) into shared memory; two EPILOGUE warps add the three groups' sums (fixed order), write the
//     CTA's partial tile to L2 and meet the other contributors at the arrival counter -- while dequant / MMA are
//     already on the CTA's next tile.  The fused finalisers of wna16.cuh (run by the last arriver of a tile) are
//     deferred to the CTA's last segment, where sixteen warps (dequant, epilogue and producer warps, idle by
//     then) run them: 32 token rows in one pass.
//   * history (each step measured, profiles/r2_j_sk_vs_dec.txt): the first versions kept EXACT integers (q - z) in
//     TMEM and folded the scale in fp32 after the MMA with a per-unit accumulator drain -- closer to the fp32 truth
//     than the reference, but not the reference's arithmetic (one runner parity case fell outside the bar) and the
//     drain chain cost more than the multiplies it saved; ptxas re-derived the shared-memory base and the thread
//     index from S2R at every barrier until both were pinned; a 32-row finaliser on two warps was the longest
//     piece of the tail.
//   * what still limits it (ncu source view): ~10 cycles per instruction per dequant warp with three dequant warps
//     per scheduler and a ~2 us dequant -> MMA chain per unit and group that small shapes cannot hide.
#include <cstdlib>
#include <cstring>

#include "wna16.cuh"

namespace b200 {
namespace sk {

constexpr int kDG = 3;                          // dequant groups (4 warps = 128 channels each)
constexpr int kDqWarps = 4 * kDG;               // warps 0 .. 11 (warp % 4 == TMEM lane quarter)
constexpr int kEpiWarps = 2;                    // (four would push the CTA past 20 warps = below 96 registers per thread)
constexpr int kEpiWarp0 = kDqWarps;             // 12 .. 13
constexpr int kWProdWarp = kEpiWarp0 + kEpiWarps;   // 14
constexpr int kXProdWarp = kWProdWarp + 1;      // 15: activations, up to kDG * kSXS units ahead of the MMAs (fetching
                                                // them from the issuers, two units ahead, left the kernel bound by the
                                                // TMA latency: gate|up 25.5 us at one token, 37.2 us at 32)
constexpr int kIssWarp0 = kXProdWarp + 1;       // 16 .. 18 (one MMA issuer per dequant group)
constexpr int kWarps = kIssWarp0 + kDG;
constexpr int kThreads = kWarps * 32;           // 608: 19 warps leave 96 registers per thread
constexpr int kThreadsMma = kIssWarp0 * 32;     // 512: the mma.sync flavour has no issuer warps (128 registers per thread)
constexpr int kAS = 3;                          // A stages (one 64-K chunk = 32 TMEM columns) per dequant group
constexpr int kDB = 2;                          // accumulators per dequant group (unit i + 1 runs while unit i is drained)
constexpr int kMaxSeg = 8;                      // tile segments per CTA (one-shot barriers)
constexpr int kUnitK = 2 * kChunkK;             // 128
constexpr int kEpiNoneBf16 = 5;                  // template-only: plain epilogue with bf16 activations / scales / output
constexpr int kPF = 3;                          // zero-point / scale prefetch depth (units of one dequant group)

template <int BITS, int NT>
struct Cfg {
  static constexpr int kWBytes = kTileN * kChunkK * BITS / 8;   // one (tile, 64-K chunk): 4 KB / 8 KB
  static constexpr int kUnitW = 2 * kWBytes;
  static constexpr int kWStage = kDG * kUnitW;                  // one bulk copy: 24 KB / 48 KB
  static constexpr int kSWS = BITS == 4 ? 4 : 2;
  static constexpr int kXBytes = NT * kChunkK * 2;              // activations of one chunk: [NT][128 B] swizzled
  static constexpr int kXUnit = 2 * kXBytes;
  static constexpr int kSXS = 3;                                // activation slots per dequant group
  static constexpr int kXOff = 0;
  static constexpr int kWOff = kDG * kSXS * kXUnit;
  static constexpr int kHOff = kWOff + kSWS * kWStage;          // segment sums of the three groups: [group][NT][128] fp32
  static constexpr int kHBytes = NT * kTileN * 4;
  static constexpr int kBarOff = kHOff + kDG * kHBytes;
  static constexpr int kNumBars = 2 * kSWS + 2 * kDG * kSXS + 2 * kDG * kAS + 2 * kDG * kDB + 2 * kMaxSeg;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kFlagOff = kTmemPtrOff + 16;
  static constexpr int kPendOff = kFlagOff + 16;                // finaliser runs deferred to the end: count + tile ids
  static constexpr int kTidOff = kPendOff + 16 + kMaxSeg * 4;   // one word per thread (see the kernel prologue)
  static constexpr int kAlloc = kTidOff + kThreads * 4 + 1024;
  static_assert(kAlloc <= 227 * 1024, "shared memory budget");
  static_assert(kXBytes % 1024 == 0, "SWIZZLE_128B tiles are 1024-byte aligned");
  static constexpr int kDCol0 = 0;                              // accumulator (g, buf): columns [32 (2 g + buf), + NT)
  static constexpr int kACol0 = kDG * kDB * 32;
  static constexpr int kTmemCols = 512;
  static_assert(NT <= 32 && kACol0 + kDG * kAS * 32 <= 512, "TMEM budget");
};

struct Args {
  const uint32_t* bp;
  const __half* scales;
  const uint32_t* qzeros;
  const __half* bias;
  __half* c;
  float* partials;       // [cta][S_max][M][128] fp32
  uint32_t* counters;
  int M, K, N, gshift, qfmt, KP, units, share, smax;   // quantisation group of unit `pair`: pair >> gshift
  FusedEpi fe;
};

struct Pos { int tile, pair; };
__device__ __forceinline__ Pos pos_of(int u, int KP) {
  Pos p;
  p.tile = u / KP;
  p.pair = u - p.tile * KP;
  return p;
}
__device__ __forceinline__ void pos_advance(Pos& p, int d, int KP) {
  p.pair += d;
  while (p.pair >= KP) { p.pair -= KP; ++p.tile; }
}

// Which CTAs hold a piece of tile t, and where their partial tiles live.
struct Contrib { int first, cnt, slot0; };
__device__ __forceinline__ Contrib contrib_of(int t, int KP, int share, int smax) {
  Contrib c;
  c.first = (t * KP) / share;
  const int last = ((t + 1) * KP - 1) / share;
  c.cnt = last - c.first + 1;
  c.slot0 = c.first * smax + (t - (c.first * share) / KP);     // the first contributor may hold t as a later segment
  return c;
}
__device__ __forceinline__ const float* part_ptr(const float* partials, const Contrib& c, int s, int smax,
                                                 size_t slot_floats) {
  const int slot = s == 0 ? c.slot0 : (c.first + s) * smax;    // every later contributor STARTS inside t
  return partials + static_cast<size_t>(slot) * slot_floats;
}

__device__ __forceinline__ void wait_bar(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_addr(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_addr(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) __trap();               // ~2 s: a protocol bug fails the launch, never hangs
  }
}
// long waits of warps that are off the critical path (epilogue warps): back off instead of competing for issue slots
__device__ __forceinline__ void wait_bar_sleepy(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  const long long t0 = clock64();
  for (;;) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (ok) return;
    __nanosleep(1000);
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ void arrive_bar(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ float4 ld_cg_f4(const float* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// exact (q - z) as fp16 pairs: nibbles at mantissa bits 0..3 read as 1024 + q, at bits 4..7 as 64 + q
__device__ __forceinline__ void dq4_exact(uint32_t x, __half2 zlo, __half2 zhi, uint32_t (&out)[4]) {
  const uint32_t y = x >> 8;
  out[0] = h2_as_u32(__hsub2(u32_as_h2(lop3_or_and(x, 0x000F000Fu, 0x64006400u)), zlo));
  out[1] = h2_as_u32(__hsub2(u32_as_h2(lop3_or_and(x, 0x00F000F0u, 0x54005400u)), zhi));
  out[2] = h2_as_u32(__hsub2(u32_as_h2(lop3_or_and(y, 0x000F000Fu, 0x64006400u)), zlo));
  out[3] = h2_as_u32(__hsub2(u32_as_h2(lop3_or_and(y, 0x00F000F0u, 0x54005400u)), zhi));
}
// bf16 flavour (b200_marlin_*_bf16): 7 mantissa bits hold a nibble only at bits 0..3, so every pair is shifted down
// first; 128 + q and 128 + z are exact in bf16, and so is their difference
__device__ __forceinline__ void dq4_exact_bf16(uint32_t x, uint32_t zb, uint32_t (&out)[4]) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const uint32_t v = lop3_or_and(x >> (4 * p), 0x000F000Fu, 0x43004300u);
    asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(out[p]) : "r"(v), "r"(zb));
  }
}
// bf16 flavour of the reference arithmetic: w = bf16((q - z) * s), the difference exact, one rounding in the multiply
__device__ __forceinline__ void dequant4_word_bf16(uint32_t x, uint32_t zb, uint32_t sb, uint32_t (&out)[4]) {
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    const uint32_t v = lop3_or_and(x >> (4 * p), 0x000F000Fu, 0x43004300u);
    uint32_t d;
    asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(v), "r"(zb));
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(out[p]) : "r"(d), "r"(sb));
  }
}
__device__ __forceinline__ void dq8_exact(uint32_t x, __half2 zlo, uint32_t (&out)[2]) {
  uint32_t lo, hi;
  asm("prmt.b32 %0, %1, %2, 0x4140;" : "=r"(lo) : "r"(x), "r"(0x64646464u));
  asm("prmt.b32 %0, %1, %2, 0x4342;" : "=r"(hi) : "r"(x), "r"(0x64646464u));
  out[0] = h2_as_u32(__hsub2(u32_as_h2(lo), zlo));
  out[1] = h2_as_u32(__hsub2(u32_as_h2(hi), zlo));
}

// ---- tile finaliser over per-CTA partial slots (the row / column logic is wna16.cuh's; only the addressing of
// the partials differs from split_reduce_rows).  nphase warps: lane = 4 channels, warp = row phase.
template <int EPI>
__device__ __forceinline__ void reduce_tile(const Args& a, const SplitOut& o, int nt, int tid, int nphase) {
  const int M = a.M, KP = a.KP;
  const int lane = tid & 31, phase = tid >> 5;
  constexpr int R = 2;
  constexpr bool kSilu = EPI == B200_EPI_SILU_MUL;
  constexpr bool kMayPair = kSilu || EPI == kEpiRopeKv;
  constexpr int U = kMayPair ? 2 : 4;
  const EpiCol c = epi_col<EPI>(o, nt, lane * 4);
  const size_t slot_floats = static_cast<size_t>(M) * kTileN;
  const int half_tiles = (a.N / kTileN) >> 1;
  const int t0 = kSilu ? nt % half_tiles : nt;
  const Contrib k0 = contrib_of(t0, KP, a.share, a.smax);
  const Contrib k1 = kSilu ? contrib_of(t0 + half_tiles, KP, a.share, a.smax) : k0;
  const int c0 = lane * 4;
  const int c1 = kSilu ? c0 : (c.pair ? c.n1 - nt * kTileN : c0);
  const int smaxc = k0.cnt > k1.cnt ? k0.cnt : k1.cnt;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int mb = phase; mb < M; mb += R * nphase) {
    float a0[R][4], a1[R][4];
    EpiRow er[R];
#pragma unroll
    for (int j = 0; j < R; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        a0[j][e] = 0.f;
        a1[j][e] = 0.f;
      }
      epi_row_prefetch<EPI>(o, c, mb + j * nphase, mb + j * nphase < M, er[j]);
    }
    for (int s0 = 0; s0 < smaxc; s0 += U) {
      float4 v0[U][R], v1[U][R];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int s = s0 + u;
        const float* p0 = part_ptr(a.partials, k0, s, a.smax, slot_floats);
        const float* p1 = kSilu ? part_ptr(a.partials, k1, s, a.smax, slot_floats) : p0;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const int m = mb + j * nphase;
          const bool ok0 = s < k0.cnt && m < M;
          const bool ok1 = s < k1.cnt && m < M && c.pair;
          v0[u][j] = ok0 ? ld_cg_f4(p0 + m * kTileN + c0) : zero4;
          if (kMayPair) v1[u][j] = ok1 ? ld_cg_f4(p1 + m * kTileN + c1) : zero4;
        }
      }
      if (s0 == 0) {
#pragma unroll
        for (int j = 0; j < R; ++j) epi_row_prefetch2<EPI>(o, c, mb + j * nphase < M, er[j]);
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
#pragma unroll
    for (int j = 0; j < R; ++j)
      epi_row_apply<EPI>(o, c, mb + j * nphase, mb + j * nphase < M, a0[j], a1[j], er[j], lane);
  }
}

// MMA = true: the register flavour.  Same work list, rings, segment sums, meeting point and finalisers; but a
// dequant warp keeps its weights: lane (g, c) of warp q4 dequantises words {2c, 2c+1} of channels 32 q4 + 8 t + g
// (t = 0..3) of a 64-K chunk -- K values [16 c, 16 c + 16) -- straight into the A fragments of
// mma.sync.m16n8k16 (16 channels are the M rows, 8 tokens the N columns: the dequantised pairs of channels g and
// g + 8 ARE the register quad, no moves; token tiles beyond the batch are skipped), takes the matching activations
// from the swizzled activation tile with two 16-byte loads per token row (the k-steps see K in the order
// 16 c + 4 s + {0..3}, the same permutation on both operands), and accumulates in registers.  No TMEM, no issuer warps, no
// dequant -> store -> commit -> MMA hand-off chain per chunk: the only barriers left in the hot loop are the two
// TMA rings.  4-bit, fp16.
template <int BITS, int NT, int EPI, bool MMA>
__global__ void __launch_bounds__(MMA ? kThreadsMma : kThreads, 1)
wna16_sk_kernel(const __grid_constant__ CUtensorMap xmap, const Args a) {
  using C = Cfg<BITS, NT>;
  static_assert(!MMA || (BITS == 4 && EPI != kEpiNoneBf16), "the mma.sync flavour is 4-bit fp16");
  constexpr bool kBf16 = EPI == kEpiNoneBf16;
  static_assert(!kBf16 || BITS == 4, "bf16 entry points are 4-bit only (gptq.rs:150-178)");
  const int b = blockIdx.x;
  const int KP = a.KP, N = a.N, M = a.M;
  const int u0 = b * a.share;
  const int n = min(a.share, a.units - u0);            // units of this CTA (>= 1 by construction of the grid)

  extern __shared__ uint8_t smem_raw[];
  // The 1024-aligned shared-space base as ONE pinned 32-bit register: left to itself the compiler rebuilds it
  // (S2R CgaCtaId / SWINHI + align, ~10 instructions) at every barrier of the hot loops (seen in the ncu source
  // view of the first versions).  Every shared address below is s_base + offset.
  const uint32_t s_raw = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
  uint32_t s_base = (s_raw + 1023u) & ~1023u;
  asm volatile("" : "+r"(s_base));
  uint8_t* smem = smem_raw + (s_base - s_raw);
  // The thread index makes a round trip through shared memory: ptxas otherwise re-derives warp / lane / TMEM lane
  // offsets from S2R SR_TID.X (long, variable latency) at every use inside the hot loops instead of keeping them.
  uint32_t tid_o = threadIdx.x;
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(s_base + C::kTidOff + tid_o * 4), "r"(tid_o) : "memory");
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(tid_o) : "r"(s_base + C::kTidOff + threadIdx.x * 4) : "memory");
  const int warp = static_cast<int>(tid_o >> 5), lane = static_cast<int>(tid_o & 31);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* wfull = bars;
  uint64_t* wempty = wfull + C::kSWS;
  uint64_t* xfull = wempty + C::kSWS;                  // [unit % (kDG * kSXS)]
  uint64_t* xempty = xfull + kDG * C::kSXS;
  uint64_t* afree = xempty + kDG * C::kSXS;            // [group][stage]
  uint64_t* stored = afree + kDG * kAS;                // [group][stage]
  uint64_t* dfull = stored + kDG * kAS;                // [group][buf]: the unit's eight MMAs are complete
  uint64_t* dfree = dfull + kDG * kDB;                 // [group][buf]: the dequant warps have read the accumulator
  uint64_t* segdone = dfree + kDG * kDB;               // [segment]: all twelve dequant warps left their sums in smem
  uint64_t* hfree = segdone + kMaxSeg;                 // [segment]: the epilogue warps have read them
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + C::kTmemPtrOff);
  volatile int* s_flag = reinterpret_cast<volatile int*>(smem + C::kFlagOff);
  volatile int* s_pend = reinterpret_cast<volatile int*>(smem + C::kPendOff);      // [0] = count, [1 ..] = tiles
  // the same barriers as shared-space addresses
  const uint32_t s_wfull = s_base + C::kBarOff, s_wempty = s_wfull + C::kSWS * 8;
  const uint32_t s_xfull = s_wempty + C::kSWS * 8, s_xempty = s_xfull + kDG * C::kSXS * 8;
  const uint32_t s_afree = s_xempty + kDG * C::kSXS * 8;
  const uint32_t s_stored = s_afree + kDG * kAS * 8, s_dfull = s_stored + kDG * kAS * 8;
  const uint32_t s_dfree = s_dfull + kDG * kDB * 8, s_segdone = s_dfree + kDG * kDB * 8;
  const uint32_t s_hfree = s_segdone + kMaxSeg * 8;

  if (tid_o == 0) {
    s_pend[0] = 0;
    for (int i = 0; i < C::kSWS; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wempty[i], kDqWarps);
    }
    for (int i = 0; i < kDG * C::kSXS; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], MMA ? 4 : 1);                // the group's four warps / the issuer's commit
    }
    for (int i = 0; i < kDG * kAS; ++i) {
      mbar_init(&afree[i], 1);
      mbar_init(&stored[i], 4);
    }
    for (int i = 0; i < kDG * kDB; ++i) {
      mbar_init(&dfull[i], 1);
      mbar_init(&dfree[i], 4);
    }
    for (int i = 0; i < kMaxSeg; ++i) {
      mbar_init(&segdone[i], kDqWarps);
      mbar_init(&hfree[i], kEpiWarps);
    }
    fence_barrier_init();
  }
  if (!MMA && warp == kWProdWarp) {
    tmem_alloc(tmem_ptr, C::kTmemCols);
    tmem_relinquish();
  }
  if (warp == kXProdWarp && lane == 0) prefetch_tmap(&xmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  const uint32_t tmem_base = MMA ? 0u : *tmem_ptr;

  // ---- one tile segment's epilogue, run by `nw` cooperating warps (warp `wid`, thread `tid_c` of them, named barrier
  // `bar_id`): add the three groups' sums (fixed order), store the CTA's partial tile (or the output itself when
  // the CTA holds the whole tile and the epilogue is plain), meet the other contributors at the arrival counter and,
  // if last, run the fused finaliser.  Intermediate segments are done by the two epilogue warps while the next tile
  // streams; the CTA's LAST segment is done by the dequant and producer warps as well (16 warps: they are idle by then,
  // and a 32-row finaliser on two warps was the longest piece of the kernel's tail).
  constexpr int EPXL = kBf16 ? B200_EPI_NONE : EPI;
  auto segment_epilogue = [&](int t, int seg, int wid, int nw, int tid_c, int bar_id, bool free_slots) {
    constexpr bool kPlainL = EPXL == B200_EPI_NONE;
    constexpr bool kSiluL = EPXL == B200_EPI_SILU_MUL;
    SplitOut o;
    o.partials = a.partials; o.counters = a.counters; o.bias = a.bias; o.c = a.c;
    o.M = M; o.N = N; o.splits = 1; o.epilogue = EPXL; o.fe = a.fe;
    const int half_tiles = (N / kTileN) >> 1;
    const uint32_t s_hl = s_base + C::kHOff + lane * 16;
    const Contrib k = contrib_of(t, KP, a.share, a.smax);
    const bool direct = kPlainL && a.fe.ssq_in == nullptr && k.cnt == 1;
    if (free_slots) wait_bar_sleepy(s_segdone + seg * 8, 0); else wait_bar(s_segdone + seg * 8, 0);
    float* P = a.partials + (static_cast<size_t>(b) * a.smax + seg) * (static_cast<size_t>(M) * kTileN);
    float bv[4] = {0.f, 0.f, 0.f, 0.f};
    if (direct && a.bias && !kBf16) half4_to_float(a.bias + t * kTileN + lane * 4, bv);
    for (int m = wid; m < M; m += nw) {
      float4 v[kDG];
#pragma unroll
      for (int g = 0; g < kDG; ++g)
        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                     : "=f"(v[g].x), "=f"(v[g].y), "=f"(v[g].z), "=f"(v[g].w)
                     : "r"(s_hl + g * C::kHBytes + m * (kTileN * 4)));
      float sum[4] = {v[0].x, v[0].y, v[0].z, v[0].w};
#pragma unroll
      for (int g = 1; g < kDG; ++g) {
        sum[0] += v[g].x; sum[1] += v[g].y; sum[2] += v[g].z; sum[3] += v[g].w;
      }
      if (direct) {
#pragma unroll
        for (int e = 0; e < 4; ++e) sum[e] += bv[e];
        if constexpr (kBf16)
          store_bf16x4(a.c + static_cast<size_t>(m) * N + t * kTileN + lane * 4, sum);
        else
          store_half4(a.c + static_cast<size_t>(m) * N + t * kTileN + lane * 4, sum);
      } else {
        *reinterpret_cast<float4*>(P + m * kTileN + lane * 4) = make_float4(sum[0], sum[1], sum[2], sum[3]);
      }
    }
    if (free_slots) {
      __syncwarp();
      if (lane == 0) arrive_bar(s_hfree + seg * 8);
    }
    if (direct) return;
    named_bar_sync(bar_id, nw * 32);                     // every cooperating warp's rows are stored
    if (tid_c == 0) {
      int cidx = t;
      unsigned target = static_cast<unsigned>(k.cnt);
      if (kSiluL) {
        cidx = t % half_tiles;
        const Contrib kg = contrib_of(cidx, KP, a.share, a.smax);
        const Contrib ku = contrib_of(cidx + half_tiles, KP, a.share, a.smax);
        target = static_cast<unsigned>(kg.cnt + ku.cnt);
      }
      // acq_rel RMW at gpu scope: releases this CTA's partial tile (the barrier above orders the other
      // threads' stores before it), acquires the other contributors'
      unsigned old;
      asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(a.counters + cidx) : "memory");
      const int last = (old == target - 1) ? 1 : 0;
      if (last) a.counters[cidx] = 0;                   // self-reset (graph replay safe)
      *s_flag = last;
    }
    named_bar_sync(bar_id, nw * 32);
    const bool last = *s_flag != 0;
    if (free_slots) {
      // intermediate segment (two warps): a finaliser run here would hold them for 32 rows x L2 round trips while
      // later segments queue up behind; note the tile and let all sixteen warps run it at the end
      if (last && tid_c == 0) {
        const int np = s_pend[0];
        s_pend[1 + np] = t;
        s_pend[0] = np + 1;
      }
      named_bar_sync(bar_id, nw * 32);                   // s_flag is rewritten by the next segment
      return;
    }
    if (last) reduce_tile<EPXL>(a, o, t, tid_c, nw);
    const int np = s_pend[0];                            // written before the epilogue warps joined this barrier
    for (int q = 0; q < np; ++q) reduce_tile<EPXL>(a, o, s_pend[1 + q], tid_c, nw);
    if (EPXL == kEpiPushPeers && (last || np > 0)) {
      // one signal per finished tile: the consumer counts tiles (comm.cu), so bump once per tile
      const int bumps = (last ? 1 : 0) + np;
      for (int q = 0; q < bumps; ++q) push_signal(o.fe, tid_c, nw * 32, bar_id + 1);
    }
  };
  constexpr int kAllEpiWarps = kIssWarp0;              // the CTA's last segment: warps 0 .. 15 (dequant, epilogue, the two producers)
  static_assert(kAllEpiWarps == 16, "32 token rows = one pass of the finaliser (2 rows per warp)");

  if (warp == kWProdWarp) {
    // ------------------------------------------------------------ weight producer: one contiguous stream
    if (lane == 0) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a.bp) + static_cast<size_t>(u0) * C::kUnitW;
      const int nst = (n + kDG - 1) / kDG;
      for (int s = 0; s < nst; ++s) {
        const int slot = s % C::kSWS;
        wait_bar(s_wempty + slot * 8, ((s / C::kSWS) & 1) ^ 1);
        const int left = n - s * kDG;
        const uint32_t bytes = static_cast<uint32_t>(left < kDG ? left : kDG) * C::kUnitW;
        mbar_arrive_expect_tx(&wfull[slot], bytes);
        bulk_load_1d(smem + C::kWOff + slot * C::kWStage, src + static_cast<size_t>(s) * C::kWStage, bytes,
                     &wfull[slot]);
      }
    }
    __syncwarp();
    {
      const int t_first = pos_of(u0, KP).tile, t_last = pos_of(u0 + n - 1, KP).tile;
      pdl_wait();
      segment_epilogue(t_last, t_last - t_first, warp, kAllEpiWarps, warp * 32 + lane, 3, false);
    }
  } else if (warp == kXProdWarp) {
    // ------------------------------------------------------------ activation producer (ring of kDG * kSXS units)
    if (lane == 0) {
      pdl_wait();      // activations are the previous kernel's output; weights are not
      constexpr int R = kDG * C::kSXS;
      Pos p = pos_of(u0, KP);
      int slot = 0;
      uint32_t ph = 1;                                 // xempty: the first round passes
      for (int j = 0; j < n; ++j) {
        wait_bar(s_xempty + slot * 8, ph);
        mbar_arrive_expect_tx(&xfull[slot], C::kXUnit);          // TMA zero-fills rows >= M
        tma_load_3d(smem + C::kXOff + slot * C::kXUnit, &xmap, &xfull[slot], 0, 0, 2 * p.pair);
        pos_advance(p, 1, KP);
        if (++slot == R) { slot = 0; ph ^= 1; }
      }
    }
    __syncwarp();
    {
      const int t_first = pos_of(u0, KP).tile, t_last = pos_of(u0 + n - 1, KP).tile;
      pdl_wait();
      segment_epilogue(t_last, t_last - t_first, warp, kAllEpiWarps, warp * 32 + lane, 3, false);
    }
  } else if (warp >= kIssWarp0) {
    // ------------------------------------------------------------ MMA issuers (one per dequant group)
    if (lane == 0) {
      const int grp = warp - kIssWarp0;
      // kind::f16 takes bf16 operands with a_format = b_format = 1 (bits [7,10) and [10,13))
      const uint32_t idesc = make_idesc_f16(128, NT < 16 ? 16 : NT) | (kBf16 ? ((1u << 7) | (1u << 10)) : 0u);
      int ai = 0;
      uint32_t aphase = 0;
      int it = 0;
      int k = -1;                                      // this group's tile segments so far (those it has units in)
      Pos p = pos_of(u0 + grp, KP);
      int tile_prev = -1;
      uint32_t d_tmem = 0;
      for (int j = grp; j < n; j += kDG, ++it) {
        const int xs = (it % C::kSXS) * kDG + grp;               // unit j lives in ring slot j % (kDG * kSXS)
        bool first = false;
        if (p.tile != tile_prev) {
          // a new tile segment of this group: fresh accumulator (two per group, alternating); the one it takes
          // was last used by the group's segment k - 2 and must have been read out by the dequant warps
          ++k;
          tile_prev = p.tile;
          d_tmem = tmem_base + C::kDCol0 + (grp * kDB + (k & 1)) * 32;
          wait_bar(s_dfree + (grp * kDB + (k & 1)) * 8, ((k >> 1) & 1) ^ 1);
          first = true;
        }
        Pos pn = p;
        pos_advance(pn, kDG, KP);
        const bool seg_last = j + kDG >= n || pn.tile != p.tile;
        wait_bar(s_xfull + xs * 8, (it / C::kSXS) & 1);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int as = grp * kAS + ai;
          wait_bar(s_stored + as * 8, aphase);
          tc_fence_after();
          const uint32_t a_tmem = tmem_base + C::kACol0 + as * 32;
          const uint64_t bdesc = make_smem_desc_sw128(s_base + C::kXOff + xs * C::kXUnit + cc * C::kXBytes);
#pragma unroll
          for (int ks = 0; ks < kChunkK / 16; ++ks)
            tc_mma_ts(d_tmem, a_tmem + ks * 8, bdesc + static_cast<uint64_t>(ks * 2), idesc,
                      (first && cc == 0 && ks == 0) ? 0u : 1u);
          tc_commit(&afree[as]);
          if (++ai == kAS) { ai = 0; aphase ^= 1; }
        }
        tc_commit(&xempty[xs]);
        if (seg_last) tc_commit(&dfull[grp * kDB + (k & 1)]);    // the group's sums of this tile are complete
        p = pn;
      }
    }
  } else if (MMA && warp < kDqWarps) {
    // ------------------------------------------------------------ dequant + mma.sync groups (register flavour)
    if constexpr (MMA) {
    const int grp = warp >> 2, q4 = warp & 3;
    const int g = lane >> 2, c = lane & 3;
    constexpr int NT8 = NT / 8;                          // 8-token column tiles
    const int ch = q4 * 32 + lane;                       // the channel whose zero point / scale this lane FETCHES
    constexpr int ZWN = kTileN * 4 / 32;
    const int zword = ch >> 3;
    const int zshift = (a.qfmt == B200_QFMT_AWQ) ? 4 * awq_nibble_of(ch & 7) : 4 * (ch & 7);
    const uint32_t zadd = (a.qfmt == B200_QFMT_AWQ) ? 0u : 1u;
    const uint32_t zrow = static_cast<uint32_t>(N) * 4 / 32;
    const bool has_z = a.qzeros != nullptr;
    const unsigned short* scales_u16 = reinterpret_cast<const unsigned short*>(a.scales);
    auto zload = [&](const Pos& q) -> uint32_t {
      if (!has_z) return 0u;
      return __ldg(a.qzeros + (static_cast<uint32_t>(q.pair >> a.gshift) * zrow + q.tile * ZWN + zword));
    };
    auto sload = [&](const Pos& q) -> uint32_t {
      return __ldg(scales_u16 + (static_cast<uint32_t>(q.pair >> a.gshift) * N + q.tile * kTileN + ch));
    };
    // this lane's words of channel 32 q4 + g (+ 8 t), chunk 0, stage 0: half (c >> 1) of the chunk, words 2 (c & 1) ..
    uint32_t s_w = s_base + C::kWOff + grp * C::kUnitW + (c >> 1) * (kTileN * 16) + (q4 * 32 + g) * 16 + (c & 1) * 8;
    // activations: token row r of a chunk is 128 swizzled bytes; rows g, g + 8, .. all have r & 7 == g
    uint32_t s_x = s_base + C::kXOff + g * 128;
    const uint32_t xo0 = static_cast<uint32_t>((2 * c) ^ g) << 4, xo1 = static_cast<uint32_t>((2 * c + 1) ^ g) << 4;
    // segment sums [token][channel]: tokens 2 c (+ 1, + 8 nt), channels 32 q4 + g (+ 8, + 16 mt)
    uint32_t s_h = s_base + C::kHOff + grp * C::kHBytes + (2 * c) * (kTileN * 4) + (q4 * 32 + g) * 4;
    uint32_t lane0 = lane == 0 ? 1u : 0u;
    asm volatile("" : "+r"(s_w), "+r"(s_x), "+r"(s_h), "+r"(lane0));
    const int t_first = pos_of(u0, KP).tile, t_last = pos_of(u0 + n - 1, KP).tile;
    const int ntiles = (M + 7) >> 3;                     // token tiles that hold a row (the others stay zero)
    float acc[2][NT8][4];                                // [16-channel tile][8-token tile]
    auto zero_acc = [&]() {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[mt][nt][e] = 0.f;
    };
    zero_acc();
    // a tile segment's sums of this group go to shared memory (zeros when the group had no unit in the segment)
    auto flush = [&](int seg) {
      if (seg >= 1) wait_bar(s_hfree + (seg - 1) * 8, 0);        // the previous segment's sums have been read
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT8; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e)
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_h + (8 * nt + (e & 1)) * (kTileN * 4) +
                                                         (16 * mt + 8 * (e >> 1)) * 4), "f"(acc[mt][nt][e]) : "memory");
      __syncwarp();
      if (lane0) arrive_bar(s_segdone + seg * 8);
      zero_acc();
    };
    int next_seg = 0;                                  // segments [0, next_seg) have been flushed by this warp
    if (grp < n) {
      Pos p = pos_of(u0 + grp, KP);
      uint32_t zq[kPF], sq[kPF];                       // zero words / scale bits of this group's next kPF units
      Pos pz = p;
#pragma unroll
      for (int k = 0; k < kPF; ++k) {
        zq[k] = 0;
        sq[k] = 0;
        if (grp + k * kDG < n) {
          zq[k] = zload(pz);
          sq[k] = sload(pz);
          pos_advance(pz, kDG, KP);
        }
      }
      int wslot = 0;
      uint32_t wphase = 0;
      int it = 0;
      int cur_seg = p.tile - t_first;
      while (next_seg < cur_seg) flush(next_seg++);    // segments before this group's first unit
      for (int j = grp; j < n; j += kDG, ++it) {
        const int seg = p.tile - t_first;
        if (seg != cur_seg) {
          flush(cur_seg);
          next_seg = cur_seg + 1;
          while (next_seg < seg) flush(next_seg++);
          cur_seg = seg;
        }
        // zero point and scale of the lane's four channels: fetched by lane 8 t + g, handed over by shuffle
        const uint32_t zmine = has_z ? ((zq[0] >> zshift) & 15u) + zadd : 8u;
        DqConst dq[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const uint32_t z = __shfl_sync(0xffffffffu, zmine, 8 * t + g);
          const uint32_t sb = __shfl_sync(0xffffffffu, sq[0], 8 * t + g);
          dq[t].s2 = u32_as_h2(sb * 0x00010001u);
          dq[t].zlo = u32_as_h2((0x6400u | z) * 0x00010001u);
          dq[t].zhi = u32_as_h2((0x5400u | (z << 4)) * 0x00010001u);
        }
#pragma unroll
        for (int k = 0; k + 1 < kPF; ++k) {
          zq[k] = zq[k + 1];
          sq[k] = sq[k + 1];
        }
        if (j + kPF * kDG < n) {
          zq[kPF - 1] = zload(pz);
          sq[kPF - 1] = sload(pz);
          pos_advance(pz, kDG, KP);
        }
        const int xs = (it % C::kSXS) * kDG + grp;               // unit j lives in ring slot j % (kDG * kSXS)
        wait_bar(s_wfull + wslot * 8, wphase);
        wait_bar(s_xfull + xs * 8, (it / C::kSXS) & 1);
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          // activations of the chunk (the B fragments): token rows g + 8 nt, K [16 c, 16 c + 16) as two 16-byte
          // pieces = the (b0, b1) pairs of four k-steps
          uint32_t xb[NT8][2][4];
          const uint32_t xa = s_x + xs * C::kXUnit + cc * C::kXBytes;
#pragma unroll
          for (int nt = 0; nt < NT8; ++nt) {
            if (nt < ntiles) {
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                           : "=r"(xb[nt][0][0]), "=r"(xb[nt][0][1]), "=r"(xb[nt][0][2]), "=r"(xb[nt][0][3])
                           : "r"(xa + nt * (8 * 128) + xo0));
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                           : "=r"(xb[nt][1][0]), "=r"(xb[nt][1][1]), "=r"(xb[nt][1][2]), "=r"(xb[nt][1][3])
                           : "r"(xa + nt * (8 * 128) + xo1));
            }
          }
          const uint32_t wb = s_w + wslot * C::kWStage + cc * C::kWBytes;
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            // the A fragments: channels g and g + 8 of the 16-channel tile, dequantised into place
            uint32_t wlo[2], whi[2];
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(wlo[0]), "=r"(wlo[1]) : "r"(wb + (2 * mt) * (8 * 16)));
            asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(whi[0]), "=r"(whi[1]) : "r"(wb + (2 * mt + 1) * (8 * 16)));
            uint32_t olo[2][4], ohi[2][4];
            dequant4_word(wlo[0], dq[2 * mt], olo[0]);
            dequant4_word(wlo[1], dq[2 * mt], olo[1]);
            dequant4_word(whi[0], dq[2 * mt + 1], ohi[0]);
            dequant4_word(whi[1], dq[2 * mt + 1], ohi[1]);
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int hh = 0; hh < 2; ++hh) {
                const uint32_t af[4] = {olo[i][2 * hh], ohi[i][2 * hh], olo[i][2 * hh + 1], ohi[i][2 * hh + 1]};
#pragma unroll
                for (int nt = 0; nt < NT8; ++nt)
                  if (nt < ntiles) mma_16816(acc[mt][nt], af, xb[nt][i][2 * hh], xb[nt][i][2 * hh + 1]);
              }
          }
        }
        __syncwarp();
        if (lane0) {
          arrive_bar(s_wempty + wslot * 8);              // this warp's part of the stage has been consumed
          arrive_bar(s_xempty + xs * 8);
        }
        if (++wslot == C::kSWS) { wslot = 0; wphase ^= 1; }
        pos_advance(p, kDG, KP);
      }
      flush(cur_seg);
      next_seg = cur_seg + 1;
    }
    while (next_seg <= t_last - t_first) flush(next_seg++);
    pdl_wait();                                          // first global writes of this role
    segment_epilogue(t_last, t_last - t_first, warp, kAllEpiWarps, warp * 32 + lane, 3, false);
    }
  } else if (warp < kDqWarps) {
    // ------------------------------------------------------------ dequant groups
    const int grp = warp >> 2, q4 = warp & 3;
    const int ch = q4 * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    constexpr int W = 2 * BITS;                        // packed words per thread per chunk (8 / 16)
    constexpr int ZWN = kTileN * BITS / 32;            // zero-point words per tile row
    const int zword = BITS == 4 ? (ch >> 3) : (ch >> 2);
    const int zshift = BITS == 4 ? ((a.qfmt == B200_QFMT_AWQ) ? 4 * awq_nibble_of(ch & 7) : 4 * (ch & 7))
                                 : 8 * (ch & 3);
    const uint32_t zadd = (a.qfmt == B200_QFMT_AWQ && BITS == 4) ? 0u : 1u;
    const uint32_t zmask = (1u << BITS) - 1;
    const uint32_t zrow = static_cast<uint32_t>(N) * BITS / 32;      // 32-bit indices: the tensors are far below 2^31 elements
    const bool has_z = a.qzeros != nullptr;
    uint32_t s_h = s_base + C::kHOff + grp * C::kHBytes + ch * 4;
    uint32_t s_w = s_base + C::kWOff + grp * C::kUnitW + ch * 16;        // this thread's words of (stage 0, chunk 0)
    uint32_t lane0 = lane == 0 ? 1u : 0u;
    asm volatile("" : "+r"(s_h), "+r"(s_w), "+r"(lane0));
    const unsigned short* scales_u16 = reinterpret_cast<const unsigned short*>(a.scales);
    auto zload = [&](const Pos& q) -> uint32_t {
      if (!has_z) return 0u;
      return __ldg(a.qzeros + (static_cast<uint32_t>(q.pair >> a.gshift) * zrow + q.tile * ZWN + zword));
    };
    auto sload = [&](const Pos& q) -> uint32_t {
      return __ldg(scales_u16 + (static_cast<uint32_t>(q.pair >> a.gshift) * N + q.tile * kTileN + ch));
    };
    auto load_chunk = [&](uint32_t (&dst)[W], int wslot, int cc) {
      const uint32_t addr = s_w + wslot * C::kWStage + cc * C::kWBytes;
#pragma unroll
      for (int sub = 0; sub < W / 4; ++sub)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(dst[sub * 4]), "=r"(dst[sub * 4 + 1]), "=r"(dst[sub * 4 + 2]), "=r"(dst[sub * 4 + 3])
                     : "r"(addr + sub * kTileN * 16));
    };
    const int t_first = pos_of(u0, KP).tile, t_last = pos_of(u0 + n - 1, KP).tile;
    // ---- a tile segment's sums of this group go to shared memory: read out of the group's accumulator (its
    // segment ordinal kk picks the buffer and the barrier phase), or zeros when the group had no unit in the segment
    auto flush = [&](int seg, int kk) {
      if (seg >= 1) wait_bar(s_hfree + (seg - 1) * 8, 0);        // the previous segment's sums have been read
      if (kk >= 0) {
        const int gb = grp * kDB + (kk & 1);
        wait_bar(s_dfull + gb * 8, (kk >> 1) & 1);
        tc_fence_after();
        const uint32_t d_addr = tmem_base + C::kDCol0 + gb * 32 + lane_base;
#pragma unroll
        for (int h0 = 0; h0 < NT; h0 += 16) {
          uint32_t v[16];
          tmem_ld_x16(d_addr + h0, v);
          tc_wait_ld();
#pragma unroll
          for (int m = 0; m < 16; ++m)
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(s_h + (h0 + m) * (kTileN * 4)), "r"(v[m]) : "memory");
        }
        tc_fence_before();
        __syncwarp();
        if (lane0) arrive_bar(s_dfree + gb * 8);
      } else {
#pragma unroll
        for (int m = 0; m < NT; ++m)
          asm volatile("st.shared.u32 [%0], %1;" ::"r"(s_h + m * (kTileN * 4)), "r"(0u) : "memory");
        __syncwarp();
      }
      if (lane0) arrive_bar(s_segdone + seg * 8);
    };
    int next_seg = 0;                                  // segments [0, next_seg) have been flushed by this warp
    if (grp < n) {
      Pos p = pos_of(u0 + grp, KP);
      uint32_t zq[kPF], sq[kPF];                       // zero words / scale bits of this group's next kPF units
      Pos pz = p;
#pragma unroll
      for (int k = 0; k < kPF; ++k) {
        zq[k] = 0;
        sq[k] = 0;
        if (grp + k * kDG < n) {
          zq[k] = zload(pz);
          sq[k] = sload(pz);
          pos_advance(pz, kDG, KP);
        }
      }
      int wslot = 0;
      uint32_t wphase = 0;
      int ai = 0;
      uint32_t aphase = 1;                             // afree: the first round passes
      int pending = -1;
      uint32_t cur[W];
      int kk = -1;                                     // ordinal of the group's current segment (as the issuer counts)
      int cur_seg = -1;
      int pend_seg = -1, pend_kk = -1;                 // a finished segment whose read-out is deferred by one unit
      while (next_seg < p.tile - t_first) flush(next_seg++, -1);   // segments before this group's first unit
      wait_bar(s_wfull, 0);
      load_chunk(cur, 0, 0);
      for (int j = grp; j < n; j += kDG) {
        const int seg = p.tile - t_first;
        if (seg != cur_seg) {
          if (cur_seg >= 0) { pend_seg = cur_seg; pend_kk = kk; }
          cur_seg = seg;
          ++kk;
        }
        const uint32_t z = has_z ? ((zq[0] >> zshift) & zmask) + zadd : (1u << (BITS - 1));
        // the constants of make_dq (wna16.cuh) as bit patterns: 1024 + z, 64 + z, (s, s) -- the SAME arithmetic as
        // the round-1 kernel and the reference: w = fp16((q - z) * s), one rounding
        DqConst dq;
        dq.s2 = u32_as_h2(sq[0] * 0x00010001u);
        dq.zlo = u32_as_h2((0x6400u | z) * 0x00010001u);
        dq.zhi = u32_as_h2((0x5400u | (z << 4)) * 0x00010001u);
        const uint32_t zb16 = (0x4300u | z) * 0x00010001u;                          // bf16 128 + z
        const uint32_t sb16 = sq[0] * 0x00010001u;                                  // bf16 (s, s)
        const bool more = j + kDG < n;
#pragma unroll
        for (int k = 0; k + 1 < kPF; ++k) {
          zq[k] = zq[k + 1];
          sq[k] = sq[k + 1];
        }
        if (j + kPF * kDG < n) {
          zq[kPF - 1] = zload(pz);
          sq[kPF - 1] = sload(pz);
          pos_advance(pz, kDG, KP);
        }
        auto dq_half = [&](int hf, uint32_t (&dst)[16]) {
          if (BITS == 4) {
#pragma unroll
            for (int w = 0; w < 4; ++w) {
              uint32_t o4[4];
              if constexpr (kBf16) dequant4_word_bf16(cur[hf * 4 + w], zb16, sb16, o4);
              else dequant4_word(cur[hf * 4 + w], dq, o4);
#pragma unroll
              for (int e = 0; e < 4; ++e) dst[w * 4 + e] = o4[e];
            }
          } else {
#pragma unroll
            for (int w = 0; w < 8; ++w) {
              uint32_t o2[2];
              dequant8_word(cur[hf * 8 + w], dq, o2);
              dst[w * 2] = o2[0];
              dst[w * 2 + 1] = o2[1];
            }
          }
        };
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          uint32_t lo[16], hi[16];
          dq_half(0, lo);
          if (pending >= 0) {
            // deferred completion of the previous chunk's TMEM stores: their latency ran under the ALU work above
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane0) arrive_bar(s_stored + pending * 8);
          }
          const int as = grp * kAS + ai;
          const uint32_t a_tmem = tmem_base + C::kACol0 + as * 32 + lane_base;
          wait_bar(s_afree + as * 8, aphase);
          tc_fence_after();
          tmem_st_x16(a_tmem, lo);
          dq_half(1, hi);
          // ---- packed words of the next chunk (cur is consumed)
          if (cc == 0) {
            load_chunk(cur, wslot, 1);
          } else {
            __syncwarp();
            if (lane0) arrive_bar(s_wempty + wslot * 8);   // this warp's quarter of the stage is in registers
            if (++wslot == C::kSWS) { wslot = 0; wphase ^= 1; }
            if (more) {
              wait_bar(s_wfull + wslot * 8, wphase);
              load_chunk(cur, wslot, 0);
            }
          }
          tmem_st_x16(a_tmem + 16, hi);
          pending = as;
          if (++ai == kAS) { ai = 0; aphase ^= 1; }
        }
        // ---- a segment that ended one unit ago: its MMAs have had a unit's time to finish
        if (pend_seg >= 0) {
          flush(pend_seg, pend_kk);
          next_seg = pend_seg + 1;
          pend_seg = -1;
          // segments between that one and the current one hold no unit of this group: their zeros go out NOW (a
          // group that sat on them until its next crossing could stall the others' flushes behind hfree)
          while (next_seg < cur_seg) flush(next_seg++, -1);
        }
        pos_advance(p, kDG, KP);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane0) arrive_bar(s_stored + pending * 8);
      while (next_seg < cur_seg) flush(next_seg++, -1);
      flush(cur_seg, kk);
      next_seg = cur_seg + 1;
    }
    while (next_seg <= t_last - t_first) flush(next_seg++, -1);
    pdl_wait();                                          // first global writes of this role
    segment_epilogue(t_last, t_last - t_first, warp, kAllEpiWarps, warp * 32 + lane, 3, false);
  } else {
    // ------------------------------------------------------------ epilogue warps: group sums -> partial tile -> meeting point
    const int tid = static_cast<int>(tid_o) - kEpiWarp0 * 32;
    const int ewarp = tid >> 5;
    pdl_wait();
    const int t_first = pos_of(u0, KP).tile, t_last = pos_of(u0 + n - 1, KP).tile;
    for (int t = t_first; t < t_last; ++t) segment_epilogue(t, t - t_first, ewarp, kEpiWarps, tid, 1, true);
    segment_epilogue(t_last, t_last - t_first, kDqWarps + ewarp, kAllEpiWarps, kDqWarps * 32 + tid, 3, false);
  }
  __syncthreads();
  if (!MMA && warp == kWProdWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// B200_DEC_SK: 0 = never, 1 = every shape the kernel takes (M <= 32), 2 (default) = where it measured faster than the
// round-1 kernel on B200 IN THE DECODE STEP (profiles/r2_ai_sk_policy_ab.txt, same-box A/B): at most 16 tokens on
// projections with at least 32 units per CTA (gate|up): b=1 2.55 -> 2.44 ms per step.  At 32 tokens the kernel is
// faster in isolation (gate|up 31.2 vs 32.4 us) but the step gets SLOWER (5.05 -> 5.20 ms): a persistent grid ends on
// all SMs at once, while the round-1 kernel's half-empty second wave lets the next GEMM's CTAs start streaming their
// weights early under programmatic dependent launch.  The small-N projections (qkv, o, down: 10-25 units per CTA)
// stay on the round-1 kernel at every batch: 12.4 / 8.7 / 16.2 vs 9.6 / 8.2 / 15.4 us at 32 tokens.
static int sk_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_DEC_SK");
    v = e ? atoi(e) : 2;
  }
  return v;
}

template <int BITS, int NT, int EPI, bool MMA = false>
static bool launch_epi(const WnaParams& p, cudaStream_t st) {
  using C = Cfg<BITS, NT>;
  const int KC = p.K / kChunkK, KP = p.K / kUnitK;
  const int tiles = p.N / kTileN;
  const long units = static_cast<long>(tiles) * KP;
  // one CTA per SM; never less than two units (four chunks) per CTA, as in the round-1 split rule
  long G = sm_count();
  if (G > units / 2) G = units / 2 > 0 ? units / 2 : 1;
  int share = static_cast<int>((units + G - 1) / G);
  // Small-N projections (qkv, o, down): cut every tile into the SAME s <= 4 aligned pieces (s | KP), as the round-1
  // kernel does -- the last arriver then sums at most four partial tiles, one load batch of the finaliser; a free-
  // running share cut `o` into six and `down` into five or six.  Big-N projections keep the pure stream-K share.
  if (static_cast<long>(tiles) * 4 <= sm_count() + sm_count() / 8) {
    for (int sp = 4; sp >= 1; --sp)
      if (KP % sp == 0 && static_cast<long>(tiles) * sp <= sm_count() && KP / sp >= 2) {
        share = KP / sp;
        break;
      }
  }
  G = (units + share - 1) / share;
  const int smax = (share + KP - 2) / KP + 1;            // tiles a range of `share` units can touch
  if (smax > kMaxSeg) return false;
  if (G * smax * kTileN > kSkSlotCols) return false;     // partial slots must fit b200_wna16_workspace_bytes
  CUtensorMap xmap;
  {
    const uint64_t dims[3] = {kChunkK, static_cast<uint64_t>(p.M), static_cast<uint64_t>(KC)};
    const uint64_t strides[2] = {static_cast<uint64_t>(p.K) * 2, kChunkK * 2};
    const uint32_t box[3] = {kChunkK, static_cast<uint32_t>(NT), 2};
    const uint32_t es[3] = {1, 1, 1};
    if (encode_tensor_map(&xmap, EPI == kEpiNoneBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                          3, const_cast<__half*>(p.a), dims, strides,
                          box, es, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B) != CUDA_SUCCESS) {
      set_error(2, "wna16_sk: cuTensorMapEncodeTiled (activations) failed");
      return true;
    }
  }
  Args a;
  a.bp = p.bp; a.scales = p.scales; a.qzeros = p.qzeros; a.bias = p.bias; a.c = p.c;
  a.counters = reinterpret_cast<uint32_t*>(p.workspace);
  a.partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p.workspace) + kCounterBytes);
  a.M = p.M; a.K = p.K; a.N = p.N; a.qfmt = p.qfmt;
  a.gshift = 30;                                          // one group: every unit reads row 0
  if (p.group_size < p.K)
    for (int sft = 0; sft < 24; ++sft)
      if ((kUnitK << sft) == p.group_size) a.gshift = sft;
  a.KP = KP; a.units = static_cast<int>(units); a.share = share; a.smax = smax;
  if (p.fe) a.fe = *p.fe; else memset(&a.fe, 0, sizeof(a.fe));
  if (EPI == kEpiNoneBf16) a.fe.out_bf16 = 1;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(wna16_sk_kernel<BITS, NT, EPI, MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kAlloc);
    attr_set = true;
  }
  launch_pdl(wna16_sk_kernel<BITS, NT, EPI, MMA>, dim3(static_cast<unsigned>(G)), dim3(MMA ? kThreadsMma : kThreads),
             C::kAlloc, st, xmap, a);
  B200_CUDA_LAUNCH_CHECK();
  return true;
}

template <int BITS, int NT, bool MMA = false>
static bool launch_nt(const WnaParams& p, cudaStream_t st) {
  switch (p.epilogue) {
    case B200_EPI_SILU_MUL: return launch_epi<BITS, NT, B200_EPI_SILU_MUL, MMA>(p, st);
    case kEpiResidNorm: return launch_epi<BITS, NT, kEpiResidNorm, MMA>(p, st);
    case kEpiRopeKv: return launch_epi<BITS, NT, kEpiRopeKv, MMA>(p, st);
    case kEpiPushPeers: return launch_epi<BITS, NT, kEpiPushPeers, MMA>(p, st);
    default: return launch_epi<BITS, NT, B200_EPI_NONE, MMA>(p, st);
  }
}

// B200_DEC_SKM: which of the two flavours runs where the stream-K kernel is taken: 0 = tcgen05 (TMEM) flavour,
// 1 = mma.sync register flavour (4-bit fp16 only).
static int skm_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_DEC_SKM");
    v = e ? atoi(e) : 0;
  }
  return v;
}

}  // namespace sk

// Decode-shape entry (M <= 32): false when the shape needs another kernel.  `force` skips the B200_DEC_SK switch
// (b200_wna16_gemm_ex variant 5, the bf16 entry points).
bool launch_wna16_sk(const WnaParams& p, cudaStream_t st, bool force, bool bf16, int flavour) {
  if (!force && sk::sk_mode() == 0) return false;
  if (p.M > 32) return false;
  if (!force && sk::sk_mode() == 2) {
    const long units = static_cast<long>(p.N / kTileN) * (p.K / sk::kUnitK);
    if (p.M > 16 || units < 32L * sm_count()) return false;
  }
  if (p.K % sk::kUnitK != 0) return false;
  if (p.group_size < p.K) {                             // a 128-K unit must lie inside one quantisation group,
    bool ok = false;                                    // and the group of a unit is a shift of its index
    for (int sft = 0; sft < 24; ++sft) ok = ok || (sk::kUnitK << sft) == p.group_size;
    if (!ok) return false;
  }
  if (p.epilogue == B200_EPI_SILU_MUL && ((p.N / kTileN) & 1)) return false;
  if (p.epilogue == kEpiRopeKv && p.fe && (kTileN % p.fe->D) != 0) return false;
  if (bf16) {
    if (p.bits != 4 || p.epilogue != B200_EPI_NONE || p.bias || p.fe) return false;
    return p.M <= 16 ? sk::launch_epi<4, 16, sk::kEpiNoneBf16>(p, st) : sk::launch_epi<4, 32, sk::kEpiNoneBf16>(p, st);
  }
  static const bool force_nt32 = getenv("B200_SK_FORCE_NT32") != nullptr;      // experiment: 32-column tiles for any M
  const bool mma = flavour < 0 ? sk::skm_mode() == 1 : flavour == 1;
  if (p.bits == 4 && mma) return p.M <= 16 ? sk::launch_nt<4, 16, true>(p, st) : sk::launch_nt<4, 32, true>(p, st);
  if (p.bits == 4) return (p.M <= 16 && !force_nt32) ? sk::launch_nt<4, 16>(p, st) : sk::launch_nt<4, 32>(p, st);
  return p.M <= 16 ? sk::launch_nt<8, 16>(p, st) : sk::launch_nt<8, 32>(p, st);
}

}  // namespace b200
