// W4A16 / W8A16 GEMM on tcgen05 for DECODE batches (M <= 64): latency-optimised variant.
//
// Same math and data layout as wna16_tc.cu (swap-AB: weight tile = MMA A operand written to TMEM
// by the dequant threads, activations = B operand through TMA), but organised so that no chunk
// ever waits on another role's wake-up:
//   * 4 dequant groups (4 warps = all 128 channels each); group g owns chunks g, g+4, ... with its
//     OWN fp32 accumulator in TMEM and its OWN A stages, so the four groups' latency chains
//     (smem read -> dequant -> tcgen05.st -> wait::st) never wait for each other; a warp only
//     arrives on its stage's `stored` mbarrier and moves on (no intra-group barrier: that
//     barrier was 20% of all stall samples when a dequant thread also issued the MMAs).  Each
//     group has its own issuer warp that issues the four tcgen05.mma of a chunk into the group's
//     accumulator and commits; the four partial accumulators are summed in the epilogue.
//   * scales / zero points of the CTA's K range arrive through cp.async.bulk issued at t = 0
//     (no synchronous global loads on the critical path); group indices are shifts.
//   * the weight producer never waits for the previous kernel (weights do not depend on it), so
//     under PDL the 48 KB ring is already full when the activations become visible.
// Measured motivation (profiles/r1_gemm_experiments.txt): v3 spent ~10 us per launch in prologue
// + epilogue and ~900 cycles per 4 KB chunk with the tensor pipe < 10% busy.
#include <cstdlib>
#include <cstring>

#include "wna16.cuh"

namespace b200 {
namespace dec {

#ifdef B200_DEC_TRACE
// Pipeline trace (tools/dec_trace.py): SM clock of each role's events for CTA (0,0), one slot per
// (event, chunk).  Compiled only with -DB200_DEC_TRACE; production builds carry none of it.
__device__ uint32_t* g_dec_trace = nullptr;
#define DEC_TR(ev, i)                                                                        \
  do {                                                                                       \
    if (g_dec_trace && blockIdx.x == 0 && blockIdx.y == 0 && (i) < 256)                      \
      g_dec_trace[(ev) * 256 + (i)] = static_cast<uint32_t>(clock64());                      \
  } while (0)
// epilogue events of tile 0, one row per split: row 8 + blockIdx.y, column = event id
#define DEC_TRE(k)                                                                           \
  do {                                                                                       \
    if (g_dec_trace && blockIdx.x == 0 && blockIdx.y < 8 && threadIdx.x == 0)               \
      g_dec_trace[(8 + blockIdx.y) * 256 + (k)] = static_cast<uint32_t>(clock64());          \
  } while (0)
#else
#define DEC_TR(ev, i) do { } while (0)
#define DEC_TRE(k) do { } while (0)
#endif

#ifndef B200_DEC_EXP
#define B200_DEC_EXP 0            // timing experiments (side builds only; results are wrong on purpose): bit 0 no TMEM
#endif                            // stores, bit 1 no MMAs, bit 2 no activation loads
#ifndef B200_DEC_PROBE
#define B200_DEC_PROBE 0          // 1 = early non-blocking probes of afree / wfull: measured neutral (profiles/r2_v_*)
#endif
#ifndef B200_DEC_UNROLL
#define B200_DEC_UNROLL 0         // 1 = ring-unrolled dequant loop: -5 % in tools/gemm_bench.py, nothing in the decode
#endif                            // step (cold instruction fetch of a 50 % larger kernel): the round-1 loop stays the default
#ifndef B200_DEC_SXS
#define B200_DEC_SXS 3            // activation super-stages
#endif
#ifndef B200_DEC_GROUPS
#define B200_DEC_GROUPS 4          // dequant groups (experiment: -DB200_DEC_GROUPS=5 via B200_NVCC_EXTRA)
#endif
constexpr int kGroups = B200_DEC_GROUPS;
constexpr int kDqWarps = 4 * kGroups;
constexpr int kDqThreads = kDqWarps * 32;
constexpr int kThreads = (kDqWarps + 2 + kGroups) * 32;   // + weight producer, activation producer, one MMA issuer warp per group
constexpr int kSZGroups = 64;        // quantisation groups of the CTA's K range staged in smem

template <int BITS, int NT>
struct Cfg {
  static constexpr int kWBytes = kTileN * kChunkK * BITS / 8;            // one (tile, 64-K chunk)
  static constexpr int kXBytes = NT * kChunkK * 2;
  // Copies move SUPER-chunks of 4 consecutive K chunks (one per dequant group): the pipeline
  // trace (tools/dec_trace.py) showed one producer thread needs ~400 cycles per copy it issues,
  // which with 4 KB copies WAS the chunk period of the whole kernel.
  static constexpr int kSuper = kGroups;
  static constexpr int kWSBytes = kSuper * kWBytes;                       // 16 KB (4-bit) / 32 KB
  static constexpr int kXSBytes = kSuper * kXBytes;
  static constexpr int kSWS = 96 * 1024 / kWSBytes;                       // weight super-stages (96 KB)
  static constexpr int kSXS = NT <= 32 ? B200_DEC_SXS : 3;                // activation super-stages
  static constexpr int kXOff = 0;
  static constexpr int kWOff = kSXS * kXSBytes;
  static constexpr int kScaleOff = kWOff + kSWS * kWSBytes;               // [groups][128] fp16
  static constexpr int kZeroOff = kScaleOff + kSZGroups * kTileN * 2;     // [groups][ZW] u32
  static constexpr int kBarOff = kZeroOff + kSZGroups * 32 * 4;
  static constexpr int kNumBars = 2 * kSWS + 2 * kSXS + 2 * kGroups * 3 + kGroups + 1;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kAlloc = kTmemPtrOff + 16 + 1024;
  static_assert(kAlloc <= 227 * 1024, "shared memory budget");
  static constexpr int kDCols = NT < 32 ? 32 : NT;                        // per-group accumulator
  static constexpr int kAS = (NT <= 32 && kGroups <= 4) ? 3 : 2;          // A stages per group (32 columns each)
  static constexpr int kACol0 = kGroups * kDCols;
  static constexpr int kTmemCols = 512;
  static_assert(kACol0 + kGroups * kAS * 32 <= 512, "TMEM budget");
};

struct Args {
  const uint32_t* bp;
  const __half* scales;
  const uint32_t* qzeros;
  const __half* bias;
  __half* c;
  float* partials;
  uint32_t* counters;
  int M, K, N, g_shift, qfmt, chunks_per_split, splits, epilogue;
  int sz_rows;   // quantisation-group rows in the scale / zero TMA boxes   // g_shift: log2(group_size), -1 = one group
  FusedEpi fe;
};

template <int BITS, int NT, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
wna16_tc_dec_kernel(const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap smap,
                    const __grid_constant__ CUtensorMap zmap, const Args a) {
  using C = Cfg<BITS, NT>;
  const int nt = blockIdx.x, split = blockIdx.y;
  const int KC = a.K / kChunkK;
  const int c0 = split * a.chunks_per_split;
  const int c1 = min(c0 + a.chunks_per_split, KC);
  const int nchunks = c1 - c0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, M = a.M;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* wfull = bars;
  uint64_t* wempty = wfull + C::kSWS;
  uint64_t* xfull = wempty + C::kSWS;
  uint64_t* xempty = xfull + C::kSXS;
  uint64_t* afree = xempty + C::kSXS;                 // [group][stage]
  constexpr int kAS = C::kAS;
  uint64_t* gdone = afree + kGroups * kAS;           // [group]
  uint64_t* szfull = gdone + kGroups;
  uint64_t* stored = szfull + 1;                     // [group][stage]: the 4 warps' lanes are in TMEM
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + C::kTmemPtrOff);
  const __half* sm_scale = reinterpret_cast<const __half*>(smem + C::kScaleOff);
  const uint32_t* sm_zero = reinterpret_cast<const uint32_t*>(smem + C::kZeroOff);
  __shared__ int s_is_last;

  const int g_first = a.g_shift >= 0 ? (c0 * kChunkK) >> a.g_shift : 0;
  constexpr int ZW = kTileN * BITS / 32;             // zero-point words per tile row (16 / 32)

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::kSWS; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wempty[i], kDqWarps);             // every dequant warp read its quarter
    }
    for (int i = 0; i < C::kSXS; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], kGroups);              // one commit per chunk of the super-chunk
    }
    for (int i = 0; i < kGroups * kAS; ++i) mbar_init(&afree[i], 1);
    for (int i = 0; i < kGroups; ++i) mbar_init(&gdone[i], 1);
    mbar_init(szfull, 1);
    for (int i = 0; i < kGroups * kAS; ++i) mbar_init(&stored[i], 4);
    fence_barrier_init();
  }
  if (warp == kDqWarps) {
    tmem_alloc(tmem_ptr, C::kTmemCols);
    tmem_relinquish();
  }
  if (warp == kDqWarps + 1 && lane == 0) {
    prefetch_tmap(&xmap);
    prefetch_tmap(&smap);
    prefetch_tmap(&zmap);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == kDqWarps) {
    // ------------------------------------------------------------ weight / scale producer
    if (lane == 0) {
      // scales / zero points of the CTA's K range: one 2-D TMA box each ([groups][128 channels])
      const uint32_t zbytes = a.qzeros ? ZW * 4 : 0;
      mbar_arrive_expect_tx(szfull, static_cast<uint32_t>(a.sz_rows) * (kTileN * 2 + zbytes));
      tma_load_2d(smem + C::kScaleOff, &smap, szfull, nt * kTileN, g_first);
      if (a.qzeros) tma_load_2d(smem + C::kZeroOff, &zmap, szfull, nt * ZW, g_first);
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a.bp) +
                           (static_cast<size_t>(nt) * KC + c0) * C::kWBytes;
      const int nsuper = (nchunks + C::kSuper - 1) / C::kSuper;
      for (int j = 0; j < nsuper; ++j) {
        const int ss = j % C::kSWS;
        mbar_wait(&wempty[ss], ((j / C::kSWS) & 1) ^ 1);
        const int left = nchunks - j * C::kSuper;
        const uint32_t bytes = (left < C::kSuper ? left : C::kSuper) * C::kWBytes;
        DEC_TR(0, j * C::kSuper);
        mbar_arrive_expect_tx(&wfull[ss], bytes);
        bulk_load_1d(smem + C::kWOff + ss * C::kWSBytes, src + static_cast<size_t>(j) * C::kWSBytes, bytes,
                     &wfull[ss]);
      }
    }
  } else if (warp == kDqWarps + 1) {
    // ------------------------------------------------------------ activation producer
    if (lane == 0) {
      pdl_wait();      // activations are the previous kernel's output; weights are not
      const int nsuper = (nchunks + C::kSuper - 1) / C::kSuper;
      for (int j = 0; j < nsuper; ++j) {
        const int xs = j % C::kSXS;
        mbar_wait(&xempty[xs], ((j / C::kSXS) & 1) ^ 1);
        DEC_TR(6, j * C::kSuper);
#if B200_DEC_EXP & 4           // TIMING EXPERIMENT bit 2: the stage is declared full without loading anything
        mbar_arrive(&xfull[xs]);
#else
        mbar_arrive_expect_tx(&xfull[xs], C::kXSBytes);       // TMA zero-fills rows >= M / chunks >= KC
        tma_load_3d(smem + C::kXOff + xs * C::kXSBytes, &xmap, &xfull[xs], 0, 0, c0 + j * C::kSuper);
#endif
      }
    }
  } else if (warp >= kDqWarps + 2) {
    // ------------------------------------------------------------ MMA issuers (one warp per group)
    // A single in-order issuer for all groups was measured 30% slower (its wait -> issue ->
    // commit loop takes longer per chunk than the dequant of a chunk): four independent issue
    // loops keep the four accumulators busy.
    if (lane == 0) {
      const int grp = warp - (kDqWarps + 2);
      const uint32_t idesc = make_idesc_f16(128, NT < 16 ? 16 : NT);
      const uint32_t d_tmem = tmem_base + grp * C::kDCols;
      int it = 0;
      for (int i = grp; i < nchunks; i += kGroups, ++it) {
        const int as = grp * kAS + (it % kAS);
        const int xs = it % C::kSXS;                 // chunk i = 4 * it + grp: slot grp of super-chunk it
        DEC_TR(7, i);
#if B200_DEC_PROBE
        mbar_wait(&xfull[xs], (it / C::kSXS) & 1);   // long since complete: its round trip runs before the dequant
        mbar_wait(&stored[as], (it / kAS) & 1);      // warps arrive, not after
#else
        mbar_wait(&stored[as], (it / kAS) & 1);
        mbar_wait(&xfull[xs], (it / C::kSXS) & 1);
#endif
        tc_fence_after();
        DEC_TR(5, i);
        const uint32_t a_tmem = tmem_base + C::kACol0 + as * 32;
        const uint64_t bdesc =
            make_smem_desc_sw128(smem_u32(smem + C::kXOff + xs * C::kXSBytes + grp * C::kXBytes));
#if !(B200_DEC_EXP & 2)      // TIMING EXPERIMENT bit 1 (side builds): no MMAs (commits only)
#pragma unroll
        for (int ks = 0; ks < kChunkK / 16; ++ks)
          tc_mma_ts(d_tmem, a_tmem + ks * 8, bdesc + static_cast<uint64_t>(ks * 2), idesc,
                    (it > 0 || ks > 0) ? 1u : 0u);
#endif
        tc_commit(&xempty[xs]);
        tc_commit(&afree[as]);
        if (i + kGroups >= nchunks) tc_commit(&gdone[grp]);      // the group's last chunk
      }
    }
  } else {
    // ------------------------------------------------------------ dequant groups
    const int grp = warp >> 2, q4 = warp & 3;
    const int ch = q4 * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    // zero-point field of this channel inside its packed word
    const int zword = BITS == 4 ? (ch >> 3) : (ch >> 2);
    const int zshift = BITS == 4 ? ((a.qfmt == B200_QFMT_AWQ) ? 4 * awq_nibble_of(ch & 7) : 4 * (ch & 7))
                                 : 8 * (ch & 3);
    const int zadd = (a.qfmt == B200_QFMT_AWQ && BITS == 4) ? 0 : 1;
    constexpr int W = 2 * BITS;
    // Everything below addresses shared memory through 32-bit shared-space addresses (ld.shared):
    // plain C++ pointers into the manually aligned buffer compile to generic LD.E.
    const uint32_t s_base = smem_u32(smem);
    const uint32_t s_wfull = smem_u32(wfull), s_wempty = smem_u32(wempty);
    const uint32_t s_afree = smem_u32(afree), s_stored = smem_u32(stored);
    auto bar_wait = [](uint32_t bar, uint32_t parity) {
      while (!mbar_try_wait_addr(bar, parity)) {
      }
    };
    auto bar_arrive = [](uint32_t bar) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
    };
    int ss = 0;                         // weight super-stage of the group's current chunk
    uint32_t ss_phase = 0;
    auto load_words = [&](uint32_t (&dst)[W], int ci, uint32_t probed = 0) {
      if (q4 == 0 && lane == 0) DEC_TR(1, ci);
      if (!probed) bar_wait(s_wfull + ss * 8, ss_phase);
      if (q4 == 0 && lane == 0) DEC_TR(2, ci);
      const uint32_t addr = s_base + C::kWOff + ss * C::kWSBytes + grp * C::kWBytes + ch * 16;
#pragma unroll
      for (int sub = 0; sub < W / 4; ++sub)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(dst[sub * 4]), "=r"(dst[sub * 4 + 1]), "=r"(dst[sub * 4 + 2]), "=r"(dst[sub * 4 + 3])
                     : "r"(addr + sub * kTileN * 16));
      __syncwarp();
      if (lane == 0) bar_arrive(s_wempty + ss * 8);   // packed bytes are in registers
      if (++ss == C::kSWS) { ss = 0; ss_phase ^= 1; }
    };
    uint32_t cur[W];
    if (grp < nchunks) load_words(cur, grp);
    mbar_wait(szfull, 0);
    int cur_g = -1;
    DqConst dq = make_dq(__float2half(0.f), 0);
    int as_l = 0;                       // A stage of this group, 0..kAS-1
    uint32_t as_phase = 1;              // parity to wait for on afree (first round passes)
    const uint32_t s_scale = s_base + C::kScaleOff + ch * 2;
    const uint32_t s_zero = s_base + C::kZeroOff + zword * 4;
    const uint32_t zmask = (1u << BITS) - 1;
    auto dequant_half = [&](int i, int hf, uint32_t (&dst)[16]) {
      const int g = a.g_shift >= 0 ? (((c0 + i) * kChunkK + hf * 32) >> a.g_shift) - g_first : 0;
      if (g != cur_g) {
        cur_g = g;
        unsigned short sbits;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(sbits) : "r"(s_scale + g * (kTileN * 2)));
        int z = 1 << (BITS - 1);
        if (a.qzeros) {
          uint32_t zw;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(zw) : "r"(s_zero + g * (ZW * 4)));
          z = static_cast<int>((zw >> zshift) & zmask) + zadd;
        }
        dq = make_dq(__ushort_as_half(sbits), z);
      }
      if (BITS == 4) {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          uint32_t o4[4];
          dequant4_word(cur[hf * 4 + w], dq, o4);
#pragma unroll
          for (int e = 0; e < 4; ++e) dst[w * 4 + e] = o4[e];
        }
      } else {
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          uint32_t o2[2];
          dequant8_word(cur[(hf * 8 + w) % W], dq, o2);
          dst[w * 2] = o2[0];
          dst[w * 2 + 1] = o2[1];
        }
      }
    };
#if B200_DEC_UNROLL
    {
      // Ring-unrolled form of the loop below (round 2).  The ncu source view says the loop is bound by its serial
      // bookkeeping, not by the dequant arithmetic (removing 60 % of the arithmetic bought 10 %): ring-stage indices,
      // barrier addresses, phase bits and TMEM columns recomputed per chunk with dependent ISETP / SEL / IMAD chains.
      // Unrolling by P = lcm(weight stages, A stages) chunks makes every one of them a compile-time constant of the
      // unrolled copy; what is left per chunk is the waits, the loads, the stores and the arithmetic.
      constexpr int kP = (C::kSWS % kAS == 0) ? C::kSWS : C::kSWS * kAS;       // lcm for (6,3) (3,3) (6,2) (3,2)
      static_assert(kP % C::kSWS == 0 && kP % kAS == 0, "ring period");
      const uint32_t s_wq = s_base + C::kWOff + grp * C::kWBytes + ch * 16;    // this thread's words of stage 0
      const uint32_t a_tq = tmem_base + C::kACol0 + grp * kAS * 32 + lane_base;
      const uint32_t s_af = s_afree + grp * kAS * 8, s_st = s_stored + grp * kAS * 8;
      const bool lane0 = lane == 0;
      uint32_t wpar = 0;                // parity of wfull for the stages of this pass over the ring
      uint32_t apar = 1;                // parity of afree (the first round passes)
      uint32_t zlo = 0, zhi = 0, s2 = 0;
      auto set_group = [&](int g) {
        cur_g = g;
        uint32_t sbits;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(sbits) : "r"(s_scale + g * (kTileN * 2)));
        uint32_t z = 1u << (BITS - 1);
        if (a.qzeros) {
          uint32_t zw;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(zw) : "r"(s_zero + g * (ZW * 4)));
          z = ((zw >> zshift) & zmask) + static_cast<uint32_t>(zadd);
        }
        // the constants of make_dq as bit patterns (no int -> half conversions): 1024 + z, 64 + z, (s, s)
        s2 = sbits * 0x00010001u;
        zlo = (0x6400u | z) * 0x00010001u;
        zhi = (0x5400u | (z << 4)) * 0x00010001u;
        dq.s2 = u32_as_h2(s2);
        dq.zlo = u32_as_h2(zlo);
        dq.zhi = u32_as_h2(zhi);
      };
      auto dq_half = [&](int hf, uint32_t (&dst)[16]) {
        if (BITS == 4) {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            uint32_t o4[4];
            dequant4_word(cur[hf * 4 + w], dq, o4);
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[w * 4 + e] = o4[e];
          }
        } else {
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            uint32_t o2[2];
            dequant8_word(cur[(hf * 8 + w) % W], dq, o2);
            dst[w * 2] = o2[0];
            dst[w * 2 + 1] = o2[1];
          }
        }
      };
      const bool per_half = a.g_shift == 5;            // 32-wide groups: the two halves of a chunk differ
      bool have_pending = false;
      bool running = grp < nchunks;
      int i = grp;
      while (running) {
#pragma unroll
        for (int u = 0; u < kP; ++u) {
          if (i >= nchunks) { running = false; break; }
          // every index below is a constant of this unrolled copy
          const int ssn = (u + 1) % C::kSWS;                                   // stage of the group's NEXT chunk
          const uint32_t wphn = wpar ^ static_cast<uint32_t>(((u + 1) / C::kSWS) & 1);
          const int asl = u % kAS, aslp = (u + kAS - 1) % kAS;
          const uint32_t aph = apar ^ static_cast<uint32_t>((u / kAS) & 1);
          const bool more = i + kGroups < nchunks;
          uint32_t lo[16], hi[16];
          const uint32_t af_ok = mbar_probe_addr(s_af + asl * 8, aph);
          const int g = a.g_shift >= 0 ? (((c0 + i) * kChunkK) >> a.g_shift) - g_first : 0;
          if (g != cur_g) set_group(g);
          dq_half(0, lo);
          if (have_pending) {
            tc_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane0) bar_arrive(s_st + aslp * 8);
            if (q4 == 0 && lane0) DEC_TR(4, i - kGroups);
          }
          if (!af_ok) bar_wait(s_af + asl * 8, aph);
          tc_fence_after();
          if (q4 == 0 && lane0) DEC_TR(3, i);
          tmem_st_x16(a_tq + asl * 32, lo);
          const uint32_t wf_ok = more ? mbar_probe_addr(s_wfull + ssn * 8, wphn) : 0u;
          if (per_half) {
            const int g1 = (((c0 + i) * kChunkK + 32) >> 5) - g_first;
            if (g1 != cur_g) set_group(g1);
          }
          dq_half(1, hi);
          if (more) {
            if (q4 == 0 && lane0) DEC_TR(1, i + kGroups);
            if (!wf_ok) bar_wait(s_wfull + ssn * 8, wphn);
            if (q4 == 0 && lane0) DEC_TR(2, i + kGroups);
#pragma unroll
            for (int sub = 0; sub < W / 4; ++sub)
              asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                           : "=r"(cur[sub * 4]), "=r"(cur[sub * 4 + 1]), "=r"(cur[sub * 4 + 2]), "=r"(cur[sub * 4 + 3])
                           : "r"(s_wq + ssn * C::kWSBytes + sub * kTileN * 16));
            __syncwarp();
            if (lane0) bar_arrive(s_wempty + ssn * 8);
          }
          tmem_st_x16(a_tq + asl * 32 + 16, hi);
          have_pending = true;
          i += kGroups;
        }
        wpar ^= static_cast<uint32_t>((kP / C::kSWS) & 1);
        apar ^= static_cast<uint32_t>((kP / kAS) & 1);
      }
      if (have_pending) {
        // the last chunk's stage: chunk ordinal (count - 1) % kAS
        const int cnt = (nchunks - grp + kGroups - 1) / kGroups;
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane0) bar_arrive(s_st + ((cnt - 1) % kAS) * 8);
      }
    }
#else
    {
      // Deferred completion: the TMEM stores of chunk i are awaited (and the issuer signalled)
      // only after the first half of chunk i+1 has been dequantised, so tcgen05.st latency
      // overlaps ALU work instead of idling the warp.
      int pending_as = -1;
      for (int i = grp; i < nchunks; i += kGroups) {
        uint32_t lo[16], hi[16];
        const int as = grp * kAS + as_l;
#if B200_DEC_PROBE
        // is the A stage free?  Asked before the ALU work whose result goes there (a passing mbarrier wait costs a
        // ~200-cycle round trip: profiles/r2_t_dec_trace_floor.txt)
        const uint32_t af_ok = mbar_probe_addr(s_afree + as * 8, as_phase);
#else
        const uint32_t af_ok = 0;
#endif
        dequant_half(i, 0, lo);
        if (pending_as >= 0) {
          tc_wait_st();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) bar_arrive(s_stored + pending_as * 8);
          if (q4 == 0 && lane == 0) DEC_TR(4, i - kGroups);
        }
        const uint32_t a_tmem = tmem_base + C::kACol0 + as * 32;
        if (!af_ok) bar_wait(s_afree + as * 8, as_phase);
        tc_fence_after();
        if (q4 == 0 && lane == 0) DEC_TR(3, i);
#if !(B200_DEC_EXP & 1)      // TIMING EXPERIMENT bit 0 (side builds): no TMEM stores
        tmem_st_x16(a_tmem + lane_base, lo);
#else
        asm volatile("" ::"r"(lo[0]), "r"(lo[5]), "r"(lo[10]), "r"(lo[15]));
#endif
#if B200_DEC_PROBE
        const uint32_t wf_ok = (i + kGroups < nchunks) ? mbar_probe_addr(s_wfull + ss * 8, ss_phase) : 0;
#else
        const uint32_t wf_ok = 0;
#endif
        dequant_half(i, 1, hi);
        if (i + kGroups < nchunks) load_words(cur, i + kGroups, wf_ok);
#if !(B200_DEC_EXP & 1)
        tmem_st_x16(a_tmem + lane_base + 16, hi);
#else
        asm volatile("" ::"r"(hi[0]), "r"(hi[5]), "r"(hi[10]), "r"(hi[15]));
#endif
        pending_as = as;
        if (++as_l == kAS) { as_l = 0; as_phase ^= 1; }
      }
      if (pending_as >= 0) {
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(s_stored + pending_as * 8);
      }
    }
#endif
    // ------------------------------------------------------------ epilogue
    // (the kernel is templated on the epilogue kind: one lean finaliser per kernel -- code
    // size and register allocation of this tail were measured to move the whole GEMM by 2-5 us)
    constexpr bool kPlain = EPI == B200_EPI_NONE;    // EPI_NONE kernels never carry a row scale
    const bool direct = kPlain && a.splits == 1;
    DEC_TRE(0);
    pdl_wait();        // first global reads / writes below (c / partials / counters / residual)
    if (warp < 4) {
      const int n = nt * kTileN + ch;
      const int active = nchunks < kGroups ? nchunks : kGroups;
      for (int g = 0; g < active; ++g) mbar_wait(&gdone[g], 0);
      tc_fence_after();
      DEC_TRE(1);
      const float bv = (direct && a.bias) ? __half2float(a.bias[n]) : 0.f;
      float* part = a.partials + (static_cast<size_t>(split) * M) * N;
#pragma unroll
      for (int j0 = 0; j0 < NT; j0 += 16) {
        float acc[16];
        {
          uint32_t v[16];
          tmem_ld_x16(tmem_base + lane_base + j0, v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]);
        }
        for (int g = 1; g < active; ++g) {
          uint32_t v[16];
          tmem_ld_x16(tmem_base + g * C::kDCols + lane_base + j0, v);
          tc_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] += __uint_as_float(v[j]);
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int m = j0 + j;
          if (m < M) {
            if (direct)
              a.c[static_cast<size_t>(m) * N + n] = __float2half_rn(acc[j] + bv);
            else
              part[static_cast<size_t>(m) * N + n] = acc[j];
          }
        }
      }
      tc_fence_before();
      DEC_TRE(2);
    }
    if (!direct) {
      SplitOut o;
      o.partials = a.partials; o.counters = a.counters; o.bias = a.bias; o.c = a.c;
      o.M = M; o.N = N; o.splits = a.splits; o.epilogue = EPI; o.fe = a.fe;
      if (kPlain) o.fe.ssq_in = nullptr;
      const bool last = split_arrive(o, nt, threadIdx.x, kDqThreads, 0, &s_is_last, 1);
      DEC_TRE(3);
      if (last) {
        split_reduce_rows<EPI>(o, nt, threadIdx.x, kDqThreads, 0, M);
        if (EPI == kEpiPushPeers) push_signal(o.fe, threadIdx.x, kDqThreads, 1);
        DEC_TRE(4);
      }
    }
  }
  __syncthreads();
  DEC_TRE(5);
  if (warp == kDqWarps) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

template <int BITS, int NT, int EPI>
static bool launch_epi(const WnaParams& p, cudaStream_t st) {
  using C = Cfg<BITS, NT>;
  const int KC = p.K / kChunkK;
  const int tiles = p.N / kTileN;
  int splits = sm_count() / tiles;
  if (splits < 1) splits = 1;
  if (splits > KC / 4) splits = KC / 4 > 0 ? KC / 4 : 1;
  {
    // more than eight K pieces per tile cost more at the meeting point than they save in the main loop (TP-rank
    // shapes, profiles/r2_ad_tp_shape_splits.txt: qkv at TP = 8, 6 tiles x 16 pieces 7.7 us -> x 8 pieces 6.5 us)
    static const int cap = getenv("B200_DEC_MAX_SPLITS") ? atoi(getenv("B200_DEC_MAX_SPLITS")) : 8;
    if (cap > 0 && splits > cap) splits = cap;
  }
  while (splits > 1 && static_cast<long>(splits) * p.N > kMaxSplitCols) --splits;
  if (p.no_split) splits = 1;
  int cps = (KC + splits - 1) / splits;
  splits = (KC + cps - 1) / cps;
  const bool grouped = p.group_size < p.K;
  const int total_groups = grouped ? p.K / p.group_size : 1;
  int sz_rows = grouped ? (cps * kChunkK + p.group_size - 1) / p.group_size + 1 : 1;
  if (sz_rows > kSZGroups) sz_rows = kSZGroups;
  // activations as a 3-D tensor {64 K elements, M rows, K/64 chunks}: one TMA box brings 4 chunks,
  // each landing as its own [NT rows][128 B] swizzled tile (the MMA B operand of one chunk)
  CUtensorMap xmap, smap, zmap;
  {
    const uint64_t dims[3] = {kChunkK, static_cast<uint64_t>(p.M), static_cast<uint64_t>(KC)};
    const uint64_t strides[2] = {static_cast<uint64_t>(p.K) * 2, kChunkK * 2};
    const uint32_t box[3] = {kChunkK, static_cast<uint32_t>(NT), static_cast<uint32_t>(C::kSuper)};
    const uint32_t es[3] = {1, 1, 1};
    if (encode_tensor_map(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<__half*>(p.a), dims,
                          strides, box, es, CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B) != CUDA_SUCCESS) {
      set_error(2, "wna16_tc_dec: cuTensorMapEncodeTiled (activations) failed");
      return true;
    }
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(p.N), static_cast<uint64_t>(total_groups)};
    const uint64_t strides[1] = {static_cast<uint64_t>(p.N) * 2};
    const uint32_t box[2] = {kTileN, static_cast<uint32_t>(sz_rows)};
    const uint32_t es[2] = {1, 1};
    if (encode_tensor_map(&smap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(p.scales), dims,
                          strides, box, es, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B) != CUDA_SUCCESS) {
      set_error(2, "wna16_tc_dec: cuTensorMapEncodeTiled (scales) failed");
      return true;
    }
  }
  zmap = smap;
  if (p.qzeros) {
    constexpr int ZW = kTileN * BITS / 32;
    const uint64_t zcols = static_cast<uint64_t>(p.N) * BITS / 32;
    const uint64_t dims[2] = {zcols, static_cast<uint64_t>(total_groups)};
    const uint64_t strides[1] = {zcols * 4};
    const uint32_t box[2] = {ZW, static_cast<uint32_t>(sz_rows)};
    const uint32_t es[2] = {1, 1};
    if (encode_tensor_map(&zmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<uint32_t*>(p.qzeros), dims,
                          strides, box, es, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B) != CUDA_SUCCESS) {
      set_error(2, "wna16_tc_dec: cuTensorMapEncodeTiled (zero points) failed");
      return true;
    }
  }
  Args a;
  a.bp = p.bp; a.scales = p.scales; a.qzeros = p.qzeros; a.bias = p.bias; a.c = p.c;
  a.counters = reinterpret_cast<uint32_t*>(p.workspace);
  a.partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p.workspace) + kCounterBytes);
  a.M = p.M; a.K = p.K; a.N = p.N; a.qfmt = p.qfmt;
  a.g_shift = -1;
  if (grouped)
    for (int s = 5; s < 31; ++s)
      if ((1 << s) == p.group_size) a.g_shift = s;
  a.chunks_per_split = cps; a.splits = splits; a.epilogue = p.epilogue;
  a.sz_rows = sz_rows;
  if (p.fe) a.fe = *p.fe; else memset(&a.fe, 0, sizeof(a.fe));
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(wna16_tc_dec_kernel<BITS, NT, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         C::kAlloc);
    attr_set = true;
  }
  launch_pdl(wna16_tc_dec_kernel<BITS, NT, EPI>, dim3(tiles, splits, 1), dim3(kThreads), C::kAlloc, st,
             xmap, smap, zmap, a);
  B200_CUDA_LAUNCH_CHECK();
  return true;
}

template <int BITS, int NT>
static bool launch_nt(const WnaParams& p, cudaStream_t st) {
  // a row scale (folded RMSNorm) on a plain GEMM has no lean kernel: the general kernel takes it
  if (p.epilogue == B200_EPI_NONE && p.fe && p.fe->ssq_in) return false;
  switch (p.epilogue) {
    case B200_EPI_SILU_MUL: return launch_epi<BITS, NT, B200_EPI_SILU_MUL>(p, st);
    case kEpiResidNorm: return launch_epi<BITS, NT, kEpiResidNorm>(p, st);
    case kEpiRopeKv: return launch_epi<BITS, NT, kEpiRopeKv>(p, st);
    case kEpiPushPeers: return launch_epi<BITS, NT, kEpiPushPeers>(p, st);
    default: return launch_epi<BITS, NT, B200_EPI_NONE>(p, st);
  }
}

}  // namespace dec

#ifdef B200_DEC_TRACE
extern "C" void b200_debug_set_dec_trace(void* buf) {
  uint32_t* p = static_cast<uint32_t*>(buf);
  cudaMemcpyToSymbol(dec::g_dec_trace, &p, sizeof(p));
}
#endif

// Decode-shape entry: false when the shape needs the general kernel (wna16_tc.cu).
bool launch_wna16_tc_dec(const WnaParams& p, cudaStream_t st) {
  if (p.M > 64) return false;
  const bool pow2 = p.group_size >= p.K || (p.group_size & (p.group_size - 1)) == 0;
  if (!pow2) return false;
  // scale / zero staging covers 64 quantisation groups of the CTA's K range
  const int KC = p.K / kChunkK;
  int splits = sm_count() / (p.N / kTileN);
  if (splits < 1) splits = 1;
  if (splits > KC / 4) splits = KC / 4 > 0 ? KC / 4 : 1;
  if (splits > 8) splits = 8;                 // same cap as launch_epi
  if (p.no_split) splits = 1;
  const int cps = (KC + splits - 1) / splits;
  const int groups_per_cta = p.group_size >= p.K ? 1 : (cps * kChunkK + p.group_size - 1) / p.group_size + 1;
  if (groups_per_cta > dec::kSZGroups) return false;
  if (p.epilogue == B200_EPI_SILU_MUL && ((p.N / kTileN) & 1)) return false;
  if (p.epilogue == kEpiRopeKv && p.fe && (kTileN % p.fe->D) != 0) return false;
#if B200_DEC_GROUPS > 4
  if (p.M > 32 || p.bits != 4) return false;   // five groups: TMEM / shared memory only fit the 4-bit, 32-token tiles
  if (p.M <= 16) return dec::launch_nt<4, 16>(p, st);
  return dec::launch_nt<4, 32>(p, st);
#else
  if (p.bits == 4) {
    if (p.M <= 16) return dec::launch_nt<4, 16>(p, st);
    if (p.M <= 32) return dec::launch_nt<4, 32>(p, st);
    return dec::launch_nt<4, 64>(p, st);
  }
  if (p.M <= 16) return dec::launch_nt<8, 16>(p, st);
  if (p.M <= 32) return dec::launch_nt<8, 32>(p, st);
  return dec::launch_nt<8, 64>(p, st);
#endif
}

}  // namespace b200
