// W4A16 / W8A16 GEMM on tcgen05 tensor cores with register dequant feeding TMEM.
//
// Replaces marlin_awq_4bit_f16 / marlin_4bit_f16 / gemm_half_q_half_alt
// (/root/reference/src/utils/gptq.rs:118-195) for every M > 4 (decode batches and prefill).
//
// "Swap-AB" formulation: D^T[n, m] = sum_k W^T[n, k] * X[m, k].  The WEIGHT tile is the MMA
// A operand (M_mma = 128 output channels = 128 TMEM lanes), the activations are the B operand
// (N_mma = NT tokens, 16..256), so a decode batch of 32 runs 128x32x16 MMAs with no wasted
// rows, and prefill simply uses NT = 256.
//
// Per CTA (one 128-channel weight tile x one K split x one token tile), warp roles:
//   warps 0-15 dequant, as 4 groups of 4 warps; group j owns chunks j, j+4, ... (4 chunk
//              latency chains in flight per SM).  Packed weights arrive in shared memory through
//              a deep ring of 4 KB cp.async.bulk copies (the tile-major prepack makes every stage
//              one contiguous block); thread = output channel: it reads its 32 bytes of the chunk,
//              dequantises 64 K values to 32 half2 registers (lop3 / prmt magic numbers, exact
//              (q - z), one rounding on the scale multiply; scales and zero points of the CTA's
//              K range are staged in shared memory once) and writes them with two
//              tcgen05.st 32x32b.x16 into its TMEM lane: the A operand never touches shared
//              memory again.
//   warp 16    weight producer (cp.async.bulk, 1-D)
//   warp 17    activation producer (TMA 2-D tensor map, SWIZZLE_128B, OOB rows zero-filled)
//   warp 18    MMA issuer: tcgen05.mma.kind::f16 with A from TMEM, B from the smem descriptor,
//              fp32 accumulators in TMEM; tcgen05.commit releases the rings.
//   epilogue   warps 0-3 read the accumulators with tcgen05.ld.  A full-K CTA stores fp16
//              directly.  K splits store fp32 partials (plain coalesced stores, no atomics); the
//              LAST CTA of a tile (self-resetting arrival counter) sums the partials in split
//              order -- deterministic -- and applies bias / SiLU*mul, so there is no second
//              kernel and nothing to zero.
//
// Roofline: decode (NT <= 64) is HBM-bound on the packed weights (K*N/2 bytes per launch);
// prefill (NT = 256) is tensor-bound: 2*M*N*K flop.
#include <cstdlib>
#include <cstring>

#include "wna16.cuh"

namespace b200 {

constexpr int kDqWarps = 16;
constexpr int kDqThreads = kDqWarps * 32;
constexpr int kTcThreads = (kDqWarps + 3) * 32;
constexpr int kGroups = 4;          // dequant warp groups (each owns every 4th chunk)
constexpr int kSA = 8;              // TMEM A stages (32 columns each)
constexpr int kSZGroups = 64;       // scale/zero groups staged in smem per CTA

template <int BITS, int NT>
struct TcCfg {
  static constexpr int kWBytes = kTileN * kChunkK * BITS / 8;      // packed weights per stage
  static constexpr int kXBytes = NT * kChunkK * 2;                 // activations per stage
  static constexpr int kSW = NT <= 64 ? (BITS == 4 ? 12 : 8) : (BITS == 4 ? 8 : 4);
  static constexpr int kSX = NT <= 64 ? 6 : (NT == 128 ? 5 : 4);
  static constexpr int kXOff = 0;
  static constexpr int kWOff = kSX * kXBytes;                      // 1024-aligned (kXBytes % 2048 == 0)
  static constexpr int kSZOff = kWOff + kSW * kWBytes;
  static constexpr int kBarOff = kSZOff + kSZGroups * kTileN * 4;
  static constexpr int kNumBars = 2 * kSW + 2 * kSX + 2 * kSA + 1;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kTotal = kTmemPtrOff + 16;
  static constexpr int kAlloc = kTotal + 1024;
  static constexpr int kDCols = NT < 32 ? 32 : NT;
  static constexpr int kTmemCols = (kDCols + kSA * 32) <= 256 ? 256 : 512;
};

struct TcArgs {
  const uint32_t* bp;
  const __half* scales;
  const uint32_t* qzeros;
  const __half* bias;
  __half* c;
  float* partials;        // [splits][M][N] fp32
  uint32_t* counters;     // tile arrival counters (zero at rest)
  int M, K, N, group_size, qfmt, chunks_per_split, splits, epilogue;
  FusedEpi fe;
  int dbg;                // B200_TC_DBG experiment switches (0 in production)
};

template <int BITS, int NT>
__global__ void __launch_bounds__(kTcThreads, 1)
wna16_tc_kernel(const __grid_constant__ CUtensorMap xmap, const TcArgs a) {
  using Cfg = TcCfg<BITS, NT>;
  const int nt = blockIdx.x, split = blockIdx.y;
  const int KC = a.K / kChunkK;
  const int c0 = split * a.chunks_per_split;
  const int c1 = min(c0 + a.chunks_per_split, KC);
  const int nchunks = c1 - c0;
  const int m0 = blockIdx.z * NT;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = a.N, M = a.M;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kBarOff);
  uint64_t* wfull = bars;
  uint64_t* wempty = wfull + Cfg::kSW;
  uint64_t* xfull = wempty + Cfg::kSW;
  uint64_t* xempty = xfull + Cfg::kSX;
  uint64_t* afull = xempty + Cfg::kSX;
  uint64_t* aempty = afull + kSA;
  uint64_t* dfull = aempty + kSA;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + Cfg::kTmemPtrOff);
  uint32_t* sz = reinterpret_cast<uint32_t*>(smem + Cfg::kSZOff);     // [group][128] (scale | z << 16)
  __shared__ int s_is_last;

  if (threadIdx.x == 0) {
    for (int i = 0; i < Cfg::kSW; ++i) {
      mbar_init(&wfull[i], 1);
      mbar_init(&wempty[i], 4);
    }
    for (int i = 0; i < Cfg::kSX; ++i) {
      mbar_init(&xfull[i], 1);
      mbar_init(&xempty[i], 1);
    }
    for (int i = 0; i < kSA; ++i) {
      mbar_init(&afull[i], 4);
      mbar_init(&aempty[i], 1);
    }
    mbar_init(dfull, 1);
    fence_barrier_init();
  }
  if (warp == kDqWarps + 2) {          // MMA warp owns the TMEM allocation
    tmem_alloc(tmem_ptr, Cfg::kTmemCols);
    tmem_relinquish();
  }
  if (warp == kDqWarps + 1 && lane == 0) prefetch_tmap(&xmap);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();     // let the next kernel's prologue / weight stream start under our tail
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t d_tmem = tmem_base;                       // columns [0, kDCols)
  const uint32_t a_tmem = tmem_base + Cfg::kDCols;         // kSA stages x 32 columns

  if (warp < kDqWarps) {
    // ------------------------------------------------------------ dequant warps
    const int q4 = warp & 3;                 // TMEM lane quarter
    const int ch = q4 * 32 + lane;
    const int n = nt * kTileN + ch;
    // stage scales / zero points of this CTA's K range
    const int g_first = (c0 * kChunkK) / a.group_size;
    const int g_last = (c1 * kChunkK - 1) / a.group_size;
    const int ngroups = g_last - g_first + 1;
    const bool sz_in_smem = ngroups <= kSZGroups;
    if (sz_in_smem) {
#pragma unroll 4
      for (int idx = threadIdx.x; idx < ngroups * kTileN; idx += kDqThreads) {
        const int g = g_first + idx / kTileN, cc = idx % kTileN, nn = nt * kTileN + cc;
        const uint32_t sbits = __half_as_ushort(a.scales[static_cast<size_t>(g) * N + nn]);
        const uint32_t z = static_cast<uint32_t>(load_zero<BITS>(a.qzeros, a.qfmt, g, nn, N));
        sz[idx] = sbits | (z << 16);
      }
      named_bar_sync(1, kDqThreads);
    }
    // Chunk-interleaved pipeline: warp group `grp` (4 warps = all 128 channels) owns chunks
    // grp, grp + 4, ... entirely, so four chunk latency chains (smem read -> dequant ->
    // tcgen05.st -> wait::st -> arrive) are in flight per SM instead of one; the packed words of
    // the group's next chunk are fetched before the current one is dequantised.
    constexpr int W = 2 * BITS;                       // packed 32-bit words per thread per chunk
    const int grp = warp >> 2;
    auto load_words = [&](int ci, uint32_t (&dst)[W]) {
      const int sw = ci % Cfg::kSW;
      mbar_wait(&wfull[sw], (ci / Cfg::kSW) & 1);
      // explicit ld.shared: a plain pointer into the manually aligned buffer compiles to generic LD.E
      const uint32_t addr = smem_u32(smem) + Cfg::kWOff + sw * Cfg::kWBytes + ch * 16;
#pragma unroll
      for (int sub = 0; sub < W / 4; ++sub)
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(dst[sub * 4]), "=r"(dst[sub * 4 + 1]), "=r"(dst[sub * 4 + 2]), "=r"(dst[sub * 4 + 3])
                     : "r"(addr + sub * kTileN * 16));
      __syncwarp();
      if (lane == 0) mbar_arrive(&wempty[sw]);        // packed bytes are in registers
    };
    int cur_g = -1;
    DqConst dq = make_dq(__float2half(0.f), 0);
    uint32_t cur[W];
    if (grp < nchunks) load_words(grp, cur);
    for (int i = grp; i < nchunks; i += kGroups) {
      uint32_t nxt[W];
      const bool has_next = i + kGroups < nchunks;
      if (has_next) load_words(i + kGroups, nxt);
      const int sa = i % kSA;
      uint32_t av[32];
      if (a.dbg & 4) {
#pragma unroll
        for (int e = 0; e < 32; ++e) av[e] = cur[e % W];
      } else
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {                // two 32-K halves (group boundary safe)
        const int g = ((c0 + i) * kChunkK + hf * 32) / a.group_size;
        if (g != cur_g) {
          cur_g = g;
          if (sz_in_smem) {
            const uint32_t v = sz[(g - g_first) * kTileN + ch];
            dq = make_dq(__ushort_as_half(static_cast<unsigned short>(v & 0xFFFF)), static_cast<int>(v >> 16));
          } else {
            dq = make_dq(a.scales[static_cast<size_t>(g) * N + n], load_zero<BITS>(a.qzeros, a.qfmt, g, n, N));
          }
        }
        if (BITS == 4) {
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            uint32_t o4[4];
            dequant4_word(cur[hf * 4 + w], dq, o4);
#pragma unroll
            for (int e = 0; e < 4; ++e) av[hf * 16 + w * 4 + e] = o4[e];
          }
        } else {
#pragma unroll
          for (int w = 0; w < 8; ++w) {
            uint32_t o2[2];
            dequant8_word(cur[(hf * 8 + w) % W], dq, o2);
            av[hf * 16 + w * 2] = o2[0];
            av[hf * 16 + w * 2 + 1] = o2[1];
          }
        }
      }
      mbar_wait(&aempty[sa], ((i / kSA) & 1) ^ 1);     // MMA of chunk i - kSA has retired
      tc_fence_after();
      {
        uint32_t lo[16], hi[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) { lo[e] = av[e]; hi[e] = av[16 + e]; }
        tmem_st_x16(a_tmem + ((q4 * 32) << 16) + sa * 32, lo);
        tmem_st_x16(a_tmem + ((q4 * 32) << 16) + sa * 32 + 16, hi);
      }
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&afull[sa]);
      if (has_next) {
#pragma unroll
        for (int e = 0; e < W; ++e) cur[e] = nxt[e];
      }
    }
    // ------------------------------------------------------------ epilogue
    // warps 0-3 drain the accumulators; all 16 dequant warps then meet for the split reduce
    const bool direct = (a.splits == 1 && a.epilogue == B200_EPI_NONE && a.fe.ssq_in == nullptr);
    pdl_wait();        // first global writes below (c / partials / counters)
    if (warp < 4) {
      mbar_wait(dfull, 0);
      tc_fence_after();
      const float bv = (direct && a.bias) ? __half2float(a.bias[n]) : 0.f;
      float* part = a.partials + (static_cast<size_t>(split) * M) * N;
#pragma unroll
      for (int j0 = 0; j0 < NT; j0 += 16) {
        uint32_t v[16];
        tmem_ld_x16(d_tmem + ((q4 * 32) << 16) + j0, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int m = m0 + j0 + j;
          if (m < M) {
            if (direct)
              a.c[static_cast<size_t>(m) * N + n] = __float2half_rn(__uint_as_float(v[j]) + bv);
            else
              part[static_cast<size_t>(m) * N + n] = __uint_as_float(v[j]);
          }
        }
      }
      tc_fence_before();
    }
    if (!direct) {
      SplitOut o;
      o.partials = a.partials; o.counters = a.counters; o.bias = a.bias; o.c = a.c;
      o.M = M; o.N = N; o.splits = a.splits; o.epilogue = a.epilogue; o.fe = a.fe;
      split_arrive_and_reduce(o, nt, threadIdx.x, kDqThreads, m0, min(M, m0 + NT),
                              blockIdx.z * (N / kTileN), &s_is_last, 1);
    }
  } else if (warp == kDqWarps) {
    // ------------------------------------------------------------ weight producer
    if (lane == 0) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(a.bp) +
                           (static_cast<size_t>(nt) * KC + c0) * Cfg::kWBytes;
      for (int i = 0; i < nchunks; ++i) {
        const int sw = i % Cfg::kSW;
        mbar_wait(&wempty[sw], ((i / Cfg::kSW) & 1) ^ 1);
        if ((a.dbg & 2) && i >= Cfg::kSW) { mbar_arrive(&wfull[sw]); continue; }   // experiment: no weight stream
        mbar_arrive_expect_tx(&wfull[sw], Cfg::kWBytes);
        bulk_load_1d(smem + Cfg::kWOff + sw * Cfg::kWBytes, src + static_cast<size_t>(i) * Cfg::kWBytes,
                     Cfg::kWBytes, &wfull[sw]);
      }
    }
  } else if (warp == kDqWarps + 1) {
    // ------------------------------------------------------------ activation producer
    if (lane == 0) {
      pdl_wait();      // activations are the previous kernel's output; weights are not
      for (int i = 0; i < nchunks; ++i) {
        const int sx = i % Cfg::kSX;
        mbar_wait(&xempty[sx], ((i / Cfg::kSX) & 1) ^ 1);
        if ((a.dbg & 1) && i >= Cfg::kSX) { mbar_arrive(&xfull[sx]); continue; }   // experiment: no activation stream
        mbar_arrive_expect_tx(&xfull[sx], Cfg::kXBytes);
        tma_load_2d(smem + Cfg::kXOff + sx * Cfg::kXBytes, &xmap, &xfull[sx], (c0 + i) * kChunkK, m0);
      }
    }
  } else {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(128, NT < 16 ? 16 : NT);
      for (int i = 0; i < nchunks; ++i) {
        const int sx = i % Cfg::kSX, sa = i % kSA;
        mbar_wait(&xfull[sx], (i / Cfg::kSX) & 1);
        mbar_wait(&afull[sa], (i / kSA) & 1);
        tc_fence_after();
        const uint64_t bdesc = make_smem_desc_sw128(smem_u32(smem + Cfg::kXOff + sx * Cfg::kXBytes));
        if (!(a.dbg & 8))
#pragma unroll
        for (int ks = 0; ks < kChunkK / 16; ++ks)
          tc_mma_ts(d_tmem, a_tmem + sa * 32 + ks * 8, bdesc + static_cast<uint64_t>(ks * 2), idesc,
                    (i > 0 || ks > 0) ? 1u : 0u);
        tc_commit(&xempty[sx]);
        tc_commit(&aempty[sa]);
      }
      tc_commit(dfull);
    }
  }
  __syncthreads();
  if (warp == kDqWarps + 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

static bool make_x_map(CUtensorMap* map, const __half* a, int M, int K, int NT) {
  const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(M)};
  const uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
  const uint32_t box[2] = {kChunkK, static_cast<uint32_t>(NT)};
  const uint32_t es[2] = {1, 1};
  return encode_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(a), dims,
                           strides, box, es, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B) == CUDA_SUCCESS;
}

template <int BITS, int NT>
static bool launch_tc_nt(const WnaParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BITS, NT>;
  CUtensorMap xmap;
  if (!make_x_map(&xmap, p.a, p.M, p.K, NT)) {
    set_error(2, "wna16_tc: cuTensorMapEncodeTiled failed");
    return true;
  }
  const int KC = p.K / kChunkK;
  const int tiles = p.N / kTileN;
  const int m_tiles = (p.M + NT - 1) / NT;
  const int sms = sm_count();
  int splits = sms / (tiles * m_tiles);
  if (splits < 1) splits = 1;
  if (splits > KC / 4) splits = KC / 4 > 0 ? KC / 4 : 1;
  while (splits > 1 && static_cast<long>(splits) * p.N > kMaxSplitCols) --splits;
  if (p.no_split) splits = 1;
  int cps = (KC + splits - 1) / splits;
  splits = (KC + cps - 1) / cps;
  if (p.epilogue == B200_EPI_SILU_MUL && (tiles & 1)) {
    set_error(1, "wna16_tc: SILU_MUL needs an even number of 128-channel tiles");
    return true;
  }
  if (m_tiles * tiles > kCounterBytes / 4) {
    if (!(splits == 1 && p.epilogue == B200_EPI_NONE && !(p.fe && p.fe->ssq_in))) {
      set_error(1, "wna16_tc: too many tiles for the arrival counters");
      return true;
    }
  }
  TcArgs a;
  a.bp = p.bp; a.scales = p.scales; a.qzeros = p.qzeros; a.bias = p.bias; a.c = p.c;
  a.counters = reinterpret_cast<uint32_t*>(p.workspace);
  a.partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p.workspace) + kCounterBytes);
  a.M = p.M; a.K = p.K; a.N = p.N; a.group_size = p.group_size; a.qfmt = p.qfmt;
  a.chunks_per_split = cps; a.splits = splits; a.epilogue = p.epilogue;
  if (p.fe) a.fe = *p.fe; else memset(&a.fe, 0, sizeof(a.fe));
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("B200_TC_DBG"); dbg = e ? atoi(e) : 0; }
    a.dbg = dbg;
  }
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(wna16_tc_kernel<BITS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         Cfg::kAlloc);
    attr_set = true;
  }
  dim3 grid(tiles, splits, m_tiles);
  launch_pdl(wna16_tc_kernel<BITS, NT>, grid, dim3(kTcThreads), Cfg::kAlloc, st, xmap, a);
  B200_CUDA_LAUNCH_CHECK();
  return true;
}

template <int BITS>
static bool launch_tc_bits(const WnaParams& p, cudaStream_t st) {
  if (p.M <= 16) return launch_tc_nt<BITS, 16>(p, st);
  if (p.M <= 32) return launch_tc_nt<BITS, 32>(p, st);
  if (p.M <= 64) return launch_tc_nt<BITS, 64>(p, st);
  if (p.M <= 128) return launch_tc_nt<BITS, 128>(p, st);
  return launch_tc_nt<BITS, 256>(p, st);
}

bool launch_wna16_tc(const WnaParams& p, cudaStream_t st) {
  if (p.bits == 4) return launch_tc_bits<4>(p, st);
  return launch_tc_bits<8>(p, st);
}

}  // namespace b200
