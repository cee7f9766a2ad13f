// Dense fp16 GEMM on tcgen05 (SS mode): C[m, n] = A[m, k] x W[n, k]^T, fp32 accumulation in TMEM,
// fp16 or fp32 output, optional bias.
//
// Replaces the cuBLAS call behind Linear::forward (/root/reference/src/models/layers/linear.rs:75-123)
// for fp16 checkpoints (BASELINE.json configs[2]: Llama-3-8B fp16 chunked prefill), the lm_head
// (/root/reference/src/models/llama.rs:311-320) and -- together with wna16_dequant_kernel below --
// the W4A16 / W8A16 linears of large prefill steps, where dequantising a layer's weights ONCE into an
// fp16 scratch and running this kernel beats re-dequantising them for every 256-token tile
// (the Marlin call of /root/reference/src/utils/gptq.rs:118-146 at M = thousands of tokens).
//
// Both operands are K-major, so both arrive through TMA (SWIZZLE_128B, 64-element K blocks) and
// feed tcgen05.mma.kind::f16 through shared-memory descriptors.  "P" is the operand whose rows
// become TMEM lanes (128 per CTA), "Q" the one whose rows become accumulator columns (BN per MMA):
//   prefill (M > 128):  P = activations, Q = weights, BN = 256; with CG = 2 a CTA PAIR computes a
//                       256 x 256 tile with one cta_group::2 MMA (each CTA stages its 128 P rows and
//                       HALF of the Q rows: 128 flop per shared-memory byte instead of 85, which is
//                       what the ~12 TB/s L2 -> SM path can feed at cuBLAS speed).
//   decode  (M <= 128): swapped: P = weights (vocab rows), Q = activations (BN = 32 / 64 / 128), so
//                       a batch of 32 runs 128 x 32 MMAs with no wasted rows and the kernel is a
//                       pure HBM stream over W (lm_head: 1.05 GB / step).
// Persistent CTAs (one per SM, or one pair per 2 SMs), warp roles:
//   warp 0   TMA producer (STAGES-deep ring, full / empty mbarriers)
//   warp 1   MMA issuer (leader CTA of a pair only); tcgen05.commit releases ring slots and
//            publishes the accumulator; owns the TMEM allocation
//   warps 2-5 epilogue: tcgen05.ld of the finished accumulator (double-buffered in TMEM, so the
//            MMAs of tile i+1 overlap the drain of tile i), bias, convert, vectorised stores
// Roofline: prefill = tensor (2*M*N*K flop); decode = HBM (N*K*2 bytes).
#include <cstdlib>
#include <cstring>

#include "wna16.cuh"

namespace b200 {
namespace dg {

constexpr int kThreads = 192;
constexpr int kBM = 128;       // P rows per CTA = TMEM lanes
constexpr int kBK = 64;        // K elements per stage (one 128-byte swizzle row)
// P tiles per rasterisation group: a group's P panel (group_p * 128 * CG rows x K) stays in L2 while
// the Q panels stream past it once per group.  B200_DENSE_GROUP_P overrides (A/B testing).
static int group_p_env() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_DENSE_GROUP_P");
    v = e ? atoi(e) : 0;
  }
  return v;
}

template <int BN, int CG>
struct Cfg {
  static constexpr int kPBytes = kBM * kBK * 2;                 // 16 KB
  static constexpr int kQRows = BN / CG;                        // Q rows staged by one CTA
  static constexpr int kQBytes = kQRows * kBK * 2;
  static constexpr int kStageBytes = kPBytes + kQBytes;
  static constexpr int kStages = (196 * 1024) / kStageBytes;    // 6 (pair, BN 256) / 4 / ... / 9 (BN 32)
  static constexpr int kBarOff = kStages * kStageBytes;
  static constexpr int kNumBars = 2 * kStages + 4;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kAlloc = kTmemPtrOff + 16 + 1024;
  static_assert(kStageBytes % 1024 == 0, "stage bases must stay 1024-byte aligned (SWIZZLE_128B)");
  static_assert(kAlloc <= 227 * 1024, "shared memory budget");
  static_assert(2 * BN <= 512, "two accumulators must fit TMEM");
};

struct Args {
  void* c;
  const __half* bias;
  int M, N;            // logical output [M, N], row stride N
  int p_rows, q_rows;  // rows of the P / Q operand (swap: N / M, else M / N)
  int KB;              // K / 64
  int tiles_p, tiles_q, group_p;
};

__device__ __forceinline__ void tile_coords(int t, const Args& a, int& p, int& q) {
  const int per_group = a.group_p * a.tiles_q;
  const int g = t / per_group, r = t - g * per_group;
  const int rows = min(a.group_p, a.tiles_p - g * a.group_p);
  q = r / rows;
  p = g * a.group_p + (r - q * rows);
}

// SILU (prefill gate|up projection, BN = 256, fp16 out): the weight rows arrive interleaved in blocks of
// 128 (gate channels [128t, 128t+128) then up channels of the same range, see wna16_dequant_kernel), so
// accumulator columns c and c + 128 of a tile are gate / up of ONE output channel and the epilogue
// writes silu(gate) * up straight into [M, N/2] (mlp.rs:451-469) -- no [M, N] round trip through HBM.
template <int BN, int CG, bool SWAP, bool OUT_F32, bool SILU = false>
__global__ void __launch_bounds__(kThreads, 1)
dense_gemm_tc_kernel(const __grid_constant__ CUtensorMap pmap, const __grid_constant__ CUtensorMap qmap,
                     const Args a) {
  using C = Cfg<BN, CG>;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0;
  const int unit = blockIdx.x / CG, nunits = gridDim.x / CG;      // a unit = a CTA or a CTA pair
  const int ntiles = a.tiles_p * a.tiles_q;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + C::kBarOff);
  uint64_t* empty = full + C::kStages;
  uint64_t* tfull = empty + C::kStages;        // [2] accumulator published
  uint64_t* tempty = tfull + 2;                // [2] accumulator drained (leader's copy is the one waited on)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + C::kTmemPtrOff);

  if (threadIdx.x == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4 * CG);           // every epilogue warp of the unit
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CG == 2) {
      tmem_alloc_cg2(tmem_ptr, 512);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_ptr, 512);
      tmem_relinquish();
    }
  }
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&pmap);
    prefetch_tmap(&qmap);
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) {          // the peer's barriers must be initialised before anything signals them
    cluster_arrive();
    cluster_wait();
  }
  tc_fence_after();
  pdl_trigger();
  pdl_wait();             // activations (and, after a dequant pass, the weights) come from earlier kernels
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (every CTA)
    if (lane == 0) {
      uint32_t kit = 0;
      for (int t = unit; t < ntiles; t += nunits) {
        int p, q;
        tile_coords(t, a, p, q);
        const int prow = (p * CG + static_cast<int>(rank)) * kBM;
        const int qrow = q * BN + static_cast<int>(rank) * C::kQRows;
        for (int kb = 0; kb < a.KB; ++kb, ++kit) {
          const int s = kit % C::kStages;
          mbar_wait_bounded(&empty[s], ((kit / C::kStages) & 1) ^ 1);
          uint8_t* pdst = smem + s * C::kStageBytes;
          uint8_t* qdst = pdst + C::kPBytes;
          if (CG == 1) {
            mbar_arrive_expect_tx(&full[s], C::kStageBytes);
            tma_load_2d(pdst, &pmap, &full[s], kb * kBK, prow);
            tma_load_2d(qdst, &qmap, &full[s], kb * kBK, qrow);
          } else {
            // both CTAs' bytes complete on the LEADER's barrier, which expects the pair's total
            if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * C::kStageBytes);
            const uint32_t lbar = mapa_shared(smem_u32(&full[s]), 0);
            tma_load_2d_cg2(pdst, &pmap, lbar, kb * kBK, prow);
            tma_load_2d_cg2(qdst, &qmap, lbar, kb * kBK, qrow);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (leader CTA)
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = make_idesc_f16(kBM * CG, BN);
      uint32_t kit = 0;
      int it = 0;
      for (int t = unit; t < ntiles; t += nunits, ++it) {
        const int buf = it & 1;
        mbar_wait_bounded(&tempty[buf], ((it >> 1) & 1) ^ 1);     // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * BN;
        for (int kb = 0; kb < a.KB; ++kb, ++kit) {
          const int s = kit % C::kStages;
          mbar_wait_bounded(&full[s], (kit / C::kStages) & 1);
          tc_fence_after();
          const uint32_t pbase = smem_u32(smem + s * C::kStageBytes);
          const uint64_t adesc = make_smem_desc_sw128(pbase);
          const uint64_t bdesc = make_smem_desc_sw128(pbase + C::kPBytes);
#pragma unroll
          for (int ks = 0; ks < kBK / 16; ++ks) {
            const uint32_t acc = (kb > 0 || ks > 0) ? 1u : 0u;
            if (CG == 2)
              tc_mma_ss_cg2(d_tmem, adesc + static_cast<uint64_t>(ks * 2), bdesc + static_cast<uint64_t>(ks * 2),
                            idesc, acc);
            else
              tc_mma_ss(d_tmem, adesc + static_cast<uint64_t>(ks * 2), bdesc + static_cast<uint64_t>(ks * 2),
                        idesc, acc);
          }
          if (CG == 2) tc_commit_cg2(&empty[s], 3); else tc_commit(&empty[s]);
        }
        if (CG == 2) tc_commit_cg2(&tfull[buf], 3); else tc_commit(&tfull[buf]);
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps (every CTA)
    const int q4 = warp & 3;                          // TMEM lane quarter this warp may read
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    int it = 0;
    for (int t = unit; t < ntiles; t += nunits, ++it) {
      int p, q;
      tile_coords(t, a, p, q);
      const int buf = it & 1;
      mbar_wait_bounded(&tfull[buf], (it >> 1) & 1);
      tc_fence_after();
      const int prow = (p * CG + static_cast<int>(rank)) * kBM + q4 * 32 + lane;    // this thread's P row
      if (SILU) {
        const int half_n = a.N >> 1;
#pragma unroll 1
        for (int c0 = 0; c0 < BN / 2; c0 += 32) {
          uint32_t g[32], u[32];
          tmem_ld_x32(tmem_base + lane_base + buf * BN + c0, g);
          tmem_ld_x32(tmem_base + lane_base + buf * BN + BN / 2 + c0, u);
          tc_wait_ld();
          const int col = q * (BN / 2) + c0;          // output channel of g[0]
          if (prow < a.M && col < half_n) {
            __half* dst = reinterpret_cast<__half*>(a.c) + static_cast<size_t>(prow) * half_n + col;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                // the two GEMM outputs are rounded to fp16 first: the values the unfused pipeline sees
                const float2 gg = __half22float2(u32_as_h2(pack_half2(__uint_as_float(g[j + 2 * e]), __uint_as_float(g[j + 2 * e + 1]))));
                const float2 uu = __half22float2(u32_as_h2(pack_half2(__uint_as_float(u[j + 2 * e]), __uint_as_float(u[j + 2 * e + 1]))));
                w[e] = pack_half2(gg.x / (1.f + __expf(-gg.x)) * uu.x, gg.y / (1.f + __expf(-gg.y)) * uu.y);
              }
              if (col + j + 8 <= half_n) *reinterpret_cast<uint4*>(dst + j) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      } else
#pragma unroll 1
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t v[32];
        tmem_ld_x32(tmem_base + lane_base + buf * BN + c0, v);
        tc_wait_ld();
        const int qcol = q * BN + c0;                 // first Q row (= accumulator column) of this block
        if (!SWAP) {
          // thread = token row m, 32 consecutive output channels
          const int m = prow;
          if (m < a.M && qcol < a.N) {
            if (a.bias) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (qcol + j < a.N) v[j] = __float_as_uint(__uint_as_float(v[j]) + __half2float(a.bias[qcol + j]));
            }
            if (OUT_F32) {
              float* dst = reinterpret_cast<float*>(a.c) + static_cast<size_t>(m) * a.N + qcol;
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                if (qcol + j + 4 <= a.N)
                  *reinterpret_cast<uint4*>(dst + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
              __half* dst = reinterpret_cast<__half*>(a.c) + static_cast<size_t>(m) * a.N + qcol;
#pragma unroll
              for (int j = 0; j < 32; j += 8)
                if (qcol + j + 8 <= a.N) {
                  uint4 o;
                  o.x = pack_half2(__uint_as_float(v[j]), __uint_as_float(v[j + 1]));
                  o.y = pack_half2(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                  o.z = pack_half2(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
                  o.w = pack_half2(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
                  *reinterpret_cast<uint4*>(dst + j) = o;
                }
            }
          }
        } else {
          // thread = output channel n (a weight row), columns = tokens: a warp writes 32
          // consecutive channels of one token row per store
          const int n = prow;
          if (n < a.N) {
            const float bv = a.bias ? __half2float(a.bias[n]) : 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int m = qcol + j;
              if (m < a.M) {
                const float x = __uint_as_float(v[j]) + bv;
                if (OUT_F32)
                  reinterpret_cast<float*>(a.c)[static_cast<size_t>(m) * a.N + n] = x;
                else
                  reinterpret_cast<__half*>(a.c)[static_cast<size_t>(m) * a.N + n] = __float2half_rn(x);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2 && rank != 0)
          mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[buf]), 0));
        else
          mbar_arrive(&tempty[buf]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) {          // neither CTA may leave (or free TMEM) while its peer can still touch it
    cluster_arrive();
    cluster_wait();
  }
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

static bool make_map(CUtensorMap* map, const void* base, int rows, int K, int box_rows) {
  const uint64_t dims[2] = {static_cast<uint64_t>(K), static_cast<uint64_t>(rows)};
  const uint64_t strides[1] = {static_cast<uint64_t>(K) * 2};
  const uint32_t box[2] = {kBK, static_cast<uint32_t>(box_rows)};
  const uint32_t es[2] = {1, 1};
  return encode_tensor_map(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides,
                           box, es, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B) == CUDA_SUCCESS;
}

template <int BN, int CG, bool SWAP, bool OUT_F32, bool SILU = false>
static bool launch_cfg(const void* x, const void* w, void* c, const __half* bias, int M, int N, int K,
                       cudaStream_t st) {
  using C = Cfg<BN, CG>;
  const void* pbase = SWAP ? w : x;
  const void* qbase = SWAP ? x : w;
  Args a;
  a.c = c; a.bias = bias; a.M = M; a.N = N;
  a.p_rows = SWAP ? N : M;
  a.q_rows = SWAP ? M : N;
  a.KB = K / kBK;
  a.tiles_p = (a.p_rows + kBM * CG - 1) / (kBM * CG);
  a.tiles_q = (a.q_rows + BN - 1) / BN;
  // 16 P tiles per group measured best on all four Llama-3-8B prefill shapes (M = 8192: 1390-1450 TF/s
  // with CTA pairs; 8 re-reads the weight panels from HBM 6x, 32 thrashes L2 at K = 14336)
  {
    int gp = 16;
    if (group_p_env() > 0) gp = group_p_env();
    a.group_p = gp;
  }
  CUtensorMap pmap, qmap;
  if (!make_map(&pmap, pbase, a.p_rows, K, kBM) || !make_map(&qmap, qbase, a.q_rows, K, C::kQRows)) {
    set_error(2, "dense_gemm_tc: cuTensorMapEncodeTiled failed");
    return true;
  }
  auto kern = dense_gemm_tc_kernel<BN, CG, SWAP, OUT_F32, SILU>;
  static bool attr_set = false;
  static int max_units = 0;
  if (!attr_set) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kAlloc);
    if (CG == 2) {
      max_units = max_active_clusters(kern, dim3(sm_count()), dim3(kThreads), C::kAlloc, dim3(2, 1, 1));
      if (max_units <= 0) max_units = sm_count() / 2;
    } else {
      max_units = sm_count();
    }
    attr_set = true;
  }
  const int ntiles = a.tiles_p * a.tiles_q;
  const int units = ntiles < max_units ? ntiles : max_units;
  const cudaError_t e = launch_pdl_cluster(kern, dim3(units * CG), dim3(kThreads), C::kAlloc, st,
                                           dim3(CG, 1, 1), pmap, qmap, a);
  if (e != cudaSuccess) set_error(2, cudaGetErrorString(e));
  B200_CUDA_LAUNCH_CHECK();
  return true;
}

template <bool OUT_F32>
static bool launch_out(const void* x, const void* w, void* c, const __half* bias, int M, int N, int K,
                       int mode, cudaStream_t st) {
  if (M <= 128) {             // decode shapes: weights stream once, tokens are accumulator columns
    if (M <= 32) return launch_cfg<32, 1, true, OUT_F32>(x, w, c, bias, M, N, K, st);
    if (M <= 64) return launch_cfg<64, 1, true, OUT_F32>(x, w, c, bias, M, N, K, st);
    return launch_cfg<128, 1, true, OUT_F32>(x, w, c, bias, M, N, K, st);
  }
  if (mode == 1) return launch_cfg<256, 1, false, OUT_F32>(x, w, c, bias, M, N, K, st);
  return launch_cfg<256, 2, false, OUT_F32>(x, w, c, bias, M, N, K, st);
}

}  // namespace dg

// mode: 1 = single-CTA tiles, 2 = CTA pairs (cta_group::2).  false: shape not supported here.
bool launch_dense_gemm_tc(const void* x, const void* w, void* c, const void* bias, int M, int N, int K,
                          int out_f32, int mode, cudaStream_t st) {
  if (K % dg::kBK != 0 || N % 8 != 0) return false;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(c)) & 15)
    return false;
  const __half* b = static_cast<const __half*>(bias);
  return out_f32 ? dg::launch_out<true>(x, w, c, b, M, N, K, mode, st)
                 : dg::launch_out<false>(x, w, c, b, M, N, K, mode, st);
}

// gate|up projection with the SiLU * mul epilogue: w rows interleaved in 128-blocks (gate, up, gate, ...),
// out [M, N/2] fp16.  false: shape not supported (the caller runs GEMM + silu_mul instead).
bool launch_dense_gemm_tc_silu(const void* x, const void* w, void* out, int M, int N, int K, int mode,
                               cudaStream_t st) {
  if (K % dg::kBK != 0 || N % 256 != 0 || M <= 128) return false;
  if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(out)) & 15)
    return false;
  if (mode == 1) return dg::launch_cfg<256, 1, false, false, true>(x, w, out, nullptr, M, N, K, st);
  return dg::launch_cfg<256, 2, false, false, true>(x, w, out, nullptr, M, N, K, st);
}

// ---------------------------------------------------------------- W4 / W8 -> fp16 [N, K]
// One pass over a prepacked (tile-major) quantised weight: thread = output channel, it reads the
// 32 / 64 packed bytes of a 64-K chunk and writes the chunk's 128 bytes of fp16((q - z) * s) into
// row n of the K-major scratch -- the Q operand of the kernel above.  Same rounding as the
// register dequant of wna16_tc.cu (one rounding, on the scale multiply).
template <int BITS>
__global__ void __launch_bounds__(kTileN)
wna16_dequant_kernel(const uint32_t* __restrict__ bp, const __half* __restrict__ scales,
                     const uint32_t* __restrict__ qzeros, __half* __restrict__ out, int K, int N,
                     int group_size, int qfmt, int chunks_per_cta, int interleave) {
  const int nt = blockIdx.x, ch = threadIdx.x, n = nt * kTileN + ch;
  // interleave: the fused gate|up weight ([gate tiles | up tiles] along N) is written as
  // gate tile 0, up tile 0, gate tile 1, ... so a 256-row GEMM tile holds both halves of 128 channels
  const int half_tiles = gridDim.x >> 1;
  const int dst_tile = interleave ? (nt < half_tiles ? 2 * nt : 2 * (nt - half_tiles) + 1) : nt;
  const int n_dst = dst_tile * kTileN + ch;
  const int KC = K / kChunkK;
  const int c0 = blockIdx.y * chunks_per_cta, c1 = min(c0 + chunks_per_cta, KC);
  constexpr int W = 2 * BITS;                                  // packed words per thread per chunk
  constexpr int kWBytes = kTileN * kChunkK * BITS / 8;
  pdl_trigger();
  int cur_g = -1;
  DqConst dq = make_dq(__float2half(0.f), 0);
  for (int kc = c0; kc < c1; ++kc) {
    const uint8_t* src = reinterpret_cast<const uint8_t*>(bp) + (static_cast<size_t>(nt) * KC + kc) * kWBytes;
    uint32_t words[W];
#pragma unroll
    for (int sub = 0; sub < W / 4; ++sub) {
      const uint4 u = ld_nc_v4(src + sub * kTileN * 16 + ch * 16);
      words[sub * 4] = u.x; words[sub * 4 + 1] = u.y; words[sub * 4 + 2] = u.z; words[sub * 4 + 3] = u.w;
    }
    uint32_t o[32];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int g = (kc * kChunkK + hf * 32) / group_size;
      if (g != cur_g) {
        cur_g = g;
        dq = make_dq(scales[static_cast<size_t>(g) * N + n], load_zero<BITS>(qzeros, qfmt, g, n, N));
      }
      if (BITS == 4) {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          uint32_t o4[4];
          dequant4_word(words[hf * 4 + w], dq, o4);
#pragma unroll
          for (int e = 0; e < 4; ++e) o[hf * 16 + w * 4 + e] = o4[e];
        }
      } else {
#pragma unroll
        for (int w = 0; w < 8; ++w) {
          uint32_t o2[2];
          dequant8_word(words[(hf * 8 + w) % W], dq, o2);
          o[hf * 16 + w * 2] = o2[0];
          o[hf * 16 + w * 2 + 1] = o2[1];
        }
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(n_dst) * K + kc * kChunkK);
#pragma unroll
    for (int e = 0; e < 8; ++e) dst[e] = make_uint4(o[e * 4], o[e * 4 + 1], o[e * 4 + 2], o[e * 4 + 3]);
  }
}

void launch_wna16_dequant(const uint32_t* bp, const __half* scales, const uint32_t* qzeros, __half* out,
                          int K, int N, int bits, int group_size, int qfmt, cudaStream_t st, int interleave) {
  const int KC = K / kChunkK;
  const int cpc = 8;
  dim3 grid(N / kTileN, (KC + cpc - 1) / cpc);
  if (bits == 4)
    wna16_dequant_kernel<4><<<grid, kTileN, 0, st>>>(bp, scales, qzeros, out, K, N, group_size, qfmt, cpc, interleave);
  else
    wna16_dequant_kernel<8><<<grid, kTileN, 0, st>>>(bp, scales, qzeros, out, K, N, group_size, qfmt, cpc, interleave);
  B200_CUDA_LAUNCH_CHECK();
}

}  // namespace b200

// Dequantise a prepacked W4 / W8 weight to fp16 [n, k] (row n = output channel, K-major).
extern "C" void b200_wna16_dequant_f16(const void* b_packed, const void* scales, const void* qzeros,
                                       void* out, int32_t k, int32_t n, int32_t bits, int32_t group_size,
                                       int32_t qfmt, int64_t stream) {
  using namespace b200;
  B200_CHECK_ARG(b_packed && scales && out, "wna16_dequant: null pointer");
  B200_CHECK_ARG((bits == 4 || bits == 8) && k % kChunkK == 0 && n % kTileN == 0, "wna16_dequant: bad shape");
  const int g = group_size <= 0 ? k : group_size;
  B200_CHECK_ARG(g % 32 == 0 && k % g == 0, "wna16_dequant: bad group size");
  launch_wna16_dequant(static_cast<const uint32_t*>(b_packed), static_cast<const __half*>(scales),
                       static_cast<const uint32_t*>(qzeros), static_cast<__half*>(out), k, n, bits, g, qfmt,
                       reinterpret_cast<cudaStream_t>(stream), 0);
}
