// Varlen causal prefill attention over the paged cache on tcgen05 tensor cores (head_dim 128).
//
// Same contract as attn_prefill.cu (the prefill branch of PagedAttention::forward,
// /root/reference/src/models/layers/attention.rs:808-820, metadata of prepare_prefill,
// /root/reference/src/core/runner.rs:978-1241) -- that file's mma.sync kernel stays as the head_dim 64
// path and as the A/B baseline (B200_PREFILL_ATTN=mma).
//
// A CTA owns 128 rows = 128/G query tokens x the G query heads of ONE kv head (row = token * G + head,
// staged by one 3-D TMA box per 64-wide half of D, SWIZZLE_128B) and walks the sequence's visible K/V
// pages (64 tokens, gathered by block-table index with the 4-D TMA boxes of the decode kernel):
//   S  = Q K^T     tcgen05.mma SS: A = Q (K-major), B = K page (K-major), fp32 [128 x 64] in TMEM
//   P  = softmax   thread = row: tcgen05.ld S, scale, causal mask (diagonal pages only), running max
//                  with LAZY rescaling (the accumulator is rescaled only when the max grows by > 2^8,
//                  exact because the final normalisation uses the same reference max), exp2, fp16 P
//                  written back to TMEM with tcgen05.st
//   O += P V       tcgen05.mma TS: A = P from TMEM, B = V page as an MN-major operand (D contiguous),
//                  fp32 [128 x 128] in TMEM for the whole walk
// Warp roles: warp 0 TMA producer (2-stage K|V ring), warp 1 MMA issuer (QK^T of page j+1 is issued
// before PV of page j, so it runs under the softmax of page j), warps 2-5 softmax / correction /
// epilogue.  96 KB of shared memory and 256 TMEM columns per CTA: two CTAs share an SM and fill each
// other's softmax bubbles.  Tensor-bound (4 * 128 * ctx flop per row) with MUFU.EX2 as co-limit.
#include <cstdlib>
#include <cstring>

#include "../../include/b200_vllm.h"
#include "attn.cuh"

namespace b200 {
namespace pft {

constexpr int kD = 128;
constexpr int kRows = 128;
constexpr int kTok = 64;
constexpr int kStages = 2;
constexpr int kThreads = 192;
constexpr int kQBytes = kRows * kD * 2;          // 32 KB: [2 D-halves][128 rows][128 B]
constexpr int kTileBytes = kTok * kD * 2;        // 16 KB: [2 D-halves][64 tokens][128 B]
constexpr int kStageBytes = 2 * kTileBytes;      // K | V
constexpr int kBarOff = kQBytes + kStages * kStageBytes;
constexpr int kTmemPtrOff = kBarOff + 16 * 8;
constexpr int kAlloc = kTmemPtrOff + 16 + 1024;
constexpr int kTmemCols = 256;
constexpr int kSCol = 0, kPCol = 64, kOCol = 128;
constexpr float kRescaleThreshold = 8.f;         // log2 units

// MN-major SWIZZLE_128B operand: rows of 128 B hold 64 consecutive MN elements of ONE K index;
// 8 K rows form a 1024-byte swizzle atom.  SBO = distance between 8-row K groups, LBO = distance
// between 64-element MN chunks (cute mma_traits_sm100: ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units).
__device__ __forceinline__ uint64_t make_smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  d |= uint64_t(2) << 61;
  return d;
}

__device__ __forceinline__ void tmem_st_x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
      "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]),
      "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]),
      "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

struct Args {
  __half* out;
  const uint32_t* cu_seqlens_q;
  const uint32_t* context_lens;
  const uint32_t* block_tables;
  int H, Hkv, max_blocks;
  float scale_log2;
  int dbg;           // bit 0: swap LBO / SBO of the V descriptor (bring-up switch)
};

// K-major operand without a swizzle ("interleave" layout, cute mma_traits_sm100: ((8,m),(T,2)):((1T,SBO),(1,LBO))):
// a core matrix is 8 rows x 16 bytes stored as 128 contiguous bytes; LBO = distance between the two core matrices of
// one k-step, SBO = distance between 8-row groups.  This is how a (block, head) K tile of the reference's
// default-build cache lies in memory ([D/8][64 tokens][8]: LBO = 1024, SBO = 128), so the tile is consumed as copied.
__device__ __forceinline__ uint64_t make_smem_desc_k_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFF) >> 4);
  d |= uint64_t((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= uint64_t((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= uint64_t(1) << 46;
  return d;
}

// L1 = the reference's default-build cache layout (K [nb, Hkv, D/8, 64, 8], V [nb, Hkv, D, 64]): K as above, V is a
// K-major operand of the P V MMA (rows = d, 64 tokens = 128 swizzled bytes) instead of the flash layout's MN-major.
template <bool BF, bool L1>
__global__ void __launch_bounds__(kThreads, 2)
attn_prefill_tc_kernel(const __grid_constant__ CUtensorMap qmap, const __grid_constant__ CUtensorMap kmap,
                       const __grid_constant__ CUtensorMap vmap, const Args a) {
  const int qt = gridDim.x - 1 - blockIdx.x;       // longest (latest) query tiles first
  const int hkv = blockIdx.y, b = blockIdx.z;
  const int G = a.H / a.Hkv;
  const int TQ = kRows / G;
  const int row0 = static_cast<int>(a.cu_seqlens_q[b]);
  const int nq = static_cast<int>(a.cu_seqlens_q[b + 1]) - row0;
  const int q0 = qt * TQ;
  if (q0 >= nq) return;
  const int ctx = a.context_lens ? static_cast<int>(a.context_lens[b]) : nq;
  const int cached = ctx - nq;
  const int q_end = min(q0 + TQ, nq);
  const int npages = (cached + q_end - 1) / kTok + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* q_full = bars;
  uint64_t* kv_full = bars + 1;        // [2]
  uint64_t* kv_empty = bars + 3;       // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* s_free = bars + 6;
  uint64_t* p_full = bars + 7;
  uint64_t* pv_done = bars + 8;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(smem + kTmemPtrOff);

  if (threadIdx.x == 0) {
    mbar_init(q_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_barrier_init();
    prefetch_tmap(&qmap);
    prefetch_tmap(&kmap);
    prefetch_tmap(&vmap);
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  pdl_wait();                    // q and this step's K/V rows come from the previous kernels
  const uint32_t tmem_base = *tmem_ptr;
  uint8_t* q_smem = smem;
  uint8_t* kv_smem = smem + kQBytes;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kQBytes);
      tma_load_3d(q_smem, &qmap, q_full, 0, hkv * G, row0 + q0);
      tma_load_3d(q_smem + kRows * 128, &qmap, q_full, 64, hkv * G, row0 + q0);
    }
    const uint32_t* bt = a.block_tables + static_cast<size_t>(b) * a.max_blocks;
    for (int base = 0; base < npages; base += 32) {
      const int mine = base + lane < npages ? static_cast<int>(bt[base + lane]) : 0;
      const int cnt = min(32, npages - base);
      for (int jj = 0; jj < cnt; ++jj) {
        const int blk = __shfl_sync(0xffffffffu, mine, jj);
        if (lane == 0) {
          const int j = base + jj, s = j & 1;
          mbar_wait_bounded(&kv_empty[s], ((j >> 1) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], kStageBytes);
          uint8_t* kdst = kv_smem + s * kStageBytes;
          uint8_t* vdst = kdst + kTileBytes;
          if (L1) {
            tma_load_4d(kdst, &kmap, &kv_full[s], 0, 0, hkv, blk);
            tma_load_4d(vdst, &vmap, &kv_full[s], 0, 0, hkv, blk);
          } else {
            tma_load_4d(kdst, &kmap, &kv_full[s], 0, hkv, 0, blk);
            tma_load_4d(kdst + kTok * 128, &kmap, &kv_full[s], 64, hkv, 0, blk);
            tma_load_4d(vdst, &vmap, &kv_full[s], 0, hkv, 0, blk);
            tma_load_4d(vdst + kTok * 128, &vmap, &kv_full[s], 64, hkv, 0, blk);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t kFmt = BF ? ((1u << 7) | (1u << 10)) : 0u;          // a_format / b_format = BF16
      const uint32_t idesc_s = make_idesc_f16(kRows, kTok) | kFmt;
      const uint32_t idesc_o = make_idesc_f16(kRows, kD) | (L1 ? 0u : (1u << 16)) | kFmt;   // flash V: MN-major B
      const uint32_t s_tmem = tmem_base + kSCol, p_tmem = tmem_base + kPCol, o_tmem = tmem_base + kOCol;
      const uint32_t q_addr = smem_u32(q_smem);
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        mbar_wait_bounded(&kv_full[s], (j >> 1) & 1);
        if (j > 0) mbar_wait_bounded(s_free, (j - 1) & 1);      // softmax holds S of page j-1 in registers
        tc_fence_after();
        const uint32_t k_addr = smem_u32(kv_smem + s * kStageBytes);
#pragma unroll
        for (int ks = 0; ks < kD / 16; ++ks) {
          const uint32_t off = (ks >> 2) * 0 + (ks & 3) * 32;
          const uint64_t adesc = make_smem_desc_sw128(q_addr + (ks >> 2) * (kRows * 128) + off);
          const uint64_t bdesc = L1 ? ((a.dbg & 1) ? make_smem_desc_k_nosw(k_addr + ks * 2048, 128, 1024)
                                                   : make_smem_desc_k_nosw(k_addr + ks * 2048, 1024, 128))
                                    : make_smem_desc_sw128(k_addr + (ks >> 2) * (kTok * 128) + off);
          tc_mma_ss(s_tmem, adesc, bdesc, idesc_s, ks > 0 ? 1u : 0u);
        }
        tc_commit(s_full);
      };
      mbar_wait_bounded(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < npages; ++j) {
        if (j + 1 < npages) issue_qk(j + 1);
        const int s = j & 1;
        mbar_wait_bounded(p_full, j & 1);
        tc_fence_after();
        const uint32_t v_addr = smem_u32(kv_smem + s * kStageBytes + kTileBytes);
#pragma unroll
        for (int ks = 0; ks < kTok / 16; ++ks) {
          const uint64_t bdesc = L1 ? make_smem_desc_sw128(v_addr + ks * 32)
                                 : (a.dbg & 1) ? make_smem_desc_mn_sw128(v_addr + ks * 2048, 1024, kTok * 128)
                                               : make_smem_desc_mn_sw128(v_addr + ks * 2048, kTok * 128, 1024);
          tc_mma_ts(o_tmem, p_tmem + ks * 8, bdesc, idesc_o, (j > 0 || ks > 0) ? 1u : 0u);
        }
        tc_commit(&kv_empty[s]);
        tc_commit(pv_done);
      }
    }
  } else {
    // ------------------------------------------------------------ softmax / correction / epilogue
    const int q4 = warp & 3;
    const int r = q4 * 32 + lane;                         // accumulator row = TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(q4 * 32) << 16;
    const int t_local = r / G, g = r - t_local * G;
    const int tq = q0 + t_local;
    const bool valid = tq < nq;
    const int qpos = cached + tq;                         // last visible key of this row
    const int first_limit = cached + q0;                  // smallest limit in the CTA
    float m_ref = -INFINITY, l_run = 0.f;
    for (int j = 0; j < npages; ++j) {
      const int page_base = j * kTok;
      mbar_wait_bounded(s_full, j & 1);
      tc_fence_after();
      float x[kTok];
      {
        uint32_t v0[32], v1[32];
        tmem_ld_x32(tmem_base + lane_base + kSCol, v0);
        tmem_ld_x32(tmem_base + lane_base + kSCol + 32, v1);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          x[i] = __uint_as_float(v0[i]) * a.scale_log2;
          x[32 + i] = __uint_as_float(v1[i]) * a.scale_log2;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);                 // S may be overwritten by QK^T of page j+1
      if (page_base + kTok - 1 > first_limit) {           // diagonal page: causal mask
#pragma unroll
        for (int i = 0; i < kTok; ++i) x[i] = (page_base + i <= qpos) ? x[i] : -INFINITY;
      }
      float mx = x[0];
#pragma unroll
      for (int i = 1; i < kTok; ++i) mx = fmaxf(mx, x[i]);
      // lazy rescale: keep the reference max unless the true max outgrew it by 2^8
      float corr = 1.f;
      bool grow = false;
      if (j == 0) {
        m_ref = (mx == -INFINITY) ? 0.f : mx;
      } else if (mx - m_ref > kRescaleThreshold) {
        corr = exp2f(m_ref - mx);
        m_ref = mx;
        grow = true;
      }
      float psum = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float p0 = exp2f(x[2 * i] - m_ref), p1 = exp2f(x[2 * i + 1] - m_ref);
        psum += p0 + p1;
        pk[i] = pack16x2<BF>(p0, p1);
      }
      l_run = l_run * corr + psum;
      // stale V rows past the context (uninitialised cache memory) must not meet the MMA: 0 * NaN
      if (page_base + kTok > ctx) {
        mbar_wait_bounded(&kv_full[j & 1], (j >> 1) & 1);
        uint8_t* vt = kv_smem + (j & 1) * kStageBytes + kTileBytes;
        const int first_bad = ctx - page_base;             // rows [first_bad, 64) of both D halves
        if (L1) {
          // rows = d, tokens along the swizzled 128-byte row: 16-byte chunk c holds tokens 8c .. 8c+7
          for (int idx = (warp - 2) * 32 + lane; idx < kD * 8; idx += 128) {
            const int row = idx >> 3, c = idx & 7;
            if (8 * c + 7 < first_bad) continue;
            uint8_t* p = vt + row * 128 + ((c ^ (row & 7)) << 4);
            if (8 * c >= first_bad) {
              *reinterpret_cast<uint4*>(p) = make_uint4(0, 0, 0, 0);
            } else {
              for (int e = first_bad - 8 * c; e < 8; ++e) reinterpret_cast<uint16_t*>(p)[e] = 0;
            }
          }
        } else
        for (int idx = (warp - 2) * 32 + lane; idx < (kTok - first_bad) * 16; idx += 128) {
          const int row = first_bad + (idx >> 4), c16 = idx & 15;
          *reinterpret_cast<uint4*>(vt + (c16 >> 3) * (kTok * 128) + row * 128 + (c16 & 7) * 16) =
              make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async();                               // generic-proxy writes -> tensor-core reads
      }
      if (j > 0) {
        mbar_wait_bounded(pv_done, (j - 1) & 1);           // PV of page j-1 retired: P and O are ours
        tc_fence_after();
        if (__any_sync(0xffffffffu, grow)) {
#pragma unroll 1
          for (int c0 = 0; c0 < kD; c0 += 32) {
            uint32_t o[32];
            tmem_ld_x32(tmem_base + lane_base + kOCol + c0, o);
            tc_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * corr);
            tmem_st_x32(tmem_base + lane_base + kOCol + c0, o);
          }
        }
      }
      tmem_st_x32(tmem_base + lane_base + kPCol, pk);
      tc_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> fp16
    mbar_wait_bounded(pv_done, (npages - 1) & 1);
    tc_fence_after();
    const float inv_l = valid ? 1.f / l_run : 0.f;
    __half* dst = a.out + (static_cast<size_t>(row0 + tq) * a.H + hkv * G + g) * kD;
#pragma unroll 1
    for (int c0 = 0; c0 < kD; c0 += 32) {
      uint32_t o[32];
      tmem_ld_x32(tmem_base + lane_base + kOCol + c0, o);
      tc_wait_ld();
      if (valid) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack16x2<BF>(__uint_as_float(o[i]) * inv_l, __uint_as_float(o[i + 1]) * inv_l);
          w.y = pack16x2<BF>(__uint_as_float(o[i + 2]) * inv_l, __uint_as_float(o[i + 3]) * inv_l);
          w.z = pack16x2<BF>(__uint_as_float(o[i + 4]) * inv_l, __uint_as_float(o[i + 5]) * inv_l);
          w.w = pack16x2<BF>(__uint_as_float(o[i + 6]) * inv_l, __uint_as_float(o[i + 7]) * inv_l);
          *reinterpret_cast<uint4*>(dst + c0 + i) = w;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, kTmemCols);
  }
}

}  // namespace pft

static int prefill_attn_mode() {       // 0 = mma.sync kernel, 1 = tcgen05 kernel (default), 2 = tcgen05 with dbg bit 0
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("B200_PREFILL_ATTN");
    mode = 1;
    if (e && !strcmp(e, "mma")) mode = 0;
    if (e && !strcmp(e, "tc_swapdesc")) mode = 2;
  }
  return mode;
}

// false: shape not taken here (head_dim != 128, or the mma.sync kernel was asked for)
bool launch_attn_prefill_tc(const AttnMaps& maps, void* out, const void* q, int q_stride, const uint32_t* cu_q,
                            const uint32_t* ctx_lens, const uint32_t* block_tables, int B, int H, int Hkv, int D,
                            int max_blocks, int max_seqlen_q, int total_rows, float scale, int mode,
                            cudaStream_t st, bool bf16) {
  using namespace pft;
  if (mode < 0) mode = prefill_attn_mode();
  if (mode == 0 || D != kD) return false;
  const int G = H / Hkv;
  if (kRows % G != 0 || (q_stride * 2) % 16 != 0 || (reinterpret_cast<uintptr_t>(q) & 15)) return false;
  const int TQ = kRows / G;
  // q as a 3-D tensor {D, heads, rows}: one box = 64 D-elements x the G heads of a kv head x TQ tokens,
  // landing as 128 swizzled 128-byte rows (row = token * G + head)
  CUtensorMap qmap;
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(kD), static_cast<uint64_t>(H), static_cast<uint64_t>(total_rows)};
    const uint64_t strides[2] = {static_cast<uint64_t>(kD) * 2, static_cast<uint64_t>(q_stride) * 2};
    const uint32_t box[3] = {64, static_cast<uint32_t>(G), static_cast<uint32_t>(TQ)};
    const uint32_t es[3] = {1, 1, 1};
    if (encode_tensor_map(&qmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(q), dims, strides, box, es,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B) != CUDA_SUCCESS) {
      set_error(2, "attn_prefill_tc: cuTensorMapEncodeTiled (q) failed");
      return true;
    }
  }
  Args a;
  a.out = reinterpret_cast<__half*>(out);
  a.cu_seqlens_q = cu_q; a.context_lens = ctx_lens; a.block_tables = block_tables;
  a.H = H; a.Hkv = Hkv; a.max_blocks = max_blocks;
  a.scale_log2 = scale * 1.4426950408889634f;
  a.dbg = mode == 2 ? 1 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_prefill_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAlloc);
    cudaFuncSetAttribute(attn_prefill_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAlloc);
    cudaFuncSetAttribute(attn_prefill_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAlloc);
    cudaFuncSetAttribute(attn_prefill_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAlloc);
    attr_set = true;
  }
  dim3 grid((max_seqlen_q + TQ - 1) / TQ, Hkv, B);
  auto kern = maps.l1 ? (bf16 ? attn_prefill_tc_kernel<true, true> : attn_prefill_tc_kernel<false, true>)
                      : (bf16 ? attn_prefill_tc_kernel<true, false> : attn_prefill_tc_kernel<false, false>);
  launch_pdl(kern, grid, dim3(kThreads), kAlloc, st, qmap, maps.k, maps.v, a);
  B200_CUDA_LAUNCH_CHECK();
  return true;
}

}  // namespace b200
