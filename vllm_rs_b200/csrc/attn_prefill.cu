// Varlen causal attention for a prefill step over the PAGED cache (fresh prompts and later
// chunks alike: the KV write of this step's tokens happens first, attention.rs:808-820, so every
// key of [0, context_len) is in the cache and is read through the block table).
//
// Replaces the prefill branch of attention_rs::PagedAttention::forward
// (/root/reference/src/models/layers/attention.rs:808-820) with the metadata of
// ModelRunner::prepare_prefill (/root/reference/src/core/runner.rs:978-1241): cu_seqlens_q u32
// [B+1], context_lens u32 [B] (= num_cached_tokens + chunk tokens), block_tables u32 [B, max_blocks].
//
// Design (flash-attention-2 dataflow on mma.sync, K/V pages gathered by TMA like the decode
// kernel): grid = (query tile, kv head, sequence).  A CTA owns 128 rows = G query heads x 128/G
// query tokens of ONE kv head, so a K/V page is staged once for the whole GQA group.  8 consumer
// warps x 16 rows; a producer warp walks the block table 32 entries at a time and issues the
// 4-D TMA boxes into a 4-stage mbarrier ring.  fp32 online softmax, causal mask per row,
// P rounded to fp16 for the PV MMA.  Compute-bound on the legacy tensor path (HMMA); a tcgen05
// version is the next step for this row of SURVEY section 8.
#include "../../include/b200_vllm.h"
#include "attn.cuh"

namespace b200 {

namespace pf {
constexpr int kTok = 64;
constexpr int kStages = 4;
constexpr int kWarps = 8;
constexpr int kThreads = (kWarps + 1) * 32;

template <int D>
struct Smem {
  static constexpr int kTileBytes = kTok * D * 2;
  static constexpr int kStageBytes = 2 * kTileBytes;
  static constexpr int kBarOff = kStages * kStageBytes;
  static constexpr int kTotal = kBarOff + 128;
  static constexpr int kAlloc = kTotal + 1024;
};
}  // namespace pf

// L1: the reference's default-build cache layout, staged and read as in paged_attn_decode.cu (K tile as it lies,
// V tile = D rows of 64 swizzled tokens through a non-transposed ldmatrix).
template <int D, bool BF, bool L1>
__global__ void __launch_bounds__(pf::kThreads, 1)
attn_prefill_kernel(const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap,
                    const __half* __restrict__ q, __half* __restrict__ out,
                    const uint32_t* __restrict__ cu_seqlens_q,
                    const uint32_t* __restrict__ context_lens,
                    const uint32_t* __restrict__ block_tables, int q_stride, int H, int Hkv,
                    int max_blocks, float scale_log2) {
  using namespace pf;
  using SM = Smem<D>;
  const int qt = blockIdx.x, hkv = blockIdx.y, b = blockIdx.z;
  const int G = H / Hkv;
  const int wph = kWarps / G;              // warps per query head
  const int TQ = wph * 16;                 // query tokens per CTA
  const int row0 = static_cast<int>(cu_seqlens_q[b]);
  const int nq = static_cast<int>(cu_seqlens_q[b + 1]) - row0;
  const int q0 = qt * TQ;
  if (q0 >= nq) return;
  const int ctx = context_lens ? static_cast<int>(context_lens[b]) : nq;
  const int cached = ctx - nq;
  const int q_end = min(q0 + TQ, nq);
  const int last_key = cached + q_end - 1;               // last visible key of the tile
  const int npages = last_key / kTok + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
  uint64_t* empty_bar = full_bar + kStages;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kWarps);
    }
    fence_barrier_init();
    prefetch_tmap(&kmap);
    prefetch_tmap(&vmap);
  }
  __syncthreads();
  pdl_trigger();
  pdl_wait();

  if (warp == kWarps) {
    // ------------------------------------------------------------ producer
    const uint32_t* bt = block_tables + static_cast<size_t>(b) * max_blocks;
    for (int base = 0; base < npages; base += 32) {
      const int mine = base + lane < npages ? static_cast<int>(bt[base + lane]) : 0;
      const int cnt = min(32, npages - base);
      for (int j = 0; j < cnt; ++j) {
        const int blk = __shfl_sync(0xffffffffu, mine, j);
        if (lane == 0) {
          const int i = base + j;
          const int s = i % kStages;
          mbar_wait(&empty_bar[s], ((i / kStages) & 1) ^ 1);
          mbar_arrive_expect_tx(&full_bar[s], SM::kStageBytes);
          uint8_t* kdst = smem + s * SM::kStageBytes;
          uint8_t* vdst = kdst + SM::kTileBytes;
          if (L1) {
            tma_load_4d(kdst, &kmap, &full_bar[s], 0, 0, hkv, blk);
            tma_load_4d(vdst, &vmap, &full_bar[s], 0, 0, hkv, blk);
          } else {
#pragma unroll
            for (int hb = 0; hb < D / 64; ++hb) {
              tma_load_4d(kdst + hb * (kTok * 128), &kmap, &full_bar[s], hb * 64, hkv, 0, blk);
              tma_load_4d(vdst + hb * (kTok * 128), &vmap, &full_bar[s], hb * 64, hkv, 0, blk);
            }
          }
        }
      }
    }
    return;
  }
  // -------------------------------------------------------------- consumers
  const int g = lane >> 2, c = lane & 3;
  const int head = hkv * G + warp / wph;
  const int tq0 = q0 + (warp % wph) * 16;                 // first query token of this warp
  const int tqa = tq0 + g, tqb = tq0 + g + 8;             // the two rows of this thread
  const bool va = tqa < nq, vb = tqb < nq;
  const int qpos_a = cached + tqa, qpos_b = cached + tqb; // causal limit per row
  const int warp_last_key = cached + min(tq0 + 16, nq) - 1;
  const bool warp_active = tq0 < nq;

  uint32_t qa[D / 16][4];
  {
    const __half* qra = q + static_cast<size_t>(row0 + tqa) * q_stride + head * D;
    const __half* qrb = q + static_cast<size_t>(row0 + tqb) * q_stride + head * D;
#pragma unroll
    for (int ks = 0; ks < D / 16; ++ks) {
      const int d0 = ks * 16 + 2 * c;
      qa[ks][0] = va ? *reinterpret_cast<const uint32_t*>(qra + d0) : 0u;
      qa[ks][1] = vb ? *reinterpret_cast<const uint32_t*>(qrb + d0) : 0u;
      qa[ks][2] = va ? *reinterpret_cast<const uint32_t*>(qra + d0 + 8) : 0u;
      qa[ks][3] = vb ? *reinterpret_cast<const uint32_t*>(qrb + d0 + 8) : 0u;
    }
  }
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  float o_acc[D / 8][4];
#pragma unroll
  for (int j = 0; j < D / 8; ++j) o_acc[j][0] = o_acc[j][1] = o_acc[j][2] = o_acc[j][3] = 0.f;

  const int k_row = (lane >> 4) * 8 + (lane & 7);
  const int k_chunk_sel = (lane >> 3) & 1;
  const int v_row = ((lane >> 3) & 1) * 8 + (lane & 7);
  const int v_chunk_sel = lane >> 4;

  for (int i = 0; i < npages; ++i) {
    const int s = i % kStages;
    mbar_wait(&full_bar[s], (i / kStages) & 1);
    const int page_base = i * kTok;
    if (warp_active && page_base <= warp_last_key) {
      const uint32_t kbase = smem_u32(smem + s * SM::kStageBytes);
      const uint32_t vbase = kbase + SM::kTileBytes;
      float sc[8][4];
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        const int hb = ks >> 2;
        const int chunk = ((ks & 3) << 1) + k_chunk_sel;
#pragma unroll
        for (int np = 0; np < 4; ++np) {
          const int tok = np * 16 + k_row;
          uint32_t r[4];
          if (L1) ldsm_x4(r, kbase + (2 * ks + k_chunk_sel) * (kTok * 16) + tok * 16);
          else ldsm_x4(r, kbase + hb * (kTok * 128) + tok * 128 + ((chunk ^ (tok & 7)) << 4));
          mma_16816_t<BF>(sc[2 * np], qa[ks], r[0], r[1]);
          mma_16816_t<BF>(sc[2 * np + 1], qa[ks], r[2], r[3]);
        }
      }
      // ---- causal mask + online softmax (rows a: elements 0,1; rows b: elements 2,3)
      const bool need_mask = page_base + kTok - 1 > cached + tq0;   // some (row, key) pair masked
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          float v = sc[nt][e] * scale_log2;
          if (need_mask) {
            const int kt = page_base + nt * 8 + 2 * c + (e & 1);
            const int lim = (e < 2) ? qpos_a : qpos_b;
            v = kt <= lim ? v : -INFINITY;
          }
          sc[nt][e] = v;
          mx[e >> 1] = fmaxf(mx[e >> 1], v);
        }
      }
      float alpha[2], base[2];
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) {
        mx[rh] = fmaxf(mx[rh], __shfl_xor_sync(0xffffffffu, mx[rh], 1));
        mx[rh] = fmaxf(mx[rh], __shfl_xor_sync(0xffffffffu, mx[rh], 2));
        const float m_new = fmaxf(m_run[rh], mx[rh]);
        base[rh] = (m_new == -INFINITY) ? 0.f : m_new;
        alpha[rh] = exp2f(m_run[rh] - base[rh]);
        m_run[rh] = m_new;
      }
      float psum[2] = {0.f, 0.f};
#pragma unroll
      for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float p = exp2f(sc[nt][e] - base[e >> 1]);
          sc[nt][e] = p;
          psum[e >> 1] += p;
        }
      }
#pragma unroll
      for (int rh = 0; rh < 2; ++rh) l_run[rh] = l_run[rh] * alpha[rh] + psum[rh];
#pragma unroll
      for (int j = 0; j < D / 8; ++j) {
        o_acc[j][0] *= alpha[0];
        o_acc[j][1] *= alpha[0];
        o_acc[j][2] *= alpha[1];
        o_acc[j][3] *= alpha[1];
      }
      // ---- O += P V
      const bool partial = page_base + kTok > ctx;       // stale rows past the context: zero V
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t pa[4];
        pa[0] = pack16x2<BF>(sc[2 * kk][0], sc[2 * kk][1]);
        pa[1] = pack16x2<BF>(sc[2 * kk][2], sc[2 * kk][3]);
        pa[2] = pack16x2<BF>(sc[2 * kk + 1][0], sc[2 * kk + 1][1]);
        pa[3] = pack16x2<BF>(sc[2 * kk + 1][2], sc[2 * kk + 1][3]);
        const int tb = page_base + kk * 16;
        const uint32_t vmask_lo = (tb + 2 * c < ctx ? 0x0000FFFFu : 0u) | (tb + 2 * c + 1 < ctx ? 0xFFFF0000u : 0u);
        const uint32_t vmask_hi = (tb + 8 + 2 * c < ctx ? 0x0000FFFFu : 0u) | (tb + 9 + 2 * c < ctx ? 0xFFFF0000u : 0u);
        const int tok = kk * 16 + v_row;
#pragma unroll
        for (int jj = 0; jj < D / 16; ++jj) {
          const int chunk = 2 * jj + v_chunk_sel;
          const int hb = chunk >> 3, c16 = chunk & 7;
          uint32_t r[4];
          if (L1) {
            const int dd = (2 * jj + (lane >> 4)) * 8 + (lane & 7);
            const int tch = 2 * kk + ((lane >> 3) & 1);
            ldsm_x4(r, vbase + dd * 128 + ((tch ^ (lane & 7)) << 4));
          } else {
            ldsm_x4_t(r, vbase + hb * (kTok * 128) + tok * 128 + ((c16 ^ (tok & 7)) << 4));
          }
          if (partial) {
            r[0] &= vmask_lo; r[2] &= vmask_lo;
            r[1] &= vmask_hi; r[3] &= vmask_hi;
          }
          mma_16816_t<BF>(o_acc[2 * jj], pa, r[0], r[1]);
          mma_16816_t<BF>(o_acc[2 * jj + 1], pa, r[2], r[3]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[s]);
  }
  // ---- normalise and store
#pragma unroll
  for (int rh = 0; rh < 2; ++rh) {
    l_run[rh] += __shfl_xor_sync(0xffffffffu, l_run[rh], 1);
    l_run[rh] += __shfl_xor_sync(0xffffffffu, l_run[rh], 2);
  }
  const float inv_a = va ? 1.f / l_run[0] : 0.f, inv_b = vb ? 1.f / l_run[1] : 0.f;
  __half* oa = out + (static_cast<size_t>(row0 + tqa) * H + head) * D;
  __half* ob = out + (static_cast<size_t>(row0 + tqb) * H + head) * D;
#pragma unroll
  for (int j = 0; j < D / 8; ++j) {
    if (va) *reinterpret_cast<uint32_t*>(oa + j * 8 + 2 * c) = pack16x2<BF>(o_acc[j][0] * inv_a, o_acc[j][1] * inv_a);
    if (vb) *reinterpret_cast<uint32_t*>(ob + j * 8 + 2 * c) = pack16x2<BF>(o_acc[j][2] * inv_b, o_acc[j][3] * inv_b);
  }
}

template <int D, bool BF, bool L1>
static void launch_pf(const AttnMaps& maps, void* out, const void* q, int q_stride,
                      const uint32_t* cu_q, const uint32_t* ctx_lens, const uint32_t* block_tables,
                      int B, int H, int Hkv, int max_blocks, int max_seqlen_q, float scale,
                      cudaStream_t st) {
  const int G = H / Hkv;
  const int TQ = (pf::kWarps / G) * 16;
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(attn_prefill_kernel<D, BF, L1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         pf::Smem<D>::kAlloc);
    attr_set = true;
  }
  dim3 grid((max_seqlen_q + TQ - 1) / TQ, Hkv, B);
  launch_pdl(attn_prefill_kernel<D, BF, L1>, grid, dim3(pf::kThreads), pf::Smem<D>::kAlloc, st, maps.k, maps.v,
             reinterpret_cast<const __half*>(q), reinterpret_cast<__half*>(out), cu_q, ctx_lens,
             block_tables, q_stride, H, Hkv, max_blocks, scale * 1.4426950408889634f);
  B200_CUDA_LAUNCH_CHECK();
}

void launch_attn_prefill(const AttnMaps& maps, void* out, const void* q, int q_stride,
                         const uint32_t* cu_q, const uint32_t* ctx_lens,
                         const uint32_t* block_tables, int B, int H, int Hkv, int D, int max_blocks,
                         int max_seqlen_q, float scale, cudaStream_t st, int total_rows, int mode, bool bf16) {
  // total_rows > 0: the caller knows the extent of the q buffer, which the tcgen05 kernel's q tensor map needs
  if (total_rows > 0 && launch_attn_prefill_tc(maps, out, q, q_stride, cu_q, ctx_lens, block_tables, B, H, Hkv, D,
                                               max_blocks, max_seqlen_q, total_rows, scale, mode, st, bf16))
    return;
#define B200_PF_ARGS maps, out, q, q_stride, cu_q, ctx_lens, block_tables, B, H, Hkv, max_blocks, max_seqlen_q, scale, st
  if (maps.l1) {
    if (D == 128 && bf16) launch_pf<128, true, true>(B200_PF_ARGS);
    else if (D == 128) launch_pf<128, false, true>(B200_PF_ARGS);
    else if (bf16) launch_pf<64, true, true>(B200_PF_ARGS);
    else launch_pf<64, false, true>(B200_PF_ARGS);
  } else if (D == 128 && bf16) launch_pf<128, true, false>(B200_PF_ARGS);
  else if (D == 128) launch_pf<128, false, false>(B200_PF_ARGS);
  else if (bf16) launch_pf<64, true, false>(B200_PF_ARGS);
  else launch_pf<64, false, false>(B200_PF_ARGS);
#undef B200_PF_ARGS
}

}  // namespace b200

static void attn_prefill_ex_any(void* out, const void* q, const void* k_cache,
                                const void* v_cache, const uint32_t* cu_seqlens_q,
                                const uint32_t* context_lens, const uint32_t* block_tables,
                                int32_t batch, int32_t num_heads, int32_t num_kv_heads,
                                int32_t head_dim, int32_t block_size,
                                int32_t max_blocks_per_seq, int32_t num_blocks,
                                int32_t max_seqlen_q, int32_t num_tokens, int32_t mode, float scale,
                                int64_t stream, bool bf16, bool l1 = false) {
  using namespace b200;
  B200_CHECK_ARG(block_size == 64, "attn_prefill: block_size must be 64");
  B200_CHECK_ARG(head_dim == 64 || head_dim == 128, "attn_prefill: head_dim must be 64/128");
  const int G = num_kv_heads > 0 ? num_heads / num_kv_heads : 0;
  B200_CHECK_ARG(num_kv_heads > 0 && num_heads % num_kv_heads == 0 &&
                     (G == 1 || G == 2 || G == 4 || G == 8),
                 "attn_prefill: H/Hkv must be 1, 2, 4 or 8");
  B200_CHECK_ARG(batch > 0 && max_seqlen_q > 0 && num_tokens > 0 && block_tables && cu_seqlens_q,
                 "attn_prefill: bad sizes or NULL metadata");
  AttnMaps maps;
  if (!make_attn_maps(&maps, k_cache, v_cache, num_blocks, num_kv_heads, head_dim, false, l1)) {
    set_error(2, "attn_prefill: cuTensorMapEncodeTiled failed");
    return;
  }
  launch_attn_prefill(maps, out, q, num_heads * head_dim, cu_seqlens_q, context_lens, block_tables,
                      batch, num_heads, num_kv_heads, head_dim, max_blocks_per_seq, max_seqlen_q,
                      scale, reinterpret_cast<cudaStream_t>(stream), num_tokens, mode, bf16);
}

extern "C" void b200_attn_prefill_ex_f16(void* out, const void* q, const void* k_cache,
                                         const void* v_cache, const uint32_t* cu_seqlens_q,
                                         const uint32_t* context_lens, const uint32_t* block_tables,
                                         int32_t batch, int32_t num_heads, int32_t num_kv_heads,
                                         int32_t head_dim, int32_t block_size,
                                         int32_t max_blocks_per_seq, int32_t num_blocks,
                                         int32_t max_seqlen_q, int32_t num_tokens, int32_t mode, float scale,
                                         int64_t stream) {
  attn_prefill_ex_any(out, q, k_cache, v_cache, cu_seqlens_q, context_lens, block_tables, batch, num_heads,
                      num_kv_heads, head_dim, block_size, max_blocks_per_seq, num_blocks, max_seqlen_q, num_tokens,
                      mode, scale, stream, false);
}

extern "C" void b200_attn_prefill_l1(void* out, const void* q, const void* k_cache,
                                     const void* v_cache, const uint32_t* cu_seqlens_q,
                                     const uint32_t* context_lens, const uint32_t* block_tables,
                                     int32_t batch, int32_t num_heads, int32_t num_kv_heads,
                                     int32_t head_dim, int32_t block_size,
                                     int32_t max_blocks_per_seq, int32_t num_blocks,
                                     int32_t max_seqlen_q, int32_t num_tokens, int32_t mode, float scale,
                                     int32_t bf16, int64_t stream) {
  attn_prefill_ex_any(out, q, k_cache, v_cache, cu_seqlens_q, context_lens, block_tables, batch, num_heads,
                      num_kv_heads, head_dim, block_size, max_blocks_per_seq, num_blocks, max_seqlen_q, num_tokens,
                      mode, scale, stream, bf16 != 0, true);
}

extern "C" void b200_attn_prefill_ex_bf16(void* out, const void* q, const void* k_cache,
                                          const void* v_cache, const uint32_t* cu_seqlens_q,
                                          const uint32_t* context_lens, const uint32_t* block_tables,
                                          int32_t batch, int32_t num_heads, int32_t num_kv_heads,
                                          int32_t head_dim, int32_t block_size,
                                          int32_t max_blocks_per_seq, int32_t num_blocks,
                                          int32_t max_seqlen_q, int32_t num_tokens, int32_t mode, float scale,
                                          int64_t stream) {
  attn_prefill_ex_any(out, q, k_cache, v_cache, cu_seqlens_q, context_lens, block_tables, batch, num_heads,
                      num_kv_heads, head_dim, block_size, max_blocks_per_seq, num_blocks, max_seqlen_q, num_tokens,
                      mode, scale, stream, true);
}

extern "C" void b200_attn_prefill_f16(void* out, const void* q, const void* k_cache,
                                      const void* v_cache, const uint32_t* cu_seqlens_q,
                                      const uint32_t* context_lens, const uint32_t* block_tables,
                                      int32_t batch, int32_t num_heads, int32_t num_kv_heads,
                                      int32_t head_dim, int32_t block_size,
                                      int32_t max_blocks_per_seq, int32_t num_blocks,
                                      int32_t max_seqlen_q, float scale, int64_t stream) {
  using namespace b200;
  B200_CHECK_ARG(block_size == 64, "attn_prefill: block_size must be 64");
  B200_CHECK_ARG(head_dim == 64 || head_dim == 128, "attn_prefill: head_dim must be 64/128");
  const int G = num_kv_heads > 0 ? num_heads / num_kv_heads : 0;
  B200_CHECK_ARG(num_kv_heads > 0 && num_heads % num_kv_heads == 0 &&
                     (G == 1 || G == 2 || G == 4 || G == 8),
                 "attn_prefill: H/Hkv must be 1, 2, 4 or 8");
  B200_CHECK_ARG(batch > 0 && max_seqlen_q > 0 && block_tables && cu_seqlens_q,
                 "attn_prefill: bad sizes or NULL metadata");
  AttnMaps maps;
  if (!make_attn_maps(&maps, k_cache, v_cache, num_blocks, num_kv_heads, head_dim)) {
    set_error(2, "attn_prefill: cuTensorMapEncodeTiled failed");
    return;
  }
  launch_attn_prefill(maps, out, q, num_heads * head_dim, cu_seqlens_q, context_lens, block_tables,
                      batch, num_heads, num_kv_heads, head_dim, max_blocks_per_seq, max_seqlen_q,
                      scale, reinterpret_cast<cudaStream_t>(stream));
}
