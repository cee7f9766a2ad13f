// Paged-KV attention, decode step (one query token per sequence).
//
// Replaces the decode branch of attention_rs::PagedAttention::forward as called from
// /root/reference/src/models/layers/attention.rs:808-820 with the metadata built by
// ModelRunner::prepare_decode (/root/reference/src/core/runner.rs:1243-1388):
//   block_tables u32 [B, max_blocks] (zero padded), context_lens u32 [B] -- both read
//   on the device so the launch is CUDA-graph safe (graph.rs:741-767 pads rows with
//   context_len 0).
//
// B200 design
//   * KV cache layout [num_blocks, 64, Hkv, D] fp16 (reference layout L2).  One 4-D
//     TMA tensor map per cache; a (page, kv-head) tile is fetched by block-table index
//     as D/64 boxes of [64 tokens x 64 d] (128-byte rows, SWIZZLE_128B) straight into
//     shared memory by a dedicated producer warp; 3-stage mbarrier ring, 2 CTAs/SM.
//   * grid = (kv split, kv head, sequence): every CTA streams `pages_per_split` pages
//     of ONE kv head and serves the whole GQA group (H/Hkv query heads share each byte
//     fetched).  4 consumer warps each own 16 of the 64 tokens of a page:
//     S = Q K^T and O += P V run on mma.sync m16n8k16 (rows = query heads, padded to
//     16) with ldmatrix on the swizzled tile; online softmax in fp32 with quad
//     shuffles.  HBM-bound: 4 flop/B of useful math, the tensor pipe is ~15% busy.
//   * split partials (fp32 O, m, l) go to the caller's workspace and are merged in fixed split
//     order (deterministic) by a small second kernel launched with PDL; single-split sequences
//     write fp16 output directly.
//
// Algorithmic bytes per launch: B * ctx * Hkv * D * 2 (K) * 2 (V) * 2 B  (+ q/out).
//
// FP8 KV cache (`--fp8-kvcache`, kvcache_allocator.rs:188-193,776: cache dtype U8, same
// [num_blocks, 64, Hkv, D] shape; attention.rs:615 hands only the flag to attention-rs, so the
// storage format is this library's: e4m3, scale 1, saturating round-to-nearest from the fp16 K/V).
// The F8 instantiation streams half the bytes: a (page, head) tile is ONE box of 64 rows x 128 B,
// 6 stages keep the same bytes in flight, and the consumers expand e4m3 -> fp16 in registers:
//   * K: the same ldmatrix addresses now return 4 BYTES of one token row per lane; the contraction
//     index is permuted (lane c's k-slots {2c,2c+1,2c+8,2c+9} <-> d = 16 ks + 4c + {0,1,2,3}) and the
//     Q fragments are loaded with the same permutation, so no shuffles are needed;
//   * V: ldmatrix.trans on the byte tile yields (d even, d odd) byte pairs of two tokens; one prmt
//     regroups them into the {token 2c, 2c+1, 2c+8, 2c+9} bytes of an even-d and an odd-d n-tile, whose
//     output columns are un-permuted when O is written out.
#include <cuda_fp8.h>
#include <cstdio>
#include <type_traits>
#include <cstdlib>
#include "../../include/b200_vllm.h"
#include "attn.cuh"

namespace b200 {

constexpr int kTok = 64;        // tokens per page (block_size)
constexpr int kStages = 3;
constexpr int kConsumerWarps = 4;
constexpr int kThreads = (kConsumerWarps + 1) * 32;
constexpr int kMaxPagesPerSplit = 128;

// NS = ring depth.  fp16: 3 stages x 32 KB, 2 CTAs per SM.  F8 (16 KB stages): 6 stages, 2 CTAs per SM -- the same
// 192 KB in flight per SM.  (4 stages with 3 CTAs per SM was measured slower: 128 registers spill inside the
// page loop, 136 do not fit three CTAs -- profiles/r1_v7_attn_fp8_variants.txt.)
template <int D, bool F8 = false, int NS = kStages>
struct AttnSmem {
  static constexpr int kNumStages = NS;
  static constexpr int kTileBytes = kTok * D * (F8 ? 1 : 2);      // one K (or V) page tile
  static constexpr int kStageBytes = 2 * kTileBytes;
  static constexpr int kDataBytes = kNumStages * kStageBytes;
  static constexpr int kBarOff = kDataBytes;
  static constexpr int kIdsOff = kBarOff + 128;
  static constexpr int kMlOff = kIdsOff + kMaxPagesPerSplit * 4;
  static constexpr int kTotal = kMlOff + kConsumerWarps * 16 * 2 * 4;
  static constexpr int kAlloc = kTotal + 1024;          // manual 1024-B alignment slack
};

// four e4m3 bytes -> half2 of bytes (0,1) and half2 of bytes (2,3).  Same instruction cuda_fp8.hpp's
// __nv_cvt_fp8x2_to_halfraw2 emits (low byte -> low half), written inline so the two f16x2 results stay
// whole 32-bit registers (the struct round trip cost two PRMTs per conversion).
__device__ __forceinline__ void e4m3x4_to_half2x2(uint32_t w, uint32_t& lo, uint32_t& hi) {
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.rn.f16x2.e4m3x2 %0, l;\n\t"
      "cvt.rn.f16x2.e4m3x2 %1, h;\n\t}\n"
      : "=r"(lo), "=r"(hi)
      : "r"(w));
}

// 2^x in one MUFU (exp2f without fast-math wraps it in a denormal-range rescale); ex2.approx(-inf) = +0
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ int first_page_of(int ctx, int window) {
  return (window > 0 && ctx > window) ? (ctx - window) / kTok : 0;
}

// HI = false promises H / Hkv <= 8: rows 8..15 of every A fragment are the constant 0, which frees the 16
// registers the fp8 instantiation needs to fit 3 CTAs per SM without spilling.
// L1 = the reference's default-build cache layout (kvcache_allocator.rs:150-173, no flashattn / flashinfer feature):
// K [num_blocks, Hkv, D/8, 64, 8], V [num_blocks, Hkv, D, 64].  A (block, head) tile is 64*D*2 contiguous bytes in
// both: K lands as it lies (16-byte chunks of 8 d-values, chunk-major, token-minor: eight consecutive token rows of
// one chunk are 128 contiguous bytes, conflict-free for ldmatrix without a swizzle); V lands as D rows of 64 tokens
// = 128 bytes under SWIZZLE_128B and feeds the P V MMA through a NON-transposed ldmatrix (rows = d, 16 bytes = 8
// consecutive tokens, which is the B-fragment's k pairing).
template <int D, bool F8, bool HI, int NS, bool BF, bool L1>
__device__ __forceinline__ void
paged_attn_decode_body(const CUtensorMap& kmap, const CUtensorMap& vmap,
                       const __half* __restrict__ q, __half* __restrict__ out,
                       const uint32_t* __restrict__ block_tables,
                       const uint32_t* __restrict__ context_lens, float* __restrict__ part_o,
                       float* __restrict__ part_ml, unsigned* __restrict__ counters,
                       int q_stride, int H, int Hkv, int max_blocks,
                       int pages_per_split, int max_splits, float scale_log2, float softcap,
                       int window) {
  using SM = AttnSmem<D, F8, NS>;
  constexpr int kNS = SM::kNumStages;
  static_assert(!F8 || D == 128, "fp8 tiles are one 128-byte row per token");
  static_assert(!(F8 && BF), "the e4m3 cache is expanded to fp16 operands");
  static_assert(!(F8 && L1), "the e4m3 cache is built for the flash layout only");
  const int split = blockIdx.x, hkv = blockIdx.y, b = blockIdx.z;
  const int ctx = static_cast<int>(context_lens[b]);
  if (ctx <= 0) {
    // padded graph rows (graph.rs:741-767): define the output, do no work
    pdl_trigger();
    pdl_wait();
    if (split == 0) {
      const int Gz = H / Hkv;
      for (int i = threadIdx.x; i < Gz * D; i += kThreads)
        out[(static_cast<size_t>(b) * H + static_cast<size_t>(hkv) * Gz) * D + i] = __float2half_rn(0.f);
    }
    return;
  }
  const int npages = (ctx + kTok - 1) / kTok;
  const int fpage = first_page_of(ctx, window);
  const int page0 = fpage + split * pages_per_split;
  const int page1 = min(page0 + pages_per_split, npages);
  if (page0 >= page1) return;
  const int n_cta_pages = page1 - page0;
  const int nsplits_seq = (npages - fpage + pages_per_split - 1) / pages_per_split;
  const int G = H / Hkv;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
  uint64_t* empty_bar = full_bar + kNS;
  uint32_t* sm_ids = reinterpret_cast<uint32_t*>(smem + SM::kIdsOff);
  float* sm_ml = reinterpret_cast<float*>(smem + SM::kMlOff);   // [warp][16][2]

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kNS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerWarps);
    }
    fence_barrier_init();
  }
  if (warp == kConsumerWarps) {
    for (int i = lane; i < n_cta_pages; i += 32)
      sm_ids[i] = block_tables[static_cast<size_t>(b) * max_blocks + page0 + i];
    if (lane == 0) {
      prefetch_tmap(&kmap);
      prefetch_tmap(&vmap);
    }
  }
  __syncthreads();
  // prologue above touched only step-constant inputs (context_lens, block_tables): it overlaps
  // the previous kernel's tail.  q and the newest KV rows come from the previous kernel: the
  // consumers wait here, the producer streams every page EXCEPT the sequence's last one (the only
  // page this step writes) before it waits -- the ring is full when q becomes visible.
  pdl_trigger();
  if (warp != kConsumerWarps) pdl_wait();

  // per-thread softmax / output state (consumer warps)
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  float o_acc[D / 8][4];
#pragma unroll
  for (int j = 0; j < D / 8; ++j) o_acc[j][0] = o_acc[j][1] = o_acc[j][2] = o_acc[j][3] = 0.f;

  const int g = lane >> 2, c = lane & 3;
  const bool hi_rows = HI && G > 8;

  if (warp == kConsumerWarps) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int i = 0; i < n_cta_pages; ++i) {
        const int s = i % kNS;
        const uint32_t ph = (i / kNS) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (page0 + i == npages - 1) pdl_wait();       // the page holding this step's new token
        mbar_arrive_expect_tx(&full_bar[s], SM::kStageBytes);
        const int blk = static_cast<int>(sm_ids[i]);
        uint8_t* kdst = smem + s * SM::kStageBytes;
        uint8_t* vdst = kdst + SM::kTileBytes;
        if (F8) {
          tma_load_4d(kdst, &kmap, &full_bar[s], 0, hkv, 0, blk);       // 64 rows x 128 e4m3 bytes
          tma_load_4d(vdst, &vmap, &full_bar[s], 0, hkv, 0, blk);
        } else if (L1) {
          tma_load_4d(kdst, &kmap, &full_bar[s], 0, 0, hkv, blk);       // the contiguous tile, as 512-byte rows
          tma_load_4d(vdst, &vmap, &full_bar[s], 0, 0, hkv, blk);       // D rows x 64 tokens, swizzled
        } else {
#pragma unroll
          for (int hb = 0; hb < D / 64; ++hb) {
            tma_load_4d(kdst + hb * (kTok * 128), &kmap, &full_bar[s], hb * 64, hkv, 0, blk);
            tma_load_4d(vdst + hb * (kTok * 128), &vmap, &full_bar[s], hb * 64, hkv, 0, blk);
          }
        }
      }
    }
  } else {
    // ------------------------------------------------------------ consumers
    // Q fragments (A operand, rows = query heads of this kv group, zero padded to 16)
    uint32_t qa[D / 16][4];
    {
      const __half* qb = q + static_cast<size_t>(b) * q_stride + static_cast<size_t>(hkv) * G * D;
#pragma unroll
      for (int ks = 0; ks < D / 16; ++ks) {
        // F8: k-slots {2c, 2c+1 | 2c+8, 2c+9} of step ks stand for d = 16 ks + 4c + {0, 1 | 2, 3}
        const int d0 = ks * 16 + (F8 ? 4 : 2) * c;
        constexpr int kHi = F8 ? 2 : 8;
        qa[ks][0] = (g < G) ? *reinterpret_cast<const uint32_t*>(qb + g * D + d0) : 0u;
        qa[ks][1] = (HI && g + 8 < G) ? *reinterpret_cast<const uint32_t*>(qb + (g + 8) * D + d0) : 0u;
        qa[ks][2] = (g < G) ? *reinterpret_cast<const uint32_t*>(qb + g * D + d0 + kHi) : 0u;
        qa[ks][3] =
            (HI && g + 8 < G) ? *reinterpret_cast<const uint32_t*>(qb + (g + 8) * D + d0 + kHi) : 0u;
      }
    }
    const int t0 = warp * 16;
    const int win_lo = (window > 0) ? ctx - window : 0;
    const float inv_cap = softcap > 0.f ? 1.f / softcap : 0.f;
    const float LOG2E = 1.4426950408889634f;
    const float scale_nat = scale_log2 / LOG2E;

    // ldmatrix lane address components
    const int k_tok = t0 + (lane >> 4) * 8 + (lane & 7);    // K: matrix = lane/8
    const int k_chunk_sel = (lane >> 3) & 1;
    const int v_tok = t0 + ((lane >> 3) & 1) * 8 + (lane & 7);
    const int v_chunk_sel = lane >> 4;

    if constexpr (F8) {
      // ---------------------------------------------------------------- e4m3 cache
      // Half the bytes per page means half the time per page: the ldmatrix -> cvt -> mma -> softmax -> mma
      // chain of ONE page no longer hides under the stream with two warps per scheduler.  Pages are
      // therefore processed in PAIRS: both pages' S = Q K^T run as four independent accumulation chains,
      // one online-softmax update (one max, one rescale of O) covers the 32 tokens, then both P V.
      auto process = [&](auto np_tag, int i) {
        constexpr int NP = decltype(np_tag)::value;
        int st[NP], tok_base[NP];
        bool any = false;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
          st[p] = (i + p) % kNS;
          mbar_wait(&full_bar[st[p]], ((i + p) / kNS) & 1);
          tok_base[p] = (page0 + i + p) * kTok + t0;
          any = any || (tok_base[p] < ctx && tok_base[p] + 16 > win_lo);
        }
        if (any) {
          // ---- S = Q K^T: one 16-byte chunk of a token row = one whole k-step; matrices = (token half) x (chunk pair)
          float sc[NP][2][4];
#pragma unroll
          for (int p = 0; p < NP; ++p)
#pragma unroll
            for (int e = 0; e < 4; ++e) sc[p][0][e] = sc[p][1][e] = 0.f;
#pragma unroll
          for (int kp = 0; kp < D / 32; ++kp) {
            const int chunk = 2 * kp + k_chunk_sel;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              uint32_t r[4], b0, b1;
              ldsm_x4(r, smem_u32(smem + st[p] * SM::kStageBytes) + k_tok * 128 + ((chunk ^ (k_tok & 7)) << 4));
              e4m3x4_to_half2x2(r[0], b0, b1);
              mma_16816(sc[p][0], qa[2 * kp], b0, b1);
              e4m3x4_to_half2x2(r[1], b0, b1);
              mma_16816(sc[p][0], qa[2 * kp + 1], b0, b1);
              e4m3x4_to_half2x2(r[2], b0, b1);
              mma_16816(sc[p][1], qa[2 * kp], b0, b1);
              e4m3x4_to_half2x2(r[3], b0, b1);
              mma_16816(sc[p][1], qa[2 * kp + 1], b0, b1);
            }
          }
          // ---- online softmax over the NP x 16 tokens (rows g and, when G > 8, g + 8)
          uint32_t pa[NP][4];
#pragma unroll
          for (int p = 0; p < NP; ++p) pa[p][0] = pa[p][1] = pa[p][2] = pa[p][3] = 0u;
#pragma unroll
          for (int rh = 0; rh < 2; ++rh) {
            if (rh == 1 && !hi_rows) break;
            float t[NP][4];
            float mx = -INFINITY;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              t[p][0] = sc[p][0][2 * rh]; t[p][1] = sc[p][0][2 * rh + 1];
              t[p][2] = sc[p][1][2 * rh]; t[p][3] = sc[p][1][2 * rh + 1];
              const int tk[4] = {tok_base[p] + 2 * c, tok_base[p] + 2 * c + 1, tok_base[p] + 8 + 2 * c,
                                 tok_base[p] + 9 + 2 * c};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                float v = softcap > 0.f ? softcap * tanhf(t[p][e] * scale_nat * inv_cap) * LOG2E
                                        : t[p][e] * scale_log2;
                v = (tk[e] < ctx && tk[e] >= win_lo) ? v : -INFINITY;
                t[p][e] = v;
                mx = fmaxf(mx, v);
              }
            }
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
            mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
            const float m_new = fmaxf(m_run[rh], mx);
            const float base = (m_new == -INFINITY) ? 0.f : m_new;
            const float alpha = ex2_approx(m_run[rh] - base);
            float psum = 0.f;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
              float pe[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                pe[e] = ex2_approx(t[p][e] - base);
                psum += pe[e];
              }
              pa[p][rh] = pack_half2(pe[0], pe[1]);
              pa[p][2 + rh] = pack_half2(pe[2], pe[3]);
            }
            l_run[rh] = l_run[rh] * alpha + psum;
            m_run[rh] = m_new;
            // skip the 2 * D/8 multiplies while the running maximum stands (alpha == 1 exactly; the vote is
            // warp-uniform, so the result is bit-identical either way)
            if (!__all_sync(0xffffffffu, alpha == 1.f)) {
#pragma unroll
              for (int j = 0; j < D / 8; ++j) {
                o_acc[j][2 * rh] *= alpha;
                o_acc[j][2 * rh + 1] *= alpha;
              }
            }
          }
          // ---- O += P V.  Transposed 8x8 b16 matrices of the byte tile: lane (g, c) gets the byte pairs
          // (d = 2g, 2g+1) of tokens 2c and 2c+1; prmt gathers {tok 2c, 2c+1, 2c+8, 2c+9} for d even / d odd.
#pragma unroll
          for (int p = 0; p < NP; ++p) {
            const uint32_t vbase = smem_u32(smem + st[p] * SM::kStageBytes) + SM::kTileBytes;
            const bool partial = tok_base[p] + 16 > ctx;       // never let stale cache rows (possibly NaN) meet p == 0
            const uint32_t vmask8 = (tok_base[p] + 2 * c < ctx ? 0x000000FFu : 0u) |
                                    (tok_base[p] + 2 * c + 1 < ctx ? 0x0000FF00u : 0u) |
                                    (tok_base[p] + 8 + 2 * c < ctx ? 0x00FF0000u : 0u) |
                                    (tok_base[p] + 9 + 2 * c < ctx ? 0xFF000000u : 0u);
#pragma unroll
            for (int jp = 0; jp < D / 32; ++jp) {
              const int chunk = 2 * jp + v_chunk_sel;     // 16-byte chunk along d (0..7)
              uint32_t r[4];
              ldsm_x4_t(r, vbase + v_tok * 128 + ((chunk ^ (v_tok & 7)) << 4));
#pragma unroll
              for (int hc = 0; hc < 2; ++hc) {            // chunk 2 jp + hc: r[2 hc] = tokens 0-7, r[2 hc + 1] = 8-15
                uint32_t ev = __byte_perm(r[2 * hc], r[2 * hc + 1], 0x6420);
                uint32_t od = __byte_perm(r[2 * hc], r[2 * hc + 1], 0x7531);
                if (partial) {
                  ev &= vmask8;
                  od &= vmask8;
                }
                uint32_t b0, b1;
                e4m3x4_to_half2x2(ev, b0, b1);
                mma_16816(o_acc[4 * jp + 2 * hc], pa[p], b0, b1);        // columns d = 16 chunk + 2 n
                e4m3x4_to_half2x2(od, b0, b1);
                mma_16816(o_acc[4 * jp + 2 * hc + 1], pa[p], b0, b1);    // columns d = 16 chunk + 2 n + 1
              }
            }
          }
        }
        __syncwarp();
        if (lane == 0) {
#pragma unroll
          for (int p = 0; p < NP; ++p) mbar_arrive(&empty_bar[st[p]]);
        }
      };
      int i = 0;
      for (; i + 1 < n_cta_pages; i += 2) process(std::integral_constant<int, 2>{}, i);
      if (i < n_cta_pages) process(std::integral_constant<int, 1>{}, i);
    } else {
    for (int i = 0; i < n_cta_pages; ++i) {
      const int s = i % kNS;
      const uint32_t ph = (i / kNS) & 1;
      mbar_wait(&full_bar[s], ph);
      const int tok_base = (page0 + i) * kTok + t0;   // global index of this warp's first token
      if (tok_base < ctx && tok_base + 16 > win_lo) {
        const uint32_t kbase = smem_u32(smem + s * SM::kStageBytes);
        const uint32_t vbase = kbase + SM::kTileBytes;
        // ---- S = Q K^T for 16 tokens
        float s0[4] = {0.f, 0.f, 0.f, 0.f}, s1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int ks = 0; ks < D / 16; ++ks) {
          const int hb = ks >> 2;
          const int chunk = ((ks & 3) << 1) + k_chunk_sel;
          uint32_t r[4];
          if (L1) ldsm_x4(r, kbase + (2 * ks + k_chunk_sel) * (kTok * 16) + k_tok * 16);
          else ldsm_x4(r, kbase + hb * (kTok * 128) + k_tok * 128 + ((chunk ^ (k_tok & 7)) << 4));
          mma_16816_t<BF>(s0, qa[ks], r[0], r[1]);
          mma_16816_t<BF>(s1, qa[ks], r[2], r[3]);
        }
        // ---- online softmax (rows g and, when G > 8, g + 8)
        uint32_t pa[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int rh = 0; rh < 2; ++rh) {
          if (rh == 1 && !hi_rows) break;
          float t[4] = {s0[2 * rh], s0[2 * rh + 1], s1[2 * rh], s1[2 * rh + 1]};
          const int tk[4] = {tok_base + 2 * c, tok_base + 2 * c + 1, tok_base + 8 + 2 * c,
                             tok_base + 9 + 2 * c};
          float mx = -INFINITY;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float v = softcap > 0.f ? softcap * tanhf(t[e] * scale_nat * inv_cap) * LOG2E
                                    : t[e] * scale_log2;
            v = (tk[e] < ctx && tk[e] >= win_lo) ? v : -INFINITY;
            t[e] = v;
            mx = fmaxf(mx, v);
          }
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
          mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
          const float m_new = fmaxf(m_run[rh], mx);
          const float base = (m_new == -INFINITY) ? 0.f : m_new;
          const float alpha = exp2f(m_run[rh] - base);
          float p[4], psum = 0.f;
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            p[e] = exp2f(t[e] - base);
            psum += p[e];
          }
          l_run[rh] = l_run[rh] * alpha + psum;
          m_run[rh] = m_new;
#pragma unroll
          for (int j = 0; j < D / 8; ++j) {
            o_acc[j][2 * rh] *= alpha;
            o_acc[j][2 * rh + 1] *= alpha;
          }
          pa[rh] = pack16x2<BF>(p[0], p[1]);
          pa[2 + rh] = pack16x2<BF>(p[2], p[3]);
        }
        // ---- O += P V
        const bool partial = tok_base + 16 > ctx;
        const uint32_t vmask_lo = (tok_base + 2 * c < ctx ? 0x0000FFFFu : 0u) |
                                  (tok_base + 2 * c + 1 < ctx ? 0xFFFF0000u : 0u);
        const uint32_t vmask_hi = (tok_base + 8 + 2 * c < ctx ? 0x0000FFFFu : 0u) |
                                  (tok_base + 9 + 2 * c < ctx ? 0xFFFF0000u : 0u);
#pragma unroll
        for (int jj = 0; jj < D / 16; ++jj) {
          const int chunk = 2 * jj + v_chunk_sel;       // 16-byte chunk along d (0..D/8-1)
          const int hb = chunk >> 3, c16 = chunk & 7;
          uint32_t r[4];
          if (L1) {
            // matrix m = lane / 8: d-tile 2 jj + (m >> 1), token chunk (t0 / 8) + (m & 1); row = d, 16 B = 8 tokens
            const int dd = (2 * jj + (lane >> 4)) * 8 + (lane & 7);
            const int tch = (t0 >> 3) + ((lane >> 3) & 1);
            ldsm_x4(r, vbase + dd * 128 + ((tch ^ (lane & 7)) << 4));
          } else {
            ldsm_x4_t(r, vbase + hb * (kTok * 128) + v_tok * 128 + ((c16 ^ (v_tok & 7)) << 4));
          }
          if (partial) {   // never let stale cache rows (possibly NaN) meet p == 0
            r[0] &= vmask_lo; r[2] &= vmask_lo;
            r[1] &= vmask_hi; r[3] &= vmask_hi;
          }
          mma_16816_t<BF>(o_acc[2 * jj], pa, r[0], r[1]);
          mma_16816_t<BF>(o_acc[2 * jj + 1], pa, r[2], r[3]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty_bar[s]);
    }
    }
    // quad-reduce the row sums, publish (m, l) per warp
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      l_run[rh] += __shfl_xor_sync(0xffffffffu, l_run[rh], 1);
      l_run[rh] += __shfl_xor_sync(0xffffffffu, l_run[rh], 2);
      if (c == 0) {
        sm_ml[(warp * 16 + g + 8 * rh) * 2 + 0] = m_run[rh];
        sm_ml[(warp * 16 + g + 8 * rh) * 2 + 1] = l_run[rh];
      }
    }
  }
  __syncthreads();   // pipeline drained: stage memory is reusable

  float* sm_o = reinterpret_cast<float*>(smem);   // [warp][G][D]
  if (warp < kConsumerWarps) {
#pragma unroll
    for (int rh = 0; rh < 2; ++rh) {
      const int row = g + 8 * rh;
      if (row < G) {
        float m_tot = -INFINITY;
#pragma unroll
        for (int w = 0; w < kConsumerWarps; ++w) m_tot = fmaxf(m_tot, sm_ml[(w * 16 + row) * 2]);
        const float f = exp2f(m_run[rh] - m_tot);   // m_tot finite: the CTA saw >= 1 token
        if (F8) {
          // tiles 2 ch / 2 ch + 1 hold the even / odd columns of 16-wide chunk ch: d = 16 ch + 4c + {0..3}
#pragma unroll
          for (int ch = 0; ch < D / 16; ++ch) {
            float4 v = make_float4(o_acc[2 * ch][2 * rh] * f, o_acc[2 * ch + 1][2 * rh] * f,
                                   o_acc[2 * ch][2 * rh + 1] * f, o_acc[2 * ch + 1][2 * rh + 1] * f);
            *reinterpret_cast<float4*>(&sm_o[(warp * G + row) * D + ch * 16 + 4 * c]) = v;
          }
        } else {
#pragma unroll
          for (int j = 0; j < D / 8; ++j) {
            float2 v = make_float2(o_acc[j][2 * rh] * f, o_acc[j][2 * rh + 1] * f);
            *reinterpret_cast<float2*>(&sm_o[(warp * G + row) * D + j * 8 + 2 * c]) = v;
          }
        }
      }
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < G * D; idx += kThreads) {
    const int row = idx / D, d = idx - row * D;
    float m_tot = -INFINITY;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) m_tot = fmaxf(m_tot, sm_ml[(w * 16 + row) * 2]);
    float l_tot = 0.f, o = 0.f;
#pragma unroll
    for (int w = 0; w < kConsumerWarps; ++w) {
      l_tot += sm_ml[(w * 16 + row) * 2 + 1] * exp2f(sm_ml[(w * 16 + row) * 2] - m_tot);
      o += sm_o[(w * G + row) * D + d];
    }
    const int h = hkv * G + row;
    if (nsplits_seq == 1) {
      reinterpret_cast<uint16_t*>(out)[(static_cast<size_t>(b) * H + h) * D + d] = f32_to_bits16<BF>(o / l_tot);
    } else {
      const size_t pidx = (static_cast<size_t>(b) * H + h) * max_splits + split;
      part_o[pidx * D + d] = o;
      if (d == 0) {
        part_ml[pidx * 2 + 0] = m_tot;
        part_ml[pidx * 2 + 1] = l_tot;
      }
    }
  }
}

#define B200_ATTN_DECODE_PARAMS                                                                      \
  const __grid_constant__ CUtensorMap kmap, const __grid_constant__ CUtensorMap vmap,                \
      const __half *__restrict__ q, __half *__restrict__ out, const uint32_t *__restrict__ block_tables, \
      const uint32_t *__restrict__ context_lens, float *__restrict__ part_o, float *__restrict__ part_ml, \
      unsigned *__restrict__ counters, int q_stride, int H, int Hkv, int max_blocks, int pages_per_split, \
      int max_splits, float scale_log2, float softcap, int window
#define B200_ATTN_DECODE_ARGS                                                                         \
  kmap, vmap, q, out, block_tables, context_lens, part_o, part_ml, counters, q_stride, H, Hkv, max_blocks, \
      pages_per_split, max_splits, scale_log2, softcap, window

// 2 CTAs per SM (96 KB of stages each) for both cache formats
template <int D, bool F8, bool HI, bool BF, bool L1>
__global__ void __launch_bounds__(kThreads, 2) paged_attn_decode_kernel(B200_ATTN_DECODE_PARAMS) {
  paged_attn_decode_body<D, F8, HI, F8 ? 6 : kStages, BF, L1>(B200_ATTN_DECODE_ARGS);
}
// Split merge: one CTA per (head, sequence).  Launched with PDL right behind the main kernel, so
// its launch latency hides under the main kernel's tail (a fused last-CTA merge was measured
// slower: it serialises two L2 round trips behind the last wave).
template <int D, bool BF>
__global__ void __launch_bounds__(D)
paged_attn_reduce_kernel(__half* __restrict__ out, const float* __restrict__ part_o,
                         const float* __restrict__ part_ml,
                         const uint32_t* __restrict__ context_lens, int H, int pages_per_split,
                         int max_splits, int window) {
  pdl_trigger();
  pdl_wait();
  const int h = blockIdx.x, b = blockIdx.y, d = threadIdx.x;
  const int ctx = static_cast<int>(context_lens[b]);
  if (ctx <= 0) return;          // zero-filled by the main kernel
  const int npages = (ctx + kTok - 1) / kTok;
  const int fpage = first_page_of(ctx, window);
  const int ns = (npages - fpage + pages_per_split - 1) / pages_per_split;
  if (ns <= 1) return;           // written directly by the main kernel
  const size_t base = (static_cast<size_t>(b) * H + h) * max_splits;
  float m_tot = -INFINITY;
#pragma unroll 8
  for (int s = 0; s < ns; ++s) m_tot = fmaxf(m_tot, part_ml[(base + s) * 2]);
  float l = 0.f, acc = 0.f;
#pragma unroll 8
  for (int s = 0; s < ns; ++s) {            // fixed split order => deterministic
    const float f = exp2f(part_ml[(base + s) * 2] - m_tot);
    l += part_ml[(base + s) * 2 + 1] * f;
    acc += part_o[(base + s) * D + d] * f;
  }
  reinterpret_cast<uint16_t*>(out)[(static_cast<size_t>(b) * H + h) * D + d] = f32_to_bits16<BF>(acc / l);
}

struct SplitPlan {
  int pages_per_split, max_splits;
};
// Pages per split from a wave model: CTAs = batch * kv heads * ceil(max_blocks / p) run in
// waves of 2 CTAs per SM, a CTA costs p pages plus about one page of fixed work (prologue,
// partial store).  Picks the p that minimises waves * (p + 1): e.g. 32 x 8 x 65 pages on 148 SMs
// -> p = 9 (8 splits, 6.9 waves) rather than p = 10 (7 splits, 6.05 -> 7 waves).
static SplitPlan plan_splits(int batch, int num_kv_heads, int max_blocks, int ctas_per_sm = 2) {
  const long slots = static_cast<long>(sm_count()) * ctas_per_sm;
  const long heads = static_cast<long>(batch) * num_kv_heads;
  int hi = max_blocks < kMaxPagesPerSplit ? max_blocks : kMaxPagesPerSplit;
  if (hi < 1) hi = 1;
  int lo = hi < 2 ? hi : 2;
  // keep the partial buffers bounded: at most 64 splits per sequence
  while ((max_blocks + lo - 1) / lo > 64 && lo < hi) ++lo;
  long best_cost = -1;
  int best_p = lo;
  for (int p = lo; p <= hi; ++p) {
    const long splits = (max_blocks + p - 1) / p;
    const long ctas = heads * splits;
    const long waves = (ctas + slots - 1) / slots;
    const long cost = waves * (p + 1);
    if (best_cost < 0 || cost <= best_cost) {      // ties: fewer, larger splits
      best_cost = cost;
      best_p = p;
    }
  }
  {
    static int force = -1;             // B200_ATTN_PPS: experiment override of pages per split
    if (force < 0) { const char* e = getenv("B200_ATTN_PPS"); force = e ? atoi(e) : 0; }
    if (force > 0) best_p = force < hi ? force : hi;
  }
  SplitPlan r;
  r.pages_per_split = best_p;
  r.max_splits = (max_blocks + r.pages_per_split - 1) / r.pages_per_split;
  return r;
}

static bool make_cache_map(CUtensorMap* map, const void* cache, int num_blocks, int Hkv, int D, bool fp8) {
  const uint64_t es_bytes = fp8 ? 1 : 2;
  const uint64_t dims[4] = {static_cast<uint64_t>(D), static_cast<uint64_t>(Hkv), kTok,
                            static_cast<uint64_t>(num_blocks)};
  const uint64_t strides[3] = {static_cast<uint64_t>(D) * es_bytes, static_cast<uint64_t>(Hkv) * D * es_bytes,
                               static_cast<uint64_t>(kTok) * Hkv * D * es_bytes};
  const uint32_t box[4] = {fp8 ? 128u : 64u, 1, kTok, 1};      // 128-byte rows either way
  const uint32_t es[4] = {1, 1, 1, 1};
  return encode_tensor_map(map, fp8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                           const_cast<void*>(cache), dims, strides, box, es, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B) == CUDA_SUCCESS;
}

// L1 layout (see paged_attn_decode_body): K tile = 64*D*2 contiguous bytes viewed as rows of 256 elements, no swizzle;
// V tile = D rows of 64 tokens under the 128-byte swizzle.
static bool make_cache_maps_l1(AttnMaps* m, const void* k_cache, const void* v_cache, int num_blocks, int Hkv, int D) {
  const uint64_t tile = static_cast<uint64_t>(kTok) * D * 2;
  const uint32_t es[4] = {1, 1, 1, 1};
  {
    const uint64_t dims[4] = {256, static_cast<uint64_t>(D) / 4, static_cast<uint64_t>(Hkv),
                              static_cast<uint64_t>(num_blocks)};
    const uint64_t strides[3] = {512, tile, tile * Hkv};
    const uint32_t box[4] = {256, static_cast<uint32_t>(D) / 4, 1, 1};
    if (encode_tensor_map(&m->k, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(k_cache), dims, strides, box,
                          es, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B) != CUDA_SUCCESS)
      return false;
  }
  const uint64_t dims[4] = {kTok, static_cast<uint64_t>(D), static_cast<uint64_t>(Hkv),
                            static_cast<uint64_t>(num_blocks)};
  const uint64_t strides[3] = {kTok * 2, tile, tile * Hkv};
  const uint32_t box[4] = {kTok, static_cast<uint32_t>(D), 1, 1};
  return encode_tensor_map(&m->v, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(v_cache), dims, strides, box,
                           es, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B) == CUDA_SUCCESS;
}

bool make_attn_maps(AttnMaps* m, const void* k_cache, const void* v_cache, int num_blocks,
                    int Hkv, int D, bool fp8, bool l1) {
  m->fp8 = fp8;
  m->l1 = l1;
  if (fp8 && (D != 128 || l1)) return false;
  if (l1) return make_cache_maps_l1(m, k_cache, v_cache, num_blocks, Hkv, D);
  return make_cache_map(&m->k, k_cache, num_blocks, Hkv, D, fp8) &&
         make_cache_map(&m->v, v_cache, num_blocks, Hkv, D, fp8);
}

template <int D, bool F8 = false, bool HI = true, bool BF = false, bool L1 = false>
static void launch_decode(const AttnMaps& maps, void* out, const void* q, int q_stride,
                          const uint32_t* block_tables, const uint32_t* context_lens, int B,
                          int H, int Hkv, int max_blocks, float scale, float softcap, int window,
                          void* workspace, cudaStream_t st) {
  constexpr int NS = F8 ? 6 : kStages;
  using SM = AttnSmem<D, F8, NS>;
  auto kern = paged_attn_decode_kernel<D, F8, HI, BF, L1>;
  // resident CTAs per SM as the driver computes them (registers are allocated in per-warp units): the
  // wave model of the split planner must use the real number
  static int occ = 0;
  if (occ == 0) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kAlloc);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, SM::kAlloc) != cudaSuccess || occ < 1)
      occ = 1;
    if (getenv("B200_ATTN_DEBUG")) fprintf(stderr, "paged_attn_decode<%d,%d,%d>: %d CTAs/SM\n", D, F8, HI, occ);
  }
  const SplitPlan sp = plan_splits(B, Hkv, max_blocks, occ > 3 ? 3 : occ);
  // workspace: [arrival counters B*Hkv u32, zero at rest][partial O fp32][partial (m, l)]
  unsigned* counters = reinterpret_cast<unsigned*>(workspace);
  float* part_o = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) +
                                           ((static_cast<size_t>(B) * Hkv * 4 + 255) / 256) * 256);
  float* part_ml = part_o + static_cast<size_t>(B) * H * sp.max_splits * D;
  dim3 grid(sp.max_splits, Hkv, B);
  launch_pdl(kern, grid, dim3(kThreads), SM::kAlloc, st,
      maps.k, maps.v, reinterpret_cast<const __half*>(q), reinterpret_cast<__half*>(out),
      block_tables, context_lens, part_o, part_ml, counters, q_stride, H, Hkv, max_blocks,
      sp.pages_per_split, sp.max_splits, scale * 1.4426950408889634f, softcap, window);
  B200_CUDA_LAUNCH_CHECK();
  if (sp.max_splits > 1) {      // one split per sequence: every CTA wrote its fp16 rows itself
    launch_pdl(paged_attn_reduce_kernel<D, BF>, dim3(H, B), dim3(D), 0, st, reinterpret_cast<__half*>(out),
               static_cast<const float*>(part_o), static_cast<const float*>(part_ml), context_lens, H,
               sp.pages_per_split, sp.max_splits, window);
    B200_CUDA_LAUNCH_CHECK();
  }
}

void launch_paged_attn_decode(const AttnMaps& maps, void* out, const void* q, int q_stride,
                              const uint32_t* block_tables, const uint32_t* context_lens, int B,
                              int H, int Hkv, int D, int max_blocks, float scale, float softcap,
                              int window, void* workspace, cudaStream_t st, bool bf16) {
#define B200_DEC_ARGS maps, out, q, q_stride, block_tables, context_lens, B, H, Hkv, max_blocks, scale, softcap, window, \
                      workspace, st
  if (maps.l1) {
    if (D == 128 && bf16) launch_decode<128, false, true, true, true>(B200_DEC_ARGS);
    else if (D == 128) launch_decode<128, false, true, false, true>(B200_DEC_ARGS);
    else if (bf16) launch_decode<64, false, true, true, true>(B200_DEC_ARGS);
    else launch_decode<64, false, true, false, true>(B200_DEC_ARGS);
    return;
  }
#undef B200_DEC_ARGS
  if (bf16 && !maps.fp8 && D == 128)
    launch_decode<128, false, true, true>(maps, out, q, q_stride, block_tables, context_lens, B, H, Hkv,
                                          max_blocks, scale, softcap, window, workspace, st);
  else if (bf16 && !maps.fp8)
    launch_decode<64, false, true, true>(maps, out, q, q_stride, block_tables, context_lens, B, H, Hkv,
                                         max_blocks, scale, softcap, window, workspace, st);
  else if (maps.fp8 && H / Hkv <= 8)
    launch_decode<128, true, false>(maps, out, q, q_stride, block_tables, context_lens, B, H, Hkv, max_blocks,
                                    scale, softcap, window, workspace, st);
  else if (maps.fp8)
    launch_decode<128, true, true>(maps, out, q, q_stride, block_tables, context_lens, B, H, Hkv, max_blocks,
                                   scale, softcap, window, workspace, st);
  else if (D == 128)
    launch_decode<128>(maps, out, q, q_stride, block_tables, context_lens, B, H, Hkv, max_blocks,
                       scale, softcap, window, workspace, st);
  else
    launch_decode<64>(maps, out, q, q_stride, block_tables, context_lens, B, H, Hkv, max_blocks,
                      scale, softcap, window, workspace, st);
}

}  // namespace b200

extern "C" {

int64_t b200_paged_attn_workspace_bytes(int32_t batch, int32_t num_heads, int32_t head_dim,
                                        int32_t max_blocks_per_seq) {
  // worst case over kv-head counts: plan with Hkv = 1 gives the most splits per sequence
  int max_splits = 1;
  for (int hkv = 1; hkv <= num_heads; hkv <<= 1)
    for (int occ = 1; occ <= 3; ++occ) {       // whatever occupancy the launcher will find (it caps at 3)
      b200::SplitPlan sp = b200::plan_splits(batch, hkv, max_blocks_per_seq, occ);
      if (sp.max_splits > max_splits) max_splits = sp.max_splits;
    }
  return static_cast<int64_t>(batch) * num_heads * max_splits * (head_dim + 2) * 4 +
         static_cast<int64_t>(batch) * num_heads * 4 + 1024;
}

static void paged_attn_decode_any(void* out, const void* q, const void* k_cache,
                                const void* v_cache, const uint32_t* block_tables,
                                const uint32_t* context_lens, int32_t batch, int32_t num_heads,
                                int32_t num_kv_heads, int32_t head_dim, int32_t block_size,
                                int32_t max_blocks_per_seq, int32_t num_blocks, float scale,
                                float softcap, int32_t window, void* workspace,
                                int64_t stream, bool fp8, bool bf16 = false, bool l1 = false) {
  using namespace b200;
  B200_CHECK_ARG(!fp8 || head_dim == 128, "paged_attn_decode: the fp8 KV cache needs head_dim 128");
  B200_CHECK_ARG(block_size == kTok, "paged_attn_decode: block_size must be 64");
  B200_CHECK_ARG(head_dim == 64 || head_dim == 128, "paged_attn_decode: head_dim must be 64/128");
  B200_CHECK_ARG(num_kv_heads > 0 && num_heads % num_kv_heads == 0 &&
                     num_heads / num_kv_heads <= 16,
                 "paged_attn_decode: H/Hkv must be an integer <= 16");
  B200_CHECK_ARG(batch > 0 && max_blocks_per_seq > 0 && num_blocks > 0 && workspace != nullptr,
                 "paged_attn_decode: bad sizes or NULL workspace");
  AttnMaps maps;
  if (!make_attn_maps(&maps, k_cache, v_cache, num_blocks, num_kv_heads, head_dim, fp8, l1)) {
    set_error(2, "paged_attn_decode: cuTensorMapEncodeTiled failed");
    return;
  }
  launch_paged_attn_decode(maps, out, q, num_heads * head_dim, block_tables, context_lens, batch,
                           num_heads, num_kv_heads, head_dim, max_blocks_per_seq, scale, softcap,
                           window, workspace, reinterpret_cast<cudaStream_t>(stream), bf16);
}

void b200_paged_attn_decode_f16(void* out, const void* q, const void* k_cache,
                                const void* v_cache, const uint32_t* block_tables,
                                const uint32_t* context_lens, int32_t batch, int32_t num_heads,
                                int32_t num_kv_heads, int32_t head_dim, int32_t block_size,
                                int32_t max_blocks_per_seq, int32_t num_blocks, float scale,
                                float softcap, int32_t window, void* workspace,
                                int64_t stream) {
  paged_attn_decode_any(out, q, k_cache, v_cache, block_tables, context_lens, batch, num_heads,
                        num_kv_heads, head_dim, block_size, max_blocks_per_seq, num_blocks, scale, softcap,
                        window, workspace, stream, false);
}

void b200_paged_attn_decode_bf16(void* out, const void* q, const void* k_cache,
                                 const void* v_cache, const uint32_t* block_tables,
                                 const uint32_t* context_lens, int32_t batch, int32_t num_heads,
                                 int32_t num_kv_heads, int32_t head_dim, int32_t block_size,
                                 int32_t max_blocks_per_seq, int32_t num_blocks, float scale,
                                 float softcap, int32_t window, void* workspace,
                                 int64_t stream) {
  paged_attn_decode_any(out, q, k_cache, v_cache, block_tables, context_lens, batch, num_heads,
                        num_kv_heads, head_dim, block_size, max_blocks_per_seq, num_blocks, scale, softcap,
                        window, workspace, stream, false, true);
}

void b200_paged_attn_decode_l1(void* out, const void* q, const void* k_cache,
                               const void* v_cache, const uint32_t* block_tables,
                               const uint32_t* context_lens, int32_t batch, int32_t num_heads,
                               int32_t num_kv_heads, int32_t head_dim, int32_t block_size,
                               int32_t max_blocks_per_seq, int32_t num_blocks, float scale,
                               float softcap, int32_t window, void* workspace, int32_t bf16,
                               int64_t stream) {
  paged_attn_decode_any(out, q, k_cache, v_cache, block_tables, context_lens, batch, num_heads,
                        num_kv_heads, head_dim, block_size, max_blocks_per_seq, num_blocks, scale, softcap,
                        window, workspace, stream, false, bf16 != 0, true);
}

void b200_paged_attn_decode_fp8kv_f16(void* out, const void* q, const void* k_cache,
                                      const void* v_cache, const uint32_t* block_tables,
                                      const uint32_t* context_lens, int32_t batch, int32_t num_heads,
                                      int32_t num_kv_heads, int32_t head_dim, int32_t block_size,
                                      int32_t max_blocks_per_seq, int32_t num_blocks, float scale,
                                      float softcap, int32_t window, void* workspace,
                                      int64_t stream) {
  paged_attn_decode_any(out, q, k_cache, v_cache, block_tables, context_lens, batch, num_heads,
                        num_kv_heads, head_dim, block_size, max_blocks_per_seq, num_blocks, scale, softcap,
                        window, workspace, stream, true);
}

}  // extern "C"
