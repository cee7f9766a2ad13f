// Memory-bound glue kernels of the decoder layer: RoPE fused with the KV-cache write,
// RMSNorm (optionally fused with the residual add), SiLU*mul, embedding / row gather,
// bias add and greedy argmax.  All are HBM/L2-bound elementwise or row-reduction
// kernels: 128-bit loads, fp32 math, one rounding to fp16 on store.
//
// Reference call sites they replace:
//   FusedRope::apply_inplace          /root/reference/src/models/layers/rotary_emb.rs:88-104
//   cache write in PagedAttention     /root/reference/src/models/layers/attention.rs:808-820
//   NormX::forward / rms_norm         /root/reference/src/models/layers/others.rs:11-66
//   residual adds                     /root/reference/src/models/llama.rs:107-131
//   act(gate) * up                    /root/reference/src/models/layers/mlp.rs:451-469
//   embedding / index_select          /root/reference/src/models/llama.rs:269-310
//   logits.argmax                     /root/reference/src/utils/logits_processor.rs:67-70
#include "../../include/b200_vllm.h"
#include <cuda_fp8.h>
#include <cstdlib>

#include "attn.cuh"
#include "wna16.cuh"

namespace b200 {

// fp16 -> e4m3 byte of the FP8 KV cache: round to nearest even, saturate to +-448 (scale 1)
__device__ __forceinline__ uint8_t half_to_e4m3(__half x) {
  return static_cast<uint8_t>(__nv_cvt_halfraw_to_fp8(static_cast<__half_raw>(x), __NV_SATFINITE, __NV_E4M3));
}
__device__ __forceinline__ uint32_t half2x2_to_e4m3x4(uint32_t lo, uint32_t hi) {
  __half2_raw a, b;
  a.x = static_cast<unsigned short>(lo & 0xFFFFu); a.y = static_cast<unsigned short>(lo >> 16);
  b.x = static_cast<unsigned short>(hi & 0xFFFFu); b.y = static_cast<unsigned short>(hi >> 16);
  return static_cast<uint32_t>(__nv_cvt_halfraw2_to_fp8x2(a, __NV_SATFINITE, __NV_E4M3)) |
         (static_cast<uint32_t>(__nv_cvt_halfraw2_to_fp8x2(b, __NV_SATFINITE, __NV_E4M3)) << 16);
}

// Element index in the reference's default-build cache layout (kvcache_allocator.rs:150-173): slot = block * 64 +
// offset; K [num_blocks, Hkv, D/8, 64, 8], V [num_blocks, Hkv, D, 64].
__device__ __forceinline__ size_t l1_k_index(int64_t slot, int h, int d, int Hkv, int D) {
  const size_t blk = static_cast<size_t>(slot >> 6), off = static_cast<size_t>(slot & 63);
  return (((blk * Hkv + h) * (D >> 3) + (d >> 3)) * 64 + off) * 8 + (d & 7);
}
__device__ __forceinline__ size_t l1_v_index(int64_t slot, int h, int d, int Hkv, int D) {
  const size_t blk = static_cast<size_t>(slot >> 6), off = static_cast<size_t>(slot & 63);
  return ((blk * Hkv + h) * D + d) * 64 + off;
}

// ------------------------------------------------------------------ RoPE + KV write
// grid = tokens; one thread per rotation pair / per 2 copied elements.
// Cache layout [num_blocks, block_size, Hkv, D]: slot s -> row s of [.., Hkv, D].
template <bool kCosSinF32, bool BF>
__global__ void rope_kv_write_kernel(uint16_t* __restrict__ q, uint16_t* __restrict__ k,
                                     const uint16_t* __restrict__ v, const void* __restrict__ cosp,
                                     const void* __restrict__ sinp,
                                     const int64_t* __restrict__ positions,
                                     const int64_t* __restrict__ slot_mapping,
                                     uint16_t* __restrict__ k_cache, uint16_t* __restrict__ v_cache,
                                     int q_stride, int k_stride, int v_stride, int H, int Hkv,
                                     int D, int rot_dim, int interleaved, int kv_fp8, int l1) {
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.x;
  const int64_t pos = positions ? positions[t] : 0;
  const int half_rot = rot_dim >> 1;
  const int64_t slot = (slot_mapping && k_cache) ? slot_mapping[t] : -1;
  uint16_t* qt = q + static_cast<size_t>(t) * q_stride;
  uint16_t* kt = k + static_cast<size_t>(t) * k_stride;
  const uint16_t* vt = v ? v + static_cast<size_t>(t) * v_stride : nullptr;
  uint16_t* kc = slot >= 0 ? k_cache + static_cast<size_t>(slot) * Hkv * D : nullptr;
  uint16_t* vc = (slot >= 0 && v_cache) ? v_cache + static_cast<size_t>(slot) * Hkv * D : nullptr;
  // FP8 cache: same [slot][Hkv][D] indexing over one-byte elements
  uint8_t* kc8 = slot >= 0 ? reinterpret_cast<uint8_t*>(k_cache) + static_cast<size_t>(slot) * Hkv * D : nullptr;
  uint8_t* vc8 = (slot >= 0 && v_cache) ? reinterpret_cast<uint8_t*>(v_cache) + static_cast<size_t>(slot) * Hkv * D
                                        : nullptr;

  // rotation pairs of q and k
  const int pairs_per_head = D >> 1;
  const int total_pairs = (H + Hkv) * pairs_per_head;
  for (int idx = threadIdx.x; idx < total_pairs; idx += blockDim.x) {
    const int head = idx / pairs_per_head;
    const int p = idx - head * pairs_per_head;
    const bool is_q = head < H;
    uint16_t* base = is_q ? qt + head * D : kt + (head - H) * D;
    int i0, i1;
    bool rotate = p < half_rot;
    if (interleaved) {
      i0 = 2 * p;
      i1 = 2 * p + 1;
    } else if (rotate) {
      i0 = p;
      i1 = p + half_rot;
    } else {
      // un-rotated tail (partial rotary): elements rot_dim + 2*(p-half_rot) + {0,1}
      i0 = rot_dim + 2 * (p - half_rot);
      i1 = i0 + 1;
    }
    float x0 = bits16_to_f32<BF>(base[i0]), x1 = bits16_to_f32<BF>(base[i1]);
    if (rotate) {
      float cs, sn;
      if (kCosSinF32) {
        cs = reinterpret_cast<const float*>(cosp)[pos * half_rot + p];
        sn = reinterpret_cast<const float*>(sinp)[pos * half_rot + p];
      } else {
        cs = bits16_to_f32<BF>(reinterpret_cast<const uint16_t*>(cosp)[pos * half_rot + p]);
        sn = bits16_to_f32<BF>(reinterpret_cast<const uint16_t*>(sinp)[pos * half_rot + p]);
      }
      const float y0 = x0 * cs - x1 * sn;
      const float y1 = x1 * cs + x0 * sn;
      const uint16_t h0 = f32_to_bits16<BF>(y0), h1 = f32_to_bits16<BF>(y1);
      base[i0] = h0;
      base[i1] = h1;
      if (!is_q && kc) {
        if (kv_fp8) {
          kc8[(head - H) * D + i0] = half_to_e4m3(__ushort_as_half(h0));
          kc8[(head - H) * D + i1] = half_to_e4m3(__ushort_as_half(h1));
        } else if (l1) {
          k_cache[l1_k_index(slot, head - H, i0, Hkv, D)] = h0;
          k_cache[l1_k_index(slot, head - H, i1, Hkv, D)] = h1;
        } else {
          kc[(head - H) * D + i0] = h0;
          kc[(head - H) * D + i1] = h1;
        }
      }
    } else if (!is_q && kc) {
      if (kv_fp8) {
        kc8[(head - H) * D + i0] = half_to_e4m3(__ushort_as_half(base[i0]));
        kc8[(head - H) * D + i1] = half_to_e4m3(__ushort_as_half(base[i1]));
      } else if (l1) {
        k_cache[l1_k_index(slot, head - H, i0, Hkv, D)] = base[i0];
        k_cache[l1_k_index(slot, head - H, i1, Hkv, D)] = base[i1];
      } else {
        kc[(head - H) * D + i0] = base[i0];
        kc[(head - H) * D + i1] = base[i1];
      }
    }
  }
  if (l1) {                                        // V [nb, Hkv, D, 64]: one element per (head, d) row of the block
    if (vc && vt)
      for (int i = threadIdx.x; i < Hkv * D; i += blockDim.x) v_cache[l1_v_index(slot, i / D, i % D, Hkv, D)] = vt[i];
    return;
  }
  // V copy (8 halves per thread)
  if (vc && vt) {
    const int n16 = (Hkv * D) >> 3;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) {
      const uint4 x = reinterpret_cast<const uint4*>(vt)[i];
      if (kv_fp8)
        reinterpret_cast<uint2*>(vc8)[i] = make_uint2(half2x2_to_e4m3x4(x.x, x.y), half2x2_to_e4m3x4(x.z, x.w));
      else
        reinterpret_cast<uint4*>(vc)[i] = x;
    }
  }
}

// FP8 KV cache -> fp16 scratch pages for the prefill attention kernels (which read fp16 tiles): page i
// of sequence b (block table entry) lands in scratch block b * max_blocks + i, for K and V.
// grid = (max_blocks, B); ctx = context_lens[b], or the query length when context_lens is NULL.
__global__ void kv_fp8_gather_kernel(const uint8_t* __restrict__ k8, const uint8_t* __restrict__ v8,
                                     __half* __restrict__ ks, __half* __restrict__ vs,
                                     const uint32_t* __restrict__ block_tables,
                                     const uint32_t* __restrict__ context_lens,
                                     const uint32_t* __restrict__ cu_q, int max_blocks, int page_elems) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x, b = blockIdx.y;
  const int ctx = context_lens ? static_cast<int>(context_lens[b]) : static_cast<int>(cu_q[b + 1] - cu_q[b]);
  if (i * 64 >= ctx) return;
  const size_t src = static_cast<size_t>(block_tables[static_cast<size_t>(b) * max_blocks + i]) * page_elems;
  const size_t dst = (static_cast<size_t>(b) * max_blocks + i) * page_elems;
  for (int e = threadIdx.x * 8; e < page_elems; e += blockDim.x * 8) {
#pragma unroll
    for (int kv = 0; kv < 2; ++kv) {
      const uint2 x = *reinterpret_cast<const uint2*>((kv ? v8 : k8) + src + e);
      uint32_t o[4];
      const uint32_t w[2] = {x.x, x.y};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const __half2_raw a = __nv_cvt_fp8x2_to_halfraw2(static_cast<__nv_fp8x2_storage_t>(w[j] & 0xFFFFu), __NV_E4M3);
        const __half2_raw c = __nv_cvt_fp8x2_to_halfraw2(static_cast<__nv_fp8x2_storage_t>(w[j] >> 16), __NV_E4M3);
        o[2 * j] = static_cast<uint32_t>(a.x) | (static_cast<uint32_t>(a.y) << 16);
        o[2 * j + 1] = static_cast<uint32_t>(c.x) | (static_cast<uint32_t>(c.y) << 16);
      }
      *reinterpret_cast<uint4*>((kv ? vs : ks) + dst + e) = make_uint4(o[0], o[1], o[2], o[3]);
    }
  }
}
__global__ void scratch_block_table_kernel(uint32_t* __restrict__ bt, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) bt[i] = static_cast<uint32_t>(i);
}

// ------------------------------------------------------------------ cache::swap_blocks
// One launch moves every mapped KV block (runner.rs:1626-1645 CPU swap, transfer/mod.rs:260-311 PD transfer):
// grid = (16 KB pieces of a block, mapped pairs).  Either side may be pinned host memory (UVA: the kernel
// reads / writes it over PCIe / C2C directly) or a peer GPU's cache (peer access enabled): no per-block
// cudaMemcpy calls, no staging.
template <typename V>
__global__ void __launch_bounds__(256)
swap_blocks_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst,
                   const int64_t* __restrict__ pairs, int64_t block_bytes) {
  const int64_t sb = pairs[2 * blockIdx.y], db = pairs[2 * blockIdx.y + 1];
  const V* s = reinterpret_cast<const V*>(src + sb * block_bytes);
  V* d = reinterpret_cast<V*>(dst + db * block_bytes);
  const int64_t nv = block_bytes / static_cast<int64_t>(sizeof(V));
  const int64_t lo = static_cast<int64_t>(blockIdx.x) * 1024;
  const int64_t hi = lo + 1024 < nv ? lo + 1024 : nv;
  V v[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t i = lo + threadIdx.x + j * 256;
    if (i < hi) v[j] = s[i];
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t i = lo + threadIdx.x + j * 256;
    if (i < hi) d[i] = v[j];
  }
}

void launch_kv_fp8_gather(const void* k8, const void* v8, void* ks, void* vs, uint32_t* scratch_bt,
                          const uint32_t* block_tables, const uint32_t* context_lens, const uint32_t* cu_q,
                          int B, int max_blocks, int Hkv, int D, cudaStream_t st) {
  launch_pdl(kv_fp8_gather_kernel, dim3(max_blocks, B), dim3(256), 0, st, (const uint8_t*)k8, (const uint8_t*)v8,
             (__half*)ks, (__half*)vs, block_tables, context_lens, cu_q, max_blocks, 64 * Hkv * D);
  B200_CUDA_LAUNCH_CHECK();
  if (scratch_bt) {
    const int n = B * max_blocks;
    scratch_block_table_kernel<<<(n + 255) / 256, 256, 0, st>>>(scratch_bt, n);
    B200_CUDA_LAUNCH_CHECK();
  }
}

__global__ void kv_write_kernel(const __half* __restrict__ k, const __half* __restrict__ v,
                                __half* __restrict__ k_cache, __half* __restrict__ v_cache,
                                const int64_t* __restrict__ slot_mapping, int row_halfs, int Hkv, int D, int l1) {
  const int t = blockIdx.x;
  const int64_t slot = slot_mapping[t];
  if (slot < 0) return;
  if (l1) {
    // K: 16-byte chunks of 8 d-values keep together; V: one element per (head, d) row
    const uint4* ks = reinterpret_cast<const uint4*>(k + static_cast<size_t>(t) * row_halfs);
    for (int i = threadIdx.x; i < (row_halfs >> 3); i += blockDim.x) {
      const int h = (i * 8) / D, d = (i * 8) % D;
      *reinterpret_cast<uint4*>(k_cache + l1_k_index(slot, h, d, Hkv, D)) = ks[i];
    }
    const __half* vs = v + static_cast<size_t>(t) * row_halfs;
    for (int i = threadIdx.x; i < row_halfs; i += blockDim.x) v_cache[l1_v_index(slot, i / D, i % D, Hkv, D)] = vs[i];
    return;
  }
  const int n16 = row_halfs >> 3;
  const uint4* ks = reinterpret_cast<const uint4*>(k + static_cast<size_t>(t) * row_halfs);
  const uint4* vs = reinterpret_cast<const uint4*>(v + static_cast<size_t>(t) * row_halfs);
  uint4* kd = reinterpret_cast<uint4*>(k_cache + static_cast<size_t>(slot) * row_halfs);
  uint4* vd = reinterpret_cast<uint4*>(v_cache + static_cast<size_t>(slot) * row_halfs);
  for (int i = threadIdx.x; i < n16; i += blockDim.x) {
    kd[i] = ks[i];
    vd[i] = vs[i];
  }
}

// ------------------------------------------------------------------ RMSNorm
// One CTA per token.  kAdd: residual = fp16(residual + x) first (llama.rs:126-130
// performs the add in model dtype), then out = fp16(residual * rsqrt(mean sq + eps) * w).
template <bool kAdd, bool BF>
__global__ void __launch_bounds__(256)
rmsnorm_kernel(__half* __restrict__ out, __half* __restrict__ residual,
               const __half* __restrict__ x, const __half* __restrict__ w, int hidden,
               float eps) {
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.x;
  const int n8 = hidden >> 3;
  const uint4* xs = reinterpret_cast<const uint4*>(x + static_cast<size_t>(t) * hidden);
  uint4* rs = kAdd ? reinterpret_cast<uint4*>(residual + static_cast<size_t>(t) * hidden) : nullptr;
  uint4* os = reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * hidden);
  const uint4* ws = reinterpret_cast<const uint4*>(w);
  extern __shared__ uint4 row[];   // hidden halfs (kept on chip between the two passes)
  __shared__ float red[8];
  float ss = 0.f;
  for (int i = threadIdx.x; i < n8; i += blockDim.x) {
    uint4 a = xs[i];
    if (kAdd) {
      uint4 r = rs[i];
      a.x = add16x2<BF>(a.x, r.x);
      a.y = add16x2<BF>(a.y, r.y);
      a.z = add16x2<BF>(a.z, r.z);
      a.w = add16x2<BF>(a.w, r.w);
      rs[i] = a;
    }
    row[i] = a;
    const uint32_t* ah = reinterpret_cast<const uint32_t*>(&a);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack16x2<BF>(ah[j]);
      ss += f.x * f.x + f.y * f.y;
    }
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) tot += red[i];
  const float inv = rsqrtf(tot / static_cast<float>(hidden) + eps);
  for (int i = threadIdx.x; i < n8; i += blockDim.x) {
    uint4 a = row[i];
    const uint4 wv = ws[i];
    uint32_t* ah = reinterpret_cast<uint32_t*>(&a);
    const uint32_t* wh = reinterpret_cast<const uint32_t*>(&wv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = unpack16x2<BF>(ah[j]);
      const float2 g = unpack16x2<BF>(wh[j]);
      ah[j] = pack16x2<BF>(f.x * inv * g.x, f.y * inv * g.y);
    }
    os[i] = a;
  }
}

// ------------------------------------------------------------------ folded-norm row scale
// out[t, :] = x[t, :] * rsqrt(ssq[t] / hidden + eps): finishes an RMSNorm whose weight multiply and
// sum of squares were fused into the producing GEMM's epilogue (model_runner.cpp, fused plan).
__global__ void rowscale_kernel(__half* __restrict__ out, const __half* __restrict__ x,
                                const float* __restrict__ ssq, int hidden, float inv_dim, float eps) {
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.x;
  // per-tile slots summed in the finalisers' order (wna16.cuh: ssq_lane_partial + warp_sum), by every warp
  const float rs = rsqrtf(warp_sum(ssq_lane_partial(ssq + t * kSsqStride, hidden / kTileN, threadIdx.x & 31)) *
                              inv_dim + eps);
  const uint4* xs = reinterpret_cast<const uint4*>(x + static_cast<size_t>(t) * hidden);
  uint4* os = reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * hidden);
  for (int i = threadIdx.x; i < (hidden >> 3); i += blockDim.x) {
    uint4 a = xs[i];
    __half2* ah = reinterpret_cast<__half2*>(&a);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(ah[j]);
      ah[j] = __floats2half2_rn(f.x * rs, f.y * rs);
    }
    os[i] = a;
  }
}

// ------------------------------------------------------------------ SiLU(gate) * up
template <bool BF>
__global__ void silu_mul_kernel(__half* __restrict__ out, const __half* __restrict__ gate_up,
                                int inter) {
  const int t = blockIdx.x;                              // tokens on grid.x: a prefill step may hold > 65535 of them
  const int i = blockIdx.y * blockDim.x + threadIdx.x;   // index of an 8-half vector
  if (i * 8 >= inter) return;
  const uint4 gv = reinterpret_cast<const uint4*>(gate_up + static_cast<size_t>(t) * 2 * inter)[i];
  const uint4 uv =
      reinterpret_cast<const uint4*>(gate_up + static_cast<size_t>(t) * 2 * inter + inter)[i];
  uint4 o;
  const uint32_t* gh = reinterpret_cast<const uint32_t*>(&gv);
  const uint32_t* uh = reinterpret_cast<const uint32_t*>(&uv);
  uint32_t* oh = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 g = unpack16x2<BF>(gh[j]);
    const float2 u = unpack16x2<BF>(uh[j]);
    oh[j] = pack16x2<BF>(g.x / (1.f + __expf(-g.x)) * u.x, g.y / (1.f + __expf(-g.y)) * u.y);
  }
  reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * inter)[i] = o;
}

// ------------------------------------------------------------------ row gathers
__global__ void gather_rows_kernel(__half* __restrict__ out, const __half* __restrict__ table,
                                   const uint32_t* __restrict__ ids, int hidden, int nrows_src) {
  pdl_trigger();
  pdl_wait();
  const int t = blockIdx.x;
  uint32_t id = ids[t];
  if (nrows_src > 0 && id >= static_cast<uint32_t>(nrows_src)) id = 0;
  const uint4* src = reinterpret_cast<const uint4*>(table + static_cast<size_t>(id) * hidden);
  uint4* dst = reinterpret_cast<uint4*>(out + static_cast<size_t>(t) * hidden);
  for (int i = threadIdx.x; i < (hidden >> 3); i += blockDim.x) dst[i] = src[i];
}

__global__ void add_bias_kernel(__half* __restrict__ x, const __half* __restrict__ bias, int n) {
  const int t = blockIdx.x;
  const int i = blockIdx.y * blockDim.x + threadIdx.x;
  if (i * 2 >= n) return;
  __half2* p = reinterpret_cast<__half2*>(x + static_cast<size_t>(t) * n) + i;
  *p = __hadd2(*p, reinterpret_cast<const __half2*>(bias)[i]);
}

// ------------------------------------------------------------------ argmax (first max wins)
__global__ void __launch_bounds__(1024)
argmax_kernel(uint32_t* __restrict__ out_ids, const float* __restrict__ logits, int vocab) {
  pdl_trigger();
  pdl_wait();
  const int b = blockIdx.x;
  const float* row = logits + static_cast<size_t>(b) * vocab;
  float best = -INFINITY;
  int best_i = 0x7fffffff;
  const int n4 = vocab >> 2;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(row)[i];
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (e[j] > best) {   // ascending index within a thread: strict > keeps the first
        best = e[j];
        best_i = i * 4 + j;
      }
  }
  for (int i = n4 * 4 + threadIdx.x; i < vocab; i += blockDim.x)
    if (row[i] > best || (row[i] == best && i < best_i)) {
      best = row[i];
      best_i = i;
    }
  __shared__ float sv[32];
  __shared__ int si[32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
    if (ov > best || (ov == best && oi < best_i)) {
      best = ov;
      best_i = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    sv[threadIdx.x >> 5] = best;
    si[threadIdx.x >> 5] = best_i;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    best = threadIdx.x < nw ? sv[threadIdx.x] : -INFINITY;
    best_i = threadIdx.x < nw ? si[threadIdx.x] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ov > best || (ov == best && oi < best_i)) {
        best = ov;
        best_i = oi;
      }
    }
    if (threadIdx.x == 0) out_ids[b] = best_i == 0x7fffffff ? 0u : static_cast<uint32_t>(best_i);
  }
}

void launch_rowscale(void* out, const void* x, const float* ssq, int T, int hidden, float eps,
                     cudaStream_t st) {
  launch_pdl(rowscale_kernel, dim3(T), dim3(256), 0, st, (__half*)out, (const __half*)x, ssq, hidden,
             1.f / static_cast<float>(hidden), eps);
  B200_CUDA_LAUNCH_CHECK();
}

template <bool F32, bool BF>
static void launch_rope_kv_write_t(void* q, void* k, const void* v, int q_stride, int k_stride, int v_stride,
                                   const void* cos, const void* sin, const int64_t* positions,
                                   const int64_t* slot_mapping, void* k_cache, void* v_cache, int T, int H, int Hkv,
                                   int D, int rot_dim, int interleaved, cudaStream_t st, int kv_fp8, int l1) {
  const int pairs = (H + Hkv) * D / 2;
  const int threads = pairs >= 512 ? 512 : (pairs >= 256 ? 256 : 128);
  launch_pdl(rope_kv_write_kernel<F32, BF>, dim3(T), dim3(threads), 0, st, (uint16_t*)q, (uint16_t*)k,
             (const uint16_t*)v, cos, sin, positions, slot_mapping, (uint16_t*)k_cache, (uint16_t*)v_cache,
             q_stride, k_stride, v_stride, H, Hkv, D, rot_dim, interleaved, kv_fp8, l1);
  B200_CUDA_LAUNCH_CHECK();
}

void launch_rope_kv_write(void* q, void* k, const void* v, int q_stride, int k_stride,
                          int v_stride, const void* cos, const void* sin,
                          const int64_t* positions, const int64_t* slot_mapping, void* k_cache,
                          void* v_cache, int T, int H, int Hkv, int D, int rot_dim,
                          int interleaved, int cos_sin_f32, cudaStream_t st, int kv_fp8, bool bf16, int l1) {
#define B200_ROPE_ARGS q, k, v, q_stride, k_stride, v_stride, cos, sin, positions, slot_mapping, k_cache, v_cache, T, \
                       H, Hkv, D, rot_dim, interleaved, st, kv_fp8, l1
  if ((bf16 || l1) && kv_fp8) {
    set_error(1, "rope_kv_write: the e4m3 cache write is fp16 / flash layout only");
    return;
  }
  if (bf16) {
    if (cos_sin_f32) launch_rope_kv_write_t<true, true>(B200_ROPE_ARGS);
    else launch_rope_kv_write_t<false, true>(B200_ROPE_ARGS);
  } else {
    if (cos_sin_f32) launch_rope_kv_write_t<true, false>(B200_ROPE_ARGS);
    else launch_rope_kv_write_t<false, false>(B200_ROPE_ARGS);
  }
#undef B200_ROPE_ARGS
}

template <bool kAdd, bool BF>
static void rmsnorm_any(const char* what, void* out, void* residual, const void* x, const void* weight,
                        int32_t num_tokens, int32_t hidden, float eps, int64_t stream) {
  if (num_tokens <= 0) return;
  if (hidden % 8 != 0 || hidden > 16384) {
    set_error(1, what);
    return;
  }
  launch_pdl(rmsnorm_kernel<kAdd, BF>, dim3(num_tokens), dim3(256), hidden * 2,
             reinterpret_cast<cudaStream_t>(stream), (__half*)out, (__half*)residual, (const __half*)x,
             (const __half*)weight, hidden, eps);
  B200_CUDA_LAUNCH_CHECK();
}

template <bool BF>
static void silu_mul_any(void* out, const void* gate_up, int32_t num_tokens, int32_t inter, int64_t stream) {
  if (num_tokens <= 0) return;
  B200_CHECK_ARG(inter % 8 == 0, "silu_mul: inter % 8 != 0");
  dim3 grid(num_tokens, (inter / 8 + 255) / 256);
  silu_mul_kernel<BF><<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (__half*)out, (const __half*)gate_up, inter);
  B200_CUDA_LAUNCH_CHECK();
}
}  // namespace b200

extern "C" {
using namespace b200;

void b200_rope_kv_write_f16(void* q, void* k, const void* v, const void* cos, const void* sin,
                            const int64_t* positions, const int64_t* slot_mapping,
                            void* k_cache, void* v_cache, int32_t num_tokens, int32_t num_heads,
                            int32_t num_kv_heads, int32_t head_dim, int32_t rot_dim,
                            int32_t block_size, int32_t interleaved, int32_t cos_sin_f32,
                            int64_t stream) {
  (void)block_size;
  if (num_tokens <= 0) return;
  B200_CHECK_ARG(head_dim % 8 == 0 && rot_dim % 2 == 0 && rot_dim <= head_dim && rot_dim > 0,
                 "rope_kv_write: bad head_dim / rot_dim");
  B200_CHECK_ARG(q && k && cos && sin && positions, "rope_kv_write: NULL pointer");
  launch_rope_kv_write(q, k, v, num_heads * head_dim, num_kv_heads * head_dim,
                       num_kv_heads * head_dim, cos, sin, positions, slot_mapping, k_cache,
                       v_cache, num_tokens, num_heads, num_kv_heads, head_dim, rot_dim,
                       interleaved, cos_sin_f32, reinterpret_cast<cudaStream_t>(stream));
}

void b200_rope_kv_write_bf16(void* q, void* k, const void* v, const void* cos, const void* sin,
                             const int64_t* positions, const int64_t* slot_mapping,
                             void* k_cache, void* v_cache, int32_t num_tokens, int32_t num_heads,
                             int32_t num_kv_heads, int32_t head_dim, int32_t rot_dim,
                             int32_t block_size, int32_t interleaved, int32_t cos_sin_f32,
                             int64_t stream) {
  (void)block_size;
  if (num_tokens <= 0) return;
  B200_CHECK_ARG(head_dim % 8 == 0 && rot_dim % 2 == 0 && rot_dim <= head_dim && rot_dim > 0,
                 "rope_kv_write: bad head_dim / rot_dim");
  B200_CHECK_ARG(q && k && cos && sin && positions, "rope_kv_write: NULL pointer");
  launch_rope_kv_write(q, k, v, num_heads * head_dim, num_kv_heads * head_dim,
                       num_kv_heads * head_dim, cos, sin, positions, slot_mapping, k_cache,
                       v_cache, num_tokens, num_heads, num_kv_heads, head_dim, rot_dim,
                       interleaved, cos_sin_f32, reinterpret_cast<cudaStream_t>(stream), 0, true);
}

void b200_rope_kv_write_fp8kv_f16(void* q, void* k, const void* v, const void* cos, const void* sin,
                                  const int64_t* positions, const int64_t* slot_mapping,
                                  void* k_cache, void* v_cache, int32_t num_tokens, int32_t num_heads,
                                  int32_t num_kv_heads, int32_t head_dim, int32_t rot_dim,
                                  int32_t block_size, int32_t interleaved, int32_t cos_sin_f32,
                                  int64_t stream) {
  (void)block_size;
  if (num_tokens <= 0) return;
  B200_CHECK_ARG(head_dim % 8 == 0 && rot_dim % 2 == 0 && rot_dim <= head_dim && rot_dim > 0,
                 "rope_kv_write: bad head_dim / rot_dim");
  B200_CHECK_ARG(q && k && cos && sin && positions, "rope_kv_write: NULL pointer");
  launch_rope_kv_write(q, k, v, num_heads * head_dim, num_kv_heads * head_dim,
                       num_kv_heads * head_dim, cos, sin, positions, slot_mapping, k_cache,
                       v_cache, num_tokens, num_heads, num_kv_heads, head_dim, rot_dim,
                       interleaved, cos_sin_f32, reinterpret_cast<cudaStream_t>(stream), 1);
}

void b200_swap_blocks(const void* src_cache, void* dst_cache, const int64_t* pairs_dev, int32_t num_pairs,
                      int64_t block_bytes, int64_t stream) {
  if (num_pairs <= 0) return;
  B200_CHECK_ARG(src_cache && dst_cache && pairs_dev, "swap_blocks: NULL pointer");
  B200_CHECK_ARG(block_bytes > 0 && block_bytes % 16 == 0, "swap_blocks: block bytes must be a multiple of 16");
  B200_CHECK_ARG(num_pairs <= 65535, "swap_blocks: at most 65535 blocks per call");
  // 16-byte accesses when both sides allow it; the payloads of a PD wire frame sit at 8 (mod 16) -- the bincode
  // length fields in front of them are 8 bytes -- and take the 8-byte flavour
  const uintptr_t both = reinterpret_cast<uintptr_t>(src_cache) | reinterpret_cast<uintptr_t>(dst_cache);
  B200_CHECK_ARG((both & 7) == 0, "swap_blocks: caches must be 8-byte aligned");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if ((both & 15) == 0) {
    const unsigned pieces = static_cast<unsigned>((block_bytes / 16 + 1023) / 1024);
    swap_blocks_kernel<uint4><<<dim3(pieces, num_pairs), 256, 0, st>>>(
        static_cast<const uint8_t*>(src_cache), static_cast<uint8_t*>(dst_cache), pairs_dev, block_bytes);
  } else {
    const unsigned pieces = static_cast<unsigned>((block_bytes / 8 + 1023) / 1024);
    swap_blocks_kernel<uint2><<<dim3(pieces, num_pairs), 256, 0, st>>>(
        static_cast<const uint8_t*>(src_cache), static_cast<uint8_t*>(dst_cache), pairs_dev, block_bytes);
  }
  B200_CUDA_LAUNCH_CHECK();
}

void b200_kv_write_f16(const void* k, const void* v, void* k_cache, void* v_cache,
                       const int64_t* slot_mapping, int32_t num_tokens, int32_t num_kv_heads,
                       int32_t head_dim, int32_t block_size, int64_t stream) {
  (void)block_size;
  if (num_tokens <= 0) return;
  B200_CHECK_ARG((num_kv_heads * head_dim) % 8 == 0, "kv_write: row must be a multiple of 8");
  kv_write_kernel<<<num_tokens, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __half*)k, (const __half*)v, (__half*)k_cache, (__half*)v_cache, slot_mapping,
      num_kv_heads * head_dim, num_kv_heads, head_dim, 0);
  B200_CUDA_LAUNCH_CHECK();
}

void b200_kv_write_l1(const void* k, const void* v, void* k_cache, void* v_cache,
                      const int64_t* slot_mapping, int32_t num_tokens, int32_t num_kv_heads,
                      int32_t head_dim, int32_t block_size, int64_t stream) {
  if (num_tokens <= 0) return;
  B200_CHECK_ARG(block_size == 64 && head_dim % 8 == 0, "kv_write_l1: block_size must be 64, head_dim a multiple of 8");
  kv_write_kernel<<<num_tokens, 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const __half*)k, (const __half*)v, (__half*)k_cache, (__half*)v_cache, slot_mapping,
      num_kv_heads * head_dim, num_kv_heads, head_dim, 1);
  B200_CUDA_LAUNCH_CHECK();
}

void b200_rope_kv_write_l1(void* q, void* k, const void* v, const void* cos, const void* sin,
                           const int64_t* positions, const int64_t* slot_mapping,
                           void* k_cache, void* v_cache, int32_t num_tokens, int32_t num_heads,
                           int32_t num_kv_heads, int32_t head_dim, int32_t rot_dim,
                           int32_t block_size, int32_t interleaved, int32_t cos_sin_f32, int32_t bf16,
                           int64_t stream) {
  if (num_tokens <= 0) return;
  B200_CHECK_ARG(block_size == 64, "rope_kv_write_l1: block_size must be 64");
  B200_CHECK_ARG(head_dim % 8 == 0 && rot_dim % 2 == 0 && rot_dim <= head_dim && rot_dim > 0,
                 "rope_kv_write: bad head_dim / rot_dim");
  B200_CHECK_ARG(q && k && cos && sin && positions, "rope_kv_write: NULL pointer");
  launch_rope_kv_write(q, k, v, num_heads * head_dim, num_kv_heads * head_dim,
                       num_kv_heads * head_dim, cos, sin, positions, slot_mapping, k_cache,
                       v_cache, num_tokens, num_heads, num_kv_heads, head_dim, rot_dim,
                       interleaved, cos_sin_f32, reinterpret_cast<cudaStream_t>(stream), 0, bf16 != 0, 1);
}

void b200_rmsnorm_f16(void* out, const void* x, const void* weight, int32_t num_tokens,
                      int32_t hidden, float eps, int64_t stream) {
  rmsnorm_any<false, false>("rmsnorm: hidden % 8 != 0 or > 16384", out, nullptr, x, weight, num_tokens, hidden, eps,
                            stream);
}
void b200_rmsnorm_bf16(void* out, const void* x, const void* weight, int32_t num_tokens,
                       int32_t hidden, float eps, int64_t stream) {
  rmsnorm_any<false, true>("rmsnorm: hidden % 8 != 0 or > 16384", out, nullptr, x, weight, num_tokens, hidden, eps,
                           stream);
}

void b200_add_rmsnorm_f16(void* out, void* residual, const void* x, const void* weight,
                          int32_t num_tokens, int32_t hidden, float eps, int64_t stream) {
  rmsnorm_any<true, false>("add_rmsnorm: hidden % 8 != 0 or > 16384", out, residual, x, weight, num_tokens, hidden,
                           eps, stream);
}
void b200_add_rmsnorm_bf16(void* out, void* residual, const void* x, const void* weight,
                           int32_t num_tokens, int32_t hidden, float eps, int64_t stream) {
  rmsnorm_any<true, true>("add_rmsnorm: hidden % 8 != 0 or > 16384", out, residual, x, weight, num_tokens, hidden,
                          eps, stream);
}

void b200_silu_mul_f16(void* out, const void* gate_up, int32_t num_tokens, int32_t inter,
                       int64_t stream) {
  silu_mul_any<false>(out, gate_up, num_tokens, inter, stream);
}
void b200_silu_mul_bf16(void* out, const void* gate_up, int32_t num_tokens, int32_t inter,
                        int64_t stream) {
  silu_mul_any<true>(out, gate_up, num_tokens, inter, stream);
}

void b200_embedding_f16(void* out, const void* table, const uint32_t* ids, int32_t num_tokens,
                        int32_t hidden, int32_t vocab, int64_t stream) {
  if (num_tokens <= 0) return;
  B200_CHECK_ARG(hidden % 8 == 0, "embedding: hidden % 8 != 0");
  launch_pdl(gather_rows_kernel, dim3(num_tokens), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
             (__half*)out, (const __half*)table, ids, hidden, vocab);
  B200_CUDA_LAUNCH_CHECK();
}

void b200_gather_rows_f16(void* out, const void* x, const uint32_t* rows, int32_t num_rows,
                          int32_t hidden, int64_t stream) {
  if (num_rows <= 0) return;
  B200_CHECK_ARG(hidden % 8 == 0, "gather_rows: hidden % 8 != 0");
  launch_pdl(gather_rows_kernel, dim3(num_rows), dim3(256), 0, reinterpret_cast<cudaStream_t>(stream),
             (__half*)out, (const __half*)x, rows, hidden, 0);
  B200_CUDA_LAUNCH_CHECK();
}

void b200_add_bias_f16(void* x, const void* bias, int32_t num_tokens, int32_t n,
                       int64_t stream) {
  if (num_tokens <= 0) return;
  B200_CHECK_ARG(n % 2 == 0, "add_bias: n % 2 != 0");
  dim3 grid(num_tokens, (n / 2 + 255) / 256);
  add_bias_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (__half*)x, (const __half*)bias, n);
  B200_CUDA_LAUNCH_CHECK();
}

void b200_argmax_f32(uint32_t* out_ids, const float* logits, int32_t batch, int32_t vocab,
                     int64_t stream) {
  if (batch <= 0) return;
  B200_CHECK_ARG(vocab > 0 && (reinterpret_cast<uintptr_t>(logits) & 15) == 0 &&
                     (vocab % 4 == 0 || batch == 1),
                 "argmax: logits rows must be 16-byte aligned");
  launch_pdl(argmax_kernel, dim3(batch), dim3(1024), 0, reinterpret_cast<cudaStream_t>(stream), out_ids,
             logits, vocab);
  B200_CUDA_LAUNCH_CHECK();
}

}  // extern "C"
