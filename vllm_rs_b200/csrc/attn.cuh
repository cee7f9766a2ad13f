// Internal launch interface of the paged-attention decode kernel (tensor maps prebuilt
// once per layer by the model runner; the C ABI wrapper builds them per call).
#pragma once
#include "common.cuh"

namespace b200 {
struct AttnMaps {
  CUtensorMap k, v;
  bool fp8 = false;      // e4m3 cache (one byte per element, head_dim 128): decode kernel only
  bool l1 = false;       // the reference's default-build layout: K [nb, Hkv, D/8, 64, 8], V [nb, Hkv, D, 64]
};
bool make_attn_maps(AttnMaps* m, const void* k_cache, const void* v_cache, int num_blocks,
                    int Hkv, int D, bool fp8 = false, bool l1 = false);
void launch_paged_attn_decode(const AttnMaps& maps, void* out, const void* q, int q_stride,
                              const uint32_t* block_tables, const uint32_t* context_lens, int B,
                              int H, int Hkv, int D, int max_blocks, float scale, float softcap,
                              int window, void* workspace, cudaStream_t st, bool bf16 = false);
void launch_attn_prefill(const AttnMaps& maps, void* out, const void* q, int q_stride,
                         const uint32_t* cu_q, const uint32_t* ctx_lens,
                         const uint32_t* block_tables, int B, int H, int Hkv, int D, int max_blocks,
                         int max_seqlen_q, float scale, cudaStream_t st, int total_rows = 0, int mode = -1, bool bf16 = false);
// tcgen05 version (head_dim 128); false when the shape / mode is not taken.  total_rows = rows of the q
// buffer (bounds of its tensor map); mode -1 = env B200_PREFILL_ATTN (default tcgen05), 0 = mma.sync.
bool launch_attn_prefill_tc(const AttnMaps& maps, void* out, const void* q, int q_stride, const uint32_t* cu_q,
                            const uint32_t* ctx_lens, const uint32_t* block_tables, int B, int H, int Hkv, int D,
                            int max_blocks, int max_seqlen_q, int total_rows, float scale, int mode,
                            cudaStream_t st, bool bf16 = false);
// RoPE + KV write with explicit row strides (in halfs) so q|k|v can live in one fused
// GEMM output row.
void launch_rope_kv_write(void* q, void* k, const void* v, int q_stride, int k_stride,
                          int v_stride, const void* cos, const void* sin,
                          const int64_t* positions, const int64_t* slot_mapping, void* k_cache,
                          void* v_cache, int T, int H, int Hkv, int D, int rot_dim,
                          int interleaved, int cos_sin_f32, cudaStream_t st, int kv_fp8 = 0, bool bf16 = false, int l1 = 0);
// FP8 cache pages of every sequence's context -> fp16 scratch blocks b * max_blocks + i (+ that identity
// block table when scratch_bt != NULL): what the prefill attention kernels read in --fp8-kvcache mode.
void launch_kv_fp8_gather(const void* k8, const void* v8, void* ks, void* vs, uint32_t* scratch_bt,
                          const uint32_t* block_tables, const uint32_t* context_lens, const uint32_t* cu_q,
                          int B, int max_blocks, int Hkv, int D, cudaStream_t st);
void launch_rowscale(void* out, const void* x, const float* ssq, int T, int hidden, float eps,
                     cudaStream_t st);
}  // namespace b200
