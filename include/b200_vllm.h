/*
 * b200_vllm.h -- C ABI of the B200-native (sm_100a) forward hot path for vllm.rs.
 *
 * Drop-in boundary.  Every symbol below replaces one call the
 * reference crate makes into its un-vendored native dependency `attention-rs`
 * (or composes several of them into one fused launch).  Conventions are the
 * reference's own (src/utils/gptq.rs:100-200, 314-333):
 *   - raw device pointers, dims as int32, CUDA stream passed as int64 (CUstream cast);
 *   - the caller owns every buffer; outputs may be uninitialised on entry and are
 *     fully written; the callee never allocates, frees, synchronises or retains
 *     pointers (=> CUDA-graph capturable); scratch comes in through an explicit
 *     `workspace` pointer sized by the matching *_workspace_bytes() query;
 *   - `void` return like the reference FFI; failures are latched and can be read
 *     with b200_last_error()/b200_last_error_string() (0 = none).
 *   - index dtypes are the reference's: slot_mapping / positions int64,
 *     block_tables / context_lens / cu_seqlens uint32 (src/core/runner.rs:1243-1388).
 *
 * Native KV-cache layout (reference layout "L2", src/utils/kvcache_allocator.rs:150-173):
 *   K and V, per layer:  [num_blocks, block_size, num_kv_heads, head_dim]  fp16.
 *
 * Native quantised-weight layout ("B200 tile-major", produced by the prepack
 * entry points, same byte count as the checkpoint tensor so it can live in the
 * buffer the reference allocates for its Marlin repack, src/utils/gptq.rs:286-298):
 *   4-bit: u32 [N/128][K/64][2][128][4]  -- for out-channel tile nt, K chunk kc,
 *          half h, channel c, word w: the eight K values k = kc*64 + h*32 + w*8 + j,
 *          nibble p (p<4) holds j=2p, nibble p+4 holds j=2p+1.
 *   8-bit: u32 [N/128][K/64][4][128][4]  -- quarter q: k = kc*64 + q*16 + w*4 + j,
 *          byte j of the word holds K offset j (natural order).
 *   scales stay [K/G, N] fp16 and qzeros stay in checkpoint packing (AWQ or GPTQ).
 */
#ifndef B200_VLLM_H
#define B200_VLLM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ errors */
int b200_last_error(void);                 /* 0 = ok, 1 = bad argument, 2 = CUDA error */
const char* b200_last_error_string(void);
void b200_clear_error(void);
int b200_abi_version(void);
/* kernels launched by this library since load (every launch site counts itself) */
int64_t b200_launch_count(void);

/* quantisation formats (checkpoint packing of qzeros / zero-point semantics) */
#define B200_QFMT_AWQ 0       /* zero = nibble (order 0,2,4,6,1,3,5,7), no +1           */
#define B200_QFMT_GPTQ 1      /* zero = field + 1 (sequential order); NULL qzeros => sym */

/* epilogue flags of b200_wna16_gemm_ex */
#define B200_EPI_NONE 0
#define B200_EPI_SILU_MUL 1   /* out[m, n] = silu(acc[m, n]) * acc[m, n + N/2], out is [M, N/2] */

/* ------------------------------------------------- quantised weight prepack
 * Replaces attention_rs::kernels::ffi::awq_repack (gptq.rs:316-324):
 *   in  : AWQ qweight  i32 [k, n_packed]   (n = n_packed * 32/bits)
 *   out : B200 tile-major layout, k*n_packed u32.                              */
void b200_awq_prepack(const void* in, void* out, int32_t k, int32_t n_packed, int32_t bits,
                      int64_t stream);
/* Replaces attention_rs::kernels::ffi::gptq_repack (gptq.rs:325-333):
 *   in  : GPTQ qweight i32 [k_packed, n]   (k = k_packed * 32/bits)            */
void b200_gptq_prepack(const void* in, void* out, int32_t k_packed, int32_t n, int32_t bits,
                       int64_t stream);

/* ------------------------------------------------- W4A16 / W8A16 GEMM
 * C[m, n] (fp16) = A[m, k] (fp16) x dequant(B)  with fp32 accumulation.
 * Replace marlin_awq_4bit_f16 / marlin_4bit_f16 (gptq.rs:118-146) and
 * gemm_half_q_half_alt (gptq.rs:183-195); same argument order as those.
 * `workspace`: b200_wna16_workspace_bytes() bytes, zero-initialised once by the
 * caller (the reference already allocates a zeroed workspace, wna16.rs:238-241);
 * the kernels leave it zeroed on exit.                                          */
void b200_marlin_awq_4bit_f16(const void* a, const int32_t* b, const void* scales,
                              const void* qzeros, const void* g_idx, void* c, int32_t m,
                              int32_t k, int32_t n, const void* workspace, int32_t group_size,
                              int64_t stream);
void b200_marlin_4bit_f16(const void* a, const int32_t* b, const void* scales,
                          const void* qzeros, const void* g_idx, void* c, int32_t m, int32_t k,
                          int32_t n, const void* workspace, int32_t group_size, int64_t stream);
/* bf16 flavour (replace marlin_awq_4bit_bf16 / marlin_4bit_bf16, gptq.rs:150-178; bf16 is the reference's
 * default dtype, utils/mod.rs:1919-1925): a, scales and c are bf16, same argument order and workspace.  Needs
 * k % 128 == 0 and group_size = 128 * 2^j (or <= 0 / >= k: one group per channel).  Decode-shaped calls
 * (m <= 32) are one launch of the stream-K tcgen05 kernel; larger m runs 32 rows per launch.            */
void b200_marlin_awq_4bit_bf16(const void* a, const int32_t* b, const void* scales,
                               const void* qzeros, const void* g_idx, void* c, int32_t m,
                               int32_t k, int32_t n, const void* workspace, int32_t group_size,
                               int64_t stream);
void b200_marlin_4bit_bf16(const void* a, const int32_t* b, const void* scales,
                           const void* qzeros, const void* g_idx, void* c, int32_t m, int32_t k,
                           int32_t n, const void* workspace, int32_t group_size, int64_t stream);
/* 8-bit GPTQ path (reference: gemm_half_q_half_alt, bits == 8).  Unlike the reference
 * symbol this takes the prepacked weight, an explicit group_size and a workspace.  */
void b200_gemm_half_q_half_8bit(const void* a, const uint32_t* qweight, const uint32_t* qzeros,
                                const void* scales, const int32_t* g_idx, void* c, int32_t m,
                                int32_t n, int32_t k, int32_t group_size, const void* workspace,
                                int64_t stream);
/* gemm_half_q_half_alt with the reference's EXACT signature (gptq.rs:183-195 call site, 316-324 declaration):
 * `qweight` is the checkpoint-format GPTQ tensor [k / (32/bits), n] as the loader leaves it on this path, no
 * group size, no workspace.  The group size is derived from g_idx (required; one cached 8-byte read), the
 * tile-major copy of the weight is made on first use and kept, the workspace is library-owned.  The first call
 * per tensor synchronises the stream; later calls do not.  b200_gemm_half_q_half_alt_release() frees all of it. */
void b200_gemm_half_q_half_alt(const void* a, const uint32_t* qweight, const uint32_t* qzeros,
                               const void* scales, const int32_t* g_idx, void* c, int32_t m, int32_t n,
                               int32_t k, int32_t bits, int64_t stream);
void b200_gemm_half_q_half_alt_release(void);
/* General entry: bits in {4,8}, qfmt AWQ/GPTQ, optional bias[n] (fp16, may be NULL),
 * epilogue flag.  variant: 0 = auto by m, 1 = flat GEMV (CUDA cores), 2 = tcgen05 (round-1 decode kernel for m <= 64),
 * 3 = tcgen05 with a single K split (direct store, workspace untouched; prefill),
 * 4 = the general tcgen05 kernel even for decode shapes (A/B testing),
 * 5 = the stream-K tcgen05 decode kernel (m <= 32, k % 128 == 0, group_size = 128 * 2^j). */
void b200_wna16_gemm_ex(const void* a, const void* b_packed, const void* scales,
                        const void* qzeros, const void* bias, void* c, int32_t m, int32_t k,
                        int32_t n, int32_t bits, int32_t group_size, int32_t qfmt,
                        int32_t epilogue, int32_t variant, void* workspace, int64_t stream);
int64_t b200_wna16_workspace_bytes(int32_t max_m, int32_t max_n);
/* Dequantise a prepacked weight to fp16 out[n, k] (row = output channel, K-major): the weight
 * operand of b200_dense_gemm_f16.  Large prefill steps dequantise a layer once and run the dense
 * tcgen05 GEMM instead of re-dequantising per token tile (no reference counterpart: Marlin
 * re-dequantises, gptq.rs:118-146).                                                      */
void b200_wna16_dequant_f16(const void* b_packed, const void* scales, const void* qzeros, void* out,
                            int32_t k, int32_t n, int32_t bits, int32_t group_size, int32_t qfmt,
                            int64_t stream);

/* ------------------------------------------------- dense fp16 GEMM (lm_head, fp16 linears)
 * C[m, n] = A[m, k] x W[n, k]^T ; out_f32 != 0 writes fp32 logits (llama.rs:311-320).
 * Replaces Linear::forward -> cuBLAS (linear.rs:75-123) with a tcgen05 kernel (TMA-fed SS-mode
 * MMA, CTA pairs for m > 128, weight-streaming swapped form for m <= 128).               */
void b200_dense_gemm_f16(const void* a, const void* w, void* c, int32_t m, int32_t n, int32_t k,
                         int32_t out_f32, int64_t stream);
/* Same with an optional fused bias[n] (fp16, may be NULL) and an explicit implementation:
 * mode -1 = default (env B200_DENSE), 0 = cuBLAS (the reference's own call, A/B baseline),
 * 1 = tcgen05 single-CTA tiles, 2 = tcgen05 CTA pairs.                                   */
void b200_dense_gemm_ex(const void* a, const void* w, const void* bias, void* c, int32_t m,
                        int32_t n, int32_t k, int32_t out_f32, int32_t mode, int64_t stream);

/* ------------------------------------------------- RoPE + KV-cache write
 * Replaces FusedRope::apply_inplace (rotary_emb.rs:88-104) followed by the cache
 * write inside PagedAttention::forward (attention.rs:808-820) in ONE launch.
 * q [T,H,D], k/v [T,Hkv,D] fp16, rotated in place (q,k); cos/sin [max_pos, rot_dim/2]
 * (fp16 or, when cos_sin_f32 != 0, fp32); positions int64 [T]; slot_mapping int64 [T]
 * (slot = block*block_size + offset; negative = skip the cache write).
 * k_cache/v_cache may be NULL (RoPE only).  interleaved = GPT-J style pairs.       */
void b200_rope_kv_write_f16(void* q, void* k, const void* v, const void* cos, const void* sin,
                            const int64_t* positions, const int64_t* slot_mapping,
                            void* k_cache, void* v_cache, int32_t num_tokens, int32_t num_heads,
                            int32_t num_kv_heads, int32_t head_dim, int32_t rot_dim,
                            int32_t block_size, int32_t interleaved, int32_t cos_sin_f32,
                            int64_t stream);
/* bf16 flavour (the reference's default dtype, utils/mod.rs:1919-1925): q / k / v, the caches and -- unless
 * cos_sin_f32 -- the cos / sin tables are bf16; same arithmetic (fp32 rotation, one rounding on store).            */
void b200_rope_kv_write_bf16(void* q, void* k, const void* v, const void* cos, const void* sin,
                             const int64_t* positions, const int64_t* slot_mapping,
                             void* k_cache, void* v_cache, int32_t num_tokens, int32_t num_heads,
                             int32_t num_kv_heads, int32_t head_dim, int32_t rot_dim,
                             int32_t block_size, int32_t interleaved, int32_t cos_sin_f32,
                             int64_t stream);
/* Same with an FP8 KV cache (`--fp8-kvcache`; kvcache_allocator.rs:188-193,776 allocates the caches as
 * U8 in the same [num_blocks, block_size, Hkv, D] shape): the rotated fp16 K row and the V row are
 * stored as e4m3 bytes (round to nearest even, saturating at +-448, scale 1 -- attention-rs's own
 * format is not source-visible, so this is the library's definition).  q / k are still rotated in
 * place as fp16.                                                                                   */
void b200_rope_kv_write_fp8kv_f16(void* q, void* k, const void* v, const void* cos, const void* sin,
                                  const int64_t* positions, const int64_t* slot_mapping,
                                  void* k_cache, void* v_cache, int32_t num_tokens, int32_t num_heads,
                                  int32_t num_kv_heads, int32_t head_dim, int32_t rot_dim,
                                  int32_t block_size, int32_t interleaved, int32_t cos_sin_f32,
                                  int64_t stream);
/* Cache write only (reshape_and_cache equivalent). */
void b200_kv_write_f16(const void* k, const void* v, void* k_cache, void* v_cache,
                       const int64_t* slot_mapping, int32_t num_tokens, int32_t num_kv_heads,
                       int32_t head_dim, int32_t block_size, int64_t stream);

/* attention_rs::cache::swap_blocks as called for CPU swap (runner.rs:1626-1645: once per layer for K and for
 * V, mapping = HashMap<src block, dst block>) and PD transfer (transfer/mod.rs:260-311): copies whole cache
 * blocks src[pairs[i][0]] -> dst[pairs[i][1]] in ONE launch.  block_bytes = bytes of one block of this cache
 * (block_size * Hkv * D * element size; multiple of 16).  pairs_dev: int64 [num_pairs][2] in DEVICE memory.
 * src / dst: device memory, pinned host memory (read / written by the kernel through UVA) or a peer GPU's
 * cache with peer access enabled.  Blocks must not overlap between src and dst ranges of one call.        */
void b200_swap_blocks(const void* src_cache, void* dst_cache, const int64_t* pairs_dev, int32_t num_pairs,
                      int64_t block_bytes, int64_t stream);

/* ------------------------------------------------- PD (prefill / decode) KV hand-off: wire frames + block movement
 * transfer/mod.rs:82-125 (KVTransferHandle / FinishedPrefillData / TransferMessage), comm.rs:384-410 (u32 LE length +
 * bincode 1.3.1 body) and the data legs of transfer/mod.rs:260-311,370-470.  A KV frame is built in place in a PINNED
 * host buffer: _frame_begin writes every field but the payloads and returns where layer l's K and V bytes go;
 * _export_blocks lets the copy kernel write them there from the paged caches (pairs = (server block, i));
 * _frame_parse gives the same offsets of a received frame and _import_blocks copies (i, local block) out of it.
 * _copy_blocks_peer is the LocalIpc leg: (server block, local block) between two GPUs' caches over NVLink.      */
typedef struct b200_pd_kv_header {
  uint64_t seq_id;
  uint32_t first_token;
  int32_t num_layers;
  uint64_t layer_bytes;        /* bytes of ONE layer's K (= V) payload: num_blocks * block bytes */
  uint64_t num_blocks;
  uint64_t sending_time;       /* ms since the epoch (mod.rs:383-386) */
  uint64_t num_cached_tokens;
} b200_pd_kv_header;
int64_t b200_pd_kv_frame_bytes(int32_t num_layers, int64_t layer_bytes);
int64_t b200_pd_kv_frame_begin(uint8_t* frame, int64_t cap, const b200_pd_kv_header* h, int64_t* k_off,
                               int64_t* v_off);
int32_t b200_pd_kv_frame_parse(const uint8_t* frame, int64_t len, b200_pd_kv_header* h, int64_t* k_off,
                               int64_t* v_off, int32_t max_layers);
/* The LocalIpc flavour of the same message (transfer/mod.rs:77,84-90): per layer a (K, V) pair of
 * CudaIpcMemHandle(64 handle bytes, tensor shape, SerializableDType: U8 0, U32 1, I64 2, BF16 3, F16 4, F32 5, F64 6)
 * and the server's block ids of the sequence (h->num_blocks of them; h->layer_bytes is unused).  k_handles / v_handles
 * are [num_layers][64] bytes as cudaIpcGetMemHandle returned them; every layer shares one shape and dtype.           */
int64_t b200_pd_ipc_frame(uint8_t* out, int64_t cap, const b200_pd_kv_header* h, const uint8_t* k_handles,
                          const uint8_t* v_handles, const uint64_t* shape, int32_t ndim, int32_t dtype,
                          const uint32_t* server_block_ids);
int32_t b200_pd_ipc_frame_parse(const uint8_t* frame, int64_t len, b200_pd_kv_header* h, uint8_t* k_handles,
                                uint8_t* v_handles, int32_t max_layers, uint64_t* shape, int32_t* ndim,
                                int32_t* dtype, uint32_t* server_block_ids, int32_t max_blocks);
/* kind 1 = AvailableTokenResponse(value), 3 = ReleaseKvCache(value); returns the frame length (16) */
int64_t b200_pd_frame_small(uint8_t* out, int64_t cap, int32_t kind, uint64_t value);
/* TransferMessage variant of a frame (0..3), -1 if malformed; *value for kinds 1 and 3 */
int32_t b200_pd_frame_kind(const uint8_t* frame, int64_t len, uint64_t* value);
void b200_pd_export_blocks(const void* const* k_caches, const void* const* v_caches, int32_t num_layers,
                           const int64_t* pairs_dev, int32_t n, int64_t block_bytes, uint8_t* frame,
                           const int64_t* k_off, const int64_t* v_off, int64_t stream);
void b200_pd_import_blocks(void* const* k_caches, void* const* v_caches, int32_t num_layers,
                           const int64_t* pairs_dev, int32_t n, int64_t block_bytes, const uint8_t* frame,
                           const int64_t* k_off, const int64_t* v_off, int64_t stream);
void b200_pd_copy_blocks_peer(const void* const* remote_k, const void* const* remote_v, void* const* local_k,
                              void* const* local_v, int32_t num_layers, const int64_t* pairs_dev, int32_t n,
                              int64_t block_bytes, int64_t stream);
int32_t b200_enable_peer_access(int32_t dev, int32_t peer);

/* ------------------------------------------------- engine <-> runner-process frames of the decode step
 * runner/mod.rs:168-295 (MessageType; u32 LE length + bincode 1.3.1; every frame is answered by ONE ack byte 0x01),
 * core/sequence.rs:52-83 (DecodeSequence), utils/config.rs:504-537 (SamplingParams).  A runner process built on this
 * library reads RunDecode, derives the step metadata (core/runner.rs:1243-1264), runs b200_runner_decode_host and
 * answers RunResponse; KVCacheSwap yields the (src, dst) pairs b200_swap_blocks takes.  MessageType variant indices:
 * RunDecode 4, RunResponse 5, FinishDecode 8, Heartbeat 14, KVCacheSwap 21, KVCacheSwapResponse 22, ClearBlocks 31,
 * ClearBlocksResponse 32, Shutdown 34.  Init / RunPrefill / RunEmbed carry whole Sequence / EngineConfig structs and
 * stay with the Rust host.                                                                                          */
typedef struct b200_wire_sampling {          /* the numeric fields of SamplingParams; has_* = Option is Some */
  int32_t has_temperature, has_max_tokens, ignore_eos, has_top_k, has_top_p, has_frequency_penalty,
      has_presence_penalty, reserved;
  float temperature, top_p, frequency_penalty, presence_penalty;
  int64_t max_tokens, top_k;
} b200_wire_sampling;
int32_t b200_wire_kind(const uint8_t* frame, int64_t len);
int32_t b200_wire_parse_run_decode(const uint8_t* frame, int64_t len, int32_t max_seqs, int32_t max_blocks,
                                   uint64_t* seq_ids, uint32_t* last_tokens, uint64_t* lens,
                                   uint64_t* last_block_tokens, uint32_t* block_table_last, uint32_t* block_tables,
                                   int32_t* num_blocks, b200_wire_sampling* sampling, int32_t* is_prefill);
int64_t b200_wire_frame_run_decode(uint8_t* out, int64_t cap, int32_t n, const uint64_t* seq_ids,
                                   const uint32_t* last_tokens, const uint64_t* lens,
                                   const uint64_t* last_block_tokens, const uint32_t* block_table_last,
                                   const uint32_t* block_tables, int32_t max_blocks, const int32_t* num_blocks,
                                   const b200_wire_sampling* sampling, int32_t is_prefill);
int64_t b200_wire_frame_run_response(uint8_t* out, int64_t cap, const uint32_t* ids, int32_t n);
int32_t b200_wire_parse_run_response(const uint8_t* frame, int64_t len, uint32_t* ids, int32_t cap);
int32_t b200_wire_parse_kv_swap(const uint8_t* frame, int64_t len, int64_t* pairs, int32_t cap, int32_t* swap_in);
int64_t b200_wire_frame_kv_swap(uint8_t* out, int64_t cap, const int64_t* pairs, int32_t n, int32_t swap_in);
int64_t b200_wire_frame_simple(uint8_t* out, int64_t cap, int32_t kind, int32_t payload_bytes, uint64_t value);
/* RunPrefill((Vec<Sequence>, bool)) (variant 3; Sequence as core/sequence.rs:32-50).  The engine side frames the
 * sequences `indices` of its scheduler; the runner side loads such a frame into a scheduler object used as a plain
 * sequence container (its running list is REPLACED; indices[0..n) = 0..n-1) and calls b200_prepare_prefill on it.
 * Fields the mirror does not keep travel as the reference's defaults (created_time 0, the Option fields None) and are
 * skipped on read; first_sampling (may be NULL) receives the SamplingParams of the first sequence, which pick the
 * step's strategy (core/runner.rs:1390-1570).  load returns n, or -1 malformed, -2 more than cap sequences, -3 a
 * sequence prepare_prefill cannot walk (other block size, short block table), -4 a sequence with images.           */
struct b200_scheduler;
int64_t b200_scheduler_frame_run_prefill(const struct b200_scheduler* s, const int32_t* indices, int32_t n,
                                         uint8_t* out, int64_t cap, int32_t is_prefill);
int32_t b200_scheduler_load_run_prefill(struct b200_scheduler* s, const uint8_t* frame, int64_t len,
                                        int32_t* indices, int32_t cap, int32_t* is_prefill,
                                        b200_wire_sampling* first_sampling);
/* Single-Sequence messages of the PD hand-off.  family 0 = runner IPC MessageType (runner/mod.rs:168-249): kind 15
 * TransferPrefill(Sequence), 18 ReceivePrefillResponse((bool, Option<Sequence>)), 23 KvCacheSend((Sequence, u32)),
 * 25 KvCacheReceive(Sequence); family 1 = client <-> PD server TransferMessage (transfer/mod.rs:113-122): kind 0
 * TransferPrefill(Sequence).  `extra` = the bool of kind 18 / the first token of kind 23; index < 0 with kind 18 frames
 * (extra, None).  load puts the sequence into the container as running[0]; returns 1, 0 (kind 18 with None), -1, -4.  */
int64_t b200_scheduler_frame_sequence(const struct b200_scheduler* s, int32_t index, int32_t family, int32_t kind,
                                      uint32_t extra, uint8_t* out, int64_t cap);
int32_t b200_scheduler_load_sequence(struct b200_scheduler* s, const uint8_t* frame, int64_t len, int32_t family,
                                     int32_t* kind, uint32_t* extra);
/* KvCacheReceiveResponse((bool, u32, usize, usize)) (variant 26): success, first token, the server's sending time
 * and the prefix-cache hit length recorded there (transfer/mod.rs:319).  parse returns 0, -1 malformed.            */
int64_t b200_wire_frame_kv_receive_response(uint8_t* out, int64_t cap, int32_t ok, uint32_t first_token, uint64_t sending_time,
                                            uint64_t num_cached_tokens);
int32_t b200_wire_parse_kv_receive_response(const uint8_t* frame, int64_t len, int32_t* ok, uint32_t* first_token,
                                            uint64_t* sending_time, uint64_t* num_cached_tokens);
void b200_wire_decode_metadata(int32_t n, const uint32_t* last_tokens, const uint64_t* lens,
                               const uint64_t* last_block_tokens, const uint32_t* block_table_last,
                               int32_t block_size, uint32_t* input_ids, int64_t* positions, int64_t* slot_mapping,
                               uint32_t* context_lens);

/* ------------------------------------------------- the reference's default-build cache layout ("L1")
 * kvcache_allocator.rs:150-173: without the flashattn / flashinfer features the caches are
 *   K [num_blocks, Hkv, D/x, block_size, x] (x = 16 / element size = 8),   V [num_blocks, Hkv, D, block_size].
 * These entry points read / write that layout directly (block_size 64, 16-bit elements; bf16 != 0 selects the bf16
 * arithmetic): a (block, head) tile is 64*D*2 contiguous bytes, K feeds the MMAs as it lies (8 x 16-byte core
 * matrices), V as a K-major operand.  Argument meaning as for the flash-layout functions of the same name.        */
void b200_kv_write_l1(const void* k, const void* v, void* k_cache, void* v_cache,
                      const int64_t* slot_mapping, int32_t num_tokens, int32_t num_kv_heads,
                      int32_t head_dim, int32_t block_size, int64_t stream);
void b200_rope_kv_write_l1(void* q, void* k, const void* v, const void* cos, const void* sin,
                           const int64_t* positions, const int64_t* slot_mapping,
                           void* k_cache, void* v_cache, int32_t num_tokens, int32_t num_heads,
                           int32_t num_kv_heads, int32_t head_dim, int32_t rot_dim,
                           int32_t block_size, int32_t interleaved, int32_t cos_sin_f32, int32_t bf16,
                           int64_t stream);
void b200_paged_attn_decode_l1(void* out, const void* q, const void* k_cache,
                               const void* v_cache, const uint32_t* block_tables,
                               const uint32_t* context_lens, int32_t batch, int32_t num_heads,
                               int32_t num_kv_heads, int32_t head_dim, int32_t block_size,
                               int32_t max_blocks_per_seq, int32_t num_blocks, float scale,
                               float softcap, int32_t window, void* workspace, int32_t bf16, int64_t stream);
void b200_attn_prefill_l1(void* out, const void* q, const void* k_cache, const void* v_cache,
                          const uint32_t* cu_seqlens_q, const uint32_t* context_lens,
                          const uint32_t* block_tables, int32_t batch, int32_t num_heads,
                          int32_t num_kv_heads, int32_t head_dim, int32_t block_size,
                          int32_t max_blocks_per_seq, int32_t num_blocks, int32_t max_seqlen_q,
                          int32_t num_tokens, int32_t mode, float scale, int32_t bf16, int64_t stream);

/* ------------------------------------------------- paged attention, decode
 * Replaces the decode branch of PagedAttention::forward (attention.rs:808-820).
 * out/q [B,H,D] fp16; block_tables u32 [B, max_blocks_per_seq] (zero padded);
 * context_lens u32 [B] (0 => row skipped, out row zeroed).  softcap <= 0 disables;
 * window <= 0 disables sliding window.  head_dim in {64,128}, block_size == 64,
 * H/Hkv <= 16.  Per-sequence lengths are read on the device (graph safe).  `workspace`:
 * b200_paged_attn_workspace_bytes() bytes, zeroed once by the caller (arrival counters;
 * the kernel leaves them zeroed).                                                    */
void b200_paged_attn_decode_f16(void* out, const void* q, const void* k_cache,
                                const void* v_cache, const uint32_t* block_tables,
                                const uint32_t* context_lens, int32_t batch, int32_t num_heads,
                                int32_t num_kv_heads, int32_t head_dim, int32_t block_size,
                                int32_t max_blocks_per_seq, int32_t num_blocks, float scale,
                                float softcap, int32_t window, void* workspace, int64_t stream);
/* bf16 flavour: q / out and both caches are bf16; S = Q K^T and P V run as bf16 x bf16 -> fp32 MMAs (P rounded
 * to bf16), softmax in fp32.  Same workspace.                                                                      */
void b200_paged_attn_decode_bf16(void* out, const void* q, const void* k_cache,
                                 const void* v_cache, const uint32_t* block_tables,
                                 const uint32_t* context_lens, int32_t batch, int32_t num_heads,
                                 int32_t num_kv_heads, int32_t head_dim, int32_t block_size,
                                 int32_t max_blocks_per_seq, int32_t num_blocks, float scale,
                                 float softcap, int32_t window, void* workspace, int64_t stream);
int64_t b200_paged_attn_workspace_bytes(int32_t batch, int32_t num_heads, int32_t head_dim,
                                        int32_t max_blocks_per_seq);
/* Decode attention over an FP8 (e4m3, scale 1) KV cache: same arguments, k_cache / v_cache are
 * u8 [num_blocks, 64, Hkv, 128]; half the HBM bytes per step.  head_dim must be 128.             */
void b200_paged_attn_decode_fp8kv_f16(void* out, const void* q, const void* k_cache,
                                      const void* v_cache, const uint32_t* block_tables,
                                      const uint32_t* context_lens, int32_t batch, int32_t num_heads,
                                      int32_t num_kv_heads, int32_t head_dim, int32_t block_size,
                                      int32_t max_blocks_per_seq, int32_t num_blocks, float scale,
                                      float softcap, int32_t window, void* workspace, int64_t stream);

/* ------------------------------------------------- paged attention, prefill (varlen causal)
 * Replaces the prefill branch of PagedAttention::forward (attention.rs:808-820) with the
 * metadata of prepare_prefill (runner.rs:978-1241).  q/out [T,H,D] fp16 (T = cu_seqlens_q[B]);
 * keys/values are read from the paged cache through block_tables for the whole context
 * [0, context_lens[b]) -- call b200_rope_kv_write_f16 for this step's tokens first.
 * context_lens may be NULL (no cached prefix: context = query length).  H/Hkv in {1,2,4,8}. */
void b200_attn_prefill_f16(void* out, const void* q, const void* k_cache, const void* v_cache,
                           const uint32_t* cu_seqlens_q, const uint32_t* context_lens,
                           const uint32_t* block_tables, int32_t batch, int32_t num_heads,
                           int32_t num_kv_heads, int32_t head_dim, int32_t block_size,
                           int32_t max_blocks_per_seq, int32_t num_blocks, int32_t max_seqlen_q,
                           float scale, int64_t stream);

/* Same, with the row count of q / out (num_tokens = cu_seqlens_q[batch], which the tcgen05 kernel's q
 * tensor map needs as a HOST value) and an explicit implementation: mode -1 = default (tcgen05 for
 * head_dim 128, env B200_PREFILL_ATTN), 0 = mma.sync kernel, 1 = tcgen05, 2 = tcgen05 bring-up switch. */
void b200_attn_prefill_ex_f16(void* out, const void* q, const void* k_cache, const void* v_cache,
                              const uint32_t* cu_seqlens_q, const uint32_t* context_lens,
                              const uint32_t* block_tables, int32_t batch, int32_t num_heads,
                              int32_t num_kv_heads, int32_t head_dim, int32_t block_size,
                              int32_t max_blocks_per_seq, int32_t num_blocks, int32_t max_seqlen_q,
                              int32_t num_tokens, int32_t mode, float scale, int64_t stream);
/* bf16 flavour (q / out / caches bf16; tcgen05 kind::f16 with bf16 operand formats, or the bf16 mma.sync kernel). */
void b200_attn_prefill_ex_bf16(void* out, const void* q, const void* k_cache, const void* v_cache,
                               const uint32_t* cu_seqlens_q, const uint32_t* context_lens,
                               const uint32_t* block_tables, int32_t batch, int32_t num_heads,
                               int32_t num_kv_heads, int32_t head_dim, int32_t block_size,
                               int32_t max_blocks_per_seq, int32_t num_blocks, int32_t max_seqlen_q,
                               int32_t num_tokens, int32_t mode, float scale, int64_t stream);

/* ------------------------------------------------- glue kernels (others.rs, mlp.rs, llama.rs) */
void b200_rmsnorm_f16(void* out, const void* x, const void* weight, int32_t num_tokens,
                      int32_t hidden, float eps, int64_t stream);
/* residual += x (fp16 rounding, as llama.rs:126-130); out = rmsnorm(residual) * weight */
void b200_add_rmsnorm_f16(void* out, void* residual, const void* x, const void* weight,
                          int32_t num_tokens, int32_t hidden, float eps, int64_t stream);
void b200_silu_mul_f16(void* out, const void* gate_up, int32_t num_tokens, int32_t inter,
                       int64_t stream);
/* bf16 flavours of the three (x / weight / residual / out bf16, fp32 math, the residual add rounds to bf16).
 * b200_embedding_f16, b200_gather_rows_f16, b200_kv_write_f16 and b200_swap_blocks move 16-bit elements without
 * arithmetic and serve bf16 tensors as they are.                                                                */
void b200_rmsnorm_bf16(void* out, const void* x, const void* weight, int32_t num_tokens,
                       int32_t hidden, float eps, int64_t stream);
void b200_add_rmsnorm_bf16(void* out, void* residual, const void* x, const void* weight,
                           int32_t num_tokens, int32_t hidden, float eps, int64_t stream);
void b200_silu_mul_bf16(void* out, const void* gate_up, int32_t num_tokens, int32_t inter,
                        int64_t stream);
void b200_embedding_f16(void* out, const void* table, const uint32_t* ids, int32_t num_tokens,
                        int32_t hidden, int32_t vocab, int64_t stream);
void b200_gather_rows_f16(void* out, const void* x, const uint32_t* rows, int32_t num_rows,
                          int32_t hidden, int64_t stream);
void b200_argmax_f32(uint32_t* out_ids, const float* logits, int32_t batch, int32_t vocab,
                     int64_t stream);
void b200_add_bias_f16(void* x, const void* bias, int32_t num_tokens, int32_t n, int64_t stream);

/* ------------------------------------------------- sampling (logits_processor.rs, runner.rs:1390-1570)
 * Device replacement of attention_rs::sampler::Sampler::sample_cuda (logits_processor.rs:226-229):
 * full-vocabulary softmax(logits / temperature) -> top_k (1..256) -> top-p cut on the
 * un-renormalised top-k probabilities -> multinomial (top_k = 0: Sampling::All, the whole vocabulary, no
 * cut); out_ids u32 [batch].  `seed_dev` points to
 * ONE u64 in device memory (the step seed, LogitsProcessor's rng.next_u64()), so a captured graph
 * changes it with a copy.  Row b draws u = f(seed, b) (this library's definition; the reference
 * kernel's is not source-visible).  Greedy (Sampling::ArgMax) is b200_argmax_f32.          */
void b200_sample_topk_topp_f32(uint32_t* out_ids, const float* logits, int32_t batch, int32_t vocab,
                               int32_t top_k, float top_p, float temperature,
                               const uint64_t* seed_dev, int64_t stream);
/* apply_batch_repeat_penalty (logits_processor.rs:308-343): context u32 [batch, context_cap] holds
 * each row's penalty window (the caller applies the "last 128 tokens once more than 128 exist"
 * rule of runner.rs:1522-1533), context_lens i32 [batch].  In place on fp32 logits.        */
void b200_apply_penalties_f32(float* logits, int32_t batch, int32_t vocab, const uint32_t* context,
                              const int32_t* context_lens, int32_t context_cap,
                              float frequency_penalty, float presence_penalty, int64_t stream);
/* LogitsProcessor::get_strategy + the (k, p, t) of sample_with_strategy (logits_processor.rs:48-70,
 * 208-216).  has_* = 0 is None.  Returns 0 ArgMax, 1 All, 2 TopK, 3 TopP, 4 TopKThenTopP.       */
int32_t b200_sampling_strategy(int32_t has_temperature, float temperature, int32_t has_top_k,
                               int64_t top_k, int32_t has_top_p, float top_p, int32_t* k, float* p,
                               float* t);
/* rand::rngs::StdRng::seed_from_u64 (logits_processor.rs:29-30): ChaCha12 keyed by the PCG32
 * expansion of the seed; next_u64 is the per-step kernel seed (logits_processor.rs:222-225).  */
typedef struct b200_rng b200_rng;
b200_rng* b200_rng_new(uint64_t seed);
void b200_rng_free(b200_rng*);
uint32_t b200_rng_next_u32(b200_rng*);
uint64_t b200_rng_next_u64(b200_rng*);

/* ------------------------------------------------- host side: block manager / scheduler
 * C++ mirror of src/core/{sequence,block_manager,scheduler}.rs and the input
 * preparation of src/core/runner.rs:952-1388.  Handles are opaque.               */
typedef struct b200_scheduler b200_scheduler;
b200_scheduler* b200_scheduler_new(int32_t num_blocks, int32_t block_size, int32_t max_num_seqs,
                                   int32_t max_num_batched_tokens, int32_t chunk_size,
                                   const uint32_t* eos_ids, int32_t num_eos);
void b200_scheduler_free(b200_scheduler*);
/* returns seq id (scheduler.rs:186-192) */
int64_t b200_scheduler_add(b200_scheduler*, const uint32_t* tokens, int32_t num_tokens,
                           int32_t max_tokens);
/* SamplingParams.stop_token_ids of a sequence that is still WAITING (id from add): num_stops token sequences
 * back to back in stop_tokens, lengths in stop_lens.  postprocess() then finishes the sequence when the sampled
 * token completes one of them (scheduler.rs:1309-1341: the head must equal the tail of the GENERATED tokens;
 * the stop token itself is not appended, like EOS).  Returns 0, -1 if the id is not waiting.             */
int32_t b200_scheduler_set_stop_sequences(b200_scheduler*, int64_t seq_id, const uint32_t* stop_tokens,
                                          const int32_t* stop_lens, int32_t num_stops);
/* schedule(): writes indices into `running` (<= cap); returns count, *is_prefill set.
 * Returns -1 on error (e.g. out of blocks).  (scheduler.rs:200-380)               */
int32_t b200_scheduler_schedule(b200_scheduler*, int32_t* indices, int32_t cap,
                                int32_t* is_prefill);
/* filter_prefill_finished (scheduler.rs:718-785): returns number finished; writes the
 * positions within `indices` (out_pos) and the new running indices (out_idx).      */
int32_t b200_scheduler_filter_prefill_finished(b200_scheduler*, const int32_t* indices,
                                               int32_t n, int32_t* out_pos, int32_t* out_idx);
void b200_scheduler_postprocess(b200_scheduler*, const int32_t* indices, int32_t n,
                                const uint32_t* tokens);
void b200_scheduler_clear_finished(b200_scheduler*);
int32_t b200_scheduler_num_running(const b200_scheduler*);
int32_t b200_scheduler_num_waiting(const b200_scheduler*);
int32_t b200_scheduler_num_free_blocks(const b200_scheduler*);
/* Prefix cache (src/core/prefix_cache.rs; block_manager.rs:179-221,275-299,346-442,552-604,
 * 657-672; scheduler.rs:60-104,288-337,617-620,1167-1175).  enable: max_tokens >= 0 is
 * `prefix_cache_max_tokens`, < 0 the default 65 % of the blocks; returns the capacity in blocks
 * (0 = disabled, as build_prefix_cache_config does).  Once enabled, allocate() shares the cached
 * full blocks of a matching prompt prefix (never the whole prompt), postprocess() inserts a
 * finished sequence's full blocks before dropping its references, and schedule() evicts 10 % of
 * the cache when a sequence cannot append or usage is over 95 %.  hash_block is the chain hash
 * (Rust DefaultHasher = SipHash-1-3, zero key, over parent u64 | len usize | tokens u32 LE).      */
int32_t b200_scheduler_enable_prefix_cache(b200_scheduler*, int64_t max_tokens);
int32_t b200_scheduler_prefix_cache_blocks(const b200_scheduler*);
int32_t b200_scheduler_evict_prefix_cache_blocks(b200_scheduler*, int32_t num_blocks);
int32_t b200_scheduler_block_ref_count(const b200_scheduler*, int32_t block_id);
/* free_block_ids in pop order; returns the count, copies up to cap */
int32_t b200_scheduler_free_block_ids(const b200_scheduler*, uint32_t* out, int32_t cap);
uint64_t b200_prefix_hash_block(uint64_t parent_hash, const uint32_t* tokens, int32_t n);
/* CPU swap of running sequences (block_manager.rs:69-78,255-272,878-1000; scheduler.rs:303-337,830-954), the
 * bookkeeping: enable with the number of CPU blocks (kvcache plan: num_cpu_blocks).  schedule() then swaps the
 * OLDEST sequence that cannot append out to CPU blocks when there is no prefix cache left to evict (and
 * another sequence keeps running), and brings ONE swapped sequence back per call once it has cooled down
 * 5 s, nothing is preempted and the cache has room.  The reference reads the wall clock; here the host sets it
 * (set_time_ms) before schedule().  The block copies themselves are NOT done here: after schedule() drain
 * pop_swap_op -- (src, dst) block pairs, GPU->CPU or CPU->GPU -- and run b200_swap_blocks per layer cache
 * before the step's forward.  Session ("Cached") sequences are out of scope.                              */
void b200_scheduler_enable_cpu_swap(b200_scheduler*, int32_t num_cpu_blocks);
void b200_scheduler_set_time_ms(b200_scheduler*, int64_t now_ms);
int32_t b200_scheduler_num_swapped(const b200_scheduler*);
int64_t b200_scheduler_swapped_seq_id(const b200_scheduler*, int32_t i);
int32_t b200_scheduler_free_cpu_block_ids(const b200_scheduler*, uint32_t* out, int32_t cap);
int32_t b200_scheduler_pop_swap_op(b200_scheduler*, int64_t* pairs, int32_t cap_pairs, int32_t* swap_in);
/* inspect a running sequence: returns token count; copies up to cap entries */
int32_t b200_scheduler_seq_info(const b200_scheduler*, int32_t running_idx, int64_t* seq_id,
                                int32_t* num_cached_tokens, int32_t* status,
                                int32_t* num_output_tokens);
int32_t b200_scheduler_seq_block_table(const b200_scheduler*, int32_t running_idx,
                                       uint32_t* out, int32_t cap);
int32_t b200_scheduler_seq_tokens(const b200_scheduler*, int32_t running_idx, uint32_t* out,
                                  int32_t cap);

/* prepare_decode (runner.rs:1243-1388) + prepare_block_tables (runner.rs:952-974).
 * Host output buffers sized [n] / [n * max_blocks_cap]; returns max_blocks (row
 * stride used) or -1 if max_blocks_cap is too small.  *max_context_len set.        */
int32_t b200_prepare_decode(const b200_scheduler*, const int32_t* indices, int32_t n,
                            uint32_t* input_ids, int64_t* positions, int64_t* slot_mapping,
                            uint32_t* context_lens, uint32_t* block_tables,
                            int32_t max_blocks_cap, int32_t* max_context_len);
/* prepare_prefill (runner.rs:978-1241).  Returns total tokens T (or -1 if cap too
 * small).  cu_seqlens_* are [n+1].  *has_context != 0 when block_tables /
 * context_lens were produced (some sequence has cached KV).                        */
int32_t b200_prepare_prefill(const b200_scheduler*, const int32_t* indices, int32_t n,
                             int32_t token_cap, uint32_t* input_ids, int64_t* positions,
                             int64_t* slot_mapping, uint32_t* cu_seqlens_q,
                             uint32_t* cu_seqlens_k, uint32_t* context_lens,
                             uint32_t* block_tables, int32_t max_blocks_cap,
                             int32_t* max_blocks, int32_t* max_seqlen_q, int32_t* max_seqlen_k,
                             int32_t* max_context_len, int32_t* has_context);

/* ------------------------------------------------- KV-cache planning (host, integer)
 * Mirror of KVCacheAllocator::{per_block_bytes, plan_allocation, auto_decide_params}
 * (src/utils/kvcache_allocator.rs:447-468, 553-612, 616-707) for dense (non-MLA, non-hybrid)
 * models.  `available_bytes` is what get_rank_available_memory returns: free * kv_fraction,
 * minimum over ranks.  user_max_num_seqs / user_max_model_len <= 0 mean "auto-decide".
 * Returns 0, or -1 for InvalidConfiguration, -2 for InsufficientGpuMemory.             */
typedef struct b200_kvcache_plan {
  int64_t per_block_bytes, num_gpu_blocks, num_cpu_blocks, max_num_seqs, max_model_len;
  int64_t kvcache_memory_bytes, max_num_batched_tokens;
} b200_kvcache_plan;
int32_t b200_kvcache_plan_allocation(int64_t available_bytes, int32_t num_layers,
                                     int32_t num_kv_heads, int32_t head_dim, int32_t num_shards,
                                     int32_t block_size, int32_t dtype_size,
                                     int64_t config_model_len, int64_t user_max_num_seqs,
                                     int64_t user_max_model_len, float cpu_mem_fold,
                                     b200_kvcache_plan* out);

/* ------------------------------------------------- model runner (decode step)
 * C++ mirror of ModelRunner::run for the decode branch (runner.rs:743-896) over
 * LLaMaForCausalLM::forward_inner (llama.rs:269-321): every launch of one decode
 * step, optionally captured into CUDA graphs per batch bucket (graph.rs:370-377).  */
typedef struct b200_model_config {
  int32_t hidden, num_layers, num_heads, num_kv_heads, head_dim, inter, vocab;
  int32_t bits, group_size, qfmt;      /* bits = 16 => dense fp16 weights              */
  int32_t block_size, num_blocks, max_blocks_per_seq, max_batch, max_pos;
  int32_t qkv_bias;                    /* Qwen2: q/k/v bias                            */
  int32_t cos_sin_f32;
  float rms_eps, attn_scale;
  int32_t tp_rank, tp_size;            /* shapes above are the LOCAL (per-rank) shard  */
  int32_t max_prefill_tokens;          /* 0 => decode only; else tokens per prefill step */
  int32_t kv_fp8;                      /* EngineConfig.fp8_kvcache: caches are e4m3 bytes (head_dim 128) */
} b200_model_config;

typedef struct b200_layer_weights {
  const void *input_norm, *post_norm;                 /* fp16 [hidden]                  */
  const void *qkv_w, *qkv_s, *qkv_z, *qkv_bias;       /* fused q|k|v along N            */
  const void *o_w, *o_s, *o_z;
  const void *gate_up_w, *gate_up_s, *gate_up_z;      /* fused gate|up along N          */
  const void *down_w, *down_s, *down_z;
  void *k_cache, *v_cache;
} b200_layer_weights;

typedef struct b200_runner b200_runner;
b200_runner* b200_runner_new(const b200_model_config* cfg);
void b200_runner_free(b200_runner*);
void b200_runner_set_layer(b200_runner*, int32_t layer, const b200_layer_weights* w);
void b200_runner_set_globals(b200_runner*, const void* embed, const void* final_norm,
                             const void* lm_head, const void* cos, const void* sin);
int64_t b200_runner_workspace_bytes(const b200_runner*);
void b200_runner_set_workspace(b200_runner*, void* workspace, int64_t bytes);
/* all-reduce hook for tensor parallelism: called on the stream after o_proj / down_proj
 * with (buf, count_halfs, user, stream); NULL => single GPU (distributed.rs:438-454).  */
typedef void (*b200_allreduce_fn)(void* buf, int64_t count, void* user, int64_t stream);
void b200_runner_set_allreduce(b200_runner*, b200_allreduce_fn fn, void* user);
/* One decode step from DEVICE inputs; logits (fp32 [B, vocab]) and/or greedy token
 * ids (u32 [B]) are written to device buffers (either may be NULL).                */
void b200_runner_decode_device(b200_runner*, const uint32_t* input_ids, const int64_t* positions,
                               const int64_t* slot_mapping, const uint32_t* block_tables,
                               const uint32_t* context_lens, int32_t batch, float* logits,
                               uint32_t* out_ids, int64_t stream);
/* ModelRunner::run equivalent for decode: HOST inputs (pinned), copies H2D, runs the
 * step (graph replay when captured for this bucket), copies token ids D2H and
 * synchronises the stream.  Returns 0 on success.                                   */
int32_t b200_runner_decode_host(b200_runner*, const uint32_t* input_ids,
                                const int64_t* positions, const int64_t* slot_mapping,
                                const uint32_t* block_tables, int32_t max_blocks,
                                const uint32_t* context_lens, int32_t batch,
                                uint32_t* out_ids_host, int64_t stream);
/* One prefill step (ModelRunner::run, prefill branch, runner.rs:743-896 with the metadata of
 * prepare_prefill): DEVICE inputs; T = num_tokens rows.  block_tables must cover every
 * sequence (keys are read from the paged cache after this step's KV write); context_lens may
 * be NULL when nothing is cached.  last_rows u32 [num_seqs] = cu_seqlens_q[1:] - 1
 * (llama.rs:306-310).  logits f32 [num_seqs, vocab] and/or greedy ids are written.        */
void b200_runner_prefill_device(b200_runner*, const uint32_t* input_ids, const int64_t* positions,
                                const int64_t* slot_mapping, const uint32_t* cu_seqlens_q,
                                const uint32_t* context_lens, const uint32_t* block_tables,
                                const uint32_t* last_rows, int32_t num_seqs, int32_t num_tokens,
                                int32_t max_seqlen_q, int32_t max_blocks, float* logits,
                                uint32_t* out_ids, int64_t stream);
/* Capture CUDA graphs for the given batch buckets (graph.rs:448-690). 0 on success. */
int32_t b200_runner_capture(b200_runner*, const int32_t* buckets, int32_t num_buckets,
                            int64_t stream);
int64_t b200_runner_launches_per_step(const b200_runner*);
/* Sampling of the decode / prefill steps: top_k == 0 => greedy argmax (default); otherwise the
 * device sampler above with these (k, p, t) (the strategy cached at prefill, runner.rs:1402-1510).
 * b200_runner_set_step_seed copies the step seed into the runner's device slot on `stream`
 * (call once per step before b200_runner_decode_host / _device; graph-replay safe).          */
void b200_runner_set_sampling(b200_runner*, int32_t top_k, float top_p, float temperature);
void b200_runner_set_step_seed(b200_runner*, uint64_t seed, int64_t stream);
/* The same from b200_sampling_strategy's result (0 ArgMax, 1 All, 2 TopK, 3 TopP, 4 TopKThenTopP) and its
 * (k, p, t): Sampling::All (temperature only, logits_processor.rs:199-207) draws from the whole vocabulary.
 * Any change of strategy or parameters drops the captured graphs (the values are kernel arguments).  0 / -1. */
int32_t b200_runner_set_sampling_strategy(b200_runner*, int32_t strategy, int32_t k, float p, float t);

/* ------------------------------------------------- engine step (engine.rs:1693-1757)
 * prepare_step -> run_forward -> finish_step of a DECODE batch in one native call: schedule(),
 * prepare_decode into pinned staging owned by the handle, b200_runner_decode_host, postprocess,
 * clear_finished.  b200_engine_step returns the number of sequences decoded (running indices at schedule
 * time, sequence ids and new tokens in the three output arrays, each sized >= max(max_batch, 5), any may be
 * NULL); 0 = nothing schedulable; -2 = the scheduler produced a PREFILL batch: *num_prefill indices are in
 * indices_out and the caller runs that step (b200_prepare_prefill / b200_runner_prefill_device /
 * filter_prefill_finished / postprocess); -1 = error.  The handle borrows the scheduler and the runner. */
typedef struct b200_engine b200_engine;
b200_engine* b200_engine_new(b200_scheduler* sched, b200_runner* runner, int32_t max_batch,
                             int32_t max_blocks_per_seq);
void b200_engine_free(b200_engine*);
/* CPU swap (scheduler.rs:303-337,830-954; runner.rs:1626-1645): per layer the GPU K / V caches and their
 * pinned-host mirrors.  b200_engine_step then runs every block copy schedule() queued BEFORE the step's
 * forward (a victim's GPU blocks may already belong to a sequence of the same batch); with swapping enabled
 * on the scheduler and no caches registered the step fails instead of decoding stale pages.  0 / -1.       */
int32_t b200_engine_set_swap_caches(b200_engine*, int32_t num_layers, void* const* gpu_k, void* const* gpu_v,
                                    void* const* cpu_k, void* const* cpu_v, int64_t block_bytes);
/* Seed of the sampling stream (StdRng::seed_from_u64, logits_processor.rs:29-30): once set, every decode step
 * pushes rng.next_u64() to the runner's device seed before the forward (logits_processor.rs:222-225).      */
void b200_engine_set_seed(b200_engine*, uint64_t seed);
int32_t b200_engine_step(b200_engine*, int32_t* indices_out, int64_t* seq_ids_out, uint32_t* tokens_out,
                         int32_t* num_prefill, int64_t stream);

/* ------------------------------------------------- tensor-parallel communicator
 * One process per GPU (runner/runner.rs:73-240).  NCCL communicator from a 128-byte unique id
 * (engine.rs:263, runner/mod.rs:33-58) + an optional one-shot NVLink all-reduce over
 * cudaIpc-mapped peer buffers for messages <= 1 MiB.  b200_allreduce_f16 is the in-place
 * equivalent of AllReduce::cuda_fwd (distributed.rs:335-395).                       */
typedef struct b200_comm b200_comm;
int32_t b200_nccl_unique_id(uint8_t* out128);
b200_comm* b200_comm_init_rank(const uint8_t* id128, int32_t rank, int32_t world);
int32_t b200_comm_enable_oneshot(b200_comm*, int64_t stream);
void b200_comm_set_force_nccl(b200_comm*, int32_t force);
void b200_comm_free(b200_comm*);
void b200_allreduce_f16(b200_comm*, void* buf, int64_t count, int64_t stream);
int32_t b200_comm_rank(const b200_comm*);
int32_t b200_comm_world(const b200_comm*);
/* the runner all-reduces o_proj / down_proj outputs through this communicator.  With the one-shot
 * path enabled, decode steps (batch <= 64) run the FUSED form: the row-parallel GEMM's tile
 * finalisers store their partial straight into every rank's receive buffer over NVLink and one
 * kernel waits for all ranks' tiles, sums them in rank order, adds the residual and applies the
 * next RMSNorm (3 launches and an NVLink round trip fewer per all-reduce; bit-identical to the
 * unfused GEMM -> all-reduce -> add+norm).  b200_runner_set_tp_fused: -1 env B200_TP_FUSED
 * (default on), 0 off, 1 on.                                                            */
void b200_runner_set_comm(b200_runner*, b200_comm*);
void b200_runner_set_tp_fused(b200_runner*, int32_t on);

#ifdef __cplusplus
}
#endif
#endif /* B200_VLLM_H */
