// Decode-step model runner: every launch of one LLaMA/Qwen2 decode step on one GPU (one
// TP rank), optionally captured into CUDA graphs per batch bucket.
//
// C++ mirror of the decode branch of ModelRunner::run
//   /root/reference/src/core/runner.rs:743-896   (prepare -> forward -> sample)
// over LLaMaForCausalLM::forward_inner / LLaMaDecoderLayer::forward
//   /root/reference/src/models/llama.rs:107-131,269-321
// Attention::forward_ext  /root/reference/src/models/layers/attention.rs:648-839
// MLP::forward            /root/reference/src/models/layers/mlp.rs:451-469
// TensorParallelRowLinear /root/reference/src/models/layers/distributed.rs:438-454
// GraphCapturer           /root/reference/src/utils/graph.rs:370-377,448-835
//
// Launch plan per layer (5 kernels + 2 small ones, vs 17-22 in the reference):
//   [add+]rmsnorm -> fused qkv W4A16 GEMM (+bias) -> RoPE + KV write -> paged attention
//   (+ split merge) -> o_proj GEMM [-> all-reduce] -> add+rmsnorm -> gate|up GEMM with
//   SiLU*mul epilogue -> down GEMM [-> all-reduce]
// The runner owns no device memory: weights, caches and one workspace come from the caller.
#include <cuda_runtime.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "../../../include/b200_vllm.h"
#include "../attn.cuh"
#include "../wna16.cuh"

using namespace b200;
namespace b200 {
long long launch_count();
bool comm_push_targets(b200_comm* c, int buf, int M, int N, FusedEpi* fe);
void comm_reduce_add_rmsnorm(b200_comm* c, int buf, int tiles, void* resid, const void* weight, void* out,
                             int M, int hidden, float eps, cudaStream_t st);
void launch_sample_topk_topp(uint32_t* out_ids, const float* logits, int B, int V, int k, float p, float t,
                             const uint64_t* seed_dev, cudaStream_t st);
void launch_wna16_dequant(const uint32_t* bp, const __half* scales, const uint32_t* qzeros, __half* out,
                          int K, int N, int bits, int group_size, int qfmt, cudaStream_t st, int interleave);
bool launch_dense_gemm_tc_silu(const void* x, const void* w, void* out, int M, int N, int K, int mode,
                               cudaStream_t st);
}

namespace {
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
}  // namespace

struct b200_runner {
  b200_model_config cfg;
  std::vector<b200_layer_weights> layers;
  std::vector<AttnMaps> maps;
  std::vector<bool> maps_ok;
  const void *embed = nullptr, *final_norm = nullptr, *lm_head = nullptr, *cos = nullptr,
             *sin = nullptr;
  b200_allreduce_fn allreduce = nullptr;
  void* allreduce_user = nullptr;
  b200_comm* comm = nullptr;     // set by b200_runner_set_comm: enables the fused GEMM -> all-reduce -> norm plan
  int tp_fused = -1;             // -1 = env B200_TP_FUSED (default on), 0 / 1 forced
  // workspace carve-up
  uint8_t* ws = nullptr;
  int64_t ws_bytes = 0;
  __half *x = nullptr, *xn = nullptr, *qkv = nullptr, *attn = nullptr, *proj = nullptr,
         *act = nullptr, *act2 = nullptr;
  float* logits = nullptr;
  float* ssq = nullptr;          // [2 * layers + 1][max_batch] sums of squares of the folded RMSNorms
  void *attn_ws = nullptr, *wna_ws = nullptr;
  __half* wdeq = nullptr;        // fp16 [N, K] scratch of ONE projection (large quantised prefill steps)
  // --fp8-kvcache: fp16 scratch pages (+ identity block table, tensor maps) the prefill attention reads
  __half *kv_sk = nullptr, *kv_sv = nullptr;
  uint32_t* kv_sbt = nullptr;
  AttnMaps kv_smaps;
  bool kv_smaps_ok = false;
  // static (graph) input buffers
  uint32_t *s_ids = nullptr, *s_ctx = nullptr, *s_bt = nullptr, *s_out = nullptr;
  int64_t *s_pos = nullptr, *s_slot = nullptr;
  // pinned host staging for padded rows
  uint64_t* s_seed = nullptr;    // step seed of the device sampler
  int samp_k = 0;                // 0 = greedy (or Sampling::All when samp_all)
  bool samp_all = false;         // temperature-only multinomial over the whole vocabulary
  float samp_p = 1.f, samp_t = 1.f;
  std::map<int, cudaGraphExec_t> graphs;
  int64_t launches = 0;
};

// Prefill steps of at least this many tokens run quantised linears as "dequantise the layer's weight
// once to fp16, then the dense tcgen05 GEMM" (dense_gemm_tc.cu); smaller steps keep the in-kernel
// dequant GEMM (wna16_tc.cu), which re-dequantises per 256-token tile.  B200_PREFILL_DENSE_MIN_T.
static int prefill_dense_min_tokens() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_PREFILL_DENSE_MIN_T");
    v = e ? atoi(e) : 1024;
    if (v <= 0) v = 1 << 30;
  }
  return v;
}

static bool prefill_silu_fused() {       // B200_PREFILL_SILU_FUSED=0: unfused GEMM + silu_mul (A/B arm)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_PREFILL_SILU_FUSED");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static int64_t carve(b200_runner* r, uint8_t* base) {
  const b200_model_config& c = r->cfg;
  const size_t B = c.max_batch;
  const size_t qkv_n = static_cast<size_t>(c.num_heads + 2 * c.num_kv_heads) * c.head_dim;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    uint8_t* p = base ? base + off : nullptr;
    off = align_up(off + bytes, 256);
    return p;
  };
  // activation rows: a decode batch or a whole prefill step, whichever is larger
  const size_t R = std::max<size_t>(B, c.max_prefill_tokens > 0 ? c.max_prefill_tokens : 0);
  r->x = (__half*)take(R * c.hidden * 2);
  r->xn = (__half*)take(R * c.hidden * 2);
  r->qkv = (__half*)take(R * qkv_n * 2);
  r->attn = (__half*)take(R * c.num_heads * c.head_dim * 2);
  r->proj = (__half*)take(R * c.hidden * 2);
  r->act = (__half*)take(R * 2 * static_cast<size_t>(c.inter) * 2);
  r->act2 = (__half*)take(R * static_cast<size_t>(c.inter) * 2);
  r->logits = (float*)take(B * c.vocab * 4);
  r->ssq = (float*)take((2 * static_cast<size_t>(c.num_layers) + 1) * B * kSsqStride * 4);
  r->attn_ws = take(b200_paged_attn_workspace_bytes(c.max_batch, c.num_heads, c.head_dim,
                                                    c.max_blocks_per_seq));
  const size_t max_n = std::max<size_t>(std::max<size_t>(qkv_n, 2 * c.inter), c.hidden);
  r->wna_ws = take(b200_wna16_workspace_bytes(c.max_batch, static_cast<int32_t>(max_n)));
  if (c.bits != 16 && c.max_prefill_tokens >= prefill_dense_min_tokens()) {
    const size_t kn = std::max<size_t>(static_cast<size_t>(c.hidden) * std::max<size_t>(qkv_n, 2 * c.inter),
                                       static_cast<size_t>(c.inter) * c.hidden);
    r->wdeq = (__half*)take(kn * 2);
  }
  if (c.kv_fp8 && c.max_prefill_tokens > 0) {
    const size_t pages = B * static_cast<size_t>(c.max_blocks_per_seq);
    const size_t page_bytes = 64 * static_cast<size_t>(c.num_kv_heads) * c.head_dim * 2;
    r->kv_sk = (__half*)take(pages * page_bytes);
    r->kv_sv = (__half*)take(pages * page_bytes);
    r->kv_sbt = (uint32_t*)take(pages * 4);
  }
  r->s_ids = (uint32_t*)take(B * 4);
  r->s_ctx = (uint32_t*)take(B * 4);
  r->s_out = (uint32_t*)take(B * 4);
  r->s_pos = (int64_t*)take(B * 8);
  r->s_slot = (int64_t*)take(B * 8);
  r->s_bt = (uint32_t*)take(B * c.max_blocks_per_seq * 4);
  r->s_seed = (uint64_t*)take(8);
  return static_cast<int64_t>(off);
}

extern "C" {

b200_runner* b200_runner_new(const b200_model_config* cfg) {
  if (!cfg || cfg->hidden <= 0 || cfg->num_layers <= 0 || cfg->num_heads <= 0 ||
      cfg->num_kv_heads <= 0 || cfg->max_batch <= 0 || cfg->block_size != 64 ||
      !(cfg->bits == 4 || cfg->bits == 8 || cfg->bits == 16)) {
    set_error(1, "runner_new: bad config");
    return nullptr;
  }
  if (cfg->kv_fp8 && cfg->head_dim != 128) {
    set_error(1, "runner_new: the fp8 KV cache needs head_dim 128");
    return nullptr;
  }
  auto* r = new b200_runner();
  r->cfg = *cfg;
  r->layers.resize(cfg->num_layers);
  r->maps.resize(cfg->num_layers);
  r->maps_ok.assign(cfg->num_layers, false);
  std::memset(r->layers.data(), 0, sizeof(b200_layer_weights) * cfg->num_layers);
  return r;
}

void b200_runner_free(b200_runner* r) {
  if (!r) return;
  for (auto& g : r->graphs) cudaGraphExecDestroy(g.second);
  delete r;
}

void b200_runner_set_layer(b200_runner* r, int32_t layer, const b200_layer_weights* w) {
  if (layer < 0 || layer >= r->cfg.num_layers) {
    set_error(1, "runner_set_layer: layer out of range");
    return;
  }
  r->layers[layer] = *w;
  r->maps_ok[layer] = make_attn_maps(&r->maps[layer], w->k_cache, w->v_cache, r->cfg.num_blocks,
                                     r->cfg.num_kv_heads, r->cfg.head_dim, r->cfg.kv_fp8 != 0);
  if (!r->maps_ok[layer]) set_error(2, "runner_set_layer: cuTensorMapEncodeTiled failed");
}

void b200_runner_set_globals(b200_runner* r, const void* embed, const void* final_norm,
                             const void* lm_head, const void* cos, const void* sin) {
  r->embed = embed;
  r->final_norm = final_norm;
  r->lm_head = lm_head;
  r->cos = cos;
  r->sin = sin;
}

int64_t b200_runner_workspace_bytes(const b200_runner* r) {
  b200_runner tmp;
  tmp.cfg = r->cfg;
  return carve(&tmp, nullptr) + 256;
}

void b200_runner_set_workspace(b200_runner* r, void* workspace, int64_t bytes) {
  uint8_t* base = reinterpret_cast<uint8_t*>(align_up(reinterpret_cast<size_t>(workspace), 256));
  const int64_t need = carve(r, base);
  if (need + (base - reinterpret_cast<uint8_t*>(workspace)) > bytes) {
    set_error(1, "runner_set_workspace: workspace too small");
    return;
  }
  r->ws = base;
  r->ws_bytes = bytes;
  if (r->kv_sk) {
    r->kv_smaps_ok = make_attn_maps(&r->kv_smaps, r->kv_sk, r->kv_sv, r->cfg.max_batch * r->cfg.max_blocks_per_seq,
                                    r->cfg.num_kv_heads, r->cfg.head_dim, false);
    if (!r->kv_smaps_ok) set_error(2, "runner_set_workspace: cuTensorMapEncodeTiled (fp8 scratch) failed");
  }
  // the GEMM split-K accumulators must start zeroed (they are left zeroed by every call)
  cudaMemset(workspace, 0, bytes);
}

static void comm_allreduce_thunk(void* buf, int64_t count, void* user, int64_t stream) {
  b200_allreduce_f16(reinterpret_cast<b200_comm*>(user), buf, count, stream);
}

void b200_runner_set_comm(b200_runner* r, b200_comm* c) {
  r->allreduce = c ? comm_allreduce_thunk : nullptr;
  r->allreduce_user = c;
  r->comm = c;
}

void b200_runner_set_tp_fused(b200_runner* r, int32_t on) {
  if (r->tp_fused != on) {
    for (auto& g : r->graphs) cudaGraphExecDestroy(g.second);
    r->graphs.clear();
  }
  r->tp_fused = on;
}

void b200_runner_set_allreduce(b200_runner* r, b200_allreduce_fn fn, void* user) {
  r->allreduce = fn;
  r->allreduce_user = user;
  r->comm = nullptr;
}

// ModelRunner::sample (runner.rs:1390-1570) for the cached strategy: greedy argmax or the device sampler
static void pick_tokens(b200_runner* r, uint32_t* out_ids, const float* logits, int B, int64_t stream) {
  if (r->samp_k > 0 || r->samp_all)
    launch_sample_topk_topp(out_ids, logits, B, r->cfg.vocab, r->samp_all ? 0 : std::min(r->samp_k, r->cfg.vocab),
                            r->samp_p, r->samp_t, r->s_seed, reinterpret_cast<cudaStream_t>(stream));
  else
    b200_argmax_f32(out_ids, logits, B, r->cfg.vocab, stream);
}

static void linear(b200_runner* r, const __half* in, const void* w, const void* s, const void* z,
                   const void* bias, __half* out, int M, int K, int N, int epilogue,
                   int64_t stream, int variant = 0) {
  const b200_model_config& c = r->cfg;
  if (c.bits == 16) {
    b200_dense_gemm_f16(in, w, out, M, N, K, 0, stream);
    if (bias) {
      b200_add_bias_f16(out, bias, M, N, stream);
    }
    return;
  }
  if (r->wdeq && epilogue == B200_EPI_NONE && M >= prefill_dense_min_tokens() && N % kTileN == 0 &&
      K % kChunkK == 0) {
    const int g = c.group_size <= 0 ? K : c.group_size;
    launch_wna16_dequant((const uint32_t*)w, (const __half*)s, (const uint32_t*)z, r->wdeq, K, N, c.bits, g,
                         c.qfmt, reinterpret_cast<cudaStream_t>(stream), 0);
    b200_dense_gemm_f16(in, r->wdeq, out, M, N, K, 0, stream);
    if (bias) b200_add_bias_f16(out, bias, M, N, stream);
    return;
  }
  b200_wna16_gemm_ex(in, w, s, z, bias, out, M, K, N, c.bits, c.group_size, c.qfmt, epilogue,
                     variant, r->wna_ws, stream);
}

// One quantised linear with a fused tile-finaliser epilogue (wna16.cuh, FusedEpi).
static void linear_fused(b200_runner* r, const __half* in, const void* w, const void* s, const void* z,
                         const void* bias, __half* out, int M, int K, int N, int epilogue,
                         const FusedEpi& fe, cudaStream_t st) {
  const b200_model_config& c = r->cfg;
  WnaParams p;
  p.a = in; p.bp = (const uint32_t*)w; p.scales = (const __half*)s; p.qzeros = (const uint32_t*)z;
  p.bias = (const __half*)bias; p.c = out; p.M = M; p.K = K; p.N = N; p.bits = c.bits;
  p.group_size = c.group_size <= 0 ? K : c.group_size; p.qfmt = c.qfmt; p.epilogue = epilogue;
  p.no_split = 0; p.workspace = r->wna_ws; p.fe = &fe;
  if (M > gemv_max_m() && launch_wna16_sk(p, st)) return;
  if (M > gemv_max_m() && launch_wna16_tc_dec(p, st)) return;
  if (M > gemv_max_m() && launch_wna16_tc(p, st)) return;
  launch_wna16_gemv(p, st);
}

// B200_FUSED: 1 (default) = fully fused plan (RMSNorms folded into the GEMM finalisers, RoPE + KV
// write in the qkv finaliser: 6 launches per layer), 2 = classic plan with only the RoPE + KV write
// folded into the qkv finaliser, 0 = classic plan (9 launches per layer).  Llama-3-8B b=32 ctx=4096:
// 5.19 / 5.29 / 5.33 ms per step.
static int fused_env() {
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("B200_FUSED");
    env = e ? atoi(e) : 1;
  }
  return env;
}
static bool rope_fusable(const b200_runner* r) {
  const b200_model_config& c = r->cfg;
  return c.bits != 16 && !c.cos_sin_f32 && (kTileN % c.head_dim) == 0 && c.max_batch <= 64;
}
static bool fused_plan_enabled(const b200_runner* r) {
  // one sum-of-squares slot per 128-channel tile of the residual stream (kSsqStride slots per row)
  if (r->cfg.hidden % kTileN != 0 || r->cfg.hidden / kTileN > kSsqStride) return false;
  return fused_env() == 1 && rope_fusable(r) && r->allreduce == nullptr;
}

static bool tp_fused_enabled(const b200_runner* r) {
  if (!r->comm || r->cfg.bits == 16) return false;
  if (r->tp_fused >= 0) return r->tp_fused != 0;
  static int env = -1;
  if (env < 0) {
    const char* e = getenv("B200_TP_FUSED");
    env = (e && e[0] == '0') ? 0 : 1;
  }
  return env == 1;
}

// Row-parallel linear of a TP rank followed by all-reduce, residual add and the next RMSNorm
// (distributed.rs:438-454 + llama.rs:126-130 + others.rs:11-66).  Fused form: the GEMM's tile
// finalisers push the partial into every rank's receive buffer (kEpiPushPeers) and ONE kernel waits,
// sums, adds and normalises.  Returns false when the shape needs the classic three launches.
static bool row_linear_reduce_norm(b200_runner* r, int buf, const __half* in, const void* w, const void* s,
                                   const void* z, int M, int K, const void* norm_w, cudaStream_t st) {
  const b200_model_config& c = r->cfg;
  if (!tp_fused_enabled(r) || M > 64) return false;
  FusedEpi fe;
  std::memset(&fe, 0, sizeof(fe));
  if (!comm_push_targets(r->comm, buf, M, c.hidden, &fe)) return false;
  linear_fused(r, in, w, s, z, nullptr, r->proj, M, K, c.hidden, kEpiPushPeers, fe, st);
  comm_reduce_add_rmsnorm(r->comm, buf, c.hidden / kTileN, r->x, norm_w, r->xn, M, c.hidden, c.rms_eps, st);
  return true;
}

// Fused decode plan: RMSNorm weight multiply + sum of squares folded into the producing GEMM's
// finaliser, the per-token 1/rms applied in the consuming GEMM's finaliser, RoPE + KV-cache write
// in the qkv finaliser.  Per layer: qkv, attention (+ merge), o, gate|up, down = 6 launches.
static void decode_fused(b200_runner* r, const uint32_t* input_ids, const int64_t* positions,
                         const int64_t* slot_mapping, const uint32_t* block_tables,
                         const uint32_t* context_lens, int B, float* logits, uint32_t* out_ids,
                         int64_t stream) {
  const b200_model_config& c = r->cfg;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int H = c.num_heads, Hkv = c.num_kv_heads, D = c.head_dim, L = c.num_layers;
  const int qkv_n = (H + 2 * Hkv) * D;
  const float inv_h = 1.f / static_cast<float>(c.hidden);
  const size_t MB = c.max_batch;
  b200_embedding_f16(r->x, r->embed, input_ids, B, c.hidden, c.vocab, stream);
  b200_rmsnorm_f16(r->xn, r->x, r->layers[0].input_norm, B, c.hidden, c.rms_eps, stream);
  __half* xw = r->xn;          // GEMM input buffer: x * w_norm (the 1/rms factor rides in ssq)
  for (int l = 0; l < L; ++l) {
    const b200_layer_weights& w = r->layers[l];
    FusedEpi fq;
    std::memset(&fq, 0, sizeof(fq));
    if (l > 0) {
      fq.ssq_in = r->ssq + (2 * l - 1) * MB * kSsqStride; fq.ssq_tiles = c.hidden / kTileN;
      fq.inv_dim = inv_h; fq.eps = c.rms_eps;
    }
    fq.positions = positions; fq.slots = slot_mapping;
    fq.cos = (const __half*)r->cos; fq.sin = (const __half*)r->sin;
    fq.k_cache = (__half*)w.k_cache; fq.v_cache = (__half*)w.v_cache;
    fq.H = H; fq.Hkv = Hkv; fq.D = D; fq.kv_fp8 = c.kv_fp8;
    linear_fused(r, xw, w.qkv_w, w.qkv_s, w.qkv_z, c.qkv_bias ? w.qkv_bias : nullptr, r->qkv, B,
                 c.hidden, qkv_n, kEpiRopeKv, fq, st);
    launch_paged_attn_decode(r->maps[l], r->attn, r->qkv, qkv_n, block_tables, context_lens, B, H, Hkv,
                             D, c.max_blocks_per_seq, c.attn_scale, 0.f, 0, r->attn_ws, st);
    FusedEpi fo;
    std::memset(&fo, 0, sizeof(fo));
    fo.resid = r->x; fo.norm_w = (const __half*)w.post_norm; fo.ssq_out = r->ssq + (2 * l) * MB * kSsqStride;
    linear_fused(r, r->attn, w.o_w, w.o_s, w.o_z, nullptr, xw, B, H * D, c.hidden, kEpiResidNorm, fo, st);
    FusedEpi fg;
    std::memset(&fg, 0, sizeof(fg));
    fg.ssq_in = r->ssq + (2 * l) * MB * kSsqStride; fg.ssq_tiles = c.hidden / kTileN;
    fg.inv_dim = inv_h; fg.eps = c.rms_eps;
    linear_fused(r, xw, w.gate_up_w, w.gate_up_s, w.gate_up_z, nullptr, r->act, B, c.hidden,
                 2 * c.inter, B200_EPI_SILU_MUL, fg, st);
    FusedEpi fd;
    std::memset(&fd, 0, sizeof(fd));
    fd.resid = r->x;
    fd.norm_w = (const __half*)(l + 1 < L ? r->layers[l + 1].input_norm : r->final_norm);
    fd.ssq_out = r->ssq + (2 * l + 1) * MB * kSsqStride;
    linear_fused(r, r->act, w.down_w, w.down_s, w.down_z, nullptr, xw, B, c.inter, c.hidden,
                 kEpiResidNorm, fd, st);
  }
  launch_rowscale(r->proj, xw, r->ssq + (2 * L - 1) * MB * kSsqStride, B, c.hidden, c.rms_eps, st);
  float* lg = logits ? logits : r->logits;
  b200_dense_gemm_f16(r->proj, r->lm_head, lg, B, c.vocab, c.hidden, 1, stream);
  if (out_ids) pick_tokens(r, out_ids, lg, B, stream);
}

void b200_runner_decode_device(b200_runner* r, const uint32_t* input_ids,
                               const int64_t* positions, const int64_t* slot_mapping,
                               const uint32_t* block_tables, const uint32_t* context_lens,
                               int32_t B, float* logits, uint32_t* out_ids, int64_t stream) {
  const b200_model_config& c = r->cfg;
  if (!r->ws) {
    set_error(1, "runner_decode: workspace not set");
    return;
  }
  if (B <= 0 || B > c.max_batch) {
    set_error(1, "runner_decode: batch out of range");
    return;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long launches0 = launch_count();
  if (fused_plan_enabled(r)) {
    decode_fused(r, input_ids, positions, slot_mapping, block_tables, context_lens, B, logits, out_ids,
                 stream);
    r->launches = launch_count() - launches0;
    return;
  }
  const int H = c.num_heads, Hkv = c.num_kv_heads, D = c.head_dim;
  const int qkv_n = (H + 2 * Hkv) * D;
  b200_embedding_f16(r->x, r->embed, input_ids, B, c.hidden, c.vocab, stream);
  bool normed = false;     // r->xn already holds rmsnorm(x) for the coming block (fused TP reduce wrote it)
  for (int l = 0; l < c.num_layers; ++l) {
    const b200_layer_weights& w = r->layers[l];
    if (!normed) {
      if (l == 0)
        b200_rmsnorm_f16(r->xn, r->x, w.input_norm, B, c.hidden, c.rms_eps, stream);
      else
        b200_add_rmsnorm_f16(r->xn, r->x, r->proj, w.input_norm, B, c.hidden, c.rms_eps, stream);
    }
    normed = false;
    // ---- attention block
    __half* q = r->qkv;
    if (fused_env() != 0 && rope_fusable(r)) {     // also the TP path (no collective touches q|k|v)
      FusedEpi fq;
      std::memset(&fq, 0, sizeof(fq));
      fq.positions = positions; fq.slots = slot_mapping;
      fq.cos = (const __half*)r->cos; fq.sin = (const __half*)r->sin;
      fq.k_cache = (__half*)w.k_cache; fq.v_cache = (__half*)w.v_cache;
      fq.H = H; fq.Hkv = Hkv; fq.D = D; fq.kv_fp8 = c.kv_fp8;
      linear_fused(r, r->xn, w.qkv_w, w.qkv_s, w.qkv_z, c.qkv_bias ? w.qkv_bias : nullptr, r->qkv, B,
                   c.hidden, qkv_n, kEpiRopeKv, fq, st);
    } else {
      linear(r, r->xn, w.qkv_w, w.qkv_s, w.qkv_z, c.qkv_bias ? w.qkv_bias : nullptr, r->qkv, B,
             c.hidden, qkv_n, B200_EPI_NONE, stream);
      __half* k = r->qkv + H * D;
      __half* v = r->qkv + (H + Hkv) * D;
      launch_rope_kv_write(q, k, v, qkv_n, qkv_n, qkv_n, r->cos, r->sin, positions, slot_mapping,
                           w.k_cache, w.v_cache, B, H, Hkv, D, D, 0, c.cos_sin_f32, st, c.kv_fp8);
    }
    launch_paged_attn_decode(r->maps[l], r->attn, q, qkv_n, block_tables, context_lens, B, H, Hkv,
                             D, c.max_blocks_per_seq, c.attn_scale, 0.f, 0, r->attn_ws, st);
    // o_proj (row-parallel) -> all-reduce -> residual add -> post-attention norm
    if (!row_linear_reduce_norm(r, 0, r->attn, w.o_w, w.o_s, w.o_z, B, H * D, w.post_norm, st)) {
      linear(r, r->attn, w.o_w, w.o_s, w.o_z, nullptr, r->proj, B, H * D, c.hidden, B200_EPI_NONE,
             stream);
      if (r->allreduce) r->allreduce(r->proj, static_cast<int64_t>(B) * c.hidden, r->allreduce_user, stream);
      b200_add_rmsnorm_f16(r->xn, r->x, r->proj, w.post_norm, B, c.hidden, c.rms_eps, stream);
    }
    // ---- MLP block
    const void* next_norm = l + 1 < c.num_layers ? r->layers[l + 1].input_norm : r->final_norm;
    if (c.bits == 16) {
      __half* gu = r->act;   // [B, 2*inter]
      linear(r, r->xn, w.gate_up_w, nullptr, nullptr, nullptr, gu, B, c.hidden, 2 * c.inter,
             B200_EPI_NONE, stream);
      b200_silu_mul_f16(r->act2, gu, B, c.inter, stream);
      linear(r, r->act2, w.down_w, nullptr, nullptr, nullptr, r->proj, B, c.inter, c.hidden,
             B200_EPI_NONE, stream);
      if (r->allreduce) r->allreduce(r->proj, static_cast<int64_t>(B) * c.hidden, r->allreduce_user, stream);
    } else {
      linear(r, r->xn, w.gate_up_w, w.gate_up_s, w.gate_up_z, nullptr, r->act, B, c.hidden,
             2 * c.inter, B200_EPI_SILU_MUL, stream);
      // down_proj (row-parallel) -> all-reduce -> residual add -> the NEXT block's norm
      if (row_linear_reduce_norm(r, 1, r->act, w.down_w, w.down_s, w.down_z, B, c.inter, next_norm, st)) {
        normed = true;
      } else {
        linear(r, r->act, w.down_w, w.down_s, w.down_z, nullptr, r->proj, B, c.inter, c.hidden,
               B200_EPI_NONE, stream);
        if (r->allreduce) r->allreduce(r->proj, static_cast<int64_t>(B) * c.hidden, r->allreduce_user, stream);
      }
    }
  }
  if (!normed) b200_add_rmsnorm_f16(r->xn, r->x, r->proj, r->final_norm, B, c.hidden, c.rms_eps, stream);
  float* lg = logits ? logits : r->logits;
  b200_dense_gemm_f16(r->xn, r->lm_head, lg, B, c.vocab, c.hidden, 1, stream);
  if (out_ids) pick_tokens(r, out_ids, lg, B, stream);
  r->launches = launch_count() - launches0;
}

void b200_runner_prefill_device(b200_runner* r, const uint32_t* input_ids,
                                const int64_t* positions, const int64_t* slot_mapping,
                                const uint32_t* cu_seqlens_q, const uint32_t* context_lens,
                                const uint32_t* block_tables, const uint32_t* last_rows,
                                int32_t B, int32_t T, int32_t max_seqlen_q, int32_t max_blocks,
                                float* logits, uint32_t* out_ids, int64_t stream) {
  const b200_model_config& c = r->cfg;
  if (!r->ws || B <= 0 || B > c.max_batch || T <= 0 || T > c.max_prefill_tokens) {
    set_error(1, "runner_prefill: workspace not set or sizes out of range");
    return;
  }
  if (c.kv_fp8 && (!r->kv_smaps_ok || max_blocks > c.max_blocks_per_seq)) {
    set_error(1, "runner_prefill: fp8 KV scratch not available for this step");
    return;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long launches0 = launch_count();
  const int H = c.num_heads, Hkv = c.num_kv_heads, D = c.head_dim;
  const int qkv_n = (H + 2 * Hkv) * D;
  // variant 3: tcgen05 GEMM, one K split, direct fp16 store (token tiles fill the machine)
  const int V = c.bits == 16 ? 0 : 3;
  b200_embedding_f16(r->x, r->embed, input_ids, T, c.hidden, c.vocab, stream);
  for (int l = 0; l < c.num_layers; ++l) {
    const b200_layer_weights& w = r->layers[l];
    if (l == 0)
      b200_rmsnorm_f16(r->xn, r->x, w.input_norm, T, c.hidden, c.rms_eps, stream);
    else
      b200_add_rmsnorm_f16(r->xn, r->x, r->proj, w.input_norm, T, c.hidden, c.rms_eps, stream);
    linear(r, r->xn, w.qkv_w, w.qkv_s, w.qkv_z, c.qkv_bias ? w.qkv_bias : nullptr, r->qkv, T,
           c.hidden, qkv_n, B200_EPI_NONE, stream, V);
    __half* q = r->qkv;
    __half* k = r->qkv + H * D;
    __half* v = r->qkv + (H + Hkv) * D;
    launch_rope_kv_write(q, k, v, qkv_n, qkv_n, qkv_n, r->cos, r->sin, positions, slot_mapping,
                         w.k_cache, w.v_cache, T, H, Hkv, D, D, 0, c.cos_sin_f32, st, c.kv_fp8);
    if (c.kv_fp8) {
      // every key / value of the step is read back from the e4m3 cache (the result does not depend on
      // how the prompt was chunked or which prefix came from the prefix cache): expand the context
      // pages to fp16 scratch blocks, then the fp16 prefill kernels run on those
      launch_kv_fp8_gather(w.k_cache, w.v_cache, r->kv_sk, r->kv_sv, l == 0 ? r->kv_sbt : nullptr, block_tables,
                           context_lens, cu_seqlens_q, B, max_blocks, Hkv, D, st);
      launch_attn_prefill(r->kv_smaps, r->attn, q, qkv_n, cu_seqlens_q, context_lens, r->kv_sbt, B,
                          H, Hkv, D, max_blocks, max_seqlen_q, c.attn_scale, st, T);
    } else {
      launch_attn_prefill(r->maps[l], r->attn, q, qkv_n, cu_seqlens_q, context_lens, block_tables, B,
                          H, Hkv, D, max_blocks, max_seqlen_q, c.attn_scale, st, T);
    }
    linear(r, r->attn, w.o_w, w.o_s, w.o_z, nullptr, r->proj, T, H * D, c.hidden, B200_EPI_NONE,
           stream, V);
    if (r->allreduce) r->allreduce(r->proj, static_cast<int64_t>(T) * c.hidden, r->allreduce_user, stream);
    b200_add_rmsnorm_f16(r->xn, r->x, r->proj, w.post_norm, T, c.hidden, c.rms_eps, stream);
    // gate|up with SiLU * mul in the GEMM epilogue when the step is large enough for the dequant + dense
    // route (the dequant pass interleaves gate / up rows so both land in one accumulator tile)
    bool mlp_fused = false;
    if (r->wdeq && T >= prefill_dense_min_tokens() && c.inter % kTileN == 0 && c.hidden % kChunkK == 0 &&
        prefill_silu_fused()) {
      const int g = c.group_size <= 0 ? c.hidden : c.group_size;
      launch_wna16_dequant((const uint32_t*)w.gate_up_w, (const __half*)w.gate_up_s, (const uint32_t*)w.gate_up_z,
                           r->wdeq, c.hidden, 2 * c.inter, c.bits, g, c.qfmt, st, 1);
      mlp_fused = launch_dense_gemm_tc_silu(r->xn, r->wdeq, r->act2, T, 2 * c.inter, c.hidden, 2, st);
    }
    if (!mlp_fused) {
      linear(r, r->xn, w.gate_up_w, w.gate_up_s, w.gate_up_z, nullptr, r->act, T, c.hidden,
             2 * c.inter, B200_EPI_NONE, stream, V);
      b200_silu_mul_f16(r->act2, r->act, T, c.inter, stream);
    }
    linear(r, r->act2, w.down_w, w.down_s, w.down_z, nullptr, r->proj, T, c.inter, c.hidden,
           B200_EPI_NONE, stream, V);
    if (r->allreduce) r->allreduce(r->proj, static_cast<int64_t>(T) * c.hidden, r->allreduce_user, stream);
  }
  // last token of every sequence (llama.rs:306-310), then residual add + final norm + lm_head
  __half* xs = r->xn;      // [B, hidden] gathered residual stream
  __half* ps = r->attn;    // [B, hidden] gathered last MLP output
  __half* ns = r->qkv;     // [B, hidden] normed
  b200_gather_rows_f16(xs, r->x, last_rows, B, c.hidden, stream);
  b200_gather_rows_f16(ps, r->proj, last_rows, B, c.hidden, stream);
  b200_add_rmsnorm_f16(ns, xs, ps, r->final_norm, B, c.hidden, c.rms_eps, stream);
  float* lg = logits ? logits : r->logits;
  b200_dense_gemm_f16(ns, r->lm_head, lg, B, c.vocab, c.hidden, 1, stream);
  if (out_ids) pick_tokens(r, out_ids, lg, B, stream);
  r->launches = launch_count() - launches0;
}

static int bucket_for(const b200_runner* r, int B) {
  for (auto& g : r->graphs)
    if (g.first >= B) return g.first;   // std::map iterates ascending: smallest bucket >= B
  return 0;
}

int32_t b200_runner_decode_host(b200_runner* r, const uint32_t* input_ids,
                                const int64_t* positions, const int64_t* slot_mapping,
                                const uint32_t* block_tables, int32_t max_blocks,
                                const uint32_t* context_lens, int32_t B, uint32_t* out_ids_host,
                                int64_t stream) {
  const b200_model_config& c = r->cfg;
  if (!r->ws || B <= 0 || B > c.max_batch || max_blocks > c.max_blocks_per_seq) {
    set_error(1, "runner_decode_host: bad arguments");
    return -1;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int bucket = bucket_for(r, B);
  const int rows = bucket ? bucket : B;
  if (rows > B) {
    // graph.rs:741-767: padded rows carry context_len 0; slot -1 skips their cache write
    cudaMemsetAsync(r->s_ids + B, 0, (rows - B) * 4, st);
    cudaMemsetAsync(r->s_ctx + B, 0, (rows - B) * 4, st);
    cudaMemsetAsync(r->s_pos + B, 0, (rows - B) * 8, st);
    cudaMemsetAsync(r->s_slot + B, 0xFF, (rows - B) * 8, st);
  }
  cudaMemcpyAsync(r->s_ids, input_ids, B * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(r->s_pos, positions, B * 8, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(r->s_slot, slot_mapping, B * 8, cudaMemcpyHostToDevice, st);
  cudaMemcpyAsync(r->s_ctx, context_lens, B * 4, cudaMemcpyHostToDevice, st);
  cudaMemcpy2DAsync(r->s_bt, static_cast<size_t>(c.max_blocks_per_seq) * 4, block_tables,
                    static_cast<size_t>(max_blocks) * 4, static_cast<size_t>(max_blocks) * 4, B,
                    cudaMemcpyHostToDevice, st);
  if (bucket) {
    cudaGraphLaunch(r->graphs[bucket], st);
  } else {
    b200_runner_decode_device(r, r->s_ids, r->s_pos, r->s_slot, r->s_bt, r->s_ctx, B, nullptr,
                              r->s_out, stream);
  }
  cudaMemcpyAsync(out_ids_host, r->s_out, B * 4, cudaMemcpyDeviceToHost, st);
  const cudaError_t e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) {
    set_error(2, cudaGetErrorString(e));
    return -1;
  }
  return b200_last_error() ? -1 : 0;
}

int32_t b200_runner_capture(b200_runner* r, const int32_t* buckets, int32_t num_buckets,
                            int64_t stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!r->ws) {
    set_error(1, "runner_capture: workspace not set");
    return -1;
  }
  for (int i = num_buckets - 1; i >= 0; --i) {   // descending like graph.rs:448-690
    const int bs = buckets[i];
    if (bs <= 0 || bs > r->cfg.max_batch) continue;
    // benign inputs for warm-up + capture: context_len 0 rows, no cache writes
    cudaMemsetAsync(r->s_ids, 0, bs * 4, st);
    cudaMemsetAsync(r->s_ctx, 0, bs * 4, st);
    cudaMemsetAsync(r->s_pos, 0, bs * 8, st);
    cudaMemsetAsync(r->s_slot, 0xFF, bs * 8, st);
    cudaMemsetAsync(r->s_bt, 0, static_cast<size_t>(bs) * r->cfg.max_blocks_per_seq * 4, st);
    b200_runner_decode_device(r, r->s_ids, r->s_pos, r->s_slot, r->s_bt, r->s_ctx, bs, nullptr,
                              r->s_out, stream);   // warm-up (creates cuBLAS handle etc.)
    if (cudaStreamSynchronize(st) != cudaSuccess || b200_last_error()) return -1;
    cudaGraph_t graph = nullptr;
    if (cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      set_error(2, "runner_capture: cudaStreamBeginCapture failed");
      return -1;
    }
    b200_runner_decode_device(r, r->s_ids, r->s_pos, r->s_slot, r->s_bt, r->s_ctx, bs, nullptr,
                              r->s_out, stream);
    if (cudaStreamEndCapture(st, &graph) != cudaSuccess || !graph) {
      set_error(2, "runner_capture: cudaStreamEndCapture failed");
      return -1;
    }
    cudaGraphExec_t exec = nullptr;
    if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
      cudaGraphDestroy(graph);
      set_error(2, "runner_capture: cudaGraphInstantiate failed");
      return -1;
    }
    cudaGraphDestroy(graph);
    auto it = r->graphs.find(bs);
    if (it != r->graphs.end()) cudaGraphExecDestroy(it->second);
    r->graphs[bs] = exec;
  }
  return 0;
}

int64_t b200_runner_launches_per_step(const b200_runner* r) { return r->launches; }

static void set_sampling_impl(b200_runner* r, bool all, int32_t top_k, float top_p, float temperature) {
  // k, p and 1/t are by-value kernel arguments, i.e. baked into captured graphs: ANY change drops them
  if (top_k != r->samp_k || all != r->samp_all || top_p != r->samp_p || temperature != r->samp_t) {
    for (auto& g : r->graphs) cudaGraphExecDestroy(g.second);
    r->graphs.clear();
  }
  r->samp_all = all;
  r->samp_k = top_k;
  r->samp_p = top_p;
  r->samp_t = temperature;
}

void b200_runner_set_sampling(b200_runner* r, int32_t top_k, float top_p, float temperature) {
  if (top_k < 0 || top_k > 256 || (top_k > 0 && !(temperature > 0.f))) {
    set_error(1, "runner_set_sampling: top_k must be 0 (greedy) or 1..256 with a positive temperature");
    return;
  }
  set_sampling_impl(r, false, top_k, top_p, temperature);
}

// strategy = the return value of b200_sampling_strategy (0 ArgMax, 1 All, 2 TopK, 3 TopP, 4 TopKThenTopP) with
// its (k, p, t): ArgMax -> greedy, All -> temperature-only multinomial over the whole vocabulary.
int32_t b200_runner_set_sampling_strategy(b200_runner* r, int32_t strategy, int32_t k, float p, float t) {
  if (strategy == 0) {
    set_sampling_impl(r, false, 0, 1.f, 1.f);
    return 0;
  }
  if (strategy < 0 || strategy > 4 || !(t > 0.f) || (strategy != 1 && (k < 1 || k > 256))) {
    set_error(1, "runner_set_sampling_strategy: unknown strategy or parameters out of range");
    return -1;
  }
  set_sampling_impl(r, strategy == 1, strategy == 1 ? 0 : k, p, t);
  return 0;
}

void b200_runner_set_step_seed(b200_runner* r, uint64_t seed, int64_t stream) {
  if (!r->ws) {
    set_error(1, "runner_set_step_seed: workspace not set");
    return;
  }
  cudaMemcpyAsync(r->s_seed, &seed, 8, cudaMemcpyHostToDevice, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
