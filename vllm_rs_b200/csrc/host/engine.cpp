// The engine's decode step as ONE native call: the 3-phase loop of
//   /root/reference/src/core/engine.rs:1693-1757   (prepare_step -> run_forward -> finish_step)
// for a decode batch: Scheduler::schedule (scheduler.rs:200-380), ModelRunner::prepare_decode
// (core/runner.rs:1243-1388) into pinned staging owned by this handle, ModelRunner::run
// (b200_runner_decode_host: H2D, CUDA-graph replay, D2H, stream sync), Scheduler::postprocess and
// clear_finished (scheduler.rs:500-654).  A host that drives the library from another language pays one
// FFI crossing per step instead of five.  Prefill batches are handed back to the caller (their
// metadata is built by b200_prepare_prefill and run by b200_runner_prefill_device).
//
// Built only on the public C ABI of the scheduler and the runner (no access to their internals).
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "../../../include/b200_vllm.h"

namespace b200 {
void set_error(int code, const char* msg);
}

struct b200_engine {
  b200_scheduler* sched = nullptr;
  b200_runner* runner = nullptr;
  int32_t max_batch = 0, max_blocks = 0;
  // pinned staging of one decode step (prepare_decode output / sampled ids)
  uint32_t *input_ids = nullptr, *context_lens = nullptr, *block_tables = nullptr, *out_ids = nullptr;
  int64_t *positions = nullptr, *slot_mapping = nullptr;
  std::vector<int32_t> indices;
  void* pinned = nullptr;
  // CPU swap (runner.rs:1626-1645): per-layer GPU and pinned-host caches registered by the host, and a pinned
  // list of (src, dst) block pairs the copy kernel reads over UVA
  std::vector<void*> gpu_k, gpu_v, cpu_k, cpu_v;
  int64_t block_bytes = 0;
  int64_t* pairs = nullptr;
  // non-greedy sampling: the engine owns the LogitsProcessor's StdRng (logits_processor.rs:48-70,222-225)
  b200_rng* rng = nullptr;
};
static constexpr int32_t kMaxSwapPairs = 8192;

extern "C" {

b200_engine* b200_engine_new(b200_scheduler* sched, b200_runner* runner, int32_t max_batch,
                             int32_t max_blocks_per_seq) {
  if (!sched || !runner || max_batch <= 0 || max_blocks_per_seq <= 0) {
    b200::set_error(1, "engine_new: bad arguments");
    return nullptr;
  }
  auto* e = new b200_engine();
  e->sched = sched;
  e->runner = runner;
  e->max_batch = max_batch;
  e->max_blocks = max_blocks_per_seq;
  const size_t B = static_cast<size_t>(max_batch);
  const size_t bytes = B * (4 + 4 + 4 + 8 + 8) + B * static_cast<size_t>(max_blocks_per_seq) * 4 + 64 +
                       static_cast<size_t>(kMaxSwapPairs) * 16;
  if (cudaHostAlloc(&e->pinned, bytes, cudaHostAllocDefault) != cudaSuccess) {
    cudaGetLastError();
    delete e;
    b200::set_error(2, "engine_new: cudaHostAlloc failed");
    return nullptr;
  }
  uint8_t* p = static_cast<uint8_t*>(e->pinned);
  e->positions = reinterpret_cast<int64_t*>(p);           p += B * 8;
  e->slot_mapping = reinterpret_cast<int64_t*>(p);        p += B * 8;
  e->input_ids = reinterpret_cast<uint32_t*>(p);          p += B * 4;
  e->context_lens = reinterpret_cast<uint32_t*>(p);       p += B * 4;
  e->out_ids = reinterpret_cast<uint32_t*>(p);            p += B * 4;
  e->block_tables = reinterpret_cast<uint32_t*>(p);       p += B * static_cast<size_t>(max_blocks_per_seq) * 4;
  e->pairs = reinterpret_cast<int64_t*>((reinterpret_cast<uintptr_t>(p) + 15) & ~uintptr_t(15));
  // schedule() may return up to max(max_num_seqs, 5) indices (scheduler.rs:44)
  e->indices.resize(static_cast<size_t>(max_batch < 5 ? 5 : max_batch));
  return e;
}

void b200_engine_free(b200_engine* e) {
  if (!e) return;
  if (e->pinned) cudaFreeHost(e->pinned);
  if (e->rng) b200_rng_free(e->rng);
  delete e;
}

// Register the caches the CPU swap copies between: per layer the GPU K / V cache and its pinned-host mirror
// ([num_blocks, 64, Hkv, D] each, kvcache_allocator.rs:150-173; num_cpu_blocks rows on the host side).
int32_t b200_engine_set_swap_caches(b200_engine* e, int32_t num_layers, void* const* gpu_k, void* const* gpu_v,
                                    void* const* cpu_k, void* const* cpu_v, int64_t block_bytes) {
  if (!e || num_layers <= 0 || block_bytes <= 0 || (block_bytes & 15)) {
    b200::set_error(1, "engine_set_swap_caches: bad arguments");
    return -1;
  }
  e->gpu_k.assign(gpu_k, gpu_k + num_layers);
  e->gpu_v.assign(gpu_v, gpu_v + num_layers);
  e->cpu_k.assign(cpu_k, cpu_k + num_layers);
  e->cpu_v.assign(cpu_v, cpu_v + num_layers);
  e->block_bytes = block_bytes;
  return 0;
}

// Seed of the sampling stream (SamplingParams.seed -> StdRng::seed_from_u64, logits_processor.rs:48-70): once
// set, every decode step pushes rng.next_u64() to the runner before the forward (logits_processor.rs:222-225).
void b200_engine_set_seed(b200_engine* e, uint64_t seed) {
  if (e->rng) b200_rng_free(e->rng);
  e->rng = b200_rng_new(seed);
}

// The block copies schedule() queued (try_swap_out / try_swap_in) MUST run before the step's forward: a
// swapped-out victim's GPU blocks may already belong to another sequence of this very batch, whose KV write
// would otherwise land before the GPU -> CPU copy; a swapped-in sequence is decoded in this step.
static int32_t drain_swap_ops(b200_engine* e, int64_t stream) {
  bool any = false;
  for (;;) {
    int32_t swap_in = 0;
    const int32_t np = b200_scheduler_pop_swap_op(e->sched, e->pairs, kMaxSwapPairs, &swap_in);
    if (np == 0) break;
    if (np < 0) {
      b200::set_error(1, "engine_step: a swap op has more block pairs than the staging holds");
      return -1;
    }
    if (e->gpu_k.empty()) {
      b200::set_error(1, "engine_step: the scheduler queued KV swap copies but no caches are registered "
                         "(b200_engine_set_swap_caches)");
      return -1;
    }
    if (any) cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream));   // the pair list is reused
    for (size_t l = 0; l < e->gpu_k.size(); ++l) {
      b200_swap_blocks(swap_in ? e->cpu_k[l] : e->gpu_k[l], swap_in ? e->gpu_k[l] : e->cpu_k[l], e->pairs, np,
                       e->block_bytes, stream);
      b200_swap_blocks(swap_in ? e->cpu_v[l] : e->gpu_v[l], swap_in ? e->gpu_v[l] : e->cpu_v[l], e->pairs, np,
                       e->block_bytes, stream);
    }
    any = true;
  }
  if (any && cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)) != cudaSuccess) {
    b200::set_error(2, "engine_step: KV swap copies failed");
    return -1;
  }
  return 0;
}

// Returns the number of sequences decoded this step (their running indices at schedule time in indices_out,
// their sequence ids in seq_ids_out, their new tokens in tokens_out, each sized >= max(max_batch, 5) or
// NULL); 0 when nothing is schedulable; -2 when the
// scheduler produced a PREFILL batch: *num_prefill indices are in indices_out and the caller runs the
// prefill step (and its filter_prefill_finished / postprocess) itself; -1 on error.
int32_t b200_engine_step(b200_engine* e, int32_t* indices_out, int64_t* seq_ids_out, uint32_t* tokens_out,
                         int32_t* num_prefill, int64_t stream) {
  if (num_prefill) *num_prefill = 0;
  int32_t is_prefill = 0;
  const int32_t n = b200_scheduler_schedule(e->sched, e->indices.data(), static_cast<int32_t>(e->indices.size()),
                                            &is_prefill);
  if (n < 0) return -1;
  if (drain_swap_ops(e, stream) != 0) return -1;
  if (n == 0) return 0;
  if (indices_out) std::memcpy(indices_out, e->indices.data(), sizeof(int32_t) * n);
  if (is_prefill) {
    if (num_prefill) *num_prefill = n;
    return -2;
  }
  if (seq_ids_out)
    for (int32_t i = 0; i < n; ++i)
      b200_scheduler_seq_info(e->sched, e->indices[i], &seq_ids_out[i], nullptr, nullptr, nullptr);
  if (n > e->max_batch) {
    b200::set_error(1, "engine_step: the decode batch is larger than the runner's max_batch");
    return -1;
  }
  const int32_t mb = b200_prepare_decode(e->sched, e->indices.data(), n, e->input_ids, e->positions,
                                         e->slot_mapping, e->context_lens, e->block_tables, e->max_blocks,
                                         nullptr);
  if (mb < 0) {
    b200::set_error(1, "engine_step: block table wider than max_blocks_per_seq");
    return -1;
  }
  // Reference quirk (scheduler.rs:872-876): a swapped-out sequence that can never come back is marked Finished and
  // pushed to `running` inside schedule(), so it is scheduled once more with a STALE block table.  The metadata
  // mirror stays bit-exact; this loop neutralises the row the way padded graph rows are (graph.rs:741-767):
  // slot -1 = no cache write, context 0 = no attention, so the stale slot's new owner is not overwritten.
  std::vector<int32_t> live_idx;
  std::vector<int32_t> live_pos;
  for (int32_t i = 0; i < n; ++i) {
    int32_t status = 0;
    b200_scheduler_seq_info(e->sched, e->indices[i], nullptr, nullptr, &status, nullptr);
    if (status == 2) {
      e->slot_mapping[i] = -1;
      e->context_lens[i] = 0;
    } else {
      live_idx.push_back(e->indices[i]);
      live_pos.push_back(i);
    }
  }
  if (e->rng) b200_runner_set_step_seed(e->runner, b200_rng_next_u64(e->rng), stream);
  if (b200_runner_decode_host(e->runner, e->input_ids, e->positions, e->slot_mapping, e->block_tables, mb,
                              e->context_lens, n, e->out_ids, stream) != 0)
    return -1;
  // rows that were already Finished (the quirk above) carry a stale block table: postprocess would hand it to
  // cache_sequence / deallocate if their garbage token happened to be EOS -- they are only cleared
  if (static_cast<int32_t>(live_idx.size()) == n) {
    b200_scheduler_postprocess(e->sched, e->indices.data(), n, e->out_ids);
  } else {
    std::vector<uint32_t> live_tok(live_idx.size());
    for (size_t i = 0; i < live_idx.size(); ++i) live_tok[i] = e->out_ids[live_pos[i]];
    b200_scheduler_postprocess(e->sched, live_idx.data(), static_cast<int32_t>(live_idx.size()), live_tok.data());
  }
  b200_scheduler_clear_finished(e->sched);
  if (tokens_out) std::memcpy(tokens_out, e->out_ids, sizeof(uint32_t) * n);
  return n;
}

}  // extern "C"
