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
};

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
  const size_t bytes = B * (4 + 4 + 4 + 8 + 8) + B * static_cast<size_t>(max_blocks_per_seq) * 4 + 64;
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
  e->block_tables = reinterpret_cast<uint32_t*>(p);
  // schedule() may return up to max(max_num_seqs, 5) indices (scheduler.rs:44)
  e->indices.resize(static_cast<size_t>(max_batch < 5 ? 5 : max_batch));
  return e;
}

void b200_engine_free(b200_engine* e) {
  if (!e) return;
  if (e->pinned) cudaFreeHost(e->pinned);
  delete e;
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
  for (int32_t i = 0; i < n; ++i) {
    int32_t status = 0;
    b200_scheduler_seq_info(e->sched, e->indices[i], nullptr, nullptr, &status, nullptr);
    if (status == 2) {
      e->slot_mapping[i] = -1;
      e->context_lens[i] = 0;
    }
  }
  if (b200_runner_decode_host(e->runner, e->input_ids, e->positions, e->slot_mapping, e->block_tables, mb,
                              e->context_lens, n, e->out_ids, stream) != 0)
    return -1;
  b200_scheduler_postprocess(e->sched, e->indices.data(), n, e->out_ids);
  b200_scheduler_clear_finished(e->sched);
  if (tokens_out) std::memcpy(tokens_out, e->out_ids, sizeof(uint32_t) * n);
  return n;
}

}  // extern "C"
