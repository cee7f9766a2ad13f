// Host-side integer state of the hot path: Sequence, BlockManager, Scheduler and the
// per-step input preparation.  C++ mirror of
//   /root/reference/src/core/sequence.rs           (Sequence, num_blocks, last_block_num_tokens)
//   /root/reference/src/core/block_manager.rs:51-299 (FIFO free list, allocate_fresh,
//        may_append, can_append, deallocate in reverse table order)
//   /root/reference/src/core/scheduler.rs:186-380,500-654,718-785
//        (add, schedule, postprocess, clear_finished, filter_prefill_finished)
//   /root/reference/src/core/runner.rs:952-1388 (prepare_block_tables / prefill / decode)
// Block ids, block tables, slot mappings, context lens and cu_seqlens must be BIT-EXACT
// with the reference; tests/ compare this file against oracle/ on scripted traces.
// Out of scope here: prefix cache, CPU swap, PD transfer, mamba state,
// tool-call / stop-sequence detection (EOS and max_tokens are handled).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <deque>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../../include/b200_vllm.h"

namespace b200 {
void set_error(int code, const char* msg);
}

namespace {

enum class Status : int32_t { Waiting = 0, Running = 1, Finished = 2 };

struct Sequence {
  int64_t id = 0;
  Status status = Status::Waiting;
  std::vector<uint32_t> token_ids;
  std::vector<uint32_t> output_ids;
  std::vector<uint32_t> block_table;
  size_t num_cached_tokens = 0;
  uint32_t last_token = 0;
  size_t block_size = 64;
  size_t max_tokens = 16384;   // sampling_params.max_tokens.unwrap_or(16384), scheduler.rs:605

  size_t len() const { return token_ids.size(); }
  size_t num_blocks() const { return (len() + block_size - 1) / block_size; }
  size_t last_block_num_tokens() const { return len() - (num_blocks() - 1) * block_size; }
  size_t num_cached_blocks() const { return num_cached_tokens / block_size; }
  void append_token(uint32_t t) {
    token_ids.push_back(t);
    output_ids.push_back(t);
    last_token = t;
  }
};

class BlockManager {
 public:
  BlockManager(size_t num_blocks, size_t block_size) : block_size_(block_size) {
    ref_.assign(num_blocks, 0);
    for (size_t i = 0; i < num_blocks; ++i) free_.push_back(i);
  }
  size_t num_free() const { return free_.size(); }
  bool can_allocate(const Sequence& s) const { return free_.size() >= s.num_blocks(); }
  // allocate_fresh, block_manager.rs:122-137
  bool allocate(Sequence& s) {
    s.num_cached_tokens = 0;
    const size_t need = s.num_blocks();
    for (size_t i = 0; i < need; ++i) {
      if (free_.empty()) return false;
      const size_t id = free_.front();
      free_.pop_front();
      take(id);
      s.block_table.push_back(static_cast<uint32_t>(id));
    }
    return true;
  }
  // block_manager.rs:227-231: reverse table order, freed ids go to the BACK of the list
  void deallocate(const Sequence& s) {
    for (auto it = s.block_table.rbegin(); it != s.block_table.rend(); ++it) {
      const size_t id = *it;
      if (ref_[id] > 0) --ref_[id];
      if (ref_[id] == 0 && used_.erase(id)) free_.push_back(id);
    }
  }
  // block_manager.rs:233-239
  bool can_append(const Sequence& s) const {
    size_t need = 1;
    if (s.len() % block_size_ != 0) need += 1;
    return free_.size() >= need;
  }
  // block_manager.rs:241-253
  bool may_append(Sequence& s) {
    if (s.len() % block_size_ == 1) {
      if (free_.empty()) return false;
      const size_t id = free_.front();
      free_.pop_front();
      take(id);
      s.block_table.push_back(static_cast<uint32_t>(id));
    }
    return true;
  }

 private:
  void take(size_t id) {
    ref_[id] = 1;
    used_.insert(id);
  }
  size_t block_size_;
  std::vector<uint32_t> ref_;
  std::deque<size_t> free_;
  std::unordered_set<size_t> used_;
};

constexpr size_t kMinNumScheduledReqs = 5;   // scheduler.rs:44

}  // namespace

struct b200_scheduler {
  std::deque<Sequence> waiting;
  std::vector<Sequence> running;
  BlockManager bm;
  int64_t next_seq_id = 0;
  std::vector<uint32_t> eos;
  size_t block_size, max_num_seqs, max_num_batched_tokens, chunk_size;
  bool is_last_prefill = false;
  std::string err;

  b200_scheduler(size_t nb, size_t bs, size_t mns, size_t mnbt, size_t cs)
      : bm(nb, bs), block_size(bs), max_num_seqs(mns), max_num_batched_tokens(mnbt),
        chunk_size(cs) {}
};

extern "C" {

b200_scheduler* b200_scheduler_new(int32_t num_blocks, int32_t block_size, int32_t max_num_seqs,
                                   int32_t max_num_batched_tokens, int32_t chunk_size,
                                   const uint32_t* eos_ids, int32_t num_eos) {
  if (num_blocks <= 0 || block_size <= 0 || max_num_seqs <= 0 || chunk_size <= 0) {
    b200::set_error(1, "scheduler_new: bad arguments");
    return nullptr;
  }
  auto* s = new b200_scheduler(num_blocks, block_size, max_num_seqs,
                               max_num_batched_tokens > 0
                                   ? static_cast<size_t>(max_num_batched_tokens)
                                   : static_cast<size_t>(num_blocks) * block_size,
                               chunk_size);
  for (int i = 0; i < num_eos; ++i) s->eos.push_back(eos_ids[i]);
  return s;
}

void b200_scheduler_free(b200_scheduler* s) { delete s; }

int64_t b200_scheduler_add(b200_scheduler* s, const uint32_t* tokens, int32_t num_tokens,
                           int32_t max_tokens) {
  if (num_tokens <= 0) {
    b200::set_error(1, "scheduler_add: empty prompt");
    return -1;
  }
  Sequence seq;
  seq.block_size = s->block_size;
  seq.token_ids.assign(tokens, tokens + num_tokens);
  seq.last_token = seq.token_ids.back();
  if (max_tokens > 0) seq.max_tokens = max_tokens;
  seq.id = s->next_seq_id++;
  const int64_t id = seq.id;
  s->waiting.push_back(std::move(seq));
  return id;
}

int32_t b200_scheduler_schedule(b200_scheduler* s, int32_t* indices, int32_t cap,
                                int32_t* is_prefill) {
  std::vector<size_t> scheduled;
  size_t num_tokens = 0;
  const size_t pre_existing_running = s->running.size();
  const size_t max_seqs_limit = std::max(s->max_num_seqs, kMinNumScheduledReqs);
  // ---- prefill phase (scheduler.rs:232-283)
  while (!s->waiting.empty()) {
    Sequence seq = std::move(s->waiting.front());
    s->waiting.pop_front();
    const size_t effective = std::min(s->chunk_size, seq.len() - seq.num_cached_tokens);
    if (s->running.size() >= max_seqs_limit || scheduled.size() >= max_seqs_limit ||
        num_tokens + effective >= s->max_num_batched_tokens - 1 ||
        (seq.block_table.empty() && !s->bm.can_allocate(seq)) ||
        (s->is_last_prefill && pre_existing_running > 0)) {
      s->waiting.push_front(std::move(seq));
      break;
    }
    if (seq.block_table.empty()) {
      if (!s->bm.allocate(seq)) {
        b200::set_error(1, "No free blocks available, retry later!");
        return -1;
      }
    }
    seq.status = Status::Running;
    num_tokens += effective;
    s->running.push_back(std::move(seq));
    scheduled.push_back(s->running.size() - 1);
  }
  if (!scheduled.empty()) {
    s->is_last_prefill = true;
    *is_prefill = 1;
    if (static_cast<int32_t>(scheduled.size()) > cap) return -1;
    for (size_t i = 0; i < scheduled.size(); ++i) indices[i] = static_cast<int32_t>(scheduled[i]);
    return static_cast<int32_t>(scheduled.size());
  }
  // ---- decode phase (scheduler.rs:285-378); swap / preemption is out of scope: a
  // sequence that cannot append is skipped this step exactly as the reference's filter does
  std::vector<size_t> decode;
  for (size_t idx = 0; idx < s->running.size(); ++idx) {
    if (decode.size() >= max_seqs_limit) break;
    Sequence& seq = s->running[idx];
    if (!s->bm.can_append(seq)) continue;
    if (!s->bm.may_append(seq)) {
      b200::set_error(1, "No free blocks available, retry later!");
      return -1;
    }
    decode.push_back(idx);
  }
  s->is_last_prefill = false;
  *is_prefill = 0;
  if (static_cast<int32_t>(decode.size()) > cap) return -1;
  for (size_t i = 0; i < decode.size(); ++i) indices[i] = static_cast<int32_t>(decode[i]);
  return static_cast<int32_t>(decode.size());
}

int32_t b200_scheduler_filter_prefill_finished(b200_scheduler* s, const int32_t* indices,
                                               int32_t n, int32_t* out_pos, int32_t* out_idx) {
  // scheduler.rs:718-785
  std::vector<std::pair<int32_t, int64_t>> finished;   // (position in `indices`, seq id)
  std::vector<int64_t> remove_ids;
  const size_t C = s->chunk_size;
  for (int32_t i = 0; i < n; ++i) {
    const size_t id = indices[i];
    if (id >= s->running.size()) continue;
    const Sequence& seq = s->running[id];
    if (seq.len() < C || seq.num_cached_tokens + C >= seq.len()) {
      finished.emplace_back(i, seq.id);
    } else {
      remove_ids.push_back(seq.id);
      Sequence clone = seq;
      clone.num_cached_tokens += C;
      clone.status = Status::Waiting;
      s->waiting.push_back(std::move(clone));
    }
  }
  s->running.erase(std::remove_if(s->running.begin(), s->running.end(),
                                  [&](const Sequence& q) {
                                    return std::find(remove_ids.begin(), remove_ids.end(), q.id) !=
                                           remove_ids.end();
                                  }),
                   s->running.end());
  int32_t cnt = 0;
  for (auto& f : finished) {
    for (size_t j = 0; j < s->running.size(); ++j)
      if (s->running[j].id == f.second) {
        out_pos[cnt] = f.first;
        out_idx[cnt] = static_cast<int32_t>(j);
        ++cnt;
        break;
      }
  }
  return cnt;
}

void b200_scheduler_postprocess(b200_scheduler* s, const int32_t* indices, int32_t n,
                                const uint32_t* tokens) {
  // scheduler.rs:500-629 (EOS / max_tokens / max_num_batched_tokens finish rules)
  for (int32_t i = 0; i < n; ++i) {
    const size_t idx = indices[i];
    if (idx >= s->running.size()) continue;
    Sequence& seq = s->running[idx];
    const uint32_t tok = tokens[i];
    const bool is_eos = std::find(s->eos.begin(), s->eos.end(), tok) != s->eos.end();
    if (is_eos || seq.output_ids.size() >= seq.max_tokens ||
        seq.len() > s->max_num_batched_tokens) {
      seq.status = Status::Finished;
      s->bm.deallocate(seq);
    } else {
      seq.append_token(tok);
    }
  }
}

void b200_scheduler_clear_finished(b200_scheduler* s) {
  s->running.erase(std::remove_if(s->running.begin(), s->running.end(),
                                  [](const Sequence& q) { return q.status == Status::Finished; }),
                   s->running.end());
  s->waiting.erase(std::remove_if(s->waiting.begin(), s->waiting.end(),
                                  [](const Sequence& q) { return q.status == Status::Finished; }),
                   s->waiting.end());
}

int32_t b200_scheduler_num_running(const b200_scheduler* s) {
  return static_cast<int32_t>(s->running.size());
}
int32_t b200_scheduler_num_waiting(const b200_scheduler* s) {
  return static_cast<int32_t>(s->waiting.size());
}
int32_t b200_scheduler_num_free_blocks(const b200_scheduler* s) {
  return static_cast<int32_t>(s->bm.num_free());
}

int32_t b200_scheduler_seq_info(const b200_scheduler* s, int32_t i, int64_t* seq_id,
                                int32_t* num_cached_tokens, int32_t* status,
                                int32_t* num_output_tokens) {
  if (i < 0 || static_cast<size_t>(i) >= s->running.size()) return -1;
  const Sequence& q = s->running[i];
  if (seq_id) *seq_id = q.id;
  if (num_cached_tokens) *num_cached_tokens = static_cast<int32_t>(q.num_cached_tokens);
  if (status) *status = static_cast<int32_t>(q.status);
  if (num_output_tokens) *num_output_tokens = static_cast<int32_t>(q.output_ids.size());
  return static_cast<int32_t>(q.len());
}

int32_t b200_scheduler_seq_block_table(const b200_scheduler* s, int32_t i, uint32_t* out,
                                       int32_t cap) {
  if (i < 0 || static_cast<size_t>(i) >= s->running.size()) return -1;
  const auto& bt = s->running[i].block_table;
  const int32_t n = static_cast<int32_t>(bt.size());
  if (out) std::memcpy(out, bt.data(), sizeof(uint32_t) * std::min(n, cap));
  return n;
}

int32_t b200_scheduler_seq_tokens(const b200_scheduler* s, int32_t i, uint32_t* out,
                                  int32_t cap) {
  if (i < 0 || static_cast<size_t>(i) >= s->running.size()) return -1;
  const auto& t = s->running[i].token_ids;
  const int32_t n = static_cast<int32_t>(t.size());
  if (out) std::memcpy(out, t.data(), sizeof(uint32_t) * std::min(n, cap));
  return n;
}

int32_t b200_prepare_decode(const b200_scheduler* s, const int32_t* indices, int32_t n,
                            uint32_t* input_ids, int64_t* positions, int64_t* slot_mapping,
                            uint32_t* context_lens, uint32_t* block_tables,
                            int32_t max_blocks_cap, int32_t* max_context_len) {
  // runner.rs:1243-1280 + prepare_block_tables runner.rs:952-974
  size_t max_len = 0;
  int32_t max_ctx = 0;
  for (int32_t i = 0; i < n; ++i) {
    const Sequence& q = s->running[indices[i]];
    max_len = std::max(max_len, q.block_table.size());
  }
  if (static_cast<int32_t>(max_len) > max_blocks_cap) return -1;
  for (int32_t i = 0; i < n; ++i) {
    const Sequence& q = s->running[indices[i]];
    input_ids[i] = q.last_token;
    positions[i] = static_cast<int64_t>(q.len()) - 1;
    context_lens[i] = static_cast<uint32_t>(q.len());
    max_ctx = std::max(max_ctx, static_cast<int32_t>(q.len()));
    const uint32_t slot = q.block_table.back() * static_cast<uint32_t>(s->block_size) +
                          static_cast<uint32_t>(q.last_block_num_tokens()) - 1u;
    slot_mapping[i] = static_cast<int64_t>(slot);
    uint32_t* row = block_tables + static_cast<size_t>(i) * max_len;
    std::copy(q.block_table.begin(), q.block_table.end(), row);
    std::fill(row + q.block_table.size(), row + max_len, 0u);
  }
  if (max_context_len) *max_context_len = max_ctx;
  return static_cast<int32_t>(max_len);
}

int32_t b200_prepare_prefill(const b200_scheduler* s, const int32_t* indices, int32_t n,
                             int32_t token_cap, uint32_t* input_ids, int64_t* positions,
                             int64_t* slot_mapping, uint32_t* cu_seqlens_q,
                             uint32_t* cu_seqlens_k, uint32_t* context_lens,
                             uint32_t* block_tables, int32_t max_blocks_cap,
                             int32_t* max_blocks, int32_t* max_seqlen_q, int32_t* max_seqlen_k,
                             int32_t* max_context_len, int32_t* has_context) {
  // runner.rs:978-1082
  const int64_t bs = static_cast<int64_t>(s->block_size);
  int64_t T = 0, S = 0;
  int32_t mq = 0, mk = 0, mctx = 0;
  cu_seqlens_q[0] = 0;
  cu_seqlens_k[0] = 0;
  std::vector<size_t> prefill_tokens(n);
  for (int32_t i = 0; i < n; ++i) {
    const Sequence& q = s->running[indices[i]];
    const size_t seqlen = q.len();
    const size_t num_tokens = std::min(s->chunk_size, seqlen - q.num_cached_tokens);
    if (T + static_cast<int64_t>(num_tokens) > token_cap) return -1;
    for (size_t j = 0; j < num_tokens; ++j) {
      input_ids[T + j] = q.token_ids[q.num_cached_tokens + j];
      positions[T + j] = static_cast<int64_t>(q.num_cached_tokens + j);
    }
    prefill_tokens[i] = num_tokens;
    const size_t seqlen_q = num_tokens;
    const size_t seqlen_k = q.num_cached_tokens > 0 ? q.num_cached_tokens + num_tokens : num_tokens;
    mctx = std::max(mctx, static_cast<int32_t>(q.num_cached_tokens + num_tokens));
    cu_seqlens_q[i + 1] = cu_seqlens_q[i] + static_cast<uint32_t>(seqlen_q);
    cu_seqlens_k[i + 1] = cu_seqlens_k[i] + static_cast<uint32_t>(seqlen_k);
    mq = std::max(mq, static_cast<int32_t>(seqlen_q));
    mk = std::max(mk, static_cast<int32_t>(seqlen_k));
    int64_t emitted = 0;
    for (size_t b = q.num_cached_blocks(); b < q.num_blocks(); ++b) {
      int64_t start = static_cast<int64_t>(q.block_table[b] * static_cast<uint32_t>(bs));
      if (b == q.num_cached_blocks()) start += static_cast<int64_t>(q.num_cached_tokens) % bs;
      const int64_t end = start + std::min<int64_t>(static_cast<int64_t>(num_tokens) - emitted, bs);
      for (int64_t v = start; v < end; ++v) {
        if (S >= token_cap) return -1;
        slot_mapping[S++] = v;
      }
      emitted += end - start;
      if (emitted >= static_cast<int64_t>(num_tokens)) break;
    }
    T += static_cast<int64_t>(num_tokens);
  }
  if (S != T) {
    b200::set_error(1, "input_ids and slot_mapping must have same length");
    return -1;
  }
  const bool ctx = cu_seqlens_k[n] > cu_seqlens_q[n];
  if (has_context) *has_context = ctx ? 1 : 0;
  int32_t ml = 0;
  if (ctx) {
    for (int32_t i = 0; i < n; ++i)
      ml = std::max(ml, static_cast<int32_t>(s->running[indices[i]].block_table.size()));
    if (ml > max_blocks_cap) return -1;
    for (int32_t i = 0; i < n; ++i) {
      const Sequence& q = s->running[indices[i]];
      uint32_t* row = block_tables + static_cast<size_t>(i) * ml;
      std::copy(q.block_table.begin(), q.block_table.end(), row);
      std::fill(row + q.block_table.size(), row + ml, 0u);
      context_lens[i] = static_cast<uint32_t>(q.num_cached_tokens + prefill_tokens[i]);
    }
  }
  if (max_blocks) *max_blocks = ml;
  if (max_seqlen_q) *max_seqlen_q = mq;
  if (max_seqlen_k) *max_seqlen_k = mk;
  if (max_context_len) *max_context_len = mctx;
  return static_cast<int32_t>(T);
}

// KVCacheAllocator::plan_allocation for dense models
//   /root/reference/src/utils/kvcache_allocator.rs:126-141 (kv heads per shard), 447-468
//   (per_block_bytes), 553-612 (auto_decide_params), 616-707 (plan_allocation)
int32_t b200_kvcache_plan_allocation(int64_t available, int32_t num_layers, int32_t num_kv_heads,
                                     int32_t head_dim, int32_t num_shards, int32_t block_size,
                                     int32_t dtype_size, int64_t config_model_len,
                                     int64_t user_max_num_seqs, int64_t user_max_model_len,
                                     float cpu_mem_fold, b200_kvcache_plan* out) {
  if (!out) return -1;
  std::memset(out, 0, sizeof(*out));
  const int64_t shards = num_shards <= 0 ? 1 : num_shards;
  const int64_t heads = num_kv_heads >= shards ? num_kv_heads / shards : 1;
  const int64_t bs = block_size;
  const int64_t per_block = bs * heads * head_dim * dtype_size * 2 * num_layers;
  out->per_block_bytes = per_block;
  if (per_block <= 0 || available < 0) return -1;
  const int64_t total_blocks = available / per_block;
  int64_t max_seqs = 0, max_len = 0;
  if (user_max_num_seqs > 0 && user_max_model_len > 0) {
    const int64_t blocks_per_seq = (user_max_model_len + bs - 1) / bs;
    if (user_max_num_seqs * blocks_per_seq * per_block > available) return -2;
    max_seqs = user_max_num_seqs;
    max_len = user_max_model_len;
  } else {
    if (total_blocks == 0) return -2;
    const int64_t total_tokens = total_blocks * bs;
    if (total_tokens <= config_model_len) {
      max_seqs = 1;
      max_len = total_tokens;
    } else {
      const int64_t cand[8] = {config_model_len, config_model_len / 2, config_model_len / 4,
                               config_model_len / 8, 16 * 1024, 8 * 1024, 4 * 1024, 1024};
      max_seqs = 1;
      max_len = total_tokens;
      for (int64_t len : cand) {
        if (len == 0) continue;
        const int64_t blocks_per_seq = (len + bs - 1) / bs;
        if (total_blocks >= blocks_per_seq) {
          max_seqs = std::min<int64_t>(total_blocks / blocks_per_seq, 8);
          max_len = len;
          break;
        }
      }
    }
  }
  if (total_blocks == 0 || max_seqs == 0) return -2;
  out->num_gpu_blocks = total_blocks;
  out->num_cpu_blocks = static_cast<int64_t>(static_cast<float>(total_blocks) * cpu_mem_fold);
  out->max_num_seqs = max_seqs;
  out->max_model_len = max_len;
  out->kvcache_memory_bytes = total_blocks * per_block;
  out->max_num_batched_tokens = total_blocks * bs;
  return 0;
}

}  // extern "C"
