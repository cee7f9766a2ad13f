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
//   /root/reference/src/core/prefix_cache.rs (hash chain of full blocks + leaf LRU) and
//   block_manager.rs:179-221,275-299,346-442,552-604,657-672 (allocate_with_prefix, cache_sequence,
//        reference-counted sharing, eviction); scheduler.rs:288-337,1167-1175 (eviction under pressure)
//   CPU swap of running sequences (block_manager.rs:255-272,878-1000; scheduler.rs:303-337,830-954): the
//        bookkeeping only -- the block copies are handed to the host as (src, dst) lists for b200_swap_blocks
// Out of scope here: session ("Cached") sequences, PD transfer, mamba state, image seeds,
// tool-call end detection (EOS, max_tokens and stop_token_ids sequences are handled).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <cmath>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../../include/b200_vllm.h"
#include "bincode.hpp"

namespace b200 {
void set_error(int code, const char* msg);
}

namespace {

enum class Status : int32_t { Waiting = 0, Running = 1, Finished = 2, Cached = 3, Swapped = 4, FinishSwapped = 5 };  // sequence.rs:8-15
constexpr int64_t kSwapCoolingPeriodMs = 5000;            // scheduler.rs:49
constexpr size_t kMinKvTokensLeftForSwap = 1000;          // scheduler.rs:50

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
  bool has_swapped_time = false;
  int64_t swapped_time = 0;    // ms, set by swap_out / swap-in
  std::vector<std::vector<uint32_t>> stop_token_ids;   // sampling_params.stop_token_ids

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

// SipHash-1-3 with the all-zero key = Rust's DefaultHasher::new() (prefix_cache.rs:337-349)
struct SipHasher13 {
  uint64_t v0 = 0x736F6D6570736575ull, v1 = 0x646F72616E646F6Dull, v2 = 0x6C7967656E657261ull,
           v3 = 0x7465646279746573ull;
  uint64_t tail = 0;
  size_t ntail = 0, length = 0;
  static uint64_t rotl(uint64_t x, int b) { return (x << b) | (x >> (64 - b)); }
  void round() {
    v0 += v1; v1 = rotl(v1, 13); v1 ^= v0; v0 = rotl(v0, 32);
    v2 += v3; v3 = rotl(v3, 16); v3 ^= v2;
    v0 += v3; v3 = rotl(v3, 21); v3 ^= v0;
    v2 += v1; v1 = rotl(v1, 17); v1 ^= v2; v2 = rotl(v2, 32);
  }
  void write(const uint8_t* p, size_t n) {
    length += n;
    for (size_t i = 0; i < n; ++i) {
      tail |= static_cast<uint64_t>(p[i]) << (8 * ntail);
      if (++ntail == 8) {
        v3 ^= tail; round(); v0 ^= tail;
        tail = 0; ntail = 0;
      }
    }
  }
  void write_u64(uint64_t x) {
    uint8_t b[8];
    for (int i = 0; i < 8; ++i) b[i] = static_cast<uint8_t>(x >> (8 * i));
    write(b, 8);
  }
  uint64_t finish() {
    const uint64_t b = (static_cast<uint64_t>(length & 0xFF) << 56) | tail;
    v3 ^= b; round(); v0 ^= b;
    v2 ^= 0xFF;
    round(); round(); round();
    return v0 ^ v1 ^ v2 ^ v3;
  }
};

// prefix_cache.rs:337-342: parent_hash.hash(h); tokens.hash(h) (slice = usize length + raw LE u32s)
uint64_t hash_block(uint64_t parent_hash, const uint32_t* tokens, size_t n) {
  SipHasher13 h;
  h.write_u64(parent_hash);
  h.write_u64(n);
  for (size_t i = 0; i < n; ++i) {
    uint8_t b[4] = {static_cast<uint8_t>(tokens[i]), static_cast<uint8_t>(tokens[i] >> 8),
                    static_cast<uint8_t>(tokens[i] >> 16), static_cast<uint8_t>(tokens[i] >> 24)};
    h.write(b, 4);
  }
  return h.finish();
}

// prefix_cache.rs:38-350 (no image seeds)
class PrefixCache {
 public:
  PrefixCache(size_t block_size, size_t max_cached) : block_size_(block_size), max_cached_(max_cached) {}
  bool enabled() const { return max_cached_ > 0; }
  size_t cached_blocks() const { return entries_.size(); }

  // match_prefix (:67-116): number of leading full blocks present; touches every hit
  size_t match_prefix(const std::vector<uint32_t>& tokens, bool* has_last, uint64_t* last_hash) {
    *has_last = false;
    if (!enabled()) return 0;
    const size_t full = tokens.size() / block_size_;
    size_t matched = 0;
    uint64_t parent = 0;
    for (size_t i = 0; i < full; ++i) {
      const uint64_t h = hash_block(parent, tokens.data() + i * block_size_, block_size_);
      if (!entries_.count(h)) break;
      ++matched;
      parent = h;
      *has_last = true;
      *last_hash = h;
      touch(h);
    }
    return matched;
  }
  std::vector<size_t> blocks_for_match(uint64_t last_hash) const {       // :118-132
    std::vector<size_t> blocks;
    bool has = true;
    uint64_t cur = last_hash;
    while (has) {
      auto it = entries_.find(cur);
      if (it == entries_.end()) break;
      blocks.push_back(it->second.block_id);
      has = it->second.has_parent;
      cur = it->second.parent;
    }
    std::reverse(blocks.begin(), blocks.end());
    return blocks;
  }
  // insert_prefix (:174-262)
  void insert_prefix(const std::vector<uint32_t>& tokens, const std::vector<size_t>& blocks,
                     std::vector<size_t>* inserted, std::vector<size_t>* evicted) {
    if (!enabled()) return;
    const size_t max_blocks = std::min(tokens.size() / block_size_, blocks.size());
    if (max_blocks == 0) return;
    bool has_parent = false;
    uint64_t parent = 0;
    for (size_t i = 0; i < max_blocks; ++i) {
      const uint64_t h = hash_block(has_parent ? parent : 0, tokens.data() + i * block_size_, block_size_);
      auto it = entries_.find(h);
      if (it != entries_.end()) {
        it->second.access_id = next_access_id();
        touch_leaf(h);
      } else {
        if (has_parent) {
          auto pit = entries_.find(parent);
          if (pit != entries_.end()) {
            if (pit->second.children == 0) leaf_set_.erase(parent);
            pit->second.children += 1;
          }
        }
        const uint64_t a = next_access_id();
        entries_[h] = Entry{has_parent, parent, blocks[i], 0, a};
        leaf_set_.insert(h);
        leaf_lru_.push_back({h, a});
        inserted->push_back(blocks[i]);
      }
      has_parent = true;
      parent = h;
    }
    if (entries_.size() > max_cached_) evict_blocks(entries_.size() - max_cached_, evicted);
  }
  // evict_blocks (:264-295): least-recently-used leaves first, parents become leaves
  void evict_blocks(size_t num, std::vector<size_t>* evicted) {
    while (num > 0) {
      if (leaf_lru_.empty()) break;
      const auto [h, a] = leaf_lru_.front();
      leaf_lru_.pop_front();
      if (!leaf_set_.count(h)) continue;
      auto it = entries_.find(h);
      if (it == entries_.end()) continue;
      if (it->second.access_id != a || it->second.children > 0) continue;
      const Entry e = it->second;
      entries_.erase(it);
      leaf_set_.erase(h);
      evicted->push_back(e.block_id);
      --num;
      if (e.has_parent) {
        auto pit = entries_.find(e.parent);
        if (pit != entries_.end()) {
          if (pit->second.children > 0) pit->second.children -= 1;
          if (pit->second.children == 0) {
            leaf_set_.insert(e.parent);
            leaf_lru_.push_back({e.parent, pit->second.access_id});
          }
        }
      }
    }
  }

 private:
  struct Entry {
    bool has_parent;
    uint64_t parent;
    size_t block_id, children;
    uint64_t access_id;
  };
  uint64_t next_access_id() { return ++access_counter_; }
  void touch(uint64_t h) {                                               // :297-305
    auto it = entries_.find(h);
    if (it == entries_.end()) return;
    it->second.access_id = next_access_id();
    touch_leaf(h);
  }
  void touch_leaf(uint64_t h) {                                          // :307-314
    if (leaf_set_.count(h)) {
      auto it = entries_.find(h);
      if (it != entries_.end()) leaf_lru_.push_back({h, it->second.access_id});
    }
    compact_lru_if_needed();
  }
  void compact_lru_if_needed() {                                         // :316-330
    const size_t threshold = std::max<size_t>(entries_.size(), 64) * 4;
    if (leaf_lru_.size() <= threshold) return;
    std::deque<std::pair<uint64_t, uint64_t>> kept;
    for (const auto& [h, a] : leaf_lru_) {
      if (!leaf_set_.count(h)) continue;
      auto it = entries_.find(h);
      if (it != entries_.end() && it->second.access_id == a) kept.push_back({h, a});
    }
    leaf_lru_.swap(kept);
  }
  size_t block_size_, max_cached_;
  std::unordered_map<uint64_t, Entry> entries_;
  std::unordered_set<uint64_t> leaf_set_;
  std::deque<std::pair<uint64_t, uint64_t>> leaf_lru_;
  uint64_t access_counter_ = 0;
};

class BlockManager {
 public:
  BlockManager(size_t num_blocks, size_t block_size) : block_size_(block_size), num_blocks_(num_blocks) {
    ref_.assign(num_blocks, 0);
    for (size_t i = 0; i < num_blocks; ++i) free_.push_back(i);
  }
  void enable_prefix_cache(size_t max_cached) {                          // block_manager.rs:80-84
    prefix_.reset(max_cached > 0 ? new PrefixCache(block_size_, max_cached) : nullptr);
  }
  size_t num_free() const { return free_.size(); }
  size_t num_total() const { return num_blocks_; }
  size_t ref_count(size_t id) const { return id < ref_.size() ? ref_[id] : 0; }
  size_t prefix_cache_blocks() const { return prefix_ ? prefix_->cached_blocks() : 0; }
  int32_t copy_free(uint32_t* out, int32_t cap) const {
    int32_t i = 0;
    for (size_t id : free_) {
      if (out && i < cap) out[i] = static_cast<uint32_t>(id);
      ++i;
    }
    return i;
  }
  // required_blocks / can_allocate (block_manager.rs:179-204): the prefix match touches the LRU
  bool can_allocate(const Sequence& s) {
    size_t need = s.num_blocks();
    if (prefix_) {
      bool has_last;
      uint64_t last;
      const size_t matched = adjusted(s.len(), prefix_->match_prefix(s.token_ids, &has_last, &last));
      need = need > matched ? need - matched : 0;
    }
    return free_.size() >= need;
  }
  // allocate (block_manager.rs:210-221): allocate_with_prefix (:346-442) or allocate_fresh (:122-137)
  bool allocate(Sequence& s) {
    size_t matched = 0;
    if (prefix_) {
      bool has_last;
      uint64_t last = 0;
      matched = adjusted(s.len(), prefix_->match_prefix(s.token_ids, &has_last, &last));
      if (matched > 0 && has_last) {
        std::vector<size_t> cached = prefix_->blocks_for_match(last);
        if (cached.size() > matched) cached.resize(matched);
        for (size_t id : cached) {
          increment_ref(id);
          s.block_table.push_back(static_cast<uint32_t>(id));
        }
      }
    }
    s.num_cached_tokens = matched * block_size_;
    const size_t need = s.num_blocks();
    for (size_t i = s.block_table.size(); i < need; ++i) {
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
    for (auto it = s.block_table.rbegin(); it != s.block_table.rend(); ++it) decrement_ref(*it);
  }
  // cache_sequence (block_manager.rs:552-604): the cache takes its own reference on new entries
  void cache_sequence(const Sequence& s) {
    if (!prefix_ || !prefix_->enabled()) return;
    const size_t full = s.len() / block_size_;
    if (full == 0 || s.block_table.size() < full) return;
    std::vector<size_t> blocks(s.block_table.begin(), s.block_table.begin() + full);
    std::vector<size_t> inserted, evicted;
    prefix_->insert_prefix(s.token_ids, blocks, &inserted, &evicted);
    for (size_t id : inserted) increment_ref(id);
    for (size_t id : evicted) decrement_ref(id);
  }
  size_t evict_prefix_cache_blocks(size_t n) {                           // :657-672
    if (!prefix_ || n == 0) return 0;
    std::vector<size_t> evicted;
    prefix_->evict_blocks(n, &evicted);
    for (size_t id : evicted) decrement_ref(id);
    return evicted.size();
  }
  // ---- CPU swap space (block_manager.rs:69-78, 255-272, 878-1000)
  struct SwapOp {
    bool swap_in;
    std::vector<std::pair<int64_t, int64_t>> pairs;     // (src block, dst block)
  };
  void enable_cpu_swap(size_t n) {
    free_cpu_.clear();
    for (size_t i = 0; i < n; ++i) free_cpu_.push_back(i);
  }
  int32_t copy_free_cpu(uint32_t* out, int32_t cap) const {
    int32_t i = 0;
    for (size_t id : free_cpu_) {
      if (out && i < cap) out[i] = static_cast<uint32_t>(id);
      ++i;
    }
    return i;
  }
  bool ensure_allocate(Sequence& s) {                                    // :255-272
    for (size_t i = s.block_table.size(); i < s.num_blocks(); ++i) {
      if (free_.empty()) return false;
      const size_t id = free_.front();
      free_.pop_front();
      take(id);
      s.block_table.push_back(static_cast<uint32_t>(id));
    }
    return true;
  }
  bool can_swap_out(const Sequence& s) const {                           // :878-894
    for (uint32_t id : s.block_table)
      if (ref_[id] > 1) return false;
    return free_cpu_.size() > s.num_blocks();
  }
  bool can_swap_in(const Sequence& s) const { return free_.size() > s.num_blocks(); }   // :898-906
  bool swap_out(Sequence& s, int64_t now_ms) {                           // :910-957
    const size_t n = s.block_table.size();
    if (free_cpu_.size() < n) return false;
    SwapOp op;
    op.swap_in = false;
    std::vector<size_t> cpu_ids;
    for (size_t i = 0; i < n; ++i) {
      const size_t c = free_cpu_.front();
      free_cpu_.pop_front();
      cpu_ids.push_back(c);
      op.pairs.emplace_back(static_cast<int64_t>(s.block_table[i]), static_cast<int64_t>(c));
    }
    swap_ops.push_back(std::move(op));
    s.has_swapped_time = true;
    s.swapped_time = now_ms;
    swapped_map_[s.id] = std::move(cpu_ids);
    return true;
  }
  void free_cpu_swap_for_seq(int64_t seq_id) {                           // :993-1000
    auto it = swapped_map_.find(seq_id);
    if (it == swapped_map_.end()) return;
    for (size_t c : it->second) free_cpu_.push_back(c);
    swapped_map_.erase(it);
  }
  bool swap_in(Sequence& s) {                                            // :959-990
    auto it = swapped_map_.find(s.id);
    if (it == swapped_map_.end()) return false;
    std::vector<size_t> cpu_ids = std::move(it->second);
    swapped_map_.erase(it);
    if (cpu_ids.size() > s.block_table.size()) {
      swapped_map_[s.id] = std::move(cpu_ids);
      free_cpu_swap_for_seq(s.id);
      return false;
    }
    SwapOp op;
    op.swap_in = true;
    for (size_t i = 0; i < cpu_ids.size(); ++i)
      op.pairs.emplace_back(static_cast<int64_t>(cpu_ids[i]), static_cast<int64_t>(s.block_table[i]));
    swap_ops.push_back(std::move(op));
    for (size_t c : cpu_ids) free_cpu_.push_back(c);
    return true;
  }
  size_t evict_prefix_cache_until_free(size_t min_free) {                // :674-697
    size_t total = 0;
    while (free_.size() < min_free) {
      if (!prefix_) break;
      std::vector<size_t> evicted;
      prefix_->evict_blocks(1, &evicted);
      if (evicted.empty()) break;
      total += evicted.size();
      for (size_t id : evicted) decrement_ref(id);
    }
    return total;
  }
  std::deque<SwapOp> swap_ops;      // pending block copies, oldest first

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
  size_t adjusted(size_t tokens_len, size_t matched) const {             // :291-299
    const size_t full = tokens_len / block_size_;
    if (matched == full && tokens_len % block_size_ == 0 && matched > 0) return matched - 1;
    return matched;
  }
  void take(size_t id) {                                                 // allocate_block :113-120
    ref_[id] = 1;
    used_.insert(id);
  }
  void increment_ref(size_t id) {                                        // :275-282
    if (ref_[id] == 0) {
      free_.erase(std::remove(free_.begin(), free_.end(), id), free_.end());
      used_.insert(id);
    }
    ref_[id] += 1;
  }
  void decrement_ref(size_t id) {                                        // :283-289, 139-144
    if (ref_[id] > 0) --ref_[id];
    if (ref_[id] == 0 && used_.erase(id)) free_.push_back(id);
  }
  size_t block_size_, num_blocks_;
  std::vector<uint32_t> ref_;
  std::deque<size_t> free_;
  std::unordered_set<size_t> used_;
  std::unique_ptr<PrefixCache> prefix_;
  std::deque<size_t> free_cpu_;
  std::unordered_map<int64_t, std::vector<size_t>> swapped_map_;
};

constexpr size_t kMinNumScheduledReqs = 5;   // scheduler.rs:44

// stop_sequence_match_index (scheduler.rs:1309-1341): a one-token stop matches the token itself, a longer one
// needs its head to equal the tail of the tokens generated so far (output_ids; the prompt does not count)
bool hits_stop_sequence(uint32_t token, const Sequence& seq) {
  for (const auto& stop : seq.stop_token_ids) {
    if (stop.empty()) continue;
    if (stop.size() == 1) {
      if (stop[0] == token) return true;
      continue;
    }
    const size_t prior = seq.output_ids.size();
    if (stop.size() - 1 > prior) continue;
    const size_t start = prior + 1 - stop.size();
    if (std::equal(seq.output_ids.begin() + static_cast<std::ptrdiff_t>(start), seq.output_ids.end(), stop.begin()) &&
        stop.back() == token)
      return true;
  }
  return false;
}

}  // namespace

struct b200_scheduler {
  std::deque<Sequence> waiting;
  std::vector<Sequence> running;
  std::vector<Sequence> cached;        // swapped-out sequences (scheduler.rs `cached`)
  int64_t now_ms = 0;                  // the host's clock (SystemTime::now() in the reference)
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

// try_swap_out (scheduler.rs:904-954), running sequences only
static bool try_swap_out(b200_scheduler* s, size_t idx) {
  if (idx >= s->running.size()) return false;
  Sequence& seq = s->running[idx];
  if (seq.block_table.empty() || seq.status != Status::Running || !s->bm.can_swap_out(seq)) return false;
  if (!s->bm.ensure_allocate(seq)) return false;        // same number of blocks when it comes back
  if (!s->bm.swap_out(seq, s->now_ms)) return false;
  Sequence out = std::move(s->running[idx]);
  s->running.erase(s->running.begin() + static_cast<std::ptrdiff_t>(idx));
  out.status = Status::Swapped;
  s->bm.deallocate(out);
  s->cached.push_back(std::move(out));                  // the block table is rebuilt at swap-in
  return true;
}

// try_swap_in (scheduler.rs:830-901): at most one sequence per schedule() call
static void try_swap_in(b200_scheduler* s) {
  const int64_t cur = s->now_ms;
  size_t i = 0;
  while (i < s->cached.size()) {
    Sequence& c = s->cached[i];
    const int64_t st = c.has_swapped_time ? c.swapped_time : cur;
    if (c.status != Status::Swapped || cur - st < kSwapCoolingPeriodMs) {
      ++i;
      continue;
    }
    const size_t available = s->bm.num_free() * s->block_size;
    if (!s->bm.can_swap_in(c) || available - c.len() < kMinKvTokensLeftForSwap) {
      if (!s->running.empty()) {
        ++i;
        continue;                                       // wait for a later step
      }
      if (s->bm.evict_prefix_cache_until_free(c.num_blocks() + 1) > 0) break;
      Sequence seq = std::move(s->cached[i]);
      s->cached.erase(s->cached.begin() + static_cast<std::ptrdiff_t>(i));
      seq.status = Status::Finished;                    // "No KvCache left for swap in"
      s->running.push_back(std::move(seq));
      break;
    }
    Sequence seq = std::move(s->cached[i]);
    s->cached.erase(s->cached.begin() + static_cast<std::ptrdiff_t>(i));
    seq.has_swapped_time = true;
    seq.swapped_time = cur;
    seq.block_table.clear();
    if (!s->bm.ensure_allocate(seq)) continue;          // (the reference drops the sequence here as well)
    seq.status = s->bm.swap_in(seq) ? Status::Running : Status::Finished;
    s->running.push_back(std::move(seq));
    break;
  }
}

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

// SamplingParams.stop_token_ids of a WAITING sequence (by the id add() returned): `num_stops` token sequences
// stored back to back in stop_tokens, their lengths in stop_lens.  Returns 0, or -1 if the id is not waiting.
int32_t b200_scheduler_set_stop_sequences(b200_scheduler* s, int64_t seq_id, const uint32_t* stop_tokens,
                                          const int32_t* stop_lens, int32_t num_stops) {
  for (Sequence& q : s->waiting) {
    if (q.id != seq_id) continue;
    q.stop_token_ids.clear();
    size_t off = 0;
    for (int32_t i = 0; i < num_stops; ++i) {
      const size_t n = stop_lens[i] > 0 ? static_cast<size_t>(stop_lens[i]) : 0;
      q.stop_token_ids.emplace_back(stop_tokens + off, stop_tokens + off + n);
      off += n;
    }
    return 0;
  }
  return -1;
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
  {
    // scheduler.rs:288-337: with nothing to preempt and room in the cache ONE swapped-out sequence may come
    // back; otherwise, under pressure, 10 % of the prefix cache (at least one block) is evicted
    // (scheduler.rs:1167-1175) and, if there was nothing to evict, the oldest sequence that cannot append is
    // swapped out to CPU memory (only while another one keeps running)
    std::vector<size_t> preempt;
    for (size_t i = 0; i < s->running.size(); ++i) {
      const Sequence& q = s->running[i];
      if (!s->bm.can_append(q) && q.status != Status::Swapped && q.status != Status::FinishSwapped) preempt.push_back(i);
    }
    const float usage = 1.0f - (static_cast<float>(s->bm.num_free()) * 1.0f / static_cast<float>(s->bm.num_total()));
    const bool swap_in = preempt.empty() && (usage < 0.95f * 0.9f || (s->running.empty() && usage <= 0.3f));
    if (swap_in) {
      try_swap_in(s);
    } else if (!preempt.empty() || usage > 0.95f) {
      size_t evicted = 0;
      const size_t cached = s->bm.prefix_cache_blocks();
      if (cached > 0) {
        size_t blocks = static_cast<size_t>(std::ceil(static_cast<float>(cached) * 0.1f));
        if (blocks < 1) blocks = 1;
        evicted = s->bm.evict_prefix_cache_blocks(blocks);
      }
      if (evicted == 0 && !preempt.empty() && s->running.size() > 1) {
        size_t best = preempt[0];
        for (size_t i : preempt)
          if (s->running[i].id < s->running[best].id) best = i;
        try_swap_out(s, best);
      }
    }
  }
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
    if (hits_stop_sequence(tok, seq) || is_eos || seq.output_ids.size() >= seq.max_tokens ||
        seq.len() > s->max_num_batched_tokens) {
      seq.status = Status::Finished;
      s->bm.cache_sequence(seq);     // scheduler.rs:617-620: the cache takes its references first
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

// build_prefix_cache_config (scheduler.rs:60-104) for a plain (non-PD) engine + BlockManager::new
// (block_manager.rs:80-84).  max_tokens < 0: the default 65 % of the blocks (f32 arithmetic as there).
// Returns the resulting capacity in blocks (0 = disabled).  Call before the first add().
int32_t b200_scheduler_enable_prefix_cache(b200_scheduler* s, int64_t max_tokens) {
  size_t max_cached = max_tokens >= 0
                          ? static_cast<size_t>(max_tokens) / s->block_size
                          : static_cast<size_t>(static_cast<float>(s->bm.num_total()) * 0.65f);
  max_cached = std::min(max_cached, s->bm.num_total());
  s->bm.enable_prefix_cache(max_cached);
  return static_cast<int32_t>(max_cached);
}
int32_t b200_scheduler_prefix_cache_blocks(const b200_scheduler* s) {
  return static_cast<int32_t>(s->bm.prefix_cache_blocks());
}
int32_t b200_scheduler_evict_prefix_cache_blocks(b200_scheduler* s, int32_t n) {
  return n > 0 ? static_cast<int32_t>(s->bm.evict_prefix_cache_blocks(n)) : 0;
}
int32_t b200_scheduler_block_ref_count(const b200_scheduler* s, int32_t block_id) {
  return block_id < 0 ? -1 : static_cast<int32_t>(s->bm.ref_count(block_id));
}
// free_block_ids in pop order (the next allocation takes out[0])
int32_t b200_scheduler_free_block_ids(const b200_scheduler* s, uint32_t* out, int32_t cap) {
  return s->bm.copy_free(out, cap);
}
uint64_t b200_prefix_hash_block(uint64_t parent_hash, const uint32_t* tokens, int32_t n) {
  return hash_block(parent_hash, tokens, n > 0 ? n : 0);
}

// ---- CPU swap (block_manager.rs:69-78: num_cpu_blocks CPU blocks, ids 0..n-1)
void b200_scheduler_enable_cpu_swap(b200_scheduler* s, int32_t num_cpu_blocks) {
  s->bm.enable_cpu_swap(num_cpu_blocks > 0 ? static_cast<size_t>(num_cpu_blocks) : 0);
}
void b200_scheduler_set_time_ms(b200_scheduler* s, int64_t now_ms) { s->now_ms = now_ms; }
int32_t b200_scheduler_num_swapped(const b200_scheduler* s) { return static_cast<int32_t>(s->cached.size()); }
int64_t b200_scheduler_swapped_seq_id(const b200_scheduler* s, int32_t i) {
  return (i < 0 || static_cast<size_t>(i) >= s->cached.size()) ? -1 : s->cached[i].id;
}
int32_t b200_scheduler_free_cpu_block_ids(const b200_scheduler* s, uint32_t* out, int32_t cap) {
  return s->bm.copy_free_cpu(out, cap);
}
// Oldest pending block-copy list: returns its number of (src, dst) pairs (0 = nothing pending, -1 = cap too
// small) and removes it; *swap_in = 1: CPU -> GPU block ids, 0: GPU -> CPU.
int32_t b200_scheduler_pop_swap_op(b200_scheduler* s, int64_t* pairs, int32_t cap_pairs, int32_t* swap_in) {
  if (s->bm.swap_ops.empty()) return 0;
  const auto& op = s->bm.swap_ops.front();
  const int32_t n = static_cast<int32_t>(op.pairs.size());
  if (n > cap_pairs) return -1;
  for (int32_t i = 0; i < n; ++i) {
    pairs[2 * i] = op.pairs[i].first;
    pairs[2 * i + 1] = op.pairs[i].second;
  }
  if (swap_in) *swap_in = op.swap_in ? 1 : 0;
  s->bm.swap_ops.pop_front();
  return n;
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

// ---------------------------------------------------------------- RunPrefill frames (runner/mod.rs:178, 248-295)
// MessageType::RunPrefill((Vec<Sequence>, bool)): what the engine sends a runner process for a prefill step
// (`send_local`, u32 length + bincode 1.3.1), with `Sequence` as core/sequence.rs:32-50 declares it.  The engine side
// frames the scheduled sequences; the runner side loads a frame into a scheduler object it uses as a plain sequence
// container and runs b200_prepare_prefill on it -- the same metadata the engine's own prepare_prefill would give.
// Fields the mirror does not keep are written as the reference's defaults (created_time 0 -- the engine's wall
// clock --, mamba_prefix_hash / pd_first_token / images / stop_sequence None, the two flags false) and skipped on
// read; a sequence that carries images is refused (multimodal is out of scope).
static void write_sequence(b200::wire::Wr& w, const Sequence& q) {               // core/sequence.rs:32-50, in order
  using namespace b200::wire;
  w.u64(static_cast<uint64_t>(q.id));
  w.u64(0);                                            // created_time
  w.u8(q.has_swapped_time ? 1 : 0);
  if (q.has_swapped_time) w.u64(static_cast<uint64_t>(q.swapped_time));
  w.u32(static_cast<uint32_t>(q.status));
  for (const std::vector<uint32_t>* v : {&q.token_ids, &q.output_ids, &q.block_table}) {
    w.u64(v->size());
    w.raw(v->data(), static_cast<int64_t>(v->size() * 4));
  }
  w.u64(q.num_cached_tokens);
  w.u8(0);                                             // mamba_prefix_hash
  w.u32(q.last_token);
  w.u64(q.block_size);
  b200_wire_sampling sp;
  std::memset(&sp, 0, sizeof(sp));
  sp.has_max_tokens = 1;
  sp.max_tokens = static_cast<int64_t>(q.max_tokens);
  write_sampling(w, &sp);
  w.u8(0);                                             // pd_first_token
  w.u8(0);                                             // images
  w.u8(0); w.u8(0);                                    // is_tool_call_end, hit_stop_sequence
  w.u8(0);                                             // stop_sequence
}

// 0 ok, -1 malformed, -4 carries images
static int read_sequence(b200::wire::Rd& r, Sequence& q, b200_wire_sampling* sampling) {
  using namespace b200::wire;
  q.id = static_cast<int64_t>(r.u64());
  r.u64();                                             // created_time
  q.has_swapped_time = r.opt();
  if (q.has_swapped_time) q.swapped_time = static_cast<int64_t>(r.u64());
  const uint32_t st = r.u32();
  if (!r.ok || st > 5) return -1;
  q.status = static_cast<Status>(st);
  for (std::vector<uint32_t>* v : {&q.token_ids, &q.output_ids, &q.block_table}) {
    const uint64_t m = r.u64();
    if (!r.ok || m > (1ull << 32) || !r.need(static_cast<int64_t>(m) * 4)) return -1;
    v->resize(m);
    std::memcpy(v->data(), r.p + r.off, m * 4);
    r.off += static_cast<int64_t>(m) * 4;
  }
  q.num_cached_tokens = r.u64();
  if (r.opt()) r.u64();                                // mamba_prefix_hash
  q.last_token = r.u32();
  q.block_size = r.u64();
  b200_wire_sampling sp;
  read_sampling(r, &sp);
  q.max_tokens = sp.has_max_tokens ? static_cast<size_t>(sp.max_tokens) : 16384;
  if (sampling) *sampling = sp;
  if (r.opt()) r.u32();                                // pd_first_token
  if (r.opt()) return -4;                              // images: multimodal sequences are not taken
  r.u8(); r.u8();                                      // is_tool_call_end, hit_stop_sequence
  if (r.opt()) r.skip_string();                        // stop_sequence
  return r.ok ? 0 : -1;
}

int64_t b200_scheduler_frame_run_prefill(const b200_scheduler* s, const int32_t* indices, int32_t n, uint8_t* out,
                                         int64_t cap, int32_t is_prefill) {
  using namespace b200::wire;
  Wr w{out, cap};
  w.u32(0);                                            // frame length, patched below
  w.u32(kRunPrefill);
  w.u64(static_cast<uint64_t>(n));
  for (int32_t i = 0; i < n; ++i) {
    if (indices[i] < 0 || static_cast<size_t>(indices[i]) >= s->running.size()) {
      b200::set_error(1, "scheduler_frame_run_prefill: index out of range");
      return -1;
    }
    write_sequence(w, s->running[indices[i]]);
  }
  w.u8(is_prefill ? 1 : 0);
  if (!w.ok) {
    b200::set_error(1, "scheduler_frame_run_prefill: buffer too small");
    return -1;
  }
  const uint32_t body = static_cast<uint32_t>(w.off - 4);
  std::memcpy(out, &body, 4);
  return w.off;
}

int32_t b200_scheduler_load_run_prefill(b200_scheduler* s, const uint8_t* frame, int64_t len, int32_t* indices,
                                        int32_t cap, int32_t* is_prefill, b200_wire_sampling* first_sampling) {
  using namespace b200::wire;
  Rd r{frame, len};
  const uint32_t body = r.u32();
  if (!r.ok || static_cast<int64_t>(body) + 4 != len || r.u32() != kRunPrefill) return -1;
  const uint64_t n = r.u64();
  if (!r.ok || n > static_cast<uint64_t>(cap)) return -2;
  std::vector<Sequence> seqs(n);
  for (uint64_t i = 0; i < n; ++i) {
    Sequence& q = seqs[i];
    const int rc = read_sequence(r, q, i == 0 ? first_sampling : nullptr);   // the step's strategy: its first sequence
    if (rc != 0) return rc;
    if (q.block_size != s->block_size || q.num_cached_tokens > q.token_ids.size() ||
        q.block_table.size() < q.num_blocks())
      return -3;                                       // not a sequence prepare_prefill can walk
  }
  const uint8_t pf = r.u8();
  if (!r.ok || r.off != len) return -1;
  s->running = std::move(seqs);
  for (uint64_t i = 0; i < n; ++i) indices[i] = static_cast<int32_t>(i);
  if (is_prefill) *is_prefill = pf;
  return static_cast<int32_t>(n);
}

// ---------------------------------------------------------------- single-Sequence messages of the PD hand-off
// family 0 = runner IPC `MessageType` (runner/mod.rs:168-249): 15 TransferPrefill(Sequence),
//   18 ReceivePrefillResponse((bool, Option<Sequence>)), 23 KvCacheSend((Sequence, u32)), 25 KvCacheReceive(Sequence);
// family 1 = client <-> PD server `TransferMessage` (transfer/mod.rs:113-122): 0 TransferPrefill(Sequence).
// `extra`: the bool of kind 18, the first token of kind 23.  index < 0 with kind 18 frames `(extra, None)`.
int64_t b200_scheduler_frame_sequence(const b200_scheduler* s, int32_t index, int32_t family, int32_t kind,
                                      uint32_t extra, uint8_t* out, int64_t cap) {
  using namespace b200::wire;
  const bool ok_kind = family == 1 ? kind == 0 : (kind == 15 || kind == 18 || kind == 23 || kind == 25);
  const bool none = family == 0 && kind == 18 && index < 0;
  if (!ok_kind || (!none && (index < 0 || static_cast<size_t>(index) >= s->running.size()))) {
    b200::set_error(1, "scheduler_frame_sequence: bad message kind or index");
    return -1;
  }
  Wr w{out, cap};
  w.u32(0);
  w.u32(static_cast<uint32_t>(kind));
  if (family == 0 && kind == 18) {
    w.u8(extra ? 1 : 0);
    w.u8(none ? 0 : 1);
  }
  if (!none) write_sequence(w, s->running[index]);
  if (family == 0 && kind == 23) w.u32(extra);
  if (!w.ok) {
    b200::set_error(1, "scheduler_frame_sequence: buffer too small");
    return -1;
  }
  const uint32_t body = static_cast<uint32_t>(w.off - 4);
  std::memcpy(out, &body, 4);
  return w.off;
}

// Loads the message's sequence as the container's only running sequence (index 0).  Returns 1, 0 for
// ReceivePrefillResponse((_, None)), or -1 malformed / not one of the kinds above, -4 a sequence with images.
int32_t b200_scheduler_load_sequence(b200_scheduler* s, const uint8_t* frame, int64_t len, int32_t family,
                                     int32_t* kind, uint32_t* extra) {
  using namespace b200::wire;
  Rd r{frame, len};
  const uint32_t body = r.u32();
  const uint32_t k = r.u32();
  if (!r.ok || static_cast<int64_t>(body) + 4 != len) return -1;
  if (family == 1 ? k != 0 : (k != 15 && k != 18 && k != 23 && k != 25)) return -1;
  uint32_t ex = 0;
  bool some = true;
  if (family == 0 && k == 18) {
    ex = r.u8();
    some = r.opt();
  }
  std::vector<Sequence> seqs;
  if (some) {
    seqs.resize(1);
    const int rc = read_sequence(r, seqs[0], nullptr);
    if (rc != 0) return rc;
  }
  if (family == 0 && k == 23) ex = r.u32();
  if (!r.ok || r.off != len) return -1;
  s->running = std::move(seqs);
  if (kind) *kind = static_cast<int32_t>(k);
  if (extra) *extra = ex;
  return some ? 1 : 0;
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
