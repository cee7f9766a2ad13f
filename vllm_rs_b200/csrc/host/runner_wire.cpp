// Engine <-> runner-process wire format of the decode step.
//
// Mirrors /root/reference/src/runner/mod.rs:168-295 (`MessageType`, `send_local` / `receive_local`: a frame is a
// u32 little-endian length followed by `bincode::serialize(&MessageType)`; the receiver answers every frame with ONE
// acknowledgment byte 0x01), /root/reference/src/core/sequence.rs:52-83 (`DecodeSequence`),
// /root/reference/src/utils/config.rs:504-537 (`SamplingParams`; `stop_token_ids` is `#[serde(skip)]`, `grammar` goes
// over bincode as `Option<String>` of its JSON, config.rs:128-145) and the decode metadata a runner derives from such
// a message (/root/reference/src/core/runner.rs:1243-1264).  bincode 1.3.1 defaults (Cargo.toml:37): little endian,
// fixed-width integers, usize / isize as 64 bit, enum variant as u32, Option as one tag byte, Vec / String / HashMap
// as a u64 length followed by the elements.
//
// What is here: the messages that cross the socket every decode step (RunDecode -> RunResponse), the ones that
// move or release KV blocks (KVCacheSwap, FinishDecode, ClearBlocks), and the trivial ones a runner must recognise
// (Heartbeat, Shutdown).  RunPrefill carries whole `Sequence` structs: scheduler.cpp reads / writes it
// (b200_scheduler_load_run_prefill / _frame_run_prefill), next to the Sequence mirror it fills.  Init / RunEmbed
// (EngineConfig, embedding strategies) stay with the Rust host (control plane, SURVEY 8 out of scope).
#include <cstdint>
#include <cstring>

#include "../../../include/b200_vllm.h"
#include "../common.cuh"
#include "bincode.hpp"

using namespace b200::wire;

extern "C" {

int32_t b200_wire_kind(const uint8_t* frame, int64_t len) {
  Rd r{frame, len};
  const uint32_t body = r.u32();
  const uint32_t kind = r.u32();
  if (!r.ok || static_cast<int64_t>(body) + 4 != len || kind > kShutdown) return -1;
  return static_cast<int32_t>(kind);
}

// RunDecode((Vec<DecodeSequence>, bool)) -> the arrays prepare_decode reads (runner.rs:1243-1264).
// block_tables: [max_seqs][max_blocks] row-major, rows padded with 0; returns the number of sequences, < 0 on a
// malformed frame (-2 too many sequences, -3 a block table longer than max_blocks).
int32_t b200_wire_parse_run_decode(const uint8_t* frame, int64_t len, int32_t max_seqs, int32_t max_blocks,
                                   uint64_t* seq_ids, uint32_t* last_tokens, uint64_t* lens,
                                   uint64_t* last_block_tokens, uint32_t* block_table_last, uint32_t* block_tables,
                                   int32_t* num_blocks, b200_wire_sampling* sampling, int32_t* is_prefill) {
  Rd r{frame, len};
  const uint32_t body = r.u32();
  if (!r.ok || static_cast<int64_t>(body) + 4 != len || r.u32() != kRunDecode) return -1;
  const uint64_t n = r.u64();
  if (!r.ok) return -1;
  if (n > static_cast<uint64_t>(max_seqs)) return -2;
  for (uint64_t i = 0; i < n; ++i) {
    seq_ids[i] = r.u64();
    last_tokens[i] = r.u32();
    lens[i] = r.u64();
    last_block_tokens[i] = r.u64();
    block_table_last[i] = r.u32();
    const uint64_t nb = r.u64();
    if (!r.ok) return -1;
    if (nb > static_cast<uint64_t>(max_blocks)) return -3;
    uint32_t* row = block_tables + i * static_cast<uint64_t>(max_blocks);
    for (uint64_t b = 0; b < nb; ++b) row[b] = r.u32();
    for (uint64_t b = nb; b < static_cast<uint64_t>(max_blocks); ++b) row[b] = 0;
    num_blocks[i] = static_cast<int32_t>(nb);
    read_sampling(r, sampling ? sampling + i : nullptr);
    if (!r.ok) return -1;
  }
  const uint8_t pf = r.u8();
  if (!r.ok || r.off != len) return -1;
  if (is_prefill) *is_prefill = pf;
  return static_cast<int32_t>(n);
}

// The engine side of the same message (engine.rs:857): for a non-Rust host and for the tests.
int64_t b200_wire_frame_run_decode(uint8_t* out, int64_t cap, int32_t n, const uint64_t* seq_ids,
                                   const uint32_t* last_tokens, const uint64_t* lens,
                                   const uint64_t* last_block_tokens, const uint32_t* block_table_last,
                                   const uint32_t* block_tables, int32_t max_blocks, const int32_t* num_blocks,
                                   const b200_wire_sampling* sampling, int32_t is_prefill) {
  Wr w{out, cap};
  w.u32(0);
  w.u32(kRunDecode);
  w.u64(static_cast<uint64_t>(n));
  for (int i = 0; i < n; ++i) {
    w.u64(seq_ids[i]);
    w.u32(last_tokens[i]);
    w.u64(lens[i]);
    w.u64(last_block_tokens[i]);
    w.u32(block_table_last[i]);
    w.u64(static_cast<uint64_t>(num_blocks[i]));
    for (int b = 0; b < num_blocks[i]; ++b) w.u32(block_tables[static_cast<size_t>(i) * max_blocks + b]);
    write_sampling(w, sampling ? sampling + i : nullptr);
  }
  w.u8(is_prefill ? 1 : 0);
  if (!w.ok) { b200::set_error(1, "wire_frame_run_decode: buffer too small"); return -1; }
  const uint32_t body = static_cast<uint32_t>(w.off - 4);
  std::memcpy(out, &body, 4);
  return w.off;
}

// RunResponse(Vec<u32>): the token ids of the step, runner -> engine (runner.rs:262-279)
int64_t b200_wire_frame_run_response(uint8_t* out, int64_t cap, const uint32_t* ids, int32_t n) {
  const int64_t total = 4 + 4 + 8 + 4 * static_cast<int64_t>(n);
  if (!out || total > cap) { b200::set_error(1, "wire_frame_run_response: buffer too small"); return -1; }
  Wr w{out, cap};
  w.u32(static_cast<uint32_t>(total - 4));
  w.u32(kRunResponse);
  w.u64(static_cast<uint64_t>(n));
  for (int i = 0; i < n; ++i) w.u32(ids[i]);
  return w.off;
}

int32_t b200_wire_parse_run_response(const uint8_t* frame, int64_t len, uint32_t* ids, int32_t cap) {
  Rd r{frame, len};
  const uint32_t body = r.u32();
  if (!r.ok || static_cast<int64_t>(body) + 4 != len || r.u32() != kRunResponse) return -1;
  const uint64_t n = r.u64();
  if (!r.ok || n > static_cast<uint64_t>(cap)) return -2;
  for (uint64_t i = 0; i < n; ++i) ids[i] = r.u32();
  if (!r.ok || r.off != len) return -1;
  return static_cast<int32_t>(n);
}

// KVCacheSwap((HashMap<usize, usize>, bool)): pairs as int64 [n][2] = (src block, dst block) -- the layout
// b200_swap_blocks takes (upload once, one launch per cache tensor); *swap_in = the bool (CPU -> GPU when set).
int32_t b200_wire_parse_kv_swap(const uint8_t* frame, int64_t len, int64_t* pairs, int32_t cap, int32_t* swap_in) {
  Rd r{frame, len};
  const uint32_t body = r.u32();
  if (!r.ok || static_cast<int64_t>(body) + 4 != len || r.u32() != kKVCacheSwap) return -1;
  const uint64_t n = r.u64();
  if (!r.ok || n > static_cast<uint64_t>(cap)) return -2;
  for (uint64_t i = 0; i < n; ++i) {
    pairs[2 * i] = static_cast<int64_t>(r.u64());
    pairs[2 * i + 1] = static_cast<int64_t>(r.u64());
  }
  const uint8_t flag = r.u8();
  if (!r.ok || r.off != len) return -1;
  if (swap_in) *swap_in = flag;
  return static_cast<int32_t>(n);
}

int64_t b200_wire_frame_kv_swap(uint8_t* out, int64_t cap, const int64_t* pairs, int32_t n, int32_t swap_in) {
  const int64_t total = 4 + 4 + 8 + 16 * static_cast<int64_t>(n) + 1;
  if (!out || total > cap) { b200::set_error(1, "wire_frame_kv_swap: buffer too small"); return -1; }
  Wr w{out, cap};
  w.u32(static_cast<uint32_t>(total - 4));
  w.u32(kKVCacheSwap);
  w.u64(static_cast<uint64_t>(n));
  for (int i = 0; i < 2 * n; ++i) w.u64(static_cast<uint64_t>(pairs[i]));
  w.u8(swap_in ? 1 : 0);
  return w.off;
}

// FinishDecode(usize) = 8, KVCacheSwapResponse(bool) = 22, ClearBlocksResponse(bool) = 32, Heartbeat = 14, Shutdown = 34:
// a unit variant has no payload, a bool one byte, a usize eight.  payload_bytes in {0, 1, 8}.
int64_t b200_wire_frame_simple(uint8_t* out, int64_t cap, int32_t kind, int32_t payload_bytes, uint64_t value) {
  if (!out || cap < 8 + payload_bytes || kind < 0 || kind > static_cast<int32_t>(kShutdown) ||
      !(payload_bytes == 0 || payload_bytes == 1 || payload_bytes == 8)) {
    b200::set_error(1, "wire_frame_simple: bad arguments");
    return -1;
  }
  Wr w{out, cap};
  w.u32(static_cast<uint32_t>(4 + payload_bytes));
  w.u32(static_cast<uint32_t>(kind));
  if (payload_bytes == 1) w.u8(value ? 1 : 0);
  if (payload_bytes == 8) w.u64(value);
  return w.off;
}

// The decode metadata a runner derives from a RunDecode message (runner.rs:1248-1264): input id = last token,
// position = len - 1, context length = len, slot = last block * block_size + tokens in the last block - 1.
int64_t b200_wire_frame_kv_receive_response(uint8_t* out, int64_t cap, int32_t ok, uint32_t first_token, uint64_t sending_time,
                                            uint64_t num_cached_tokens) {
  Wr w{out, cap};
  w.u32(4 + 1 + 4 + 8 + 8);
  w.u32(26);
  w.u8(ok ? 1 : 0);
  w.u32(first_token);
  w.u64(sending_time);
  w.u64(num_cached_tokens);
  if (!out || !w.ok) { b200::set_error(1, "wire_frame_kv_receive_response: buffer too small"); return -1; }
  return w.off;
}

int32_t b200_wire_parse_kv_receive_response(const uint8_t* frame, int64_t len, int32_t* ok, uint32_t* first_token,
                                            uint64_t* sending_time, uint64_t* num_cached_tokens) {
  Rd r{frame, len};
  const uint32_t body = r.u32();
  if (!r.ok || static_cast<int64_t>(body) + 4 != len || r.u32() != 26) return -1;
  const uint8_t f = r.u8();
  const uint32_t t = r.u32();
  const uint64_t x = r.u64(), y = r.u64();
  if (!r.ok || r.off != len || f > 1) return -1;
  if (ok) *ok = f;
  if (first_token) *first_token = t;
  if (sending_time) *sending_time = x;
  if (num_cached_tokens) *num_cached_tokens = y;
  return 0;
}

void b200_wire_decode_metadata(int32_t n, const uint32_t* last_tokens, const uint64_t* lens,
                               const uint64_t* last_block_tokens, const uint32_t* block_table_last, int32_t block_size,
                               uint32_t* input_ids, int64_t* positions, int64_t* slot_mapping,
                               uint32_t* context_lens) {
  for (int i = 0; i < n; ++i) {
    input_ids[i] = last_tokens[i];
    positions[i] = static_cast<int64_t>(lens[i]) - 1;
    context_lens[i] = static_cast<uint32_t>(lens[i]);
    const uint32_t slot = block_table_last[i] * static_cast<uint32_t>(block_size) +
                          static_cast<uint32_t>(last_block_tokens[i]) - 1u;       // u32 arithmetic, as the reference
    slot_mapping[i] = static_cast<int64_t>(slot);
  }
}

}  // extern "C"
