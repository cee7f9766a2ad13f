// bincode 1.3.1 (default configuration) reader / writer shared by the wire-frame code of the host mirror
// (runner_wire.cpp, scheduler.cpp): little endian, fixed-width integers, usize / isize as 64 bit, enum variant as
// u32, Option as one tag byte, Vec / String / HashMap as a u64 length followed by the elements.  SamplingParams
// follows /root/reference/src/utils/config.rs:504-537 (`stop_token_ids` is #[serde(skip)]).
#pragma once
#include <cstdint>
#include <cstring>

#include "../../../include/b200_vllm.h"

namespace b200 {
namespace wire {

// variant indices of MessageType, in declaration order (/root/reference/src/runner/mod.rs:168-249)
enum : uint32_t {
  kInit = 0, kInitAck = 1, kLoadingProgress = 2, kRunPrefill = 3, kRunDecode = 4, kRunResponse = 5, kRunEmbed = 6,
  kRunResponseEmbed = 7, kFinishDecode = 8, kHeartbeat = 14, kKVCacheSwap = 21, kKVCacheSwapResponse = 22,
  kClearBlocks = 31, kClearBlocksResponse = 32, kUsableMemoryLeft = 33, kShutdown = 34,
};

struct Rd {
  const uint8_t* p;
  int64_t len, off = 0;
  bool ok = true;
  bool need(int64_t n) {
    if (off + n > len) { ok = false; return false; }
    return true;
  }
  uint8_t u8() { if (!need(1)) return 0; return p[off++]; }
  uint32_t u32() { uint32_t v = 0; if (!need(4)) return 0; std::memcpy(&v, p + off, 4); off += 4; return v; }
  uint64_t u64() { uint64_t v = 0; if (!need(8)) return 0; std::memcpy(&v, p + off, 8); off += 8; return v; }
  float f32() { float v = 0; if (!need(4)) return 0; std::memcpy(&v, p + off, 4); off += 4; return v; }
  void skip(uint64_t n) { if (need(static_cast<int64_t>(n))) off += static_cast<int64_t>(n); }
  void skip_string() { const uint64_t n = u64(); if (ok) skip(n); }
  bool opt() { const uint8_t t = u8(); if (t > 1) ok = false; return t == 1; }
};

struct Wr {
  uint8_t* p;
  int64_t cap, off = 0;
  bool ok = true;
  void raw(const void* s, int64_t n) {
    if (off + n > cap) { ok = false; off += n; return; }
    std::memcpy(p + off, s, static_cast<size_t>(n));
    off += n;
  }
  void u8(uint8_t v) { raw(&v, 1); }
  void u32(uint32_t v) { raw(&v, 4); }
  void u64(uint64_t v) { raw(&v, 8); }
  void f32(float v) { raw(&v, 4); }
  void opt_f32(int has, float v) { u8(has ? 1 : 0); if (has) f32(v); }
};

inline void read_sampling(Rd& r, b200_wire_sampling* s) {
  b200_wire_sampling t;
  std::memset(&t, 0, sizeof(t));
  if (r.opt()) { t.has_temperature = 1; t.temperature = r.f32(); }
  if (r.opt()) { t.has_max_tokens = 1; t.max_tokens = static_cast<int64_t>(r.u64()); }
  t.ignore_eos = r.u8();
  if (r.opt()) { t.has_top_k = 1; t.top_k = static_cast<int64_t>(r.u64()); }
  if (r.opt()) { t.has_top_p = 1; t.top_p = r.f32(); }
  if (r.opt()) r.skip_string();                                   // session_id
  if (r.opt()) { t.has_frequency_penalty = 1; t.frequency_penalty = r.f32(); }
  if (r.opt()) { t.has_presence_penalty = 1; t.presence_penalty = r.f32(); }
  if (r.opt()) {                                                  // stop_sequences: Option<Vec<String>>
    const uint64_t n = r.u64();
    for (uint64_t i = 0; i < n && r.ok; ++i) r.skip_string();
  }
  if (r.opt()) r.u8();                                            // thinking / mcp_mode: two Option<bool> (their order
  if (r.opt()) r.u8();                                            // differs between the python and the plain build)
  if (r.opt()) r.skip_string();                                   // grammar (JSON string over bincode)
  if (r.opt()) r.skip_string();                                   // grammar_json
  if (r.opt()) r.u32();                                           // reasoning_effort: unit-variant enum
  if (s) *s = t;
}

inline void write_sampling(Wr& w, const b200_wire_sampling* s) {
  b200_wire_sampling z;
  std::memset(&z, 0, sizeof(z));
  if (!s) s = &z;
  w.opt_f32(s->has_temperature, s->temperature);
  w.u8(s->has_max_tokens ? 1 : 0);
  if (s->has_max_tokens) w.u64(static_cast<uint64_t>(s->max_tokens));
  w.u8(s->ignore_eos ? 1 : 0);
  w.u8(s->has_top_k ? 1 : 0);
  if (s->has_top_k) w.u64(static_cast<uint64_t>(s->top_k));
  w.opt_f32(s->has_top_p, s->top_p);
  w.u8(0);                                                        // session_id: None
  w.opt_f32(s->has_frequency_penalty, s->frequency_penalty);
  w.opt_f32(s->has_presence_penalty, s->presence_penalty);
  w.u8(0);                                                        // stop_sequences
  w.u8(0); w.u8(0);                                               // thinking, mcp_mode
  w.u8(0); w.u8(0);                                               // grammar, grammar_json
  w.u8(0);                                                        // reasoning_effort
}


}  // namespace wire
}  // namespace b200
