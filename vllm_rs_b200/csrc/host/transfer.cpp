// PD (prefill / decode) KV-cache hand-off: the wire frames and the block movement around them.
//
// Mirrors /root/reference/src/transfer/mod.rs:82-125 (KVTransferHandle, FinishedPrefillData, TransferMessage),
// transfer/comm.rs:384-410 (frame = u32 LE length + bincode 1.3.1 body: little endian, fixed-width integers,
// usize as u64, enum variant as u32, Vec length as u64 -- Cargo.toml:37) and the two data legs of
// transfer/mod.rs:260-311,370-470: RemoteTcp (server: cache blocks -> contiguous bytes per layer, client: bytes ->
// its own blocks) and LocalIpc (client copies straight out of the server's IPC-mapped caches).  Sockets, the listener
// thread and the Sequence payload of TransferPrefill stay in the host (out of scope: control plane); what is here is
// what touches KV bytes.
//
// B200 shape of it: the frame is built IN PLACE in a pinned host buffer -- the header fields are written by the
// CPU, the per-layer K / V payloads are written by swap_blocks_kernel straight from the paged cache over UVA (no
// intermediate CPU tensor, no cpu_tensor_to_bytes copy), and on the receiving side the kernel reads the payloads
// straight out of the received frame.
#include <cstdint>
#include <cstring>

#include "../../../include/b200_vllm.h"
#include "../common.cuh"

namespace {

constexpr uint32_t kMsgTransferPrefill = 0, kMsgAvailableTokens = 1, kMsgTransferKvCache = 2, kMsgReleaseKvCache = 3;
constexpr uint32_t kHandleLocalIpc = 0, kHandleRemoteTcp = 1;

struct Writer {
  uint8_t* p;
  int64_t cap, off = 0;
  bool ok = true;
  void raw(const void* src, int64_t n) {
    if (off + n > cap) { ok = false; off += n; return; }
    if (src) std::memcpy(p + off, src, static_cast<size_t>(n));
    off += n;
  }
  void u32(uint32_t v) { raw(&v, 4); }       // x86-64 / aarch64 little endian = the wire order
  void u64(uint64_t v) { raw(&v, 8); }
};
struct Reader {
  const uint8_t* p;
  int64_t len, off = 0;
  bool ok = true;
  uint32_t u32() {
    uint32_t v = 0;
    if (off + 4 > len) { ok = false; return 0; }
    std::memcpy(&v, p + off, 4);
    off += 4;
    return v;
  }
  uint64_t u64() {
    uint64_t v = 0;
    if (off + 8 > len) { ok = false; return 0; }
    std::memcpy(&v, p + off, 8);
    off += 8;
    return v;
  }
};

}  // namespace

extern "C" {

int64_t b200_pd_kv_frame_bytes(int32_t num_layers, int64_t layer_bytes) {
  //  len | msg variant | seq_id | first_token | handle variant | n layers | per layer (len K, K, len V, V) |
  //  num_blocks | sending_time | num_cached_tokens
  return 4 + 4 + 8 + 4 + 4 + 8 + static_cast<int64_t>(num_layers) * (16 + 2 * layer_bytes) + 8 + 8 + 8;
}

int64_t b200_pd_kv_frame_begin(uint8_t* frame, int64_t cap, const b200_pd_kv_header* h, int64_t* k_off,
                               int64_t* v_off) {
  if (!frame || !h || !k_off || !v_off || h->num_layers <= 0) {
    b200::set_error(1, "pd_kv_frame_begin: bad arguments");
    return -1;
  }
  const int64_t total = b200_pd_kv_frame_bytes(h->num_layers, static_cast<int64_t>(h->layer_bytes));
  if (total > cap || total - 4 > 1000000000LL) {          // comm.rs:401-404: the receiver refuses frames above 1 GB
    b200::set_error(1, "pd_kv_frame_begin: frame does not fit the buffer (or exceeds the reference's 1 GB limit)");
    return -1;
  }
  Writer w{frame, cap};
  w.u32(static_cast<uint32_t>(total - 4));
  w.u32(kMsgTransferKvCache);
  w.u64(h->seq_id);
  w.u32(h->first_token);
  w.u32(kHandleRemoteTcp);
  w.u64(static_cast<uint64_t>(h->num_layers));
  for (int l = 0; l < h->num_layers; ++l) {
    w.u64(h->layer_bytes);
    k_off[l] = w.off;
    w.raw(nullptr, static_cast<int64_t>(h->layer_bytes));   // filled by b200_pd_export_blocks
    w.u64(h->layer_bytes);
    v_off[l] = w.off;
    w.raw(nullptr, static_cast<int64_t>(h->layer_bytes));
  }
  w.u64(h->num_blocks);
  w.u64(h->sending_time);
  w.u64(h->num_cached_tokens);
  if (!w.ok || w.off != total) {
    b200::set_error(1, "pd_kv_frame_begin: internal size mismatch");
    return -1;
  }
  return total;
}

int32_t b200_pd_kv_frame_parse(const uint8_t* frame, int64_t len, b200_pd_kv_header* h, int64_t* k_off,
                               int64_t* v_off, int32_t max_layers) {
  if (!frame || !h) return -1;
  Reader r{frame, len};
  const uint32_t body = r.u32();
  if (!r.ok || static_cast<int64_t>(body) + 4 != len) return -2;
  if (r.u32() != kMsgTransferKvCache) return -3;
  h->seq_id = r.u64();
  h->first_token = r.u32();
  if (r.u32() != kHandleRemoteTcp) return -4;              // LocalIpc carries handles, not bytes
  const uint64_t nl = r.u64();
  if (!r.ok || nl == 0 || nl > static_cast<uint64_t>(max_layers)) return -5;
  h->num_layers = static_cast<int32_t>(nl);
  h->layer_bytes = 0;
  for (uint64_t l = 0; l < nl; ++l) {
    for (int kv = 0; kv < 2; ++kv) {
      const uint64_t n = r.u64();
      if (!r.ok || r.off + static_cast<int64_t>(n) > len) return -6;
      if (l == 0 && kv == 0) h->layer_bytes = n;
      if (n != h->layer_bytes) return -7;                  // every layer's K and V carry the same blocks
      (kv == 0 ? k_off : v_off)[l] = r.off;
      r.off += static_cast<int64_t>(n);
    }
  }
  h->num_blocks = r.u64();
  h->sending_time = r.u64();
  h->num_cached_tokens = r.u64();
  if (!r.ok || r.off != len) return -8;
  return 0;
}

// TransferKvCache(FinishedPrefillData { transfer_handle: LocalIpc { layer_handles, server_block_ids }, .. })
// (transfer/mod.rs:77,82-101,430-470): per layer a (K, V) pair of CudaIpcMemHandle(Vec<i8> = the 64 handle bytes,
// Vec<usize> = the tensor's shape, SerializableDType = runner/mod.rs:123-133: U8 0, U32 1, I64 2, BF16 3, F16 4,
// F32 5, F64 6), then the server's block ids of the sequence.  The handle bytes are cudaIpcGetMemHandle's, taken by
// the host; opened on the client they are the `remote_k` / `remote_v` of b200_pd_copy_blocks_peer.
int64_t b200_pd_ipc_frame(uint8_t* out, int64_t cap, const b200_pd_kv_header* h, const uint8_t* k_handles,
                          const uint8_t* v_handles, const uint64_t* shape, int32_t ndim, int32_t dtype,
                          const uint32_t* server_block_ids) {
  if (!out || !h || !k_handles || !v_handles || !shape || h->num_layers <= 0 || ndim <= 0 || ndim > 8 || dtype < 0 ||
      dtype > 6 || (h->num_blocks > 0 && !server_block_ids)) {
    b200::set_error(1, "pd_ipc_frame: bad arguments");
    return -1;
  }
  Writer w{out, cap};
  w.u32(0);
  w.u32(kMsgTransferKvCache);
  w.u64(h->seq_id);
  w.u32(h->first_token);
  w.u32(kHandleLocalIpc);
  w.u64(static_cast<uint64_t>(h->num_layers));
  for (int32_t l = 0; l < h->num_layers; ++l)
    for (int kv = 0; kv < 2; ++kv) {
      w.u64(64);
      w.raw((kv == 0 ? k_handles : v_handles) + static_cast<size_t>(l) * 64, 64);
      w.u64(static_cast<uint64_t>(ndim));
      for (int32_t d = 0; d < ndim; ++d) w.u64(shape[d]);
      w.u32(static_cast<uint32_t>(dtype));
    }
  w.u64(h->num_blocks);
  w.raw(server_block_ids, static_cast<int64_t>(h->num_blocks) * 4);
  w.u64(h->sending_time);
  w.u64(h->num_cached_tokens);
  if (!w.ok) {
    b200::set_error(1, "pd_ipc_frame: buffer too small");
    return -1;
  }
  const uint32_t body = static_cast<uint32_t>(w.off - 4);
  std::memcpy(out, &body, 4);
  return w.off;
}

// Returns 0, or < 0: malformed (-2 length, -3 not TransferKvCache, -4 not LocalIpc, -5 layer count, -6 handle / shape
// fields, -7 layers that differ in shape or dtype, -8 block list / tail).  k_handles / v_handles: [max_layers][64].
int32_t b200_pd_ipc_frame_parse(const uint8_t* frame, int64_t len, b200_pd_kv_header* h, uint8_t* k_handles,
                                uint8_t* v_handles, int32_t max_layers, uint64_t* shape, int32_t* ndim, int32_t* dtype,
                                uint32_t* server_block_ids, int32_t max_blocks) {
  if (!frame || !h || !k_handles || !v_handles || !shape || !ndim || !dtype) return -1;
  Reader r{frame, len};
  const uint32_t body = r.u32();
  if (!r.ok || static_cast<int64_t>(body) + 4 != len) return -2;
  if (r.u32() != kMsgTransferKvCache) return -3;
  h->seq_id = r.u64();
  h->first_token = r.u32();
  if (r.u32() != kHandleLocalIpc) return -4;
  const uint64_t nl = r.u64();
  if (!r.ok || nl == 0 || nl > static_cast<uint64_t>(max_layers)) return -5;
  h->num_layers = static_cast<int32_t>(nl);
  h->layer_bytes = 0;
  for (uint64_t l = 0; l < nl; ++l)
    for (int kv = 0; kv < 2; ++kv) {
      if (r.u64() != 64 || !r.ok || r.off + 64 > len) return -6;
      std::memcpy((kv == 0 ? k_handles : v_handles) + l * 64, frame + r.off, 64);
      r.off += 64;
      const uint64_t nd = r.u64();
      if (!r.ok || nd == 0 || nd > 8) return -6;
      uint64_t dims[8];
      for (uint64_t d = 0; d < nd; ++d) dims[d] = r.u64();
      const uint32_t dt = r.u32();
      if (!r.ok || dt > 6) return -6;
      if (l == 0 && kv == 0) {
        *ndim = static_cast<int32_t>(nd);
        *dtype = static_cast<int32_t>(dt);
        std::memcpy(shape, dims, nd * 8);
      } else if (static_cast<int32_t>(nd) != *ndim || static_cast<int32_t>(dt) != *dtype ||
                 std::memcmp(shape, dims, nd * 8) != 0) {
        return -7;                                         // one cache geometry for the whole model
      }
    }
  const uint64_t nb = r.u64();
  if (!r.ok || nb > static_cast<uint64_t>(max_blocks) || r.off + static_cast<int64_t>(nb) * 4 > len) return -8;
  if (nb) std::memcpy(server_block_ids, frame + r.off, nb * 4);
  r.off += static_cast<int64_t>(nb) * 4;
  h->num_blocks = nb;
  h->sending_time = r.u64();
  h->num_cached_tokens = r.u64();
  if (!r.ok || r.off != len) return -8;
  return 0;
}

int64_t b200_pd_frame_small(uint8_t* out, int64_t cap, int32_t kind, uint64_t value) {
  // AvailableTokenResponse(usize) / ReleaseKvCache(usize): len | variant | u64
  if (!out || cap < 16 || (kind != static_cast<int32_t>(kMsgAvailableTokens) && kind != static_cast<int32_t>(kMsgReleaseKvCache))) {
    b200::set_error(1, "pd_frame_small: kind must be 1 (AvailableTokenResponse) or 3 (ReleaseKvCache), 16 bytes of room");
    return -1;
  }
  Writer w{out, cap};
  w.u32(12);
  w.u32(static_cast<uint32_t>(kind));
  w.u64(value);
  return w.off;
}

int32_t b200_pd_frame_kind(const uint8_t* frame, int64_t len, uint64_t* value) {
  Reader r{frame, len};
  const uint32_t body = r.u32();
  const uint32_t kind = r.u32();
  if (!r.ok || static_cast<int64_t>(body) + 4 != len || kind > kMsgReleaseKvCache) return -1;
  if ((kind == kMsgAvailableTokens || kind == kMsgReleaseKvCache) && value) {
    *value = r.u64();
    if (!r.ok) return -1;
  }
  (void)kMsgTransferPrefill;
  return static_cast<int32_t>(kind);
}

// RemoteTcp, server leg (transfer/mod.rs:433-470: copy_blocks_to_cpu with mapping {server block -> i}, then
// cpu_tensor_to_bytes): every layer's K and V blocks of the sequence land in frame[k_off[l]..] / frame[v_off[l]..]
// in block-table order.  pairs_dev: int64 [n][2] = (server block id, i) in device memory; frame: PINNED host memory.
void b200_pd_export_blocks(const void* const* k_caches, const void* const* v_caches, int32_t num_layers,
                           const int64_t* pairs_dev, int32_t n, int64_t block_bytes, uint8_t* frame,
                           const int64_t* k_off, const int64_t* v_off, int64_t stream) {
  for (int l = 0; l < num_layers; ++l) {
    b200_swap_blocks(k_caches[l], frame + k_off[l], pairs_dev, n, block_bytes, stream);
    b200_swap_blocks(v_caches[l], frame + v_off[l], pairs_dev, n, block_bytes, stream);
  }
}

// RemoteTcp, client leg (transfer/mod.rs:292-330: bytes_to_cpu_tensor, swap_blocks with {i -> local block}): the
// kernel reads the payloads straight out of the received (pinned) frame.  pairs_dev = (i, local block id).
void b200_pd_import_blocks(void* const* k_caches, void* const* v_caches, int32_t num_layers,
                           const int64_t* pairs_dev, int32_t n, int64_t block_bytes, const uint8_t* frame,
                           const int64_t* k_off, const int64_t* v_off, int64_t stream) {
  for (int l = 0; l < num_layers; ++l) {
    b200_swap_blocks(frame + k_off[l], k_caches[l], pairs_dev, n, block_bytes, stream);
    b200_swap_blocks(frame + v_off[l], v_caches[l], pairs_dev, n, block_bytes, stream);
  }
}

// LocalIpc, client leg (transfer/mod.rs:268-290): remote caches are the server's tensors opened through cudaIpc (or,
// inside one process, a peer GPU's caches with peer access on); pairs_dev = (server block id, local block id).
void b200_pd_copy_blocks_peer(const void* const* remote_k, const void* const* remote_v, void* const* local_k,
                              void* const* local_v, int32_t num_layers, const int64_t* pairs_dev, int32_t n,
                              int64_t block_bytes, int64_t stream) {
  for (int l = 0; l < num_layers; ++l) {
    b200_swap_blocks(remote_k[l], local_k[l], pairs_dev, n, block_bytes, stream);
    b200_swap_blocks(remote_v[l], local_v[l], pairs_dev, n, block_bytes, stream);
  }
}

// cudaDeviceEnablePeerAccess for the pair (already-enabled is not an error): the LocalIpc copy above dereferences the
// server GPU's memory from a kernel running on the client GPU, over NVLink.
int32_t b200_enable_peer_access(int32_t dev, int32_t peer) {
  int can = 0;
  if (cudaDeviceCanAccessPeer(&can, dev, peer) != cudaSuccess || !can) return -1;
  int cur = 0;
  cudaGetDevice(&cur);
  cudaSetDevice(dev);
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  cudaSetDevice(cur);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return -2;
  }
  cudaGetLastError();
  return 0;
}

}  // extern "C"
