"""PD transfer wire frames, restated (TEST INFRASTRUCTURE, see oracle/__init__.py).

/root/reference/src/transfer/comm.rs:384-410: a frame is a 4-byte little-endian length followed by
`bincode::serialize(&TransferMessage)`.  bincode is a crates.io dependency (Cargo.toml:37, `bincode = "1.3.1"`),
not vendored: this restates its published default configuration -- little endian, fixed-width integers, `usize`
as u64, enum variant index as u32, `Vec<T>` / `String` as a u64 length followed by the elements, tuples and
structs as their fields in declaration order -- for the messages that carry or release KV bytes
(/root/reference/src/transfer/mod.rs:82-125):

    enum TransferMessage { TransferPrefill(Sequence) = 0, AvailableTokenResponse(usize) = 1,
                           TransferKvCache(FinishedPrefillData) = 2, ReleaseKvCache(usize) = 3 }
    struct FinishedPrefillData { seq_id: usize, first_token: u32, transfer_handle: KVTransferHandle,
                                 sending_time: usize, num_cached_tokens: usize }
    enum KVTransferHandle { LocalIpc {..} = 0,
                            RemoteTcp { layer_data: Vec<(Vec<u8>, Vec<u8>)>, num_blocks: usize } = 1 }

Pinned by tests/test_transfer_cpu.py against a frame written out by hand from those rules.
"""
from __future__ import annotations

import struct

MSG_AVAILABLE, MSG_KV, MSG_RELEASE = 1, 2, 3
HANDLE_REMOTE_TCP = 1


def frame(body: bytes) -> bytes:
    return struct.pack("<I", len(body)) + body


def kv_remote_frame(seq_id: int, first_token: int, layer_data, num_blocks: int, sending_time: int,
                    num_cached_tokens: int) -> bytes:
    """TransferKvCache(FinishedPrefillData { RemoteTcp }); layer_data = [(k_bytes, v_bytes), ...]."""
    b = struct.pack("<IQII", MSG_KV, seq_id, first_token, HANDLE_REMOTE_TCP)
    b += struct.pack("<Q", len(layer_data))
    for k, v in layer_data:
        b += struct.pack("<Q", len(k)) + bytes(k) + struct.pack("<Q", len(v)) + bytes(v)
    b += struct.pack("<QQQ", num_blocks, sending_time, num_cached_tokens)
    return frame(b)


def small_frame(kind: int, value: int) -> bytes:
    """AvailableTokenResponse(usize) / ReleaseKvCache(usize)."""
    assert kind in (MSG_AVAILABLE, MSG_RELEASE)
    return frame(struct.pack("<IQ", kind, value))


def parse_kv_remote_frame(buf: bytes):
    (n,) = struct.unpack_from("<I", buf, 0)
    assert n + 4 == len(buf)
    msg, seq_id, first_token, handle = struct.unpack_from("<IQII", buf, 4)
    assert msg == MSG_KV and handle == HANDLE_REMOTE_TCP
    off = 4 + 20
    (nl,) = struct.unpack_from("<Q", buf, off)
    off += 8
    layers = []
    for _ in range(nl):
        (lk,) = struct.unpack_from("<Q", buf, off)
        k = buf[off + 8:off + 8 + lk]
        off += 8 + lk
        (lv,) = struct.unpack_from("<Q", buf, off)
        v = buf[off + 8:off + 8 + lv]
        off += 8 + lv
        layers.append((k, v))
    num_blocks, sending_time, cached = struct.unpack_from("<QQQ", buf, off)
    assert off + 24 == len(buf)
    return dict(seq_id=seq_id, first_token=first_token, layer_data=layers, num_blocks=num_blocks,
                sending_time=sending_time, num_cached_tokens=cached)


# ---------------------------------------------------------------- engine <-> runner frames of the decode step
# /root/reference/src/runner/mod.rs:168-295 (MessageType variant order), core/sequence.rs:52-83 (DecodeSequence),
# utils/config.rs:504-537 (SamplingParams: stop_token_ids is #[serde(skip)]; grammar travels as Option<String>).
MT_RUN_DECODE, MT_RUN_RESPONSE, MT_FINISH_DECODE, MT_HEARTBEAT, MT_KV_SWAP, MT_KV_SWAP_RESPONSE = 4, 5, 8, 14, 21, 22
MT_CLEAR_BLOCKS, MT_CLEAR_BLOCKS_RESPONSE, MT_SHUTDOWN = 31, 32, 34


def _opt(fmt, v):
    return b"\x00" if v is None else b"\x01" + struct.pack(fmt, v)


def _opt_str(v):
    return b"\x00" if v is None else b"\x01" + struct.pack("<Q", len(v)) + v


def sampling_params(temperature=None, max_tokens=None, ignore_eos=False, top_k=None, top_p=None, session_id=None,
                    frequency_penalty=None, presence_penalty=None, stop_sequences=None, thinking=None, mcp_mode=None,
                    grammar=None, grammar_json=None, reasoning_effort=None) -> bytes:
    b = _opt("<f", temperature) + _opt("<Q", max_tokens) + struct.pack("<B", 1 if ignore_eos else 0)
    b += _opt("<q", top_k) + _opt("<f", top_p) + _opt_str(session_id)
    b += _opt("<f", frequency_penalty) + _opt("<f", presence_penalty)
    if stop_sequences is None:
        b += b"\x00"
    else:
        b += b"\x01" + struct.pack("<Q", len(stop_sequences)) + b"".join(
            struct.pack("<Q", len(x)) + x for x in stop_sequences)
    b += _opt("<B", thinking) + _opt("<B", mcp_mode) + _opt_str(grammar) + _opt_str(grammar_json)
    b += _opt("<I", reasoning_effort)
    return b


def run_decode_frame(seqs, is_prefill=False) -> bytes:
    """seqs: dicts with id, last_token, len, last_block_tokens, block_table_last, block_tables, sampling (bytes)."""
    b = struct.pack("<IQ", MT_RUN_DECODE, len(seqs))
    for q in seqs:
        b += struct.pack("<QIQQI", q["id"], q["last_token"], q["len"], q["last_block_tokens"], q["block_table_last"])
        b += struct.pack("<Q", len(q["block_tables"])) + struct.pack(f"<{len(q['block_tables'])}I", *q["block_tables"])
        b += q.get("sampling", sampling_params())
    b += struct.pack("<B", 1 if is_prefill else 0)
    return frame(b)


def run_response_frame(ids) -> bytes:
    return frame(struct.pack("<IQ", MT_RUN_RESPONSE, len(ids)) + struct.pack(f"<{len(ids)}I", *ids))


def kv_swap_frame(pairs, swap_in) -> bytes:
    b = struct.pack("<IQ", MT_KV_SWAP, len(pairs))
    for s, d in pairs:
        b += struct.pack("<QQ", s, d)
    return frame(b + struct.pack("<B", 1 if swap_in else 0))


def simple_frame(kind, payload=b"") -> bytes:
    return frame(struct.pack("<I", kind) + payload)


def decode_metadata(seqs, block_size=64):
    """core/runner.rs:1248-1264."""
    return dict(input_ids=[q["last_token"] for q in seqs], positions=[q["len"] - 1 for q in seqs],
                context_lens=[q["len"] for q in seqs],
                slot_mapping=[(q["block_table_last"] * block_size + q["last_block_tokens"] - 1) & 0xFFFFFFFF for q in seqs])


MT_RUN_PREFILL = 3


def sequence_bytes(q) -> bytes:
    """core/sequence.rs:32-50, field by field.  q: dict with id, status (0 Waiting .. 5 FinishSwapped), token_ids,
    output_ids, block_table, num_cached_tokens, last_token, block_size and optionally created_time, swapped_time,
    mamba_prefix_hash, sampling (bytes), pd_first_token, stop_sequence (bytes); images is None (multimodal data is
    out of scope)."""
    def vec(v):
        return struct.pack("<Q", len(v)) + struct.pack(f"<{len(v)}I", *v)
    b = struct.pack("<QQ", q["id"], q.get("created_time", 0)) + _opt("<Q", q.get("swapped_time"))
    b += struct.pack("<I", q["status"]) + vec(q["token_ids"]) + vec(q["output_ids"]) + vec(q["block_table"])
    b += struct.pack("<Q", q["num_cached_tokens"]) + _opt("<Q", q.get("mamba_prefix_hash"))
    b += struct.pack("<IQ", q["last_token"], q["block_size"]) + q.get("sampling", sampling_params())
    b += _opt("<I", q.get("pd_first_token")) + b"\x00"               # images: None
    b += struct.pack("<BB", int(q.get("is_tool_call_end", False)), int(q.get("hit_stop_sequence", False)))
    b += _opt_str(q.get("stop_sequence"))
    return b


def run_prefill_frame(seqs, is_prefill=True) -> bytes:
    """MessageType::RunPrefill((Vec<Sequence>, bool)) (runner/mod.rs:178)."""
    b = struct.pack("<IQ", MT_RUN_PREFILL, len(seqs)) + b"".join(sequence_bytes(q) for q in seqs)
    return frame(b + struct.pack("<B", 1 if is_prefill else 0))


# single-Sequence messages of the PD hand-off (runner/mod.rs:205-231, transfer/mod.rs:113-122)
MT_TRANSFER_PREFILL, MT_RECEIVE_PREFILL_RESPONSE, MT_KV_SEND, MT_KV_RECEIVE, MT_KV_RECEIVE_RESPONSE = 15, 18, 23, 25, 26


def transfer_prefill_frame(q, tcp=False) -> bytes:
    """MessageType::TransferPrefill(Sequence) (variant 15) or, tcp=True, TransferMessage::TransferPrefill (variant 0)."""
    return frame(struct.pack("<I", 0 if tcp else MT_TRANSFER_PREFILL) + sequence_bytes(q))


def receive_prefill_response_frame(ok, q=None) -> bytes:
    return frame(struct.pack("<IB", MT_RECEIVE_PREFILL_RESPONSE, int(ok)) + (b"\x00" if q is None else b"\x01" + sequence_bytes(q)))


def kv_send_frame(q, first_token) -> bytes:
    return frame(struct.pack("<I", MT_KV_SEND) + sequence_bytes(q) + struct.pack("<I", first_token))


def kv_receive_frame(q) -> bytes:
    return frame(struct.pack("<I", MT_KV_RECEIVE) + sequence_bytes(q))


def kv_receive_response_frame(ok, first_token, sending_time, num_cached_tokens) -> bytes:
    return frame(struct.pack("<IBIQQ", MT_KV_RECEIVE_RESPONSE, int(ok), first_token, sending_time, num_cached_tokens))


DTYPE_U8, DTYPE_U32, DTYPE_I64, DTYPE_BF16, DTYPE_F16, DTYPE_F32, DTYPE_F64 = range(7)   # runner/mod.rs:123-133


def kv_ipc_frame(seq_id, first_token, layer_handles, shape, dtype, server_block_ids, sending_time,
                 num_cached_tokens) -> bytes:
    """TransferKvCache(FinishedPrefillData { LocalIpc }) (transfer/mod.rs:77,84-90): layer_handles =
    [(k_handle_64_bytes, v_handle_64_bytes), ...]; CudaIpcMemHandle(Vec<i8>, Vec<usize>, SerializableDType)."""
    def handle(hb):
        assert len(hb) == 64
        return (struct.pack("<Q", 64) + bytes(hb) + struct.pack("<Q", len(shape)) +
                struct.pack(f"<{len(shape)}Q", *shape) + struct.pack("<I", dtype))
    b = struct.pack("<IQII", MSG_KV, seq_id, first_token, 0) + struct.pack("<Q", len(layer_handles))
    for k, v in layer_handles:
        b += handle(k) + handle(v)
    b += struct.pack("<Q", len(server_block_ids)) + struct.pack(f"<{len(server_block_ids)}I", *server_block_ids)
    b += struct.pack("<QQ", sending_time, num_cached_tokens)
    return frame(b)
