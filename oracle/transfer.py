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
