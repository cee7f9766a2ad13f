"""PD transfer wire frames (transfer/mod.rs:82-125, comm.rs:384-410): the native framing code vs the oracle
restatement of bincode 1.3.1 and vs a frame written out by hand.  CPU only (no copies run here)."""
import ctypes as C
import struct

import numpy as np

from oracle import transfer as ot


def _hand_frame():
    """TransferKvCache(FinishedPrefillData{seq_id 7, first_token 0x01020304, RemoteTcp{2 layers of 3-byte K / V,
    num_blocks 1}, sending_time 0x1122334455, num_cached_tokens 64}) byte by byte from the bincode rules."""
    body = bytes([2, 0, 0, 0])                                   # TransferMessage variant 2 (u32 LE)
    body += bytes([7, 0, 0, 0, 0, 0, 0, 0])                      # seq_id: usize -> u64
    body += bytes([4, 3, 2, 1])                                  # first_token u32
    body += bytes([1, 0, 0, 0])                                  # KVTransferHandle variant 1 = RemoteTcp
    body += bytes([2, 0, 0, 0, 0, 0, 0, 0])                      # layer_data: Vec len 2
    body += bytes([3, 0, 0, 0, 0, 0, 0, 0]) + b"abc"             # layer 0 K: Vec<u8> len 3
    body += bytes([3, 0, 0, 0, 0, 0, 0, 0]) + b"def"             # layer 0 V
    body += bytes([3, 0, 0, 0, 0, 0, 0, 0]) + b"ghi"             # layer 1 K
    body += bytes([3, 0, 0, 0, 0, 0, 0, 0]) + b"jkl"             # layer 1 V
    body += bytes([1, 0, 0, 0, 0, 0, 0, 0])                      # num_blocks
    body += bytes([0x55, 0x44, 0x33, 0x22, 0x11, 0, 0, 0])       # sending_time
    body += bytes([64, 0, 0, 0, 0, 0, 0, 0])                     # num_cached_tokens
    return struct.pack("<I", len(body)) + body


def test_oracle_matches_the_hand_written_frame():
    f = ot.kv_remote_frame(7, 0x01020304, [(b"abc", b"def"), (b"ghi", b"jkl")], 1, 0x1122334455, 64)
    assert f == _hand_frame()
    p = ot.parse_kv_remote_frame(f)
    assert p["layer_data"] == [(b"abc", b"def"), (b"ghi", b"jkl")] and p["seq_id"] == 7 and p["num_cached_tokens"] == 64
    assert ot.small_frame(ot.MSG_RELEASE, 9) == bytes([12, 0, 0, 0, 3, 0, 0, 0, 9, 0, 0, 0, 0, 0, 0, 0])


def test_native_framing_is_byte_exact(lib):
    from vllm_rs_b200 import _lib
    rng = np.random.default_rng(0)
    for L, layer_bytes, nblk in ((1, 16, 1), (3, 4096, 2), (32, 64 * 8 * 128 * 2 * 3, 3)):
        layers = [(rng.integers(0, 256, layer_bytes, dtype=np.uint8).tobytes(),
                   rng.integers(0, 256, layer_bytes, dtype=np.uint8).tobytes()) for _ in range(L)]
        want = ot.kv_remote_frame(12345678901, 31999, layers, nblk, 1760000000123, 192)
        total = lib.b200_pd_kv_frame_bytes(L, layer_bytes)
        assert total == len(want)
        buf = np.zeros(total, np.uint8)
        hdr = _lib.PdKvHeader(12345678901, 31999, L, layer_bytes, nblk, 1760000000123, 192)
        k_off, v_off = np.zeros(L, np.int64), np.zeros(L, np.int64)
        assert lib.b200_pd_kv_frame_begin(buf.ctypes.data, total, C.byref(hdr), k_off.ctypes.data, v_off.ctypes.data) == total
        for l, (k, v) in enumerate(layers):                     # the copy kernel's job, done by hand here
            buf[k_off[l]:k_off[l] + layer_bytes] = np.frombuffer(k, np.uint8)
            buf[v_off[l]:v_off[l] + layer_bytes] = np.frombuffer(v, np.uint8)
        assert buf.tobytes() == want
        # and back
        h2 = _lib.PdKvHeader()
        k2, v2 = np.zeros(L, np.int64), np.zeros(L, np.int64)
        assert lib.b200_pd_kv_frame_parse(buf.ctypes.data, total, C.byref(h2), k2.ctypes.data, v2.ctypes.data, L) == 0
        assert (h2.seq_id, h2.first_token, h2.num_layers, h2.layer_bytes, h2.num_blocks, h2.sending_time,
                h2.num_cached_tokens) == (12345678901, 31999, L, layer_bytes, nblk, 1760000000123, 192)
        assert np.array_equal(k2, k_off) and np.array_equal(v2, v_off)
        # malformed: truncated, wrong variant, too many layers for the receiver
        assert lib.b200_pd_kv_frame_parse(buf.ctypes.data, total - 1, C.byref(h2), k2.ctypes.data, v2.ctypes.data, L) != 0
        if L > 1:
            assert lib.b200_pd_kv_frame_parse(buf.ctypes.data, total, C.byref(h2), k2.ctypes.data, v2.ctypes.data, L - 1) != 0
        bad = buf.copy()
        bad[4] = 0
        assert lib.b200_pd_kv_frame_parse(bad.ctypes.data, total, C.byref(h2), k2.ctypes.data, v2.ctypes.data, L) != 0
    small = np.zeros(16, np.uint8)
    for kind, val in ((ot.MSG_AVAILABLE, 123456), (ot.MSG_RELEASE, 2**40 + 5)):
        assert lib.b200_pd_frame_small(small.ctypes.data, 16, kind, val) == 16
        assert small.tobytes() == ot.small_frame(kind, val)
        out = C.c_uint64(0)
        assert lib.b200_pd_frame_kind(small.ctypes.data, 16, C.byref(out)) == kind and out.value == val
    hf = np.frombuffer(_hand_frame(), np.uint8).copy()
    assert lib.b200_pd_frame_kind(hf.ctypes.data, len(hf), None) == ot.MSG_KV
    lib.b200_clear_error()


def test_local_ipc_frame_is_byte_exact(lib):
    """TransferKvCache with a LocalIpc handle (transfer/mod.rs:77,84-90): hand-written bytes for one layer, then the
    native writer / parser against the restatement for a 32-layer model."""
    import ctypes as C
    import struct
    from vllm_rs_b200 import _lib
    kh, vh = bytes(range(64)), bytes(range(64, 128))
    hand = bytes([2, 0, 0, 0]) + bytes([9, 0, 0, 0, 0, 0, 0, 0]) + bytes([7, 0, 0, 0])      # TransferKvCache, seq 9, token 7
    hand += bytes([0, 0, 0, 0]) + bytes([1, 0, 0, 0, 0, 0, 0, 0])                           # LocalIpc, one layer
    for hb in (kh, vh):
        hand += bytes([64, 0, 0, 0, 0, 0, 0, 0]) + hb                                       # Vec<i8>: the handle
        hand += bytes([2, 0, 0, 0, 0, 0, 0, 0]) + bytes([5, 0, 0, 0, 0, 0, 0, 0]) + bytes([3, 0, 0, 0, 0, 0, 0, 0])   # shape [5, 3]
        hand += bytes([4, 0, 0, 0])                                                         # SerializableDType::F16
    hand += bytes([2, 0, 0, 0, 0, 0, 0, 0]) + bytes([11, 0, 0, 0, 12, 0, 0, 0])             # server_block_ids
    hand += bytes([100, 0, 0, 0, 0, 0, 0, 0]) + bytes([64, 0, 0, 0, 0, 0, 0, 0])            # sending_time, num_cached_tokens
    hand = struct.pack("<I", len(hand)) + hand
    assert ot.kv_ipc_frame(9, 7, [(kh, vh)], [5, 3], ot.DTYPE_F16, [11, 12], 100, 64) == hand

    rng = np.random.default_rng(4)
    L, shape, blocks = 32, [2112, 64, 8, 128], rng.integers(0, 2112, size=65).astype(np.uint32)
    k = rng.integers(0, 256, size=(L, 64), dtype=np.uint8)
    v = rng.integers(0, 256, size=(L, 64), dtype=np.uint8)
    want = ot.kv_ipc_frame(41, 777, [(k[l].tobytes(), v[l].tobytes()) for l in range(L)], shape, ot.DTYPE_BF16,
                           blocks.tolist(), 1760000000000, 128)
    h = _lib.PdKvHeader(seq_id=41, first_token=777, num_layers=L, layer_bytes=0, num_blocks=len(blocks),
                        sending_time=1760000000000, num_cached_tokens=128)
    sh = np.asarray(shape, np.uint64)
    out = np.zeros(len(want) + 8, np.uint8)
    n = lib.b200_pd_ipc_frame(out.ctypes.data, len(out), C.byref(h), k.ctypes.data, v.ctypes.data, sh.ctypes.data, 4,
                              ot.DTYPE_BF16, blocks.ctypes.data)
    assert n == len(want) and out[:n].tobytes() == want
    h2 = _lib.PdKvHeader()
    k2, v2 = np.zeros((L, 64), np.uint8), np.zeros((L, 64), np.uint8)
    sh2, nd, dt, b2 = np.zeros(8, np.uint64), C.c_int32(), C.c_int32(), np.zeros(128, np.uint32)
    assert lib.b200_pd_ipc_frame_parse(out.ctypes.data, n, C.byref(h2), k2.ctypes.data, v2.ctypes.data, L,
                                       sh2.ctypes.data, C.byref(nd), C.byref(dt), b2.ctypes.data, 128) == 0
    assert (h2.seq_id, h2.first_token, h2.num_layers, h2.num_blocks, h2.sending_time, h2.num_cached_tokens) == (
        41, 777, L, 65, 1760000000000, 128)
    assert np.array_equal(k2, k) and np.array_equal(v2, v) and sh2[:nd.value].tolist() == shape
    assert dt.value == ot.DTYPE_BF16 and np.array_equal(b2[:65], blocks)
    # refused: the RemoteTcp parser on a LocalIpc frame and vice versa, a truncated frame, too few layer slots
    koff, voff = np.zeros(L, np.int64), np.zeros(L, np.int64)
    assert lib.b200_pd_kv_frame_parse(out.ctypes.data, n, C.byref(h2), koff.ctypes.data, voff.ctypes.data, L) == -4
    assert lib.b200_pd_ipc_frame_parse(out.ctypes.data, n - 1, C.byref(h2), k2.ctypes.data, v2.ctypes.data, L,
                                       sh2.ctypes.data, C.byref(nd), C.byref(dt), b2.ctypes.data, 128) == -2
    assert lib.b200_pd_ipc_frame_parse(out.ctypes.data, n, C.byref(h2), k2.ctypes.data, v2.ctypes.data, 8,
                                       sh2.ctypes.data, C.byref(nd), C.byref(dt), b2.ctypes.data, 128) == -5
    lib.b200_clear_error()
