"""Engine <-> runner-process protocol of a step, over any stream socket (harness mirror of
/root/reference/src/runner/mod.rs:248-295 `send_local` / `receive_local`, and of the runner binary's message loop,
runner/runner.rs:249-434, for the messages of the hot path).

A message is one frame (u32 little-endian length + bincode 1.3.1 body); the receiver answers every frame with ONE
acknowledgment byte 0x01 before it acts on it.  The frames themselves are built and parsed by the native library
(`b200_wire_*`, `b200_scheduler_{frame,load}_run_prefill`); what a runner does with a step's metadata is the caller's
`step_fn` -- on a GPU box `ModelRunner.prefill_device` / `run_decode`, in the CPU tests a deterministic stand-in.
The sockets, process spawning and the Init handshake (EngineConfig) are the host's business and not mirrored.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .engine import Scheduler

MT_RUN_PREFILL, MT_RUN_DECODE, MT_RUN_RESPONSE, MT_FINISH_DECODE, MT_HEARTBEAT, MT_SHUTDOWN = 3, 4, 5, 8, 14, 34
MT_KV_SWAP, MT_KV_SWAP_RESPONSE = 21, 22


def _recv_exact(sock, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(n - len(buf))
        if not chunk:
            raise ConnectionError("peer closed the stream inside a frame")
        buf += chunk
    return bytes(buf)


def send_local(sock, frame: bytes) -> None:
    """runner/mod.rs:248-275: write the frame, then wait for the acknowledgment byte."""
    sock.sendall(frame)
    if _recv_exact(sock, 1) != b"\x01":
        raise _lib.B200Error("unexpected acknowledgment value")


def receive_local(sock) -> bytes:
    """runner/mod.rs:277-295: read one frame, acknowledge it; returns the whole frame (length prefix included)."""
    head = _recv_exact(sock, 4)
    body = _recv_exact(sock, int.from_bytes(head, "little"))
    sock.sendall(b"\x01")
    return head + body


def _kind(frame: bytes) -> int:
    buf = np.frombuffer(frame, np.uint8)
    return _lib.lib().b200_wire_kind(buf.ctypes.data, len(buf))


def _simple(kind: int, payload_bytes: int = 0, value: int = 0) -> bytes:
    out = np.zeros(32, np.uint8)
    n = _lib.lib().b200_wire_frame_simple(out.ctypes.data, 32, kind, payload_bytes, value)
    return out[:n].tobytes()


def _response(ids) -> bytes:
    a = np.ascontiguousarray(ids, dtype=np.uint32)
    out = np.zeros(16 + 4 * len(a), np.uint8)
    n = _lib.lib().b200_wire_frame_run_response(out.ctypes.data, len(out), a.ctypes.data, len(a))
    return out[:n].tobytes()


class WireEngine:
    """The engine side: one step of `sched`'s schedule goes to the runner as RunPrefill / RunDecode and comes back as
    RunResponse (engine.rs:865-1010 drives `send_local` / `receive_local` the same way)."""

    def __init__(self, sched: Scheduler, sock, max_blocks=1024):
        self.sched, self.sock, self.max_blocks = sched, sock, max_blocks
        self._l = _lib.lib()

    def _recv_ids(self, n: int) -> list:
        fr = np.frombuffer(receive_local(self.sock), np.uint8)
        ids = np.zeros(max(n, 1), np.uint32)
        got = self._l.b200_wire_parse_run_response(fr.ctypes.data, len(fr), ids.ctypes.data, len(ids))
        if got != n:
            raise _lib.B200Error(f"RunResponse carries {got} ids for {n} sequences")
        return ids[:n].tolist()

    def run_prefill(self, ids) -> list:
        send_local(self.sock, self.sched.frame_run_prefill(ids, True))
        return self._recv_ids(len(ids))

    def run_decode(self, ids) -> list:
        s, n, mb = self.sched, len(ids), self.max_blocks
        sid, lt, ln, lbt = np.zeros(n, np.uint64), np.zeros(n, np.uint32), np.zeros(n, np.uint64), np.zeros(n, np.uint64)
        btl, bt, nblk = np.zeros(n, np.uint32), np.zeros((n, mb), np.uint32), np.zeros(n, np.int32)
        for r, i in enumerate(ids):                    # DecodeSequence::new (sequence.rs:64-78)
            info, table = s.seq_info(i), s.block_table(i)
            sid[r], ln[r], lt[r] = info["id"], info["len"], s.tokens(i)[-1]
            lbt[r] = info["len"] - (-(-info["len"] // 64) - 1) * 64       # last_block_num_tokens (sequence.rs:203-213)
            btl[r], nblk[r] = table[-1], len(table)
            bt[r, :len(table)] = table
        out = np.zeros(64 + n * (128 + 4 * mb), np.uint8)
        m = self._l.b200_wire_frame_run_decode(out.ctypes.data, len(out), n, sid.ctypes.data, lt.ctypes.data,
                                               ln.ctypes.data, lbt.ctypes.data, btl.ctypes.data, bt.ctypes.data, mb,
                                               nblk.ctypes.data, None, 0)
        _lib.check()
        send_local(self.sock, out[:m].tobytes())
        return self._recv_ids(n)

    def kv_swap(self, pairs, swap_in: bool) -> bool:
        """KVCacheSwap((HashMap<usize, usize>, bool)) -> KVCacheSwapResponse(bool) (runner.rs:1626-1645 runs it)."""
        a = np.ascontiguousarray(pairs, dtype=np.int64).reshape(-1, 2)
        out = np.zeros(32 + 16 * len(a), np.uint8)
        n = self._l.b200_wire_frame_kv_swap(out.ctypes.data, len(out), a.ctypes.data, len(a), int(swap_in))
        send_local(self.sock, out[:n].tobytes())
        fr = receive_local(self.sock)
        if _kind(fr) != MT_KV_SWAP_RESPONSE:
            raise _lib.B200Error("expected KVCacheSwapResponse")
        return fr[8] == 1

    def finish(self, seq_id: int) -> None:
        send_local(self.sock, _simple(MT_FINISH_DECODE, 8, seq_id))

    def shutdown(self) -> None:
        send_local(self.sock, _simple(MT_SHUTDOWN))


def serve(sock, step_fn, *, num_blocks=1 << 16, block_size=64, max_seqs=256, max_blocks=1024, on_finish=None,
          on_swap=None) -> int:
    """The runner side: answer RunPrefill / RunDecode with RunResponse(step_fn(kind, metadata)) until Shutdown; returns
    the number of steps served.  metadata is what prepare_prefill / prepare_decode give for the frame's sequences."""
    l = _lib.lib()
    box = Scheduler(num_blocks, block_size, max_seqs)          # a plain container for RunPrefill's sequences
    steps = 0
    while True:
        fr = receive_local(sock)
        buf = np.frombuffer(fr, np.uint8)
        kind = _kind(fr)
        if kind == MT_SHUTDOWN:
            return steps
        if kind == MT_RUN_PREFILL:
            ids, _, _ = box.load_run_prefill(fr, cap=max_seqs)
            meta = box.prepare_prefill(ids, max_blocks_cap=max_blocks)
            meta["seq_ids"] = [box.seq_info(i)["id"] for i in ids]
            meta["num_cached_tokens"] = [box.seq_info(i)["num_cached_tokens"] for i in ids]
            meta["seq_block_tables"] = [box.block_table(i) for i in ids]
            send_local(sock, _response(step_fn("prefill", meta)))
            steps += 1
        elif kind == MT_RUN_DECODE:
            sid, lt, ln, lbt = (np.zeros(max_seqs, np.uint64), np.zeros(max_seqs, np.uint32),
                                np.zeros(max_seqs, np.uint64), np.zeros(max_seqs, np.uint64))
            btl, bt, nblk = np.zeros(max_seqs, np.uint32), np.zeros((max_seqs, max_blocks), np.uint32), np.zeros(max_seqs, np.int32)
            pf = C.c_int32()
            n = l.b200_wire_parse_run_decode(buf.ctypes.data, len(buf), max_seqs, max_blocks, sid.ctypes.data,
                                             lt.ctypes.data, ln.ctypes.data, lbt.ctypes.data, btl.ctypes.data,
                                             bt.ctypes.data, nblk.ctypes.data, None, C.byref(pf))
            if n < 0:
                raise _lib.B200Error(f"RunDecode: {n}")
            ii, pos, slot, cl = np.zeros(n, np.uint32), np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.uint32)
            l.b200_wire_decode_metadata(n, lt.ctypes.data, ln.ctypes.data, lbt.ctypes.data, btl.ctypes.data, block_size,
                                        ii.ctypes.data, pos.ctypes.data, slot.ctypes.data, cl.ctypes.data)
            mb = int(nblk[:n].max()) if n else 0
            meta = dict(input_ids=ii, positions=pos, slot_mapping=slot, context_lens=cl,
                        block_tables=bt[:n, :mb].copy(), max_blocks=mb, seq_ids=sid[:n].copy())
            send_local(sock, _response(step_fn("decode", meta)))
            steps += 1
        elif kind == MT_KV_SWAP:
            pairs, flag = np.zeros((4096, 2), np.int64), C.c_int32()
            n = l.b200_wire_parse_kv_swap(buf.ctypes.data, len(buf), pairs.ctypes.data, 4096, C.byref(flag))
            ok = n >= 0 and on_swap is not None
            if ok:
                on_swap(pairs[:n].tolist(), bool(flag.value))
            send_local(sock, _simple(MT_KV_SWAP_RESPONSE, 1, 1 if ok else 0))
        elif kind == MT_FINISH_DECODE:
            if on_finish:
                on_finish(int.from_bytes(fr[8:16], "little"))
        elif kind == MT_HEARTBEAT:
            pass
        else:
            raise _lib.B200Error(f"runner loop: unexpected message {kind}")
