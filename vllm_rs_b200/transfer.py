"""PD (prefill / decode) KV hand-off over the C ABI: mirrors the data side of
/root/reference/src/transfer/mod.rs (Transfer::transfer_kv_cache / Transfer::receive_kv_cache) -- same names and
argument meaning -- on top of b200_pd_* (csrc/host/transfer.cpp).  Sockets and the listener thread are the host's."""
from __future__ import annotations

import ctypes as C
import time

import numpy as np
import torch

from . import _lib


def _ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])


def _block_bytes(cache):
    return cache[0].numel() * cache.element_size()


def transfer_kv_cache(k_caches, v_caches, block_table, *, seq_id: int, first_token: int, num_cached_tokens: int = 0,
                      sending_time: int | None = None, stream=None) -> torch.Tensor:
    """(Server, PdMethod::RemoteTcp, mod.rs:370-470) the frame the server sends after a prefill: every layer's K / V
    blocks of the sequence gathered in block-table order, written by the copy kernel straight into a pinned frame.
    Returns the frame (pinned uint8 tensor, complete once `stream` has run)."""
    l = _lib.lib()
    L, n = len(k_caches), len(block_table)
    bb = _block_bytes(k_caches[0])
    total = l.b200_pd_kv_frame_bytes(L, n * bb)
    frame = torch.empty(total, dtype=torch.uint8).pin_memory()
    hdr = _lib.PdKvHeader(seq_id, first_token, L, n * bb, n,
                          int(time.time() * 1000) if sending_time is None else sending_time, num_cached_tokens)
    k_off, v_off = np.zeros(L, np.int64), np.zeros(L, np.int64)
    got = l.b200_pd_kv_frame_begin(frame.data_ptr(), total, C.byref(hdr), k_off.ctypes.data, v_off.ctypes.data)
    _lib.check()
    assert got == total
    dev = k_caches[0].device
    pairs = torch.tensor([[int(b), i] for i, b in enumerate(block_table)], dtype=torch.int64).to(dev)
    st = (stream or torch.cuda.current_stream(dev)).cuda_stream
    l.b200_pd_export_blocks(_ptr_array(k_caches), _ptr_array(v_caches), L, pairs.data_ptr(), n, bb, frame.data_ptr(),
                            k_off.ctypes.data, v_off.ctypes.data, st)
    _lib.check()
    frame._keep = pairs
    return frame


def receive_kv_cache(frame: torch.Tensor, k_caches, v_caches, local_block_ids, stream=None) -> dict:
    """(Client, RemoteTcp, mod.rs:292-330) copy the i-th received block of every layer into the client's i-th
    allocated block, reading the payloads straight out of the (pinned) frame.  Returns the header fields."""
    l = _lib.lib()
    L = len(k_caches)
    assert frame.is_pinned() and frame.dtype == torch.uint8
    hdr = _lib.PdKvHeader()
    k_off, v_off = np.zeros(L, np.int64), np.zeros(L, np.int64)
    rc = l.b200_pd_kv_frame_parse(frame.data_ptr(), frame.numel(), C.byref(hdr), k_off.ctypes.data, v_off.ctypes.data, L)
    if rc != 0:
        raise ValueError(f"not a TransferKvCache / RemoteTcp frame for {L} layers (code {rc})")
    n = int(hdr.num_blocks)
    if len(local_block_ids) < n:
        raise ValueError("Not enough blocks allocated to receive KV cache")          # mod.rs:296-298
    bb = _block_bytes(k_caches[0])
    if int(hdr.layer_bytes) != n * bb:
        raise ValueError("frame block size does not match the local cache")
    dev = k_caches[0].device
    pairs = torch.tensor([[i, int(b)] for i, b in zip(range(n), local_block_ids)], dtype=torch.int64).to(dev)
    st = (stream or torch.cuda.current_stream(dev)).cuda_stream
    l.b200_pd_import_blocks(_ptr_array(k_caches), _ptr_array(v_caches), L, pairs.data_ptr(), n, bb, frame.data_ptr(),
                            k_off.ctypes.data, v_off.ctypes.data, st)
    _lib.check()
    return dict(seq_id=int(hdr.seq_id), first_token=int(hdr.first_token), num_blocks=n,
                sending_time=int(hdr.sending_time), num_cached_tokens=int(hdr.num_cached_tokens), _keep=pairs)


def receive_kv_cache_local_ipc(remote_k, remote_v, server_block_ids, k_caches, v_caches, local_block_ids, stream=None):
    """(Client, PdMethod::LocalIpc, mod.rs:268-290) copy server block i -> local block i for every layer, the kernel
    on the client GPU reading the server GPU's caches over NVLink (IPC-opened or peer-mapped tensors)."""
    l = _lib.lib()
    if len(local_block_ids) < len(server_block_ids):
        raise ValueError("Not enough blocks allocated to receive KV cache")
    dev = k_caches[0].device
    if remote_k[0].device != dev:
        rc = l.b200_enable_peer_access(dev.index, remote_k[0].device.index)
        if rc != 0:
            raise RuntimeError(f"no peer access from {dev} to {remote_k[0].device} (code {rc})")
    pairs = torch.tensor([[int(s), int(d)] for s, d in zip(server_block_ids, local_block_ids)],
                         dtype=torch.int64).to(dev)
    st = (stream or torch.cuda.current_stream(dev)).cuda_stream
    l.b200_pd_copy_blocks_peer(_ptr_array(remote_k), _ptr_array(remote_v), _ptr_array(k_caches), _ptr_array(v_caches),
                               len(k_caches), pairs.data_ptr(), len(server_block_ids), _block_bytes(k_caches[0]), st)
    _lib.check()
    return pairs
