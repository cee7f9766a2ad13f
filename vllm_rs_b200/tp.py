"""Tensor parallelism: Megatron column/row sharding of packed quantised tensors and the
communicator (NCCL + one-shot NVLink all-reduce) behind the C ABI.

Shard rules restate the reference:
  wna16.rs:35-40,56-83,127-132   logical shard(0) (out dim) -> tensor dim 1, shard(1) (in dim) ->
                                 tensor dim 0, applied alike to qweight / scales / qzeros
  distributed.rs:498-538         kv_head_shard: replicate kv heads when Hkv < world
  attention.rs:445-530, mlp.rs:367-442   q/k/v/gate/up column-parallel, o/down row-parallel
  distributed.rs:438-454         all-reduce SUM on the row-parallel output, bias afterwards
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


# ------------------------------------------------------------------ shard arithmetic (host)
def kv_head_shard(rank: int, world: int, num_kv_heads: int):
    """(shard index, number of shards) for k/v projections (distributed.rs:526-537)."""
    if num_kv_heads >= world:
        assert num_kv_heads % world == 0
        return rank, world
    assert world % num_kv_heads == 0
    ranks_per_kv_head = world // num_kv_heads
    return rank // ranks_per_kv_head, num_kv_heads


def _chunk(a, dim, idx, n):
    size = a.shape[dim]
    assert size % n == 0, f"dim {dim} of {a.shape} not divisible by {n}"
    step = size // n
    sl = [slice(None)] * a.ndim
    sl[dim] = slice(idx * step, (idx + 1) * step)
    return np.ascontiguousarray(a[tuple(sl)])


def shard_column(qweight, qzeros, scales, idx, n):
    """Out-dim shard: tensor dim 1 of qweight / qzeros / scales (AWQ and GPTQ alike)."""
    return (_chunk(qweight, 1, idx, n), None if qzeros is None else _chunk(qzeros, 1, idx, n),
            _chunk(scales, 1, idx, n))


def shard_row(qweight, qzeros, scales, idx, n, group_size):
    """In-dim shard: tensor dim 0; K/n must be a multiple of the group size (SURVEY section 7:
    Qwen2-72B inter 29568 / 4 is not -> pad to 29696)."""
    k_groups = scales.shape[0]
    assert k_groups % n == 0, f"{k_groups} groups (K={k_groups * group_size}) not divisible by {n}"
    return (_chunk(qweight, 0, idx, n), None if qzeros is None else _chunk(qzeros, 0, idx, n),
            _chunk(scales, 0, idx, n))


def padded_inter(inter: int, world: int, group_size: int) -> int:
    """Smallest intermediate size >= inter whose row shards are whole quantisation groups AND whole 128-channel
    weight tiles per rank: a multiple of world * lcm(group_size, 128).  Qwen2-72B: 29568 -> 29696 at TP=4 (the
    size Qwen2.5 checkpoints ship with), 231 -> 232 groups."""
    unit = world * int(np.lcm(group_size, 128))
    return -(-inter // unit) * unit


def pad_mlp(gate, up, down, inter_to: int, group_size: int, fmt: str, bits: int = 4):
    """Pad a quantised MLP from `inter` to `inter_to` intermediate channels without changing its output.
    gate / up get extra OUTPUT channels, down gets extra INPUT rows; every added scale is 0, so the added
    channels compute silu(0) * 0 = 0 and the added down rows dequantise to (q - z) * 0 = 0 whatever the packed
    bits hold (they are zero-filled).  Each argument is the checkpoint triple (qweight, qzeros, scales):
      AWQ   qweight [K, N/8],    qzeros [K/G, N/8],    scales [K/G, N]
      GPTQ  qweight [K/pack, N], qzeros [K/G, N/pack], scales [K/G, N]        (wna16.rs:56-83,127-152)
    `inter_to - inter` must be a multiple of the group size (added down rows = whole groups)."""
    pack = 32 // bits

    def pad_cols(t, extra_n):                               # more output channels (tensor dim 1 everywhere)
        qw, qz, s = t
        add_packed = extra_n // pack if fmt == "awq" else extra_n
        qw2 = np.concatenate([qw, np.zeros((qw.shape[0], add_packed), qw.dtype)], axis=1)
        qz2 = None if qz is None else np.concatenate([qz, np.zeros((qz.shape[0], extra_n // pack), qz.dtype)], axis=1)
        s2 = np.concatenate([s, np.zeros((s.shape[0], extra_n), s.dtype)], axis=1)
        return qw2, qz2, s2

    def pad_rows(t, extra_k):                               # more input rows (tensor dim 0 everywhere)
        qw, qz, s = t
        add_packed = extra_k if fmt == "awq" else extra_k // pack
        qw2 = np.concatenate([qw, np.zeros((add_packed, qw.shape[1]), qw.dtype)], axis=0)
        g = extra_k // group_size
        qz2 = None if qz is None else np.concatenate([qz, np.zeros((g, qz.shape[1]), qz.dtype)], axis=0)
        s2 = np.concatenate([s, np.zeros((g, s.shape[1]), s.dtype)], axis=0)
        return qw2, qz2, s2

    inter = gate[2].shape[1]
    extra = inter_to - inter
    assert extra >= 0 and extra % group_size == 0 and extra % pack == 0, (inter, inter_to, group_size)
    if extra == 0:
        return gate, up, down
    return pad_cols(gate, extra), pad_cols(up, extra), pad_rows(down, extra)


# ------------------------------------------------------------------ communicator
class Comm:
    """Comm::from_rank equivalent (runner.rs:81-89): NCCL communicator from a 128-byte id."""

    def __init__(self, id_bytes: bytes, rank: int, world: int, oneshot: bool = True, stream: int = 0):
        self._l = _lib.lib()
        buf = (C.c_uint8 * 128).from_buffer_copy(id_bytes)
        self._h = self._l.b200_comm_init_rank(buf, rank, world)
        _lib.check()
        if not self._h:
            raise _lib.B200Error("comm_init_rank failed")
        self.rank, self.world = rank, world
        self.oneshot = False
        if oneshot and world > 1:
            self.oneshot = self._l.b200_comm_enable_oneshot(self._h, stream) == 0

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_uint8 * 128)()
        rc = _lib.lib().b200_nccl_unique_id(buf)
        _lib.check()
        if rc != 0:
            raise _lib.B200Error("nccl unique id failed")
        return bytes(buf)

    def all_reduce_(self, t, stream=None):
        """In-place SUM over ranks of an fp16 CUDA tensor."""
        import torch
        sid = (stream or torch.cuda.current_stream()).cuda_stream
        self._l.b200_allreduce_f16(self._h, t.data_ptr(), t.numel(), sid)
        _lib.check()
        return t

    def force_nccl(self, on: bool):
        self._l.b200_comm_set_force_nccl(self._h, int(on))

    def close(self):
        if getattr(self, "_h", None):
            self._l.b200_comm_free(self._h)
            self._h = None


def init_comm_from_torch_dist(oneshot=True) -> Comm:
    """Rank 0 creates the NCCL id and ships it to the others (engine.rs:263 ships it inside the
    runner Init message); torch.distributed is only the courier."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    obj = [Comm.unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    comm = Comm(obj[0], rank, world, oneshot=oneshot, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    dist.barrier()
    return comm


def attach_allreduce(runner, oneshot=True) -> Comm:
    import os
    comm = init_comm_from_torch_dist(oneshot)
    if os.environ.get("B200_FORCE_NCCL") == "1":       # A/B arm: every all-reduce through ncclAllReduce
        comm.force_nccl(True)
    runner._l.b200_runner_set_comm(runner._h, comm._h)
    runner._comm = comm
    return comm
