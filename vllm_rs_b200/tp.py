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
