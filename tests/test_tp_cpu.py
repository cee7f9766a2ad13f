"""TP host logic on CPU: world_size-2 (and 4) gloo processes shard a quantised MLP + attention
projection with vllm_rs_b200.tp (the product's shard rules), run the oracle math on each shard,
all-reduce the row-parallel outputs (distributed.rs:438-454) and must reproduce the unsharded
result.  Pins: packed-tensor shard dims (wna16.rs:35-40), group alignment of row shards, kv-head
replication (distributed.rs:498-538)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import model as om
from oracle import quant as oq


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_case(fmt, seed=0):
    rng = np.random.default_rng(seed)
    hidden, inter, M = 256, 512, 5
    mk = (lambda K, N: oq.random_awq(rng, K, N)) if fmt == "awq" else (
        lambda K, N: oq.random_gptq(rng, K, N, 128, 4))
    gate, up, down = mk(hidden, inter), mk(hidden, inter), mk(inter, hidden)
    x = rng.standard_normal((M, hidden)).astype(np.float16)
    return dict(gate=gate, up=up, down=down, x=x, hidden=hidden, inter=inter)


def _deq(fmt, t):
    qw, qz, s = t
    return torch.from_numpy(oq.dequant_awq(qw, qz, s, 128) if fmt == "awq"
                            else oq.dequant_gptq(qw, qz, s, 128, 4))


def _mlp(fmt, x, gate, up, down):
    g, u = om.linear(x, _deq(fmt, gate)), om.linear(x, _deq(fmt, up))
    return om.silu_mul(g, u).float() @ _deq(fmt, down).float()      # fp32 partial (pre all-reduce)


def _worker(rank, world, port, fmt, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vllm_rs_b200 import tp
    c = _make_case(fmt)
    x = torch.from_numpy(c["x"])
    gate = tp.shard_column(*c["gate"], rank, world)
    up = tp.shard_column(*c["up"], rank, world)
    down = tp.shard_row(*c["down"], rank, world, 128)
    part = _mlp(fmt, x, gate, up, down).to(torch.float16).float()    # row-parallel output is fp16
    dist.all_reduce(part)                                            # SUM, like AllReduce::cuda_fwd
    if rank == 0:
        ret["tp"] = part.numpy()
        ret["kv"] = [tp.kv_head_shard(r, 8, 2) for r in range(8)]
    dist.destroy_process_group()


@pytest.mark.parametrize("fmt,world", [("awq", 2), ("gptq", 2), ("awq", 4)])
def test_sharded_mlp_matches_unsharded(fmt, world):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fmt, ret), nprocs=world, join=True)
    c = _make_case(fmt)
    full = _mlp(fmt, torch.from_numpy(c["x"]), c["gate"], c["up"], c["down"])
    # each rank rounds its partial to fp16 before the sum: error <= world * 1 ulp of the partials
    assert np.abs(ret["tp"] - full.numpy()).max() < 4e-3 * world
    # kv-head replication: 8 ranks, 2 kv heads -> ranks 0-3 share head 0, 4-7 head 1
    assert ret["kv"] == [(0, 2)] * 4 + [(1, 2)] * 4


def test_shard_shapes_and_group_alignment():
    from vllm_rs_b200 import tp
    rng = np.random.default_rng(0)
    qw, qz, s = oq.random_awq(rng, 512, 256)
    cw, cz, cs = tp.shard_column(qw, qz, s, 1, 2)
    assert cw.shape == (512, 16) and cz.shape == (4, 16) and cs.shape == (4, 128)
    assert np.array_equal(oq.unpack_awq(cw), oq.unpack_awq(qw)[:, 128:])
    rw, rz, rs = tp.shard_row(qw, qz, s, 1, 2, 128)
    assert rw.shape == (256, 32) and rz.shape == (2, 32) and rs.shape == (2, 256)
    assert np.array_equal(oq.dequant_awq(rw, rz, rs, 128), oq.dequant_awq(qw, qz, s, 128)[256:])
    gw, gz, gs = oq.random_gptq(rng, 512, 256, 128, 4)
    rw, rz, rs = tp.shard_row(gw, gz, gs, 0, 2, 128)
    assert rw.shape == (32, 256)                       # GPTQ packs along K: K/8 rows / 2
    assert np.array_equal(oq.dequant_gptq(rw, rz, rs, 128, 4), oq.dequant_gptq(gw, gz, gs, 128, 4)[:256])
    # Qwen2-72B TP=4 down_proj: 29568/128 = 231 groups is not divisible by 4
    bad = oq.random_awq(rng, 128 * 3, 128)
    with pytest.raises(AssertionError):
        tp.shard_row(*bad, 0, 2, 128)
    assert tp.kv_head_shard(3, 4, 8) == (3, 4)
    assert tp.padded_inter(29568, 4, 128) == 29696 and tp.padded_inter(29568, 8, 128) == 29696
    assert tp.padded_inter(14336, 8, 128) == 14336 and tp.padded_inter(28672, 8, 128) == 28672


@pytest.mark.parametrize("fmt", ["awq", "gptq"])
def test_pad_mlp_for_uneven_groups(fmt):
    """BASELINE configs[3] (Qwen2-72B GPTQ TP=4): inter / group_size is not divisible by the world size.  Scaled
    down: inter = 3 groups at TP=4 -> padded to 4 groups with zero scales: the padded MLP computes the SAME
    output as the original, its down_proj dequantises to the original rows followed by zeros, and it shards."""
    from vllm_rs_b200 import tp
    rng = np.random.default_rng(5)
    hidden, inter, world, G = 256, 3 * 128, 4, 128
    mk = (lambda K, N: oq.random_awq(rng, K, N)) if fmt == "awq" else (lambda K, N: oq.random_gptq(rng, K, N, G, 4))
    gate, up, down = mk(hidden, inter), mk(hidden, inter), mk(inter, hidden)
    inter_to = tp.padded_inter(inter, world, G)
    assert inter_to == 512
    pg, pu, pd = tp.pad_mlp(gate, up, down, inter_to, G, fmt)
    Wd, Wpd = _deq(fmt, down), _deq(fmt, pd)
    assert Wpd.shape == (inter_to, hidden) and torch.equal(Wpd[:inter], Wd) and not Wpd[inter:].any()
    Wg, Wpg = _deq(fmt, gate), _deq(fmt, pg)
    assert torch.equal(Wpg[:, :inter], Wg) and not Wpg[:, inter:].any()
    x = torch.from_numpy(rng.standard_normal((5, hidden)).astype(np.float16))
    full = _mlp(fmt, x, gate, up, down)
    padded = _mlp(fmt, x, pg, pu, pd)
    # same terms plus exact zeros; only the fp32 summation order of the (longer) K loop may differ
    assert (full - padded).abs().max() < 1e-4
    # and now every rank gets whole groups and whole 128-channel tiles
    acc = torch.zeros_like(full)
    for r in range(world):
        g_r, u_r = tp.shard_column(*pg, r, world), tp.shard_column(*pu, r, world)
        d_r = tp.shard_row(*pd, r, world, G)
        assert g_r[2].shape[1] % 128 == 0 and d_r[2].shape[0] * G == inter_to // world
        acc += _mlp(fmt, x, g_r, u_r, d_r)
    assert (acc - full).abs().max() < 2e-3
