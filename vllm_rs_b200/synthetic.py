"""Synthetic (random-init) checkpoints in the reference's on-disk formats, for benches and
parity tests (there is no network for real checkpoints).  Shapes / packing follow
src/models/layers/wna16.rs:56-83,127-152; value distributions follow SURVEY.md section 8(d)
config 2: uniform int4 codes and zero points, scales U(0.5,1.5)*(2/15)/sqrt(K).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import ops
from .engine import ModelRunner, QuantLinear, fuse_awq

SHAPES = {
    # hidden, layers, heads, kv_heads, head_dim, inter, vocab, rope_theta
    "llama3-8b": (4096, 32, 32, 8, 128, 14336, 128256, 500000.0),
    "tinyllama-1.1b": (2048, 22, 32, 4, 64, 5632, 32000, 10000.0),
    "llama3-70b": (8192, 80, 64, 8, 128, 28672, 128256, 500000.0),
    # Qwen2-72B: q/k/v bias, rope theta 1e6; inter 29568 = 231 groups of 128 (padded per TP degree, tp.padded_inter)
    "qwen2-72b": (8192, 80, 64, 8, 128, 29568, 152064, 1000000.0),
    "toy": (512, 2, 8, 2, 64, 1024, 1024, 10000.0),
}


def rope_tables_device(head_dim, max_pos, theta, device):
    inv = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.float64) / head_dim))
    inv = inv.to(torch.float32)
    t = torch.arange(max_pos, dtype=torch.float32).reshape(-1, 1)
    f = t @ inv.reshape(1, -1)
    return f.cos().to(torch.float16).to(device), f.sin().to(torch.float16).to(device)


def _rand_awq_device(gen, K, N, group, device):
    """AWQ tensors created directly on the GPU (random packed words are uniform nibbles)."""
    qw = torch.randint(-2**31, 2**31 - 1, (K, N // 8), generator=gen, device=device,
                       dtype=torch.int32)
    qz = torch.randint(-2**31, 2**31 - 1, (K // group, N // 8), generator=gen, device=device,
                       dtype=torch.int32)
    s = (torch.rand(K // group, N, generator=gen, device=device) + 0.5) * (2.0 / 15.0) / math.sqrt(K)
    return qw, qz, s.to(torch.float16)


def _fused_awq_device(gen, K, Ns, group, device):
    """Random AWQ projections sharing an input, fused along N then prepacked once."""
    N = sum(Ns)
    qw, qz, s = _rand_awq_device(gen, K, N, group, device)
    return QuantLinear(ops.awq_prepack(qw).reshape(-1), s, qz, K, N)


def _fused_gptq_sym_device(gen, K, Ns, group, device):
    """Random symmetric GPTQ-int4 projections (zero point 8: qzeros words 0x77777777, v1 checkpoints store z - 1)."""
    N = sum(Ns)
    qw = torch.randint(-2**31, 2**31 - 1, (K // 8, N), generator=gen, device=device, dtype=torch.int32)
    qz = torch.full((K // group, N // 8), 0x77777777, device=device, dtype=torch.int32)
    s = ((torch.rand(K // group, N, generator=gen, device=device) + 0.5) * (2.0 / 15.0) / math.sqrt(K)).to(torch.float16)
    return QuantLinear(ops.gptq_prepack(qw).reshape(-1), s, qz, K, N)


def _fused_dense_device(gen, K, Ns, group, device):
    """fp16 checkpoint: dense [sum N, K] (packed qkv / gate_up, attention.rs:207,224-228), N(0, 1/sqrt(K))."""
    N = sum(Ns)
    w = torch.empty(N, K, device=device, dtype=torch.float16)
    step = max(1, (1 << 26) // K)
    for i in range(0, N, step):                  # bounded fp32 temporaries
        w[i:i + step] = (torch.randn(min(step, N - i), K, generator=gen, device=device) / math.sqrt(K)).to(torch.float16)
    return QuantLinear(w, None, None, K, N)


def build_runner(name="llama3-8b", fmt="awq", *, num_blocks, max_blocks_per_seq, max_batch, seed=0,
                 group=128, num_layers=None, tp_size=1, tp_rank=0, device="cuda",
                 max_prefill_tokens=0, kv_fp8=False) -> ModelRunner:
    """Random-init model of the named architecture in checkpoint format `fmt` ("awq" | "gptq_sym" | "fp16"),
    optionally ONE tensor-parallel shard of it: q/k/v/gate/up column-sharded, o/down row-sharded, kv heads
    replicated when Hkv < tp_size (attention.rs:445-530, mlp.rs:367-442, distributed.rs:526-537); the
    intermediate size is padded to whole groups / tiles per rank (tp.padded_inter: Qwen2-72B 29568 -> 29696).
    Replicated tensors (embedding, norms, lm_head) use the same seed on every rank; the shards differ."""
    from . import tp as tpmod
    hidden, L, H, Hkv, D, inter, vocab, theta = SHAPES[name]
    L = num_layers or L
    qkv_bias = name.startswith("qwen2")
    if fmt != "fp16":
        inter = tpmod.padded_inter(inter, tp_size, group)
    Hl, Hkvl, interl = H // tp_size, max(1, Hkv // tp_size), inter // tp_size
    bits = 16 if fmt == "fp16" else 4
    r = ModelRunner(hidden=hidden, num_layers=L, num_heads=Hl, num_kv_heads=Hkvl, head_dim=D,
                    inter=interl, vocab=vocab, num_blocks=num_blocks,
                    max_blocks_per_seq=max_blocks_per_seq, max_batch=max_batch, bits=bits,
                    group_size=group, qfmt=ops.QFMT_AWQ if fmt == "awq" else ops.QFMT_GPTQ, device=device,
                    qkv_bias=qkv_bias, max_prefill_tokens=max_prefill_tokens, kv_fp8=kv_fp8)
    gen = torch.Generator(device=device).manual_seed(seed)                       # replicated tensors
    gen_r = torch.Generator(device=device).manual_seed(seed * 1000 + 17 + tp_rank)   # this rank's shards
    fused = {"awq": _fused_awq_device, "gptq_sym": _fused_gptq_sym_device, "fp16": _fused_dense_device}[fmt]

    def norm_w():
        return (1.0 + 0.05 * torch.randn(hidden, generator=gen, device=device)).to(torch.float16)

    for l in range(L):
        bias = None
        if qkv_bias:
            bias = (0.1 * torch.randn((Hl + 2 * Hkvl) * D, generator=gen_r, device=device)).to(torch.float16)
        r.set_layer(
            l, input_norm=norm_w(), post_norm=norm_w(),
            qkv=fused(gen_r, hidden, [Hl * D, Hkvl * D, Hkvl * D], group, device),
            o=fused(gen_r, Hl * D, [hidden], group, device),
            gate_up=fused(gen_r, hidden, [interl, interl], group, device),
            down=fused(gen_r, interl, [hidden], group, device), qkv_bias=bias)
    embed = torch.randn(vocab, hidden, generator=gen, device=device).to(torch.float16)
    lm_head = (0.02 * torch.randn(vocab, hidden, generator=gen, device=device)).to(torch.float16)
    cos, sin = rope_tables_device(D, 8192, theta, device)
    r.set_globals(embed, norm_w(), lm_head, cos, sin)
    r.shape = dict(hidden=hidden, L=L, H=Hl, Hkv=Hkvl, D=D, inter=interl, vocab=vocab, group=group, bits=bits)
    return r


def build_awq_runner(name="llama3-8b", **kw) -> ModelRunner:
    """Random AWQ-int4 model (BASELINE.json configs[1]); see build_runner."""
    return build_runner(name, "awq", **kw)


def fill_kv_random(r: ModelRunner, seed=1):
    """KV cache ~ N(0,1) fp16 without a huge temporary."""
    gen = torch.Generator(device=r.device).manual_seed(seed)
    for kc, vc in zip(r.k_cache, r.v_cache):
        if kc.dtype == torch.uint8:          # --fp8-kvcache: e4m3 bytes of the same N(0,1) values
            for c in (kc, vc):
                c.copy_(torch.randn(c.shape, generator=gen, device=r.device, dtype=torch.float16)
                        .to(torch.float8_e4m3fn).view(torch.uint8))
            continue
        kc.normal_(generator=gen)
        vc.normal_(generator=gen)


def decode_algorithmic_bytes(shape: dict, batch: int, ctx: int, bits=None, kv_bytes=2) -> dict:
    """Algorithmic HBM bytes of one decode step of ONE rank (SURVEY.md 8(d) / BASELINE.md section 4)."""
    h, L, H, Hkv, D, inter, V, G = (shape[k] for k in
                                    ("hidden", "L", "H", "Hkv", "D", "inter", "vocab", "group"))
    bits = bits or shape.get("bits", 4)
    per_layer_kn = h * (H + 2 * Hkv) * D + H * D * h + h * 2 * inter + inter * h
    W = L * per_layer_kn * bits // 8
    S = 0 if bits == 16 else L * per_layer_kn // G * (2 + bits / 8)
    LM = V * h * 2
    KV = batch * ctx * L * 2 * Hkv * D * kv_bytes
    KV_new = batch * L * 2 * Hkv * D * kv_bytes
    return dict(weights=W, scales_zeros=int(S), lm_head=LM, kv_read=KV, kv_write=KV_new,
                total=int(W + S + LM + KV + KV_new))
