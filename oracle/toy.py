"""Toy quantised checkpoint -> (CPU oracle, GPU runner) pair, shared by tests/ and
__graft_entry__.smoke().  TEST INFRASTRUCTURE (see oracle/__init__.py)."""
from __future__ import annotations

import numpy as np
import torch

from . import model as om
from . import quant as oq

F16 = torch.float16


def toy_cfg_d128(qkv_bias=False):
    """head_dim 128 variant (the FP8 KV cache tiles are one 128-byte row per token)."""
    return om.ModelConfig(hidden=512, num_layers=2, num_heads=4, num_kv_heads=2, head_dim=128,
                          inter=1024, vocab=1024, rope_theta=10000.0, rms_eps=1e-5, max_pos=2048,
                          qkv_bias=qkv_bias)


def toy_cfg(qkv_bias=False):
    return om.ModelConfig(hidden=512, num_layers=2, num_heads=8, num_kv_heads=2, head_dim=64,
                          inter=1024, vocab=1024, rope_theta=10000.0, rms_eps=1e-5, max_pos=2048,
                          qkv_bias=qkv_bias)


def build_pair(cfg, fmt, seed, num_blocks, max_batch, max_blocks, max_prefill_tokens=0, kv_fp8=False):
    """Random quantised checkpoint -> (oracle with dequantised weights, GPU runner)."""
    from vllm_rs_b200 import engine, ops
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    H, Hkv, D, h, I = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, cfg.hidden, cfg.inter
    bits = 8 if fmt == "gptq8" else 4

    def qlayer(K, N):
        if fmt == "awq":
            qw, qz, s = oq.random_awq(rng, K, N)
            s = (s.astype(np.float32) * 1.5).astype(np.float16)     # keep activations O(1) deep in the net
            return (qw, qz, s), torch.from_numpy(oq.dequant_awq(qw, qz, s, 128))
        qw, qz, s = oq.random_gptq(rng, K, N, 128, bits)
        s = (s.astype(np.float32) * 1.5).astype(np.float16)
        return (qw, qz, s), torch.from_numpy(oq.dequant_gptq(qw, qz, s, 128, bits))

    def norm_w():
        return (1 + 0.05 * torch.randn(h, generator=g)).to(F16)

    fuse = engine.fuse_awq if fmt == "awq" else (lambda p, d: engine.fuse_gptq(p, d, bits))
    runner = engine.ModelRunner(
        hidden=h, num_layers=cfg.num_layers, num_heads=H, num_kv_heads=Hkv, head_dim=D, inter=I,
        vocab=cfg.vocab, num_blocks=num_blocks, max_blocks_per_seq=max_blocks, max_batch=max_batch,
        bits=bits, group_size=128, qfmt=ops.QFMT_AWQ if fmt == "awq" else ops.QFMT_GPTQ,
        qkv_bias=cfg.qkv_bias, max_pos=cfg.max_pos, max_prefill_tokens=max_prefill_tokens, kv_fp8=kv_fp8)
    w = om.ModelWeights(embed=torch.randn(cfg.vocab, h, generator=g).to(F16), final_norm=norm_w(),
                        lm_head=(0.05 * torch.randn(cfg.vocab, h, generator=g)).to(F16))
    for l in range(cfg.num_layers):
        (pq, Wq), (pk, Wk), (pv, Wv) = qlayer(h, H * D), qlayer(h, Hkv * D), qlayer(h, Hkv * D)
        (po, Wo), (pg, Wg), (pu, Wu), (pd, Wd) = (qlayer(H * D, h), qlayer(h, I), qlayer(h, I),
                                                  qlayer(I, h))
        lw = om.LayerWeights(norm_w(), norm_w(), Wq, Wk, Wv, Wo, Wg, Wu, Wd)
        bias = None
        if cfg.qkv_bias:
            lw.bq, lw.bk, lw.bv = [(0.1 * torch.randn(n, generator=g)).to(F16)
                                   for n in (H * D, Hkv * D, Hkv * D)]
            bias = torch.cat([lw.bq, lw.bk, lw.bv]).cuda()
        w.layers.append(lw)
        runner.set_layer(l, input_norm=lw.input_norm.cuda(), post_norm=lw.post_norm.cuda(),
                         qkv=fuse([pq, pk, pv], "cuda"), o=fuse([po], "cuda"),
                         gate_up=fuse([pg, pu], "cuda"), down=fuse([pd], "cuda"), qkv_bias=bias)
    oracle = om.LlamaOracle(cfg, w, num_blocks, kv_fp8=kv_fp8)
    runner.set_globals(w.embed.cuda(), w.final_norm.cuda(), w.lm_head.cuda(), oracle.cos.cuda(),
                       oracle.sin.cuda())
    return oracle, runner


def sync_cache_to_gpu(oracle, runner):
    for l in range(oracle.cfg.num_layers):
        runner.k_cache[l].copy_(oracle.k_cache[l])
        runner.v_cache[l].copy_(oracle.v_cache[l])
