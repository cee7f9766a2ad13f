"""Floating-point oracle: LLaMA / Qwen2 decoder forward over a paged KV cache (torch, CPU).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, op by op,

  /root/reference/src/models/llama.rs:107-131,269-321        decoder layer / forward_inner
  /root/reference/src/models/layers/attention.rs:648-839     Attention::forward_ext
  /root/reference/src/models/layers/mlp.rs:451-469           MLP::forward
  /root/reference/src/models/layers/others.rs:11-66          NormX / RmsNorm
  /root/reference/src/models/layers/rotary_emb.rs:32-73      cos/sin tables (f32 math, cast)
  /root/reference/src/models/layers/rotary_emb.rs:208-278    llama3 rope scaling
  /root/reference/src/utils/kvcache_allocator.rs:150-173     cache layout L2 [nb, bs, Hkv, D]
  /root/reference/src/core/runner.rs:1243-1388               decode metadata semantics

Arithmetic model: tensors the reference materialises in the model dtype are rounded to fp16
at the same points (linear outputs, RoPE outputs, attention output, residual adds, norm
outputs, silu, gate*up); inside an op the math is fp32 (what cuBLAS / tensor cores / the
reference's CUDA kernels accumulate in).  Logits are fp16-rounded then cast to f32
(llama.rs:315-319).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

F16, F32 = torch.float16, torch.float32

# Rounding model: h() marks every point where the reference materialises a model-dtype (fp16)
# tensor.  `exact_mode()` turns those roundings off (fp32 everywhere, fp32 KV cache) to give the
# "infinitely precise" result of the same weights -- used by tests to size the fp16 noise floor.
_EXACT = False
_DT = F16          # the model dtype h() rounds to: fp16 (the north-star dtype) or bf16 (the reference's default)


class exact_mode:
    def __enter__(self):
        global _EXACT
        self._old, _EXACT = _EXACT, True

    def __exit__(self, *a):
        global _EXACT
        _EXACT = self._old


class model_dtype:
    """`with model_dtype(torch.bfloat16):` -- the same rounding model with bf16 as the materialised dtype
    (utils/mod.rs:1919-1925: bf16 is what the reference picks unless the user asks for f16)."""

    def __init__(self, dt):
        self.dt = dt

    def __enter__(self):
        global _DT
        self._old, _DT = _DT, self.dt

    def __exit__(self, *a):
        global _DT
        _DT = self._old


def h(x: torch.Tensor) -> torch.Tensor:
    return x.to(F32) if _EXACT else x.to(_DT)


def store_dtype():
    return F32 if _EXACT else _DT


@dataclass
class ModelConfig:
    hidden: int
    num_layers: int
    num_heads: int
    num_kv_heads: int
    head_dim: int
    inter: int
    vocab: int
    rope_theta: float = 10000.0
    rms_eps: float = 1e-5
    max_pos: int = 8192
    qkv_bias: bool = False
    block_size: int = 64
    rope_scaling: dict | None = None

    @staticmethod
    def tinyllama():
        return ModelConfig(2048, 22, 32, 4, 64, 5632, 32000, 10000.0, 1e-5, 2048)

    @staticmethod
    def llama3_8b():
        return ModelConfig(4096, 32, 32, 8, 128, 14336, 128256, 500000.0, 1e-5, 8192)


def rope_tables(cfg: ModelConfig, dtype=F16):
    """rotary_emb.rs:32-73 (default) and :208-278 (llama3)."""
    D = cfg.head_dim
    inv_freq = torch.tensor(
        [float(1.0 / (cfg.rope_theta ** (i / D))) for i in range(0, D, 2)], dtype=F32)
    rs = cfg.rope_scaling
    if rs and rs.get("rope_type", rs.get("type")) == "llama3":
        factor, lo, hi = rs["factor"], rs["low_freq_factor"], rs["high_freq_factor"]
        orig = rs["original_max_position_embeddings"]
        low_wl, high_wl = orig / lo, orig / hi
        out = []
        for f in inv_freq.tolist():
            wl = 2 * math.pi / f
            if wl < high_wl:
                out.append(f)
            elif wl > low_wl:
                out.append(f / factor)
            else:
                smooth = (orig / wl - lo) / (hi - lo)
                out.append((1 - smooth) * f / factor + smooth * f)
        inv_freq = torch.tensor(out, dtype=F32)
    t = torch.arange(cfg.max_pos, dtype=F32).reshape(-1, 1)
    freqs = t @ inv_freq.reshape(1, -1)
    return freqs.cos().to(dtype), freqs.sin().to(dtype)


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    xf = x.to(F32)
    inv = torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return h(xf * inv * w.to(F32))


def normx(x: torch.Tensor, w: torch.Tensor, eps: float, norm_dtype=None) -> torch.Tensor:
    """NormX::forward (/root/reference/src/models/layers/others.rs:11-30) with the norm weight loaded by
    rms_norm_sharded (others.rs:43-66).  llama.rs builds its norms in the model dtype (norm_dtype None): the fused
    RmsNorm runs on the fp16 tensor -- fp32 inside, ONE rounding of x * rsqrt(mean(x^2) + eps) * w to fp16.
    qwen3.rs:83,95,261 (Qwen2 / Qwen3) builds them with DType::F32: the weight is cast to f32 at load (exact:
    every fp16 is an f32), the input is cast fp16 -> f32 (exact), the norm runs in f32 and the result is cast
    back to the input dtype -- again ONE rounding of the same f32 product.  The two branches are therefore the
    same function of (x, w) up to the f32 association order of the three factors, which the final fp16 rounding
    absorbs except on exact ties; tests/test_oracle_cpu.py::test_normx_f32_round_trip_is_the_fp16_norm pins that
    on 1e6 random elements, and the CUDA kernels (glue.cu rmsnorm_kernel, the folded norms of the GEMM finalisers)
    are compared against this one definition for both model families."""
    if norm_dtype is None or norm_dtype == x.dtype:
        return rmsnorm(x, w, eps)
    in_dtype = x.dtype
    conv = x.to(norm_dtype)                                   # others.rs:15
    wf = w.to(norm_dtype)                                     # others.rs:53-55
    inv = torch.rsqrt(conv.pow(2).mean(-1, keepdim=True) + eps)
    out = conv * inv * wf                                     # RmsNorm on the converted tensor (others.rs:17)
    return out.to(F32) if _EXACT else out.to(in_dtype)        # others.rs:20


def apply_rope(q, k, cos, sin, positions, interleaved=False):
    """q [T,H,D], k [T,Hkv,D] fp16; NeoX (rotate-half) unless interleaved.  fp32 math."""
    c = cos[positions].to(F32).unsqueeze(1)      # [T,1,D/2]
    s = sin[positions].to(F32).unsqueeze(1)

    def rot(x):
        xf = x.to(F32)
        if interleaved:
            x0, x1 = xf[..., 0::2], xf[..., 1::2]
            y = torch.stack((x0 * c - x1 * s, x1 * c + x0 * s), dim=-1).flatten(-2)
        else:
            hd = xf.shape[-1] // 2
            x0, x1 = xf[..., :hd], xf[..., hd:]
            y = torch.cat((x0 * c - x1 * s, x1 * c + x0 * s), dim=-1)
        return h(y)

    return rot(q), rot(k)


E4M3_MAX = 448.0


def to_e4m3(x: torch.Tensor) -> torch.Tensor:
    """fp16 -> e4m3 bytes of the FP8 KV cache (`--fp8-kvcache`: kvcache_allocator.rs:188-193,776 makes the
    cache U8; attention.rs:615 passes only the flag on to attention-rs, whose storage format is not
    source-visible -- PARITY UNPINNED).  Defined here as: scale 1, round to nearest even, saturate."""
    return x.to(F32).clamp(-E4M3_MAX, E4M3_MAX).to(torch.float8_e4m3fn).view(torch.uint8)


def from_e4m3(b: torch.Tensor) -> torch.Tensor:
    return b.view(torch.float8_e4m3fn).to(F32)


def kv_write(k, v, k_cache, v_cache, slot_mapping):
    """cache [nb, bs, Hkv, D] (fp16, or uint8 = e4m3); slot = block*bs + offset (runner.rs:1259-1262)."""
    nb, bs = k_cache.shape[0], k_cache.shape[1]
    kc = k_cache.view(nb * bs, *k_cache.shape[2:])
    vc = v_cache.view(nb * bs, *v_cache.shape[2:])
    fp8 = k_cache.dtype == torch.uint8
    for t, slot in enumerate(slot_mapping.tolist()):
        if slot >= 0:
            kc[slot] = to_e4m3(k[t]) if fp8 else k[t]
            vc[slot] = to_e4m3(v[t]) if fp8 else v[t]


def to_l1_layout(k_cache, v_cache):
    """Flash layout [nb, bs, Hkv, D] -> the reference's default-build layout (kvcache_allocator.rs:150-173, no
    flashattn / flashinfer feature): K [nb, Hkv, D/x, bs, x] with x = 16 / element size, V [nb, Hkv, D, bs]."""
    nb, bs, Hkv, D = k_cache.shape
    x = 16 // k_cache.element_size()
    k1 = k_cache.view(nb, bs, Hkv, D // x, x).permute(0, 2, 3, 1, 4).contiguous()
    v1 = v_cache.permute(0, 2, 3, 1).contiguous()
    return k1, v1


def from_l1_layout(k1, v1):
    nb, Hkv, dx, bs, x = k1.shape
    return (k1.permute(0, 3, 1, 2, 4).reshape(nb, bs, Hkv, dx * x).contiguous(),
            v1.permute(0, 3, 1, 2).contiguous())


def gather_kv(cache, block_table, ctx_len):
    bs = cache.shape[1]
    nblk = -(-ctx_len // bs)
    blocks = cache[torch.as_tensor(block_table[:nblk], dtype=torch.long)]
    if cache.dtype == torch.uint8:
        blocks = from_e4m3(blocks)
    return blocks.reshape(nblk * bs, *cache.shape[2:])[:ctx_len]     # [ctx, Hkv, D]


def paged_attention_decode(q, k_cache, v_cache, block_tables, context_lens, scale,
                           softcap=0.0, window=0):
    """q [B,H,D] fp16 -> out [B,H,D] fp16.  One query token per sequence attends to
    tokens [0, ctx) (or the last `window`) of its block table; fp32 softmax."""
    B, H, D = q.shape
    Hkv = k_cache.shape[2]
    G = H // Hkv
    out = torch.zeros(B, H, D, dtype=store_dtype())
    for b in range(B):
        ctx = int(context_lens[b])
        if ctx == 0:
            continue
        k = gather_kv(k_cache, block_tables[b], ctx).to(F32)       # [ctx,Hkv,D]
        v = gather_kv(v_cache, block_tables[b], ctx).to(F32)
        lo = max(0, ctx - window) if window > 0 else 0
        k, v = k[lo:], v[lo:]
        qf = q[b].to(F32).view(Hkv, G, D)
        s = torch.einsum("hgd,thd->hgt", qf, k) * scale
        if softcap > 0:
            s = softcap * torch.tanh(s / softcap)
        p = torch.softmax(s, dim=-1)
        o = torch.einsum("hgt,thd->hgd", p, v)
        out[b] = h(o.reshape(H, D))
    return out


def attention_prefill(q, k, v, cu_q, k_cache=None, v_cache=None, block_tables=None,
                      context_lens=None, scale=1.0):
    """Causal varlen attention for one prefill step.  q [T,H,D]; k/v [T,Hkv,D] are this step's
    new tokens; with cached context (chunk >= 2) keys come from the paged cache, which
    already holds the new tokens (the cache write happens first, attention.rs:808-820)."""
    T, H, D = q.shape
    Hkv = k.shape[1]
    G = H // Hkv
    out = torch.zeros(T, H, D, dtype=store_dtype())
    for i in range(len(cu_q) - 1):
        a, b = cu_q[i], cu_q[i + 1]
        nq = b - a
        if context_lens is not None:
            ctx = int(context_lens[i])
            kk = gather_kv(k_cache, block_tables[i], ctx).to(F32)
            vv = gather_kv(v_cache, block_tables[i], ctx).to(F32)
        else:
            ctx = nq
            kk, vv = k[a:b].to(F32), v[a:b].to(F32)
        qf = q[a:b].to(F32).view(nq, Hkv, G, D)
        s = torch.einsum("qhgd,thd->hgqt", qf, kk) * scale
        qpos = torch.arange(ctx - nq, ctx).view(nq, 1)
        mask = torch.arange(ctx).view(1, ctx) > qpos
        s = s.masked_fill(mask, float("-inf"))
        p = torch.softmax(s, dim=-1)
        o = torch.einsum("hgqt,thd->qhgd", p, vv)
        out[a:b] = h(o.reshape(nq, H, D))
    return out


def linear(x: torch.Tensor, w_kn: torch.Tensor, bias=None) -> torch.Tensor:
    """x [T,K] fp16, w_kn [K,N] fp16 (dequantised or dense^T) -> fp16, fp32 accumulate."""
    y = x.to(F32) @ w_kn.to(F32)
    if bias is not None:
        y = y + bias.to(F32)
    return h(y)


def silu_mul(gate: torch.Tensor, up: torch.Tensor) -> torch.Tensor:
    """mlp.rs:467-468: act(gate) in fp16, then * up in fp16."""
    g = gate.to(F32)
    act = h(g * torch.sigmoid(g))
    return h(act.to(F32) * up.to(F32))


@dataclass
class LayerWeights:
    input_norm: torch.Tensor
    post_norm: torch.Tensor
    wq: torch.Tensor        # [K, N] fp16 (already dequantised for quantised checkpoints)
    wk: torch.Tensor
    wv: torch.Tensor
    wo: torch.Tensor
    wgate: torch.Tensor
    wup: torch.Tensor
    wdown: torch.Tensor
    bq: torch.Tensor | None = None
    bk: torch.Tensor | None = None
    bv: torch.Tensor | None = None


@dataclass
class ModelWeights:
    embed: torch.Tensor     # [vocab, hidden] fp16
    final_norm: torch.Tensor
    lm_head: torch.Tensor   # [vocab, hidden] fp16
    layers: list = field(default_factory=list)


class LlamaOracle:
    """LLaMaForCausalLM::forward_inner over a paged cache (one (K,V) pair per layer)."""

    def __init__(self, cfg: ModelConfig, weights: ModelWeights, num_blocks: int, kv_fp8: bool = False):
        self.cfg, self.w = cfg, weights
        self.cos, self.sin = rope_tables(cfg)
        shape = (num_blocks, cfg.block_size, cfg.num_kv_heads, cfg.head_dim)
        self.kv_fp8 = kv_fp8
        dt = torch.uint8 if kv_fp8 else store_dtype()
        self.k_cache = [torch.zeros(shape, dtype=dt) for _ in range(cfg.num_layers)]
        self.v_cache = [torch.zeros(shape, dtype=dt) for _ in range(cfg.num_layers)]
        self.scale = 1.0 / math.sqrt(cfg.head_dim)

    def _layer(self, l, x, positions, slot_mapping, attn_fn, trace=None):
        cfg, lw = self.cfg, self.w.layers[l]
        H, Hkv, D = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
        T = x.shape[0]
        residual = x
        xn = rmsnorm(x, lw.input_norm, cfg.rms_eps)
        q = linear(xn, lw.wq, lw.bq).view(T, H, D)
        k = linear(xn, lw.wk, lw.bk).view(T, Hkv, D)
        v = linear(xn, lw.wv, lw.bv).view(T, Hkv, D)
        q, k = apply_rope(q, k, self.cos, self.sin, positions)
        kv_write(k, v, self.k_cache[l], self.v_cache[l], slot_mapping)
        a = attn_fn(l, q, k, v)
        o = linear(a.reshape(T, H * D), lw.wo)
        x = h(o.to(F32) + residual.to(F32))              # llama.rs:126
        residual = x
        xn = rmsnorm(x, lw.post_norm, cfg.rms_eps)
        gate, up = linear(xn, lw.wgate), linear(xn, lw.wup)
        m = linear(silu_mul(gate, up), lw.wdown)
        x = h(residual.to(F32) + m.to(F32))              # llama.rs:130
        if trace is not None:
            trace.append(x.clone())
        return x

    def _head(self, x):
        xn = rmsnorm(x, self.w.final_norm, self.cfg.rms_eps)
        return linear(xn, self.w.lm_head.t()).to(F32)           # llama.rs:315-319

    def prefill(self, meta: dict, trace=None) -> torch.Tensor:
        """meta = oracle.sched.prepare_prefill(...).  Returns f32 logits [num_seqs, vocab]."""
        ids = torch.as_tensor(meta["input_ids"], dtype=torch.long)
        pos = torch.as_tensor(meta["positions"], dtype=torch.long)
        slots = torch.as_tensor(meta["slot_mapping"], dtype=torch.long)
        cu_q = meta["cu_seqlens_q"]
        x = h(self.w.embed[ids])
        bts, ctxs = meta["block_tables"], meta["context_lens"]
        if self.kv_fp8 and ctxs is None:
            # FP8 cache: keys / values are ALWAYS read back from the cache (written first), so the result
            # does not depend on chunking or prefix hits; needs the tables of every sequence
            bts = meta["all_block_tables"]
            ctxs = [cu_q[i + 1] - cu_q[i] for i in range(len(cu_q) - 1)]

        def attn(l, q, k, v):
            return attention_prefill(q, k, v, cu_q, self.k_cache[l], self.v_cache[l], bts, ctxs, self.scale)

        for l in range(self.cfg.num_layers):
            x = self._layer(l, x, pos, slots, attn, trace)
        last = torch.as_tensor([c - 1 for c in cu_q[1:]], dtype=torch.long)   # llama.rs:306-310
        return self._head(x[last])

    def decode(self, meta: dict, trace=None) -> torch.Tensor:
        """meta = oracle.sched.prepare_decode(...).  Returns f32 logits [B, vocab]."""
        ids = torch.as_tensor(meta["input_ids"], dtype=torch.long)
        pos = torch.as_tensor(meta["positions"], dtype=torch.long)
        slots = torch.as_tensor(meta["slot_mapping"], dtype=torch.long)
        x = h(self.w.embed[ids])

        def attn(l, q, k, v):
            return paged_attention_decode(q, self.k_cache[l], self.v_cache[l],
                                          meta["block_tables"], meta["context_lens"], self.scale)

        for l in range(self.cfg.num_layers):
            x = self._layer(l, x, pos, slots, attn, trace)
        return self._head(x)


def random_weights(cfg: ModelConfig, seed: int = 0, std: float = 0.02) -> ModelWeights:
    """Dense fp16 random-init weights (N(0, std)), norms ~ 1 + N(0, 0.05)."""
    g = torch.Generator().manual_seed(seed)

    def rnd(*shape, s=std):
        return (torch.randn(*shape, generator=g) * s).to(F16)

    def norm_w():
        return (1.0 + torch.randn(cfg.hidden, generator=g) * 0.05).to(F16)

    H, Hkv, D = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
    w = ModelWeights(embed=rnd(cfg.vocab, cfg.hidden, s=1.0), final_norm=norm_w(),
                     lm_head=rnd(cfg.vocab, cfg.hidden))
    for _ in range(cfg.num_layers):
        lw = LayerWeights(
            input_norm=norm_w(), post_norm=norm_w(),
            wq=rnd(cfg.hidden, H * D), wk=rnd(cfg.hidden, Hkv * D), wv=rnd(cfg.hidden, Hkv * D),
            wo=rnd(H * D, cfg.hidden), wgate=rnd(cfg.hidden, cfg.inter),
            wup=rnd(cfg.hidden, cfg.inter), wdown=rnd(cfg.inter, cfg.hidden))
        if cfg.qkv_bias:
            lw.bq, lw.bk, lw.bv = rnd(H * D, s=0.1), rnd(Hkv * D, s=0.1), rnd(Hkv * D, s=0.1)
        w.layers.append(lw)
    return w
