"""Host-side operator layer over the C ABI (torch tensors are only memory + streams).

Names follow the reference's layer vocabulary (src/models/layers/*.rs): WNA16 repack +
gptq_matmul, FusedRope, PagedAttention, rms_norm ...  Every function launches hand-written
sm_100a kernels from libb200_vllm.so; there is no CPU or eager-torch fallback.
"""
from __future__ import annotations

import torch

from . import _lib

QFMT_AWQ, QFMT_GPTQ = 0, 1
EPI_NONE, EPI_SILU_MUL = 0, 1
VARIANT_AUTO, VARIANT_GEMV, VARIANT_TC = 0, 1, 2


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return 0 if t is None else t.data_ptr()


def _sym(stem: str, t: torch.Tensor):
    """The `_f16` or `_bf16` flavour of an entry point, by the dtype of the activations (bf16 is the reference's
    default dtype, utils/mod.rs:1919-1925)."""
    if t.dtype == torch.bfloat16:
        return getattr(_lib.lib(), stem + "_bf16")
    if t.dtype != torch.float16:
        raise TypeError(f"{stem}: activations must be fp16 or bf16, got {t.dtype}")
    return getattr(_lib.lib(), stem + "_f16")


def _cuda(*ts):
    for t in ts:
        if t is not None and not (t.is_cuda and t.is_contiguous()):
            raise ValueError("b200 ops need contiguous CUDA tensors")


# ------------------------------------------------------------------ quantised linear
def awq_prepack(qweight: torch.Tensor, bits: int = 4) -> torch.Tensor:
    """MarlinRepack for AWQ (gptq.rs:314-324): qweight i32 [K, N/8] -> tile-major u32."""
    _cuda(qweight)
    out = torch.empty_like(qweight)
    _lib.lib().b200_awq_prepack(_p(qweight), _p(out), qweight.shape[0], qweight.shape[1], bits,
                                _stream())
    _lib.check()
    return out


def gptq_prepack(qweight: torch.Tensor, bits: int = 4) -> torch.Tensor:
    """MarlinRepack for GPTQ (gptq.rs:325-333): qweight i32 [K/pack, N] -> tile-major u32."""
    _cuda(qweight)
    out = torch.empty_like(qweight)
    _lib.lib().b200_gptq_prepack(_p(qweight), _p(out), qweight.shape[0], qweight.shape[1], bits,
                                 _stream())
    _lib.check()
    return out


def wna16_workspace(max_m: int, max_n: int, device="cuda") -> torch.Tensor:
    n = _lib.lib().b200_wna16_workspace_bytes(max_m, max_n)
    return torch.zeros(n, dtype=torch.uint8, device=device)


def wna16_gemm(x, w_packed, scales, qzeros, workspace, *, k: int, n: int, bits=4, group_size=128,
               qfmt=QFMT_AWQ, bias=None, epilogue=EPI_NONE, variant=VARIANT_AUTO, out=None):
    """gptq_matmul (gptq.rs:243-263): y[M, N] = x[M, K] @ dequant(W)."""
    _cuda(x, w_packed, scales, qzeros, bias, workspace)
    m = x.shape[0]
    n_out = n // 2 if epilogue == EPI_SILU_MUL else n
    if out is None:
        out = torch.empty(m, n_out, dtype=torch.float16, device=x.device)
    _lib.lib().b200_wna16_gemm_ex(_p(x), _p(w_packed), _p(scales), _p(qzeros), _p(bias), _p(out),
                                  m, k, n, bits, group_size, qfmt, epilogue, variant,
                                  _p(workspace), _stream())
    _lib.check()
    return out


DENSE_AUTO, DENSE_CUBLAS, DENSE_TC1, DENSE_TC2 = -1, 0, 1, 2


def dense_gemm(x, w_nk, out_f32=False, out=None, bias=None, mode=DENSE_AUTO):
    """Linear::forward (linear.rs:75-123): y = x @ W^T (+ bias), W [N, K] fp16."""
    _cuda(x, w_nk, bias)
    m, k = x.shape
    n = w_nk.shape[0]
    if out is None:
        out = torch.empty(m, n, dtype=torch.float32 if out_f32 else torch.float16, device=x.device)
    _lib.lib().b200_dense_gemm_ex(_p(x), _p(w_nk), _p(bias), _p(out), m, n, k, int(out_f32), mode, _stream())
    _lib.check()
    return out


def wna16_dequant(w_packed, scales, qzeros, *, k: int, n: int, bits=4, group_size=128, qfmt=QFMT_AWQ, out=None):
    """Prepacked W4 / W8 weight -> fp16 [N, K] (the weight operand of dense_gemm)."""
    _cuda(w_packed, scales, qzeros)
    if out is None:
        out = torch.empty(n, k, dtype=torch.float16, device=w_packed.device)
    _lib.lib().b200_wna16_dequant_f16(_p(w_packed), _p(scales), _p(qzeros), _p(out), k, n, bits, group_size,
                                      qfmt, _stream())
    _lib.check()
    return out


# ------------------------------------------------------------------ attention side
def _is_l1(k_cache, v_cache) -> bool:
    """The reference's default-build layout (kvcache_allocator.rs:150-173): K [nb, Hkv, D/8, 64, 8] is 5-D."""
    return k_cache is not None and k_cache.dim() == 5


def rope_kv_write(q, k, v, cos, sin, positions, slot_mapping=None, k_cache=None, v_cache=None,
                  rot_dim=None, interleaved=False):
    """FusedRope::apply_inplace + cache write (rotary_emb.rs:88-104, attention.rs:808-820)."""
    _cuda(q, k, v, cos, sin, positions, slot_mapping, k_cache, v_cache)
    T, H, D = q.shape
    Hkv = k.shape[1]
    if _is_l1(k_cache, v_cache):
        _lib.lib().b200_rope_kv_write_l1(
            _p(q), _p(k), _p(v), _p(cos), _p(sin), _p(positions), _p(slot_mapping), _p(k_cache), _p(v_cache), T, H,
            Hkv, D, rot_dim or D, k_cache.shape[3], int(interleaved), int(cos.dtype == torch.float32),
            int(q.dtype == torch.bfloat16), _stream())
        _lib.check()
        return
    bs = k_cache.shape[1] if k_cache is not None else 64
    fp8 = k_cache is not None and k_cache.dtype == torch.uint8        # --fp8-kvcache: e4m3 bytes
    fn = _lib.lib().b200_rope_kv_write_fp8kv_f16 if fp8 else _sym("b200_rope_kv_write", q)
    fn(
        _p(q), _p(k), _p(v), _p(cos), _p(sin), _p(positions), _p(slot_mapping), _p(k_cache),
        _p(v_cache), T, H, Hkv, D, rot_dim or D, bs, int(interleaved),
        int(cos.dtype == torch.float32), _stream())
    _lib.check()


def kv_write(k, v, k_cache, v_cache, slot_mapping):
    _cuda(k, v, k_cache, v_cache, slot_mapping)
    T, Hkv, D = k.shape
    if _is_l1(k_cache, v_cache):
        _lib.lib().b200_kv_write_l1(_p(k), _p(v), _p(k_cache), _p(v_cache), _p(slot_mapping), T, Hkv, D,
                                    k_cache.shape[3], _stream())
        _lib.check()
        return
    _lib.lib().b200_kv_write_f16(_p(k), _p(v), _p(k_cache), _p(v_cache), _p(slot_mapping), T, Hkv,
                                 D, k_cache.shape[1], _stream())
    _lib.check()


def swap_blocks(src_cache, dst_cache, mapping):
    """attention_rs::cache::swap_blocks (runner.rs:1626-1645): dst[d] = src[s] for every (s, d) of `mapping`
    (dict or list of pairs).  Caches are [num_blocks, ...] tensors on the GPU or in PINNED host memory."""
    pairs = list(mapping.items()) if isinstance(mapping, dict) else list(mapping)
    if not pairs:
        return
    for t in (src_cache, dst_cache):
        assert t.is_contiguous() and (t.is_cuda or t.is_pinned()), "caches must be CUDA or pinned host tensors"
    block_bytes = src_cache[0].numel() * src_cache.element_size()
    assert block_bytes == dst_cache[0].numel() * dst_cache.element_size()
    dev = src_cache.device if src_cache.is_cuda else dst_cache.device
    pd = torch.tensor(pairs, dtype=torch.int64).to(dev)
    _lib.lib().b200_swap_blocks(_p(src_cache), _p(dst_cache), _p(pd), len(pairs), block_bytes, _stream())
    _lib.check()
    return pd            # keep alive until the stream has run the copy


def paged_attn_workspace(batch, num_heads, head_dim, max_blocks, device="cuda"):
    n = _lib.lib().b200_paged_attn_workspace_bytes(batch, num_heads, head_dim, max_blocks)
    return torch.zeros(n, dtype=torch.uint8, device=device)     # arrival counters start at rest


def paged_attention_decode(q, k_cache, v_cache, block_tables, context_lens, scale, workspace=None,
                           softcap=0.0, window=0, out=None):
    """PagedAttention::forward, decode branch.  q [B,H,D]; caches [nb, 64, Hkv, D];
    block_tables u32/i32 [B, max_blocks]; context_lens u32/i32 [B]."""
    _cuda(q, k_cache, v_cache, block_tables, context_lens)
    B, H, D = q.shape
    l1 = _is_l1(k_cache, v_cache)
    nb, bs, Hkv = (k_cache.shape[0], k_cache.shape[3], k_cache.shape[1]) if l1 else k_cache.shape[:3]
    mb = block_tables.shape[1]
    if workspace is None:
        workspace = paged_attn_workspace(B, H, D, mb, q.device)
    if out is None:
        out = torch.empty_like(q)
    if l1:
        _lib.lib().b200_paged_attn_decode_l1(
            _p(out), _p(q), _p(k_cache), _p(v_cache), _p(block_tables), _p(context_lens), B, H, Hkv, D, bs, mb, nb,
            float(scale), float(softcap), int(window), _p(workspace), int(q.dtype == torch.bfloat16), _stream())
        _lib.check()
        return out
    fp8 = k_cache.dtype == torch.uint8                                # --fp8-kvcache: e4m3 bytes
    fn = _lib.lib().b200_paged_attn_decode_fp8kv_f16 if fp8 else _sym("b200_paged_attn_decode", q)
    fn(_p(out), _p(q), _p(k_cache), _p(v_cache), _p(block_tables), _p(context_lens), B, H, Hkv, D, bs,
       mb, nb, float(scale), float(softcap), int(window), _p(workspace), _stream())
    _lib.check()
    return out


ATTN_AUTO, ATTN_MMA, ATTN_TC, ATTN_TC_SWAPDESC = -1, 0, 1, 2


def attention_prefill(q, k_cache, v_cache, cu_seqlens_q, context_lens, block_tables, max_seqlen_q,
                      scale, out=None, mode=ATTN_AUTO):
    """PagedAttention::forward, prefill branch over the paged cache (see b200_attn_prefill_ex_f16)."""
    _cuda(q, k_cache, v_cache, cu_seqlens_q, context_lens, block_tables)
    T, H, D = q.shape
    l1 = _is_l1(k_cache, v_cache)
    nb, bs, Hkv = (k_cache.shape[0], k_cache.shape[3], k_cache.shape[1]) if l1 else k_cache.shape[:3]
    B = cu_seqlens_q.shape[0] - 1
    if out is None:
        out = torch.empty_like(q)
    if l1:
        _lib.lib().b200_attn_prefill_l1(
            _p(out), _p(q), _p(k_cache), _p(v_cache), _p(cu_seqlens_q), _p(context_lens), _p(block_tables), B, H,
            Hkv, D, bs, block_tables.shape[1], nb, int(max_seqlen_q), T, mode, float(scale),
            int(q.dtype == torch.bfloat16), _stream())
        _lib.check()
        return out
    _sym("b200_attn_prefill_ex", q)(_p(out), _p(q), _p(k_cache), _p(v_cache), _p(cu_seqlens_q),
                                    _p(context_lens), _p(block_tables), B, H, Hkv, D, bs,
                                    block_tables.shape[1], nb, int(max_seqlen_q), T, mode, float(scale),
                                    _stream())
    _lib.check()
    return out


# ------------------------------------------------------------------ glue
def rmsnorm(x, w, eps, out=None):
    _cuda(x, w)
    out = torch.empty_like(x) if out is None else out
    _sym("b200_rmsnorm", x)(_p(out), _p(x), _p(w), x.shape[0], x.shape[1], eps, _stream())
    _lib.check()
    return out


def add_rmsnorm(residual, x, w, eps, out=None):
    """residual += x (in place); returns rmsnorm(residual) * w."""
    _cuda(residual, x, w)
    out = torch.empty_like(x) if out is None else out
    _sym("b200_add_rmsnorm", x)(_p(out), _p(residual), _p(x), _p(w), x.shape[0], x.shape[1], eps, _stream())
    _lib.check()
    return out


def silu_mul(gate_up, out=None):
    _cuda(gate_up)
    T, n2 = gate_up.shape
    out = torch.empty(T, n2 // 2, dtype=gate_up.dtype, device=gate_up.device) if out is None else out
    _sym("b200_silu_mul", gate_up)(_p(out), _p(gate_up), T, n2 // 2, _stream())
    _lib.check()
    return out


def embedding(table, ids, out=None):
    _cuda(table, ids)
    T = ids.shape[0]
    out = torch.empty(T, table.shape[1], dtype=table.dtype, device=table.device) if out is None else out
    _lib.lib().b200_embedding_f16(_p(out), _p(table), _p(ids), T, table.shape[1], table.shape[0],
                                  _stream())
    _lib.check()
    return out


def argmax(logits):
    _cuda(logits)
    out = torch.empty(logits.shape[0], dtype=torch.int32, device=logits.device)
    _lib.lib().b200_argmax_f32(_p(out), _p(logits), logits.shape[0], logits.shape[1], _stream())
    _lib.check()
    return out


# ------------------------------------------------------------------ sampling
def sample_topk_topp(logits, top_k, top_p, temperature, seed, out=None):
    """Device sampler (logits_processor.rs:199-231): logits f32 [B, V] -> ids i32 [B].  `seed` is a
    python int (copied to a device u64) or a 1-element int64 CUDA tensor."""
    _cuda(logits)
    B, V = logits.shape
    if not torch.is_tensor(seed):
        seed = torch.tensor([seed - (1 << 64) if seed >= (1 << 63) else seed], dtype=torch.int64).to(logits.device)
    if out is None:
        out = torch.empty(B, dtype=torch.int32, device=logits.device)
    _lib.lib().b200_sample_topk_topp_f32(_p(out), _p(logits), B, V, top_k, float(top_p), float(temperature),
                                         _p(seed), _stream())
    _lib.check()
    return out


def apply_penalties(logits, context, context_lens, frequency_penalty, presence_penalty):
    """apply_batch_repeat_penalty (logits_processor.rs:308-343), in place.  context i32 [B, cap]."""
    _cuda(logits, context, context_lens)
    B, V = logits.shape
    _lib.lib().b200_apply_penalties_f32(_p(logits), B, V, _p(context), _p(context_lens), context.shape[1],
                                        float(frequency_penalty), float(presence_penalty), _stream())
    _lib.check()
    return logits
