"""GPU parity: every CUDA kernel through the C ABI vs the CPU oracle (oracle/).

Bars: integer/byte outputs bit-exact (prepack, argmax, cache contents); floating point within
the tolerance written in each test (north star: logits within 1e-3 in fp16 terms; single ops
are held to ~1 fp16 ulp of their output scale).
"""
import math

import numpy as np
import pytest
import torch

from oracle import model as om
from oracle import quant as oq

pytestmark = pytest.mark.gpu
F16 = torch.float16


@pytest.fixture(scope="module")
def ops(lib):
    from vllm_rs_b200 import ops as _ops
    return _ops


def _close(a, b, atol, rtol=0.0):
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    assert bool((err <= tol).all()), f"max err {err.max().item():.3e} (tol {atol}, {rtol})"


# ---------------------------------------------------------------- glue
@pytest.mark.parametrize("T,hidden", [(1, 2048), (32, 4096), (7, 8192)])
def test_rmsnorm(ops, T, hidden):
    g = torch.Generator().manual_seed(T)
    x = torch.randn(T, hidden, generator=g).to(F16)
    w = (1 + 0.1 * torch.randn(hidden, generator=g)).to(F16)
    ref = om.rmsnorm(x, w, 1e-5)
    out = ops.rmsnorm(x.cuda(), w.cuda(), 1e-5)
    _close(out, ref, 2e-3, 1e-3)          # 1 fp16 ulp at |x| ~ 4
    # fused residual add: residual rounded to fp16 first (llama.rs:126-130)
    res = torch.randn(T, hidden, generator=g).to(F16)
    res_d = res.cuda().clone()
    out2 = ops.add_rmsnorm(res_d, x.cuda(), w.cuda(), 1e-5)
    new_res = (res.float() + x.float()).to(F16)
    assert torch.equal(res_d.cpu(), new_res)
    _close(out2, om.rmsnorm(new_res, w, 1e-5), 2e-3, 1e-3)


def test_silu_mul_embedding_argmax(ops):
    g = torch.Generator().manual_seed(0)
    gu = torch.randn(5, 2 * 5632, generator=g).to(F16)
    ref = om.silu_mul(gu[:, :5632], gu[:, 5632:])
    _close(ops.silu_mul(gu.cuda()), ref, 1e-3, 2e-3)
    table = torch.randn(1000, 256, generator=g).to(F16)
    ids = torch.randint(0, 1000, (17,), generator=g, dtype=torch.int32)
    assert torch.equal(ops.embedding(table.cuda(), ids.cuda()).cpu(), table[ids.long()])
    logits = torch.randn(9, 128256, generator=g)
    logits[3, 77] = logits[3, 90000] = 50.0     # tie -> first index wins
    assert torch.equal(ops.argmax(logits.cuda()).cpu().long(), logits.argmax(-1))
    assert int(ops.argmax(logits.cuda())[3]) == 77


@pytest.mark.parametrize("H,Hkv,D", [(32, 8, 128), (32, 4, 64)])
def test_rope_kv_write(ops, H, Hkv, D):
    cfg = om.ModelConfig(H * D, 1, H, Hkv, D, 64, 64, rope_theta=500000.0, max_pos=4096)
    cos, sin = om.rope_tables(cfg)
    g = torch.Generator().manual_seed(1)
    T, nb = 19, 8
    q = torch.randn(T, H, D, generator=g).to(F16)
    k = torch.randn(T, Hkv, D, generator=g).to(F16)
    v = torch.randn(T, Hkv, D, generator=g).to(F16)
    pos = torch.randint(0, 4096, (T,), generator=g)
    slots = torch.randperm(nb * 64, generator=g)[:T]
    slots[4] = -1
    kc = torch.zeros(nb, 64, Hkv, D, dtype=F16)
    vc = torch.zeros_like(kc)
    q_ref, k_ref = om.apply_rope(q, k, cos, sin, pos)
    om.kv_write(k_ref, v, kc, vc, slots)
    qd, kd, vd = q.cuda(), k.cuda(), v.cuda()
    kcd, vcd = torch.zeros_like(kc).cuda(), torch.zeros_like(vc).cuda()
    ops.rope_kv_write(qd, kd, vd, cos.cuda(), sin.cuda(), pos.cuda(), slots.cuda(), kcd, vcd)
    _close(qd, q_ref, 2e-3, 1e-3)
    _close(kd, k_ref, 2e-3, 1e-3)
    # cache holds exactly what was written back to k / v, at exactly the mapped slots
    exp_k = torch.zeros_like(kc)
    om.kv_write(kd.cpu(), v, exp_k, torch.zeros_like(vc), slots)
    assert torch.equal(kcd.cpu(), exp_k)
    assert torch.equal(vcd.cpu(), vc)


# ---------------------------------------------------------------- paged attention
def _paged_case(B, H, Hkv, D, ctxs, seed, nb_extra=5):
    g = torch.Generator().manual_seed(seed)
    max_blocks = max(-(-c // 64) for c in ctxs) if max(ctxs) > 0 else 1
    need = sum(-(-c // 64) for c in ctxs)
    nb = need + nb_extra
    perm = torch.randperm(nb, generator=g).tolist()      # non-contiguous pages
    kc = torch.randn(nb, 64, Hkv, D, generator=g).to(F16)
    vc = torch.randn(nb, 64, Hkv, D, generator=g).to(F16)
    bt = torch.zeros(B, max_blocks, dtype=torch.int32)
    off = 0
    for b, c in enumerate(ctxs):
        n = -(-c // 64)
        bt[b, :n] = torch.tensor(perm[off:off + n], dtype=torch.int32)
        off += n
    q = torch.randn(B, H, D, generator=g).to(F16)
    return q, kc, vc, bt, torch.tensor(ctxs, dtype=torch.int32)


@pytest.mark.parametrize("H,Hkv,D,ctxs", [
    (32, 8, 128, [1, 63, 64, 65, 127, 129, 700, 4096]),     # Llama-3-8B shapes, ragged
    (32, 4, 64, [37, 1, 500, 0, 2048]),                     # TinyLlama shapes, a ctx=0 row
    (8, 1, 128, [333, 4097]),                               # TP=8 shard of L70 (G=8)
    (16, 1, 128, [130]),                                    # G=16 (both row halves)
    (4, 4, 128, [200, 77]),                                 # MHA (G=1)
])
def test_paged_attention_decode(ops, H, Hkv, D, ctxs):
    B = len(ctxs)
    q, kc, vc, bt, cl = _paged_case(B, H, Hkv, D, ctxs, seed=H + D)
    scale = 1.0 / math.sqrt(D)
    ref = om.paged_attention_decode(q, kc, vc, bt.tolist(), cl.tolist(), scale)
    out = ops.paged_attention_decode(q.cuda(), kc.cuda(), vc.cuda(), bt.cuda(), cl.cuda(), scale)
    # outputs are convex combinations of N(0,1) values: |o| <~ 1; fp16 P/V rounding ~1e-3
    _close(out, ref, 2e-3, 2e-3)


def test_paged_attention_nan_garbage_and_window(ops):
    """Stale cache rows past context_len may hold NaN; they must not leak (p == 0 * NaN)."""
    q, kc, vc, bt, cl = _paged_case(2, 32, 8, 128, [70, 5], seed=9)
    # poison everything beyond the context inside the last pages
    for b, c in enumerate(cl.tolist()):
        blk = int(bt[b, (c - 1) // 64])
        kc[blk, c % 64 if c % 64 else 64:] = float("nan")
        vc[blk, c % 64 if c % 64 else 64:] = float("nan")
    scale = 1.0 / math.sqrt(128)
    ref = om.paged_attention_decode(q, kc, vc, bt.tolist(), cl.tolist(), scale)
    out = ops.paged_attention_decode(q.cuda(), kc.cuda(), vc.cuda(), bt.cuda(), cl.cuda(), scale)
    assert not torch.isnan(out).any()
    _close(out, ref, 2e-3, 2e-3)
    # sliding window + softcap
    q, kc, vc, bt, cl = _paged_case(3, 32, 8, 128, [700, 100, 64], seed=10)
    ref = om.paged_attention_decode(q, kc, vc, bt.tolist(), cl.tolist(), scale, softcap=30.0,
                                    window=256)
    out = ops.paged_attention_decode(q.cuda(), kc.cuda(), vc.cuda(), bt.cuda(), cl.cuda(), scale,
                                     softcap=30.0, window=256)
    _close(out, ref, 2e-3, 2e-3)


def test_paged_attention_full_size_properties(ops):
    """BASELINE config 2 size (b=32, ctx=4096, L3-8B heads): too big for the torch oracle in
    seconds, so check size-independent properties: (1) V == const => output == const;
    (2) permuting pages + block table together leaves the output bit-identical."""
    B, H, Hkv, D, ctx = 32, 32, 8, 128, 4096
    g = torch.Generator(device="cuda").manual_seed(3)
    nb = B * 64 + 64
    kc = torch.randn(nb, 64, Hkv, D, generator=g, device="cuda", dtype=F16)
    vc = torch.randn(nb, 64, Hkv, D, generator=g, device="cuda", dtype=F16)
    q = torch.randn(B, H, D, generator=g, device="cuda", dtype=F16)
    perm = torch.randperm(nb, device="cuda", generator=g)
    bt = perm[:B * 64].reshape(B, 64).to(torch.int32)
    cl = torch.full((B,), ctx, dtype=torch.int32, device="cuda")
    scale = 1 / math.sqrt(D)
    out1 = ops.paged_attention_decode(q, kc, vc, bt, cl, scale)
    # (2) relocate every page
    perm2 = torch.randperm(nb, device="cuda", generator=g)
    kc2, vc2 = torch.empty_like(kc), torch.empty_like(vc)
    kc2[perm2] = kc
    vc2[perm2] = vc
    bt2 = perm2[bt.long()].to(torch.int32)
    out2 = ops.paged_attention_decode(q, kc2, vc2, bt2, cl, scale)
    assert torch.equal(out1, out2)
    # (1) constant V per (head, d)
    const = torch.randn(Hkv, D, generator=g, device="cuda", dtype=F16)
    vc[:] = const
    out3 = ops.paged_attention_decode(q, kc, vc, bt, cl, scale)
    exp = const.repeat_interleave(H // Hkv, 0).unsqueeze(0).expand(B, H, D)
    _close(out3, exp, 2e-3, 2e-3)
    # spot-check 2 sequences against the oracle
    ref = om.paged_attention_decode(q[:2].cpu(), kc.cpu(), vc.cpu(), bt[:2].cpu().tolist(),
                                    [ctx, ctx], scale)
    _close(out3[:2], ref, 2e-3, 2e-3)


# ---------------------------------------------------------------- W4A16 / W8A16
def _wna_case(K, N, M, fmt, bits, seed, group=128):
    rng = np.random.default_rng(seed)
    if fmt == "awq":
        qw, qz, s = oq.random_awq(rng, K, N, group)
        W = oq.dequant_awq(qw, qz, s, group)
    else:
        qw, qz, s = oq.random_gptq(rng, K, N, group, bits, sym=(fmt == "gptq_sym"))
        if fmt == "gptq_sym":
            qz = None
        W = oq.dequant_gptq(qw, qz, s, group, bits)
    x = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float16))
    return qw, qz, s, torch.from_numpy(W), x


# flat GEMV / round-1 tcgen05 decode kernel (M <= 64) / general tcgen05 / stream-K tcgen05 decode kernel (M <= 32)
@pytest.mark.parametrize("variant", [1, 2, 4, 5, 6])
@pytest.mark.parametrize("fmt,bits,K,N,M", [
    ("awq", 4, 4096, 1024, 1), ("awq", 4, 4096, 6144, 32), ("awq", 4, 14336, 4096, 5),
    ("gptq", 4, 2048, 256, 3), ("gptq_sym", 4, 1024, 512, 16), ("gptq", 8, 2048, 512, 7),
    ("awq", 4, 512, 128, 64), ("gptq", 4, 1792, 4096, 33),
])
def test_wna16_gemm(ops, fmt, bits, K, N, M, variant):
    if variant >= 2 and not _tc_available(ops):
        pytest.skip("tcgen05 variant not built yet")
    if variant >= 5 and M > 32:
        pytest.skip("the stream-K decode kernel takes M <= 32")
    if variant == 6 and bits != 4:
        pytest.skip("the mma.sync stream-K flavour is 4-bit only")
    qw, qz, s, W, x = _wna_case(K, N, M, fmt, bits, seed=K + N + M)
    qwd = torch.from_numpy(qw).cuda()
    packed = ops.awq_prepack(qwd) if fmt == "awq" else ops.gptq_prepack(qwd, bits)
    ws = ops.wna16_workspace(max(M, 16), N)
    y = ops.wna16_gemm(x.cuda(), packed, torch.from_numpy(s).cuda(),
                       None if qz is None else torch.from_numpy(qz).cuda(), ws, k=K, n=N,
                       bits=bits, group_size=128,
                       qfmt=ops.QFMT_AWQ if fmt == "awq" else ops.QFMT_GPTQ, variant=variant)
    ref = om.linear(x, W)
    # |y| ~ 1; fp32 accumulation both sides, one fp16 rounding on output
    _close(y, ref, 2e-3, 2e-3)
    assert int(ws[:16384].view(torch.int32).abs().sum()) == 0      # arrival counters back at rest
    # second call reuses the workspace and must agree with the first
    y2 = ops.wna16_gemm(x.cuda(), packed, torch.from_numpy(s).cuda(),
                        None if qz is None else torch.from_numpy(qz).cuda(), ws, k=K, n=N,
                        bits=bits, group_size=128,
                        qfmt=ops.QFMT_AWQ if fmt == "awq" else ops.QFMT_GPTQ, variant=variant)
    assert torch.equal(y2, y)                   # split-K partials are reduced in a fixed order


@pytest.mark.parametrize("fmt,bits,K,N,M,group,epi", [
    ("awq", 4, 4096, 28672, 32, 128, "silu"),      # gate|up of Llama-3-8B: 3 tile segments per CTA, gate / up pairs
    ("awq", 4, 14336, 4096, 17, 128, "none"),      # down: every tile cut 4-5 ways
    ("gptq", 4, 4096, 4096, 32, 4096, "none"),     # one group per channel
    ("gptq_sym", 4, 2048, 1536, 8, 256, "bias"),   # groups of two units, bias in the finaliser / direct store
    ("gptq", 8, 4096, 1024, 32, 128, "none"),      # W8A16
    ("awq", 4, 256, 128, 1, 128, "none"),          # one tile, two units, one CTA
    ("awq", 4, 8192, 7168, 32, 128, "silu"),       # Llama-3-70B gate|up shard at TP = 8
])
@pytest.mark.parametrize("variant", [5, 6])
def test_wna16_streamk_decode(ops, fmt, bits, K, N, M, group, epi, variant):
    """wna16_sk.cu (variant 5 = tcgen05 flavour, 6 = mma.sync register flavour): CTA ranges that cut tiles anywhere, per-CTA partial slots, every epilogue the
    public entry exposes; and the post-MMA scale is at least as close to the fp32 truth as the fp16-rounded
    weights the reference multiplies with."""
    if variant == 6 and bits != 4:
        pytest.skip("the mma.sync stream-K flavour is 4-bit only")
    qw, qz, s, W, x = _wna_case(K, N, M, fmt, bits, seed=K + N + M + group, group=group)
    qwd = torch.from_numpy(qw).cuda()
    packed = ops.awq_prepack(qwd) if fmt == "awq" else ops.gptq_prepack(qwd, bits)
    ws = ops.wna16_workspace(max(M, 16), N)
    sd = torch.from_numpy(s).cuda()
    zd = None if qz is None else torch.from_numpy(qz).cuda()
    qf = ops.QFMT_AWQ if fmt == "awq" else ops.QFMT_GPTQ
    bias = torch.randn(N).to(F16) if epi == "bias" else None
    kw = dict(k=K, n=N, bits=bits, group_size=group, qfmt=qf,
              epilogue=ops.EPI_SILU_MUL if epi == "silu" else ops.EPI_NONE,
              bias=None if bias is None else bias.cuda())
    y = ops.wna16_gemm(x.cuda(), packed, sd, zd, ws, variant=variant, **kw)
    full = om.linear(x, W, bias)
    ref = om.silu_mul(full[:, :N // 2], full[:, N // 2:]) if epi == "silu" else full
    _close(y, ref, 2e-3, 4e-3 if epi == "silu" else 2e-3)
    assert int(ws[:16384].view(torch.int32).abs().sum()) == 0      # arrival counters back at rest
    y2 = ops.wna16_gemm(x.cuda(), packed, sd, zd, ws, variant=variant, **kw)
    assert torch.equal(y2, y)                                      # contributor-ordered sums: deterministic
    if epi == "none":
        # the exact product x @ ((q - z) * s), no fp16 rounding of the weights: the post-MMA scale must be at least
        # as close to it as the round-1 kernel, which multiplies with the reference's fp16-rounded weights
        q = oq.unpack_awq(qw) if fmt == "awq" else oq.unpack_gptq(qw, bits)
        if qz is None:
            z = np.full((K // group, N), 1 << (bits - 1), dtype=np.int32)
        else:
            z = oq.unpack_awq(qz).astype(np.int32) if fmt == "awq" else oq.unpack_gptq_zeros(qz, bits)
        gi = np.arange(K) // group
        wexact = (q.astype(np.float64) - z[gi]) * s[gi].astype(np.float64)
        truth = torch.from_numpy(x.numpy().astype(np.float64) @ wexact)
        y_old = ops.wna16_gemm(x.cuda(), packed, sd, zd, ws, variant=2, **kw)
        e_new = (y.cpu().double() - truth).abs().max().item()
        e_old = (y_old.cpu().double() - truth).abs().max().item()
        assert e_new <= e_old + 5e-4, (e_new, e_old)


def _tc_available(ops):
    from vllm_rs_b200 import _lib
    rng = np.random.default_rng(0)
    qw, qz, s = oq.random_awq(rng, 128, 128, 128)
    x = torch.zeros(16, 128, dtype=F16, device="cuda")
    ws = ops.wna16_workspace(16, 128)
    try:
        ops.wna16_gemm(x, ops.awq_prepack(torch.from_numpy(qw).cuda()),
                       torch.from_numpy(s).cuda(), torch.from_numpy(qz).cuda(), ws, k=128, n=128,
                       variant=2)
        return True
    except _lib.B200Error:
        return False


@pytest.mark.parametrize("bits,K,N,M,group", [(8, 2048, 512, 7, 128), (4, 4096, 1024, 32, 128), (4, 1024, 256, 200, 64)])
def test_gemm_half_q_half_alt_exact_signature(ops, bits, K, N, M, group):
    """The reference's own argument list (gptq.rs:183-195): checkpoint-format qweight, g_idx, no group size, no
    workspace.  Second call hits the cached tile-major copy and must agree bit for bit."""
    from vllm_rs_b200 import _lib
    l = _lib.lib()
    qw, qz, s, W, x = _wna_case(K, N, M, "gptq", bits, seed=K + N + M + bits, group=group)
    qwd, qzd, sd = (torch.from_numpy(t).cuda() for t in (qw, qz, s))
    xd = x.cuda()
    g_idx = (torch.arange(K, dtype=torch.int32) // group + 5).cuda()       # a TP shard's g_idx starts at its first group
    outs = []
    for _ in range(2):
        y = torch.empty(M, N, dtype=F16, device="cuda")
        l.b200_gemm_half_q_half_alt(xd.data_ptr(), qwd.data_ptr(), qzd.data_ptr(), sd.data_ptr(), g_idx.data_ptr(),
                                    y.data_ptr(), M, N, K, bits, torch.cuda.current_stream().cuda_stream)
        _lib.check()
        outs.append(y)
    _close(outs[0], om.linear(x, W), 2e-3, 2e-3)
    assert torch.equal(outs[0], outs[1])
    l.b200_gemm_half_q_half_alt_release()
    _lib.check()


@pytest.mark.parametrize("fmt,K,N,M,group", [("awq", 4096, 6144, 32, 128), ("gptq", 2048, 1024, 5, 128),
                                             ("gptq_sym", 1024, 512, 70, 256), ("awq", 4096, 14336, 1, 128)])
def test_marlin_bf16_entry_points(ops, fmt, K, N, M, group):
    """marlin_awq_4bit_bf16 / marlin_4bit_bf16 (gptq.rs:150-178) with the reference's argument list: bf16 activations,
    bf16 scales, bf16 output.  Checked against the exact product of the bf16 inputs (fp64) within bf16 output
    rounding, and against the reference-style pipeline (weights rounded to bf16 first) within its own error."""
    from vllm_rs_b200 import _lib
    l = _lib.lib()
    BF = torch.bfloat16
    qw, qz, s, W, x = _wna_case(K, N, M, fmt, 4, seed=K + N + M + 1, group=group)
    s_bf = torch.from_numpy(s.astype(np.float32)).to(BF)              # the loader casts the scales to the model dtype
    x_bf = x.to(BF)
    q = oq.unpack_awq(qw) if fmt == "awq" else oq.unpack_gptq(qw, 4)
    if qz is None:
        z = np.full((K // group, N), 8, dtype=np.int32)
    else:
        z = oq.unpack_awq(qz).astype(np.int32) if fmt == "awq" else oq.unpack_gptq_zeros(qz, 4)
    gi = np.arange(K) // group
    d = torch.from_numpy((q.astype(np.int32) - z[gi]).astype(np.float64))
    sg = s_bf.double()[torch.from_numpy(gi)]
    truth = x_bf.double() @ (d * sg)
    w_ref = (d * sg).to(BF)                                           # Marlin bf16: the weight is rounded to bf16
    ref = (x_bf.double() @ w_ref.double()).to(BF)
    qwd = torch.from_numpy(qw).cuda()
    packed = ops.awq_prepack(qwd) if fmt == "awq" else ops.gptq_prepack(qwd, 4)
    ws = ops.wna16_workspace(max(M, 16), N)
    zd = None if qz is None else torch.from_numpy(qz).cuda()
    xd, sd = x_bf.cuda(), s_bf.cuda()
    y = torch.empty(M, N, dtype=BF, device="cuda")
    fn = l.b200_marlin_awq_4bit_bf16 if fmt == "awq" else l.b200_marlin_4bit_bf16
    fn(xd.data_ptr(), packed.data_ptr(), sd.data_ptr(), 0 if zd is None else zd.data_ptr(), 0, y.data_ptr(),
       M, K, N, ws.data_ptr(), group, torch.cuda.current_stream().cuda_stream)
    _lib.check()
    e_new = (y.cpu().double() - truth).abs().max().item()
    e_ref = (ref.double() - truth).abs().max().item()
    scale = truth.abs().max().item()
    assert e_new <= 2.0 ** -8 * scale + 1e-3, (e_new, scale)          # half a bf16 ulp of the largest output, + slack
    assert e_new <= e_ref + 2.0 ** -9 * scale, (e_new, e_ref)         # never worse than the reference-style pipeline
    assert int(ws[:16384].view(torch.int32).abs().sum()) == 0


def test_prepack_bit_exact(ops):
    """Prepacked words decode back to the checkpoint nibbles (layout in include/b200_vllm.h)."""
    rng = np.random.default_rng(5)
    K, N = 256, 256
    q = rng.integers(0, 16, size=(K, N), dtype=np.uint8)
    for fmt in ("awq", "gptq"):
        qw = oq.pack_awq(q) if fmt == "awq" else oq.pack_gptq(q, 4)
        d = torch.from_numpy(qw).cuda()
        p = (ops.awq_prepack(d) if fmt == "awq" else ops.gptq_prepack(d, 4)).cpu().numpy()
        p = p.view(np.uint32).reshape(N // 128, K // 64, 2, 128, 4)
        dec = np.zeros((K, N), dtype=np.uint8)
        for pnib in range(8):
            j = 2 * pnib if pnib < 4 else 2 * (pnib - 4) + 1
            vals = (p >> np.uint32(4 * pnib)) & 0xF            # [nt, kc, h, ch, w]
            nt, kc, h, ch, w = np.meshgrid(*[np.arange(s) for s in p.shape], indexing="ij")
            dec[kc * 64 + h * 32 + w * 8 + j, nt * 128 + ch] = vals
        assert np.array_equal(dec, q), fmt
    q8 = rng.integers(0, 256, size=(K, N), dtype=np.uint8)
    p = ops.gptq_prepack(torch.from_numpy(oq.pack_gptq(q8, 8)).cuda(), 8).cpu().numpy()
    p = p.view(np.uint32).reshape(N // 128, K // 64, 4, 128, 4)
    dec = np.zeros((K, N), dtype=np.uint8)
    for j in range(4):
        vals = (p >> np.uint32(8 * j)) & 0xFF
        nt, kc, h, ch, w = np.meshgrid(*[np.arange(s) for s in p.shape], indexing="ij")
        dec[kc * 64 + h * 16 + w * 4 + j, nt * 128 + ch] = vals
    assert np.array_equal(dec, q8)


def test_wna16_silu_mul_epilogue_and_bias(ops):
    K, inter, M = 1024, 512, 9
    qw, qz, s, W, x = _wna_case(K, 2 * inter, M, "awq", 4, seed=11)
    packed = ops.awq_prepack(torch.from_numpy(qw).cuda())
    ws = ops.wna16_workspace(16, 2 * inter)
    y = ops.wna16_gemm(x.cuda(), packed, torch.from_numpy(s).cuda(), torch.from_numpy(qz).cuda(),
                       ws, k=K, n=2 * inter, epilogue=ops.EPI_SILU_MUL)
    gu = om.linear(x, W)
    _close(y, om.silu_mul(gu[:, :inter], gu[:, inter:]), 2e-3, 4e-3)
    bias = torch.randn(2 * inter).to(F16)
    yb = ops.wna16_gemm(x.cuda(), packed, torch.from_numpy(s).cuda(), torch.from_numpy(qz).cuda(),
                        ws, k=K, n=2 * inter, bias=bias.cuda())
    _close(yb, om.linear(x, W, bias), 3e-3, 2e-3)


def test_dense_gemm(ops):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(32, 1024, generator=g).to(F16)
    w = (torch.randn(4096, 1024, generator=g) * 0.03).to(F16)
    ref = x.float() @ w.float().t()
    _close(ops.dense_gemm(x.cuda(), w.cuda(), out_f32=True), ref, 2e-3, 1e-3)
    _close(ops.dense_gemm(x.cuda(), w.cuda()), ref.to(F16), 2e-3, 2e-3)


def test_errors_are_latched(ops):
    from vllm_rs_b200 import _lib
    x = torch.zeros(2, 100, dtype=F16, device="cuda")
    with pytest.raises(_lib.B200Error):
        ops.rmsnorm(x, torch.zeros(100, dtype=F16, device="cuda"), 1e-5)   # hidden % 8 != 0
    ops.rmsnorm(torch.zeros(2, 128, dtype=F16, device="cuda"),
                torch.ones(128, dtype=F16, device="cuda"), 1e-5)           # state cleared


# ---------------------------------------------------------------- prefill attention
@pytest.mark.parametrize("H,Hkv,D,cases", [
    (32, 8, 128, [(37, 0), (64, 0), (65, 0), (130, 64), (1, 0), (200, 128)]),   # (new tokens, cached)
    (32, 4, 64, [(300, 0), (5, 192)]),                                           # TinyLlama shapes
    (8, 8, 128, [(100, 0), (33, 64)]),                                           # MHA
    (4, 2, 64, [(129, 0)]),
])
def test_attention_prefill_paged(ops, H, Hkv, D, cases):
    """Varlen causal prefill over the paged cache vs the oracle, fresh prompts and chunks that
    attend to cached context (runner.rs:1068-1082).  Stale rows past the context are NaN."""
    g = torch.Generator().manual_seed(H * D + len(cases))
    B = len(cases)
    ctxs = [n + c for n, c in cases]
    nblk = [-(-c // 64) for c in ctxs]
    nb = sum(nblk) + 3
    perm = torch.randperm(nb, generator=g).tolist()
    kc = torch.randn(nb, 64, Hkv, D, generator=g).to(F16)
    vc = torch.randn(nb, 64, Hkv, D, generator=g).to(F16)
    bt = torch.zeros(B, max(nblk), dtype=torch.int32)
    off = 0
    for b, n in enumerate(nblk):
        bt[b, :n] = torch.tensor(perm[off:off + n], dtype=torch.int32)
        off += n
    ref = None
    cu = [0]
    for n, _ in cases:
        cu.append(cu[-1] + n)
    T = cu[-1]
    q = torch.randn(T, H, D, generator=g).to(F16)
    scale = 1.0 / math.sqrt(D)
    dummy = torch.zeros(T, Hkv, D, dtype=F16)           # keys / values come from the paged cache
    ref = om.attention_prefill(q, dummy, dummy, cu, kc, vc, bt.tolist(), ctxs, scale)
    for b, n in enumerate(nblk):                       # poison the tail of every last page
        last = ctxs[b] % 64
        if last:
            kc[bt[b, n - 1], last:] = float("nan")
            vc[bt[b, n - 1], last:] = float("nan")
    out = ops.attention_prefill(q.cuda(), kc.cuda(), vc.cuda(),
                                torch.tensor(cu, dtype=torch.int32).cuda(),
                                torch.tensor(ctxs, dtype=torch.int32).cuda(), bt.cuda(),
                                max(n for n, _ in cases), scale)
    assert not torch.isnan(out).any()
    _close(out, ref, 2e-3, 2e-3)


@pytest.mark.parametrize("M", [100, 256, 300, 700])
def test_wna16_gemm_prefill_sizes(ops, M):
    """tcgen05 variant at prefill token counts (NT = 128 / 256, several token tiles)."""
    K, N = 1024, 512
    qw, qz, s, W, x = _wna_case(K, N, M, "awq", 4, seed=M)
    packed = ops.awq_prepack(torch.from_numpy(qw).cuda())
    ws = ops.wna16_workspace(M, N)
    y = ops.wna16_gemm(x.cuda(), packed, torch.from_numpy(s).cuda(), torch.from_numpy(qz).cuda(), ws,
                       k=K, n=N, variant=2)
    _close(y, om.linear(x, W), 2e-3, 2e-3)


# ---------------------------------------------------------------- cache::swap_blocks (data path)
@pytest.mark.parametrize("dtype", [torch.float16, torch.uint8])
def test_swap_blocks_gpu_cpu_round_trip(ops, dtype):
    """swap_kvcache (runner.rs:1626-1645): GPU -> pinned CPU swap space -> other GPU blocks, whole blocks by
    a {src: dst} mapping in one launch per cache tensor; byte-exact, untouched blocks stay untouched."""
    g = torch.Generator().manual_seed(5)
    nb_gpu, nb_cpu, Hkv, D = 24, 16, 8, 128
    gpu = torch.randint(0, 255, (nb_gpu, 64, Hkv, D), generator=g, dtype=torch.uint8)
    if dtype == torch.float16:
        gpu = torch.randn(nb_gpu, 64, Hkv, D, generator=g).to(dtype)
    ref_gpu = gpu.clone()
    ref_cpu = torch.zeros(nb_cpu, 64, Hkv, D, dtype=dtype)
    cpu = torch.zeros(nb_cpu, 64, Hkv, D, dtype=dtype).pin_memory()
    dgpu = gpu.cuda()
    out_map = {3: 7, 20: 0, 11: 15, 0: 1}                      # swap out: GPU block -> CPU block
    keep = ops.swap_blocks(dgpu, cpu, out_map)
    torch.cuda.synchronize()
    for s, d in out_map.items():
        ref_cpu[d] = ref_gpu[s]
    assert torch.equal(cpu, ref_cpu)
    in_map = {7: 22, 0: 5, 15: 3, 1: 9}                        # swap in: CPU block -> (different) GPU block
    keep = ops.swap_blocks(cpu, dgpu, in_map)
    torch.cuda.synchronize()
    for s, d in in_map.items():
        ref_gpu[d] = ref_cpu[s]
    assert torch.equal(dgpu.cpu(), ref_gpu)
    # device to device (the PD transfer path uses the same call between two GPU caches)
    dst = torch.zeros_like(dgpu)
    keep = ops.swap_blocks(dgpu, dst, [(i, nb_gpu - 1 - i) for i in range(nb_gpu)])
    torch.cuda.synchronize()
    assert torch.equal(dst.cpu(), ref_gpu.flip(0))
    del keep


# ---------------------------------------------------------------- bf16 flavours (the reference's default dtype)
BF16 = torch.bfloat16
# one bf16 ulp is 8 fp16 ulps: the fp16 bars above, times 8


def test_bf16_glue(ops):
    g = torch.Generator().manual_seed(3)
    with om.model_dtype(BF16):
        T, hidden = 9, 4096
        x = torch.randn(T, hidden, generator=g).to(BF16)
        w = (1 + 0.1 * torch.randn(hidden, generator=g)).to(BF16)
        out = ops.rmsnorm(x.cuda(), w.cuda(), 1e-5)
        assert out.dtype == BF16
        _close(out, om.rmsnorm(x, w, 1e-5), 1.6e-2, 8e-3)
        res = torch.randn(T, hidden, generator=g).to(BF16)
        res_d = res.cuda().clone()
        out2 = ops.add_rmsnorm(res_d, x.cuda(), w.cuda(), 1e-5)
        new_res = (res.float() + x.float()).to(BF16)              # the add rounds once, to bf16
        assert torch.equal(res_d.cpu(), new_res)
        _close(out2, om.rmsnorm(new_res, w, 1e-5), 1.6e-2, 8e-3)
        gu = torch.randn(5, 2 * 5632, generator=g).to(BF16)
        _close(ops.silu_mul(gu.cuda()), om.silu_mul(gu[:, :5632], gu[:, 5632:]), 8e-3, 1.6e-2)
        table = torch.randn(1000, 256, generator=g).to(BF16)
        ids = torch.randint(0, 1000, (17,), generator=g, dtype=torch.int32)
        assert torch.equal(ops.embedding(table.cuda(), ids.cuda()).cpu(), table[ids.long()])


@pytest.mark.parametrize("H,Hkv,D,f32_tables", [(32, 8, 128, False), (32, 4, 64, True)])
def test_bf16_rope_kv_write(ops, H, Hkv, D, f32_tables):
    cfg = om.ModelConfig(H * D, 1, H, Hkv, D, 64, 64, rope_theta=500000.0, max_pos=4096)
    cos, sin = om.rope_tables(cfg, dtype=torch.float32 if f32_tables else BF16)
    g = torch.Generator().manual_seed(1)
    T, nb = 19, 8
    q = torch.randn(T, H, D, generator=g).to(BF16)
    k = torch.randn(T, Hkv, D, generator=g).to(BF16)
    v = torch.randn(T, Hkv, D, generator=g).to(BF16)
    pos = torch.randint(0, 4096, (T,), generator=g)
    slots = torch.randperm(nb * 64, generator=g)[:T]
    slots[4] = -1
    kc = torch.zeros(nb, 64, Hkv, D, dtype=BF16)
    vc = torch.zeros_like(kc)
    with om.model_dtype(BF16):
        q_ref, k_ref = om.apply_rope(q, k, cos, sin, pos)
    om.kv_write(k_ref, v, kc, vc, slots)
    qd, kd, vd = q.cuda(), k.cuda(), v.cuda()
    kcd, vcd = torch.zeros_like(kc).cuda(), torch.zeros_like(vc).cuda()
    ops.rope_kv_write(qd, kd, vd, cos.cuda(), sin.cuda(), pos.cuda(), slots.cuda(), kcd, vcd)
    _close(qd, q_ref, 1.6e-2, 8e-3)
    _close(kd, k_ref, 1.6e-2, 8e-3)
    exp_k = torch.zeros_like(kc)
    om.kv_write(kd.cpu(), v, exp_k, torch.zeros_like(vc), slots)
    assert torch.equal(kcd.cpu(), exp_k)
    assert torch.equal(vcd.cpu(), vc)


@pytest.mark.parametrize("H,Hkv,D,ctxs", [
    (32, 8, 128, [1, 63, 64, 65, 127, 129, 700, 4096]),
    (32, 4, 64, [37, 1, 500, 0, 2048]),
    (16, 1, 128, [130]),
])
def test_bf16_paged_attention_decode(ops, H, Hkv, D, ctxs):
    B = len(ctxs)
    q, kc, vc, bt, cl = (t.to(BF16) if t.dtype == F16 else t for t in _paged_case(B, H, Hkv, D, ctxs, seed=H + D))
    scale = 1.0 / math.sqrt(D)
    with om.model_dtype(BF16):
        ref = om.paged_attention_decode(q, kc, vc, bt.tolist(), cl.tolist(), scale)
    out = ops.paged_attention_decode(q.cuda(), kc.cuda(), vc.cuda(), bt.cuda(), cl.cuda(), scale)
    assert out.dtype == BF16
    _close(out, ref, 1.6e-2, 1.6e-2)
    # the fp16 kernel on the same (bf16-representable) values agrees to the bf16 rounding of the output
    out16 = ops.paged_attention_decode(q.to(F16).cuda(), kc.to(F16).cuda(), vc.to(F16).cuda(), bt.cuda(), cl.cuda(),
                                       scale)
    _close(out, out16, 1.6e-2, 1.6e-2)


@pytest.mark.parametrize("H,Hkv,D,cases,mode", [
    (32, 8, 128, [(37, 0), (64, 0), (65, 0), (130, 64), (1, 0), (200, 128)], -1),     # tcgen05 kernel
    (32, 8, 128, [(70, 0), (5, 192)], 0),                                             # mma.sync kernel
    (32, 4, 64, [(300, 0), (5, 192)], -1),
])
def test_bf16_attention_prefill_paged(ops, H, Hkv, D, cases, mode):
    g = torch.Generator().manual_seed(H * D + len(cases))
    B = len(cases)
    ctxs = [n + c for n, c in cases]
    nblk = [-(-c // 64) for c in ctxs]
    nb = sum(nblk) + 3
    perm = torch.randperm(nb, generator=g).tolist()
    kc = torch.randn(nb, 64, Hkv, D, generator=g).to(BF16)
    vc = torch.randn(nb, 64, Hkv, D, generator=g).to(BF16)
    bt = torch.zeros(B, max(nblk), dtype=torch.int32)
    off = 0
    for b, n in enumerate(nblk):
        bt[b, :n] = torch.tensor(perm[off:off + n], dtype=torch.int32)
        off += n
    cu = [0]
    for n, _ in cases:
        cu.append(cu[-1] + n)
    T = cu[-1]
    q = torch.randn(T, H, D, generator=g).to(BF16)
    scale = 1.0 / math.sqrt(D)
    dummy = torch.zeros(T, Hkv, D, dtype=BF16)
    with om.model_dtype(BF16):
        ref = om.attention_prefill(q, dummy, dummy, cu, kc, vc, bt.tolist(), ctxs, scale)
    for b, n in enumerate(nblk):
        last = ctxs[b] % 64
        if last:
            kc[bt[b, n - 1], last:] = float("nan")
            vc[bt[b, n - 1], last:] = float("nan")
    out = ops.attention_prefill(q.cuda(), kc.cuda(), vc.cuda(), torch.tensor(cu, dtype=torch.int32).cuda(),
                                torch.tensor(ctxs, dtype=torch.int32).cuda(), bt.cuda(),
                                max(n for n, _ in cases), scale, mode=mode)
    assert out.dtype == BF16 and not torch.isnan(out).any()
    _close(out, ref, 1.6e-2, 1.6e-2)


# ---------------------------------------------------------------- the reference's default-build cache layout (L1)
# kvcache_allocator.rs:150-173: K [nb, Hkv, D/8, 64, 8], V [nb, Hkv, D, 64].  Same oracle, caches permuted.
@pytest.mark.parametrize("dt", [F16, BF16])
@pytest.mark.parametrize("H,Hkv,D", [(32, 8, 128), (32, 4, 64)])
def test_l1_layout_cache_writes(ops, H, Hkv, D, dt):
    cfg = om.ModelConfig(H * D, 1, H, Hkv, D, 64, 64, rope_theta=500000.0, max_pos=4096)
    cos, sin = om.rope_tables(cfg, dtype=dt)
    g = torch.Generator().manual_seed(5)
    T, nb = 23, 6
    q = torch.randn(T, H, D, generator=g).to(dt)
    k = torch.randn(T, Hkv, D, generator=g).to(dt)
    v = torch.randn(T, Hkv, D, generator=g).to(dt)
    pos = torch.randint(0, 4096, (T,), generator=g)
    slots = torch.randperm(nb * 64, generator=g)[:T]
    slots[7] = -1
    zeros = torch.zeros(nb, 64, Hkv, D, dtype=dt)
    # reshape_and_cache only
    kc1, vc1 = (t.cuda() for t in om.to_l1_layout(zeros, zeros))
    ops.kv_write(k.cuda(), v.cuda(), kc1, vc1, slots.cuda())
    exp_k, exp_v = zeros.clone(), zeros.clone()
    om.kv_write(k, v, exp_k, exp_v, slots)
    got_k, got_v = om.from_l1_layout(kc1.cpu(), vc1.cpu())
    assert torch.equal(got_k, exp_k) and torch.equal(got_v, exp_v)
    # RoPE + cache write in one launch: the same rotated values as the flash-layout call, at the L1 positions
    qd, kd = q.cuda(), k.cuda()
    kc1, vc1 = (t.cuda() for t in om.to_l1_layout(zeros, zeros))
    ops.rope_kv_write(qd, kd, v.cuda(), cos.cuda(), sin.cuda(), pos.cuda(), slots.cuda(), kc1, vc1)
    q2, k2 = q.cuda(), k.cuda()
    kc2, vc2 = zeros.clone().cuda(), zeros.clone().cuda()
    ops.rope_kv_write(q2, k2, v.cuda(), cos.cuda(), sin.cuda(), pos.cuda(), slots.cuda(), kc2, vc2)
    assert torch.equal(qd, q2) and torch.equal(kd, k2)
    got_k, got_v = om.from_l1_layout(kc1.cpu(), vc1.cpu())
    assert torch.equal(got_k, kc2.cpu()) and torch.equal(got_v, vc2.cpu())


@pytest.mark.parametrize("dt", [F16, BF16])
@pytest.mark.parametrize("H,Hkv,D,ctxs", [
    (32, 8, 128, [1, 63, 64, 65, 127, 129, 700, 4096]),
    (32, 4, 64, [37, 1, 500, 0, 2048]),
    (16, 1, 128, [130]),
])
def test_l1_layout_paged_attention_decode(ops, H, Hkv, D, ctxs, dt):
    B = len(ctxs)
    q, kc, vc, bt, cl = (t.to(dt) if t.dtype == F16 else t for t in _paged_case(B, H, Hkv, D, ctxs, seed=H + D + 1))
    scale = 1.0 / math.sqrt(D)
    with om.model_dtype(dt):
        ref = om.paged_attention_decode(q, kc, vc, bt.tolist(), cl.tolist(), scale)
    for b, c in enumerate(ctxs):                       # stale rows past the context must not leak
        if c % 64:
            kc[bt[b, (c - 1) // 64], c % 64:] = float("nan")
            vc[bt[b, (c - 1) // 64], c % 64:] = float("nan")
    k1, v1 = om.to_l1_layout(kc, vc)
    out = ops.paged_attention_decode(q.cuda(), k1.cuda(), v1.cuda(), bt.cuda(), cl.cuda(), scale)
    assert not torch.isnan(out).any()
    tol = 2e-3 if dt == F16 else 1.6e-2
    _close(out, ref, tol, tol)
    # and it is the flash-layout kernel's answer to the last bit of the dtype: same MMAs on the same fragments
    out2 = ops.paged_attention_decode(q.cuda(), kc.cuda(), vc.cuda(), bt.cuda(), cl.cuda(), scale)
    assert torch.equal(out, out2)


@pytest.mark.parametrize("dt,mode", [(F16, -1), (BF16, -1), (F16, 0)])
@pytest.mark.parametrize("H,Hkv,D,cases", [
    (32, 8, 128, [(37, 0), (64, 0), (65, 0), (130, 64), (1, 0), (200, 128)]),
    (32, 4, 64, [(300, 0), (5, 192)]),
])
def test_l1_layout_attention_prefill(ops, H, Hkv, D, cases, dt, mode):
    g = torch.Generator().manual_seed(H * D + len(cases) + 1)
    B = len(cases)
    ctxs = [n + c for n, c in cases]
    nblk = [-(-c // 64) for c in ctxs]
    nb = sum(nblk) + 3
    perm = torch.randperm(nb, generator=g).tolist()
    kc = torch.randn(nb, 64, Hkv, D, generator=g).to(dt)
    vc = torch.randn(nb, 64, Hkv, D, generator=g).to(dt)
    bt = torch.zeros(B, max(nblk), dtype=torch.int32)
    off = 0
    for b, n in enumerate(nblk):
        bt[b, :n] = torch.tensor(perm[off:off + n], dtype=torch.int32)
        off += n
    cu = [0]
    for n, _ in cases:
        cu.append(cu[-1] + n)
    T = cu[-1]
    q = torch.randn(T, H, D, generator=g).to(dt)
    scale = 1.0 / math.sqrt(D)
    dummy = torch.zeros(T, Hkv, D, dtype=dt)
    with om.model_dtype(dt):
        ref = om.attention_prefill(q, dummy, dummy, cu, kc, vc, bt.tolist(), ctxs, scale)
    for b, n in enumerate(nblk):
        last = ctxs[b] % 64
        if last:
            kc[bt[b, n - 1], last:] = float("nan")
            vc[bt[b, n - 1], last:] = float("nan")
    k1, v1 = om.to_l1_layout(kc, vc)
    out = ops.attention_prefill(q.cuda(), k1.cuda(), v1.cuda(), torch.tensor(cu, dtype=torch.int32).cuda(),
                                torch.tensor(ctxs, dtype=torch.int32).cuda(), bt.cuda(),
                                max(n for n, _ in cases), scale, mode=mode)
    assert not torch.isnan(out).any()
    tol = 2e-3 if dt == F16 else 1.6e-2
    _close(out, ref, tol, tol)
