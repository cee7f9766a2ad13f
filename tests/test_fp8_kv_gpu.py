"""FP8 KV cache (`--fp8-kvcache`; kvcache_allocator.rs:188-193,776 makes the cache U8
in the same [num_blocks, 64, Hkv, D] shape and attention.rs:615 hands only the flag to attention-rs).

PARITY UNPINNED: the reference's e4m3 arithmetic (scales, rounding) lives in attention-rs and is not
source-visible.  The format is therefore this library's definition, restated in oracle/model.py
(to_e4m3 / from_e4m3: scale 1, round to nearest even, saturate at +-448).  Bars:
  * cache BYTES written by the CUDA kernels == the oracle's bytes (bit-exact) wherever the fp16 value
    being stored is the same;
  * attention over a given e4m3 cache within 2e-3 + 2e-3|x| of the oracle reading the SAME bytes
    (the quantisation error itself is common to both and cancels);
  * the whole decode / prefill step within the usual noise-floor bar of the fp8-cache oracle.
"""
import math

import numpy as np
import pytest
import torch

from oracle import model as om
from oracle import sched as osched

pytestmark = pytest.mark.gpu
F16 = torch.float16


@pytest.fixture(scope="module")
def ops(lib):
    from vllm_rs_b200 import ops as o
    return o


def _close(a, b, atol, rtol):
    a, b = a.float().cpu(), b.float().cpu()
    err = (a - b).abs()
    bad = err > atol + rtol * b.abs()
    assert not bad.any(), f"max err {err.max().item():.3e}, {int(bad.sum())} outside tolerance"


def _paged_case(B, H, Hkv, ctxs, seed, scale_kv=1.0):
    D = 128
    g = torch.Generator().manual_seed(seed)
    max_blocks = max(-(-c // 64) for c in ctxs) if max(ctxs) > 0 else 1
    nb = sum(-(-c // 64) for c in ctxs) + 5
    perm = torch.randperm(nb, generator=g).tolist()
    kc = om.to_e4m3(scale_kv * torch.randn(nb, 64, Hkv, D, generator=g))
    vc = om.to_e4m3(scale_kv * torch.randn(nb, 64, Hkv, D, generator=g))
    bt = torch.zeros(B, max_blocks, dtype=torch.int32)
    off = 0
    for b, c in enumerate(ctxs):
        n = -(-c // 64)
        bt[b, :n] = torch.tensor(perm[off:off + n], dtype=torch.int32)
        off += n
    q = torch.randn(B, H, D, generator=g).to(F16)
    return q, kc, vc, bt, torch.tensor(ctxs, dtype=torch.int32)


@pytest.mark.parametrize("H,Hkv,ctxs", [
    (32, 8, [1, 63, 64, 65, 127, 129, 700, 4096]),          # Llama-3-8B shapes, ragged
    (8, 1, [333, 4097, 0]),                                 # TP=8 shard (G=8), a ctx=0 row
    (16, 1, [130]),                                         # G=16 (both row halves)
    (4, 4, [200, 77]),                                      # MHA (G=1)
])
def test_paged_attention_decode_fp8(ops, H, Hkv, ctxs):
    q, kc, vc, bt, cl = _paged_case(len(ctxs), H, Hkv, ctxs, seed=H + 3)
    scale = 1.0 / math.sqrt(128)
    ref = om.paged_attention_decode(q, kc, vc, bt.tolist(), cl.tolist(), scale)
    out = ops.paged_attention_decode(q.cuda(), kc.cuda(), vc.cuda(), bt.cuda(), cl.cuda(), scale)
    _close(out, ref, 2e-3, 2e-3)


def test_paged_attention_fp8_garbage_window_and_columns(ops):
    """(1) NaN bytes (0x7F / 0xFF) past context_len must not leak; (2) softcap + sliding window;
    (3) V[t, h, d] = the e4m3 value whose byte code is d makes every output column equal a value that
    identifies it: catches any mistake in the even / odd column un-permutation of the e4m3 PV path."""
    scale = 1.0 / math.sqrt(128)
    q, kc, vc, bt, cl = _paged_case(2, 32, 8, [70, 5], seed=9)
    for b, c in enumerate(cl.tolist()):
        blk = int(bt[b, (c - 1) // 64])
        kc[blk, c % 64:] = 0x7F
        vc[blk, c % 64:] = 0xFF
    ref = om.paged_attention_decode(q, kc.clone(), vc.clone(), bt.tolist(), cl.tolist(), scale)
    out = ops.paged_attention_decode(q.cuda(), kc.cuda(), vc.cuda(), bt.cuda(), cl.cuda(), scale)
    assert not torch.isnan(out).any() and not torch.isnan(ref).any()
    _close(out, ref, 2e-3, 2e-3)
    q, kc, vc, bt, cl = _paged_case(3, 32, 8, [700, 100, 64], seed=10)
    ref = om.paged_attention_decode(q, kc, vc, bt.tolist(), cl.tolist(), scale, softcap=30.0, window=256)
    out = ops.paged_attention_decode(q.cuda(), kc.cuda(), vc.cuda(), bt.cuda(), cl.cuda(), scale,
                                     softcap=30.0, window=256)
    _close(out, ref, 2e-3, 2e-3)
    q, kc, vc, bt, cl = _paged_case(2, 32, 8, [300, 64], seed=11)
    col = torch.arange(128, dtype=torch.uint8)          # column d holds the e4m3 value with byte code d ...
    col[127] = 0xFE                                     # ... (0x7F is NaN: use -448 there): 128 distinct values
    vc[:] = col
    out = ops.paged_attention_decode(q.cuda(), kc.cuda(), vc.cuda(), bt.cuda(), cl.cuda(), scale)
    exp = om.from_e4m3(col).expand(2, 32, 128)
    _close(out, exp, 2e-4, 2e-3)


def test_paged_attention_fp8_full_size_properties(ops):
    """BASELINE config 2 size with an e4m3 cache: page relocation leaves the output bit-identical, and
    two sequences are spot-checked against the oracle."""
    B, H, Hkv, D, ctx = 32, 32, 8, 128, 4096
    g = torch.Generator(device="cuda").manual_seed(3)
    nb = B * 64 + 64
    kc = om.to_e4m3(torch.randn(nb, 64, Hkv, D, generator=g, device="cuda"))
    vc = om.to_e4m3(torch.randn(nb, 64, Hkv, D, generator=g, device="cuda"))
    q = torch.randn(B, H, D, generator=g, device="cuda", dtype=F16)
    perm = torch.randperm(nb, device="cuda", generator=g)
    bt = perm[:B * 64].reshape(B, 64).to(torch.int32)
    cl = torch.full((B,), ctx, dtype=torch.int32, device="cuda")
    scale = 1 / math.sqrt(D)
    out1 = ops.paged_attention_decode(q, kc, vc, bt, cl, scale)
    perm2 = torch.randperm(nb, device="cuda", generator=g)
    kc2, vc2 = torch.empty_like(kc), torch.empty_like(vc)
    kc2[perm2] = kc
    vc2[perm2] = vc
    bt2 = perm2[bt.long()].to(torch.int32)
    out2 = ops.paged_attention_decode(q, kc2, vc2, bt2, cl, scale)
    assert torch.equal(out1, out2)
    ref = om.paged_attention_decode(q[:2].cpu(), kc.cpu(), vc.cpu(), bt[:2].cpu().tolist(), [ctx, ctx], scale)
    _close(out1[:2], ref, 2e-3, 2e-3)


def test_rope_kv_write_fp8_bytes(ops):
    """Cache bytes == oracle bytes of the fp16 rows the kernel wrote back (bit-exact), at the mapped slots only."""
    H, Hkv, D = 32, 8, 128
    cfg = om.ModelConfig(H * D, 1, H, Hkv, D, 64, 64, rope_theta=500000.0, max_pos=4096)
    cos, sin = om.rope_tables(cfg)
    g = torch.Generator().manual_seed(1)
    T, nb = 19, 8
    q = torch.randn(T, H, D, generator=g).to(F16)
    k = (3 * torch.randn(T, Hkv, D, generator=g)).to(F16)
    v = (3 * torch.randn(T, Hkv, D, generator=g)).to(F16)
    k[0, 0, :4] = torch.tensor([500.0, -600.0, 1e-4, 447.9], dtype=F16)          # saturation / underflow
    v[1, 0, :4] = torch.tensor([17.0, 19.0, -17.0, 0.0], dtype=F16)              # ties
    pos = torch.randint(0, 4096, (T,), generator=g)
    pos[0] = 0                                                                   # identity rotation for the row above
    slots = torch.randperm(nb * 64, generator=g)[:T]
    slots[4] = -1
    qd, kd, vd = q.cuda(), k.cuda(), v.cuda()
    kcd = torch.zeros(nb, 64, Hkv, D, dtype=torch.uint8, device="cuda")
    vcd = torch.zeros_like(kcd)
    ops.rope_kv_write(qd, kd, vd, cos.cuda(), sin.cuda(), pos.cuda(), slots.cuda(), kcd, vcd)
    exp_k = torch.zeros(nb, 64, Hkv, D, dtype=torch.uint8)
    exp_v = torch.zeros_like(exp_k)
    om.kv_write(kd.cpu(), v, exp_k, exp_v, slots)
    assert torch.equal(kcd.cpu(), exp_k)
    assert torch.equal(vcd.cpu(), exp_v)


def _run_pair(kv_fp8, seed=31):
    from oracle.toy import build_pair, toy_cfg_d128
    cfg = toy_cfg_d128()
    NB, MAXB, MAXBLK = 48, 4, 8
    oracle, runner = build_pair(cfg, "awq", seed=seed, num_blocks=NB, max_batch=MAXB, max_blocks=MAXBLK,
                                max_prefill_tokens=512, kv_fp8=kv_fp8)
    return cfg, NB, MAXB, oracle, runner


@pytest.mark.parametrize("plan", ["fused", "classic"])
def test_decode_and_prefill_steps_fp8_cache(lib, plan):
    """Whole steps with an e4m3 cache: a prefill step (keys / values read back from the cache through the
    fp16 scratch pages), a chunk with cached context, then decode steps through both launch plans (the
    fused qkv finaliser and the RoPE + KV write kernel store the bytes).  Cache bytes must equal the
    oracle's except where the fp16 K/V value itself differs by GEMM round-off (<= 0.5 % of the bytes, one
    code apart); logits within the noise-floor bar of the fp8-cache oracle.  The oracle's cache is copied
    to the GPU after every step (as tests/test_runner_gpu.py does), so each step is checked on its own."""
    from oracle.toy import sync_cache_to_gpu
    from vllm_rs_b200 import engine
    cfg, NB, MAXB, oracle, runner = _run_pair(True)
    if plan == "classic":       # a no-op all-reduce selects the classic (TP rank) launch plan on one GPU
        runner.set_allreduce(lambda buf, n, st: None)
    with om.exact_mode():
        exact = om.LlamaOracle(cfg, oracle.w, NB, kv_fp8=True)
    rng = np.random.default_rng(5)
    CH = 128
    osch = osched.Scheduler(NB, 64, MAXB, chunk_size=CH)
    csch = engine.Scheduler(NB, 64, MAXB, 0, CH, ())
    for L in (37, 64, 200):
        toks = rng.integers(0, cfg.vocab, size=L).tolist()
        osch.add(osched.Sequence(toks, 64, max_tokens=8))
        csch.add(toks, 8)
    steps = 0
    checked_prefill = checked_decode = 0
    while (osch.running or osch.waiting) and steps < 14:
        steps += 1
        ids, pf = osch.schedule()
        cids, cpf = csch.schedule()
        assert (ids, pf) == (cids, cpf)
        seqs = [osch.running[i] for i in ids]
        if pf:
            meta = osched.prepare_prefill(seqs, 64, CH)
            meta["all_block_tables"] = osched.prepare_block_tables(seqs)[0]
            logits = oracle.prefill(meta)
            with om.exact_mode():
                truth = exact.prefill(meta)
            m2 = dict(meta)
            if m2["block_tables"] is None:
                m2["block_tables"] = meta["all_block_tables"]
            m2 = {k: (np.asarray(v) if isinstance(v, list) else v) for k, v in m2.items()}
            got, _ = runner.prefill_device(m2)
            checked_prefill += 1
            pos, fin = osch.filter_prefill_finished(ids)
            csch.filter_prefill_finished(cids)
            toks = logits.argmax(-1)[pos].tolist() if pos else []
            for s in (osch, csch):
                s.postprocess(fin, toks)
        else:
            mo = osched.prepare_decode(seqs, 64)
            mc = csch.prepare_decode(cids, max_blocks_cap=8)
            logits = oracle.decode(mo)
            with om.exact_mode():
                truth = exact.decode(mo)
            B = len(ids)
            bt = torch.zeros(B, 8, dtype=torch.int32)
            bt[:, :mc["max_blocks"]] = torch.from_numpy(mc["block_tables"].astype(np.int32))
            got, _ = runner.decode_device(
                torch.from_numpy(mc["input_ids"].astype(np.int32)).cuda(),
                torch.from_numpy(mc["positions"]).cuda(), torch.from_numpy(mc["slot_mapping"]).cuda(),
                bt.cuda(), torch.from_numpy(mc["context_lens"].astype(np.int32)).cuda())
            checked_decode += 1
            toks = logits.argmax(-1).tolist()
            for s in (osch, csch):
                s.postprocess(ids, toks)
        torch.cuda.synchronize()
        noise = (logits - truth).abs().max().item()
        err = (got.cpu() - truth).abs().max().item()
        assert err <= 3 * noise + 2e-3, f"step {steps} ({'prefill' if pf else 'decode'}): {err} vs noise {noise}"
        for l in range(cfg.num_layers):
            for mine, ref in ((runner.k_cache[l], oracle.k_cache[l]), (runner.v_cache[l], oracle.v_cache[l])):
                a, b = mine.cpu(), ref
                diff = a != b
                assert diff.float().mean().item() <= (1e-2 if l == 0 else 3e-2), (l, diff.float().mean().item())
                if diff.any():
                    # the fp16 K/V rows differ by GEMM round-off (tests/test_runner_gpu.py bounds it by 4e-3),
                    # which moves a value across at most one e4m3 rounding boundary (step <= |x| / 8)
                    fa, fb = om.from_e4m3(a[diff]), om.from_e4m3(b[diff])
                    assert ((fa - fb).abs() <= 0.13 * fb.abs() + 6e-3).all()
        sync_cache_to_gpu(oracle, runner)
        for s in (osch, csch):
            s.clear_finished()
    assert checked_prefill >= 2 and checked_decode >= 4
