"""GPU parity of the whole decode step (C++ launch plan + every kernel) vs the CPU oracle,
driven exactly as the reference's engine loop drives it: scheduler -> prepare_decode ->
ModelRunner::run -> postprocess.  Block tables / slots bit-exact, logits within the
north-star tolerance (1e-3 in fp16 terms -> atol below scaled to the logit range).
"""
import numpy as np
import pytest
import torch

from oracle import model as om
from oracle import quant as oq
from oracle import sched as osched

pytestmark = pytest.mark.gpu
F16 = torch.float16


from oracle.toy import build_pair as _build, sync_cache_to_gpu as _sync_cache_to_gpu, toy_cfg as _toy_cfg


# Per-rank shapes of the tensor-parallel configs in BASELINE.json, run here as whole (1-layer) models
# on one GPU so that every kernel is checked at the tile counts a TP rank really launches:
#   Llama-3-8B TP=8: 4 q heads, 1 kv head, inter 1792 (qkv N = 768, o K = 512, down K = 1792)
#   Llama-3-8B TP=4: 8 q heads, 2 kv heads, inter 3584
#   Llama-3-70B TP=8 (configs[4]): hidden 8192, 8 q heads, 1 kv head, inter 3584
RANK_SHAPES = {
    "toy": None,
    "l3-8b-tp8": dict(hidden=4096, num_heads=4, num_kv_heads=1, head_dim=128, inter=1792),
    "l3-8b-tp4": dict(hidden=4096, num_heads=8, num_kv_heads=2, head_dim=128, inter=3584),
    "l3-70b-tp8": dict(hidden=8192, num_heads=8, num_kv_heads=1, head_dim=128, inter=3584),
}


def _shape_cfg(shape, bias):
    if RANK_SHAPES[shape] is None:
        return _toy_cfg(bias)
    return om.ModelConfig(num_layers=1, vocab=1024, rope_theta=500000.0, rms_eps=1e-5, max_pos=2048,
                          qkv_bias=bias, **RANK_SHAPES[shape])


@pytest.mark.parametrize("fmt,bias,shape,plan", [
    ("awq", False, "toy", "fused"), ("gptq", True, "toy", "fused"), ("gptq8", False, "toy", "fused"),
    ("awq", False, "toy", "classic"),
    ("awq", False, "l3-8b-tp8", "classic"), ("awq", False, "l3-8b-tp4", "classic"),
    ("gptq", False, "l3-70b-tp8", "classic"), ("awq", False, "l3-8b-tp8", "fused"),
])
def test_decode_steps_match_oracle(lib, fmt, bias, shape, plan):
    from vllm_rs_b200 import engine
    cfg = _shape_cfg(shape, bias)
    NB, MAXB, MAXBLK = 64, 8, 8
    oracle, runner = _build(cfg, fmt, seed=3, num_blocks=NB, max_batch=MAXB, max_blocks=MAXBLK)
    if plan == "classic":
        # a tensor-parallel rank runs the classic launch plan (an all-reduce follows o_proj / down_proj);
        # a no-op all-reduce is the world-size-1 collective and selects that plan on one GPU
        runner.set_allreduce(lambda buf, n, st: None)
    osch = osched.Scheduler(NB, 64, MAXB)
    csch = engine.Scheduler(NB, 64, MAXB)
    rng = np.random.default_rng(1)
    for L in (37, 64, 65, 1, 130, 200):
        toks = rng.integers(0, cfg.vocab, size=L).tolist()
        osch.add(osched.Sequence(toks, 64, max_tokens=40))
        csch.add(toks, 40)
    # prefill on the oracle (the CUDA prefill path is a later row of SURVEY section 8), then hand the
    # cache to the GPU -- from here on both sides run the reference's decode loop.  A second
    # oracle in exact (fp32, no fp16 rounding) mode gives the true result of the same weights:
    # the parity bar is "as close to the truth as the reference-faithful fp16 pipeline".
    with om.exact_mode():
        exact = om.LlamaOracle(cfg, oracle.w, NB)
    ids, pf = osch.schedule()
    ids_c, pf_c = csch.schedule()
    assert pf and (ids, pf) == (ids_c, pf_c)
    pmeta = osched.prepare_prefill([osch.running[i] for i in ids], 64)
    logits = oracle.prefill(pmeta)
    with om.exact_mode():
        exact.prefill(pmeta)
    first = logits.argmax(-1).tolist()
    for s in (osch, csch):
        _, fin = s.filter_prefill_finished(ids)
        s.postprocess(fin, first)
    _sync_cache_to_gpu(oracle, runner)
    max_err = 0.0
    for step in range(6):
        ids, pf = osch.schedule()
        ids_c, _ = csch.schedule()
        assert not pf and ids == ids_c
        mo = osched.prepare_decode([osch.running[i] for i in ids], 64)
        mc = csch.prepare_decode(ids, max_blocks_cap=MAXBLK)
        assert mc["block_tables"].tolist() == mo["block_tables"]          # bit-exact integers
        assert list(mc["slot_mapping"]) == mo["slot_mapping"]
        ref = oracle.decode(mo)
        with om.exact_mode():
            truth = exact.decode(mo)
        B = len(ids)
        bt = torch.zeros(B, MAXBLK, dtype=torch.int32)
        bt[:, :mc["max_blocks"]] = torch.from_numpy(mc["block_tables"].astype(np.int32))
        got, got_ids = runner.decode_device(
            torch.from_numpy(mc["input_ids"].astype(np.int32)).cuda(),
            torch.from_numpy(mc["positions"]).cuda(), torch.from_numpy(mc["slot_mapping"]).cuda(),
            bt.cuda(), torch.from_numpy(mc["context_lens"].astype(np.int32)).cuda())
        got = got.cpu()
        noise = (ref - truth).abs().max().item()       # fp16 noise floor of the reference pipeline
        err_truth = (got - truth).abs().max().item()
        err = (got - ref).abs().max().item()
        max_err = max(max_err, err)
        print(f"[{fmt}] step {step}: noise {noise:.2e} err_truth {err_truth:.2e} err_ref {err:.2e}")
        # (1) as accurate as the fp16 reference pipeline.  Both sides are a max over B x vocab
        #     logits of heavy-tailed round-off, so the ratio of two draws spreads widely: measured
        #     err_truth / noise over these 18 steps is 0.5 .. 1.7 for the unfused plan and
        #     0.7 .. 2.6 for the fused one (the oracle's own noise varies 3e-3 .. 1.5e-1 from
        #     step to step).  x3 slack, same as bar (2).
        assert err_truth <= 3 * noise + 2e-3, f"step {step}: {err_truth} vs noise floor {noise}"
        # (2) and within the combined noise of the fp16 oracle itself (north star: 1e-3 in fp16
        #     terms per op; logits |x| ~ 4 after two layers of fp16 round-off)
        assert err <= 3 * noise + 2e-3, f"step {step}: logits max err {err}, noise {noise}"
        # greedy ids: identical unless the oracle's top-2 are closer than the tolerance
        ref_ids = ref.argmax(-1)
        for b in range(B):
            if int(got_ids[b]) != int(ref_ids[b]):
                assert ref[b].max() - ref[b, int(got_ids[b])] < 3 * noise + 2e-3
        # the new token's K/V landed in the same slots on both sides
        l, slot = 0, mo["slot_mapping"][0]
        exp = oracle.k_cache[l].view(-1, cfg.num_kv_heads, cfg.head_dim)[slot]
        gk = runner.k_cache[l].view(-1, cfg.num_kv_heads, cfg.head_dim)[slot].cpu()
        assert (gk.float() - exp.float()).abs().max() < 4e-3
        toks = ref_ids.tolist()                      # teacher-force the oracle's tokens
        osch.postprocess(ids, toks)
        csch.postprocess(ids, toks)
        _sync_cache_to_gpu(oracle, runner)           # keep caches identical step to step
    print(f"[{fmt}] max |logit err| over 6 decode steps: {max_err:.2e}")
    assert runner.launches_per_step > 0


def test_run_decode_host_path_and_graphs(lib):
    """ModelRunner::run equivalent: host metadata in, token ids out, with CUDA-graph buckets
    and padded rows (graph.rs:741-767); eager and graph replays must agree exactly."""
    from vllm_rs_b200 import engine
    cfg = _toy_cfg()
    NB, MAXB, MAXBLK = 64, 8, 8
    oracle, runner = _build(cfg, "awq", seed=5, num_blocks=NB, max_batch=MAXB, max_blocks=MAXBLK)
    csch = engine.Scheduler(NB, 64, MAXB)
    osch = osched.Scheduler(NB, 64, MAXB)
    rng = np.random.default_rng(2)
    for L in (10, 70, 129):
        toks = rng.integers(0, cfg.vocab, size=L).tolist()
        osch.add(osched.Sequence(toks, 64, max_tokens=20))
        csch.add(toks, 20)
    ids, _ = osch.schedule()
    csch.schedule()
    first = oracle.prefill(osched.prepare_prefill([osch.running[i] for i in ids], 64)).argmax(-1).tolist()
    for s in (osch, csch):
        _, fin = s.filter_prefill_finished(ids)
        s.postprocess(fin, first)
    _sync_cache_to_gpu(oracle, runner)
    ids, _ = csch.schedule()
    osch.schedule()
    meta = csch.prepare_decode(ids, max_blocks_cap=MAXBLK, pinned=runner.pinned)
    eager = runner.run_decode(meta, stream=torch.cuda.Stream())
    ref = oracle.decode(osched.prepare_decode([osch.running[i] for i in ids], 64))
    for b, t in enumerate(eager.tolist()):
        assert ref[b].max() - ref[b, t] < 4e-3
    runner.capture([1, 2, 4, 8])                     # B=3 -> bucket 4 with one padded row
    _sync_cache_to_gpu(oracle, runner)               # undo the eager step's cache write
    replay = runner.run_decode(meta)
    assert replay.tolist() == eager.tolist()
    replay2 = runner.run_decode(meta)
    assert replay2.tolist() == eager.tolist()


def test_prefill_step_and_engine_loop(lib):
    """Prefill branch on the GPU (varlen causal attention over the paged cache, chunk with cached
    context included) vs the oracle, then the whole engine loop (engine.rs:1693-1757) free
    running on the GPU: schedule -> prefill / decode -> postprocess, greedy tokens compared with
    the oracle loop up to the first step whose top-2 logit margin is inside the noise."""
    from vllm_rs_b200 import engine
    cfg = _toy_cfg()
    NB, MAXB, MAXBLK = 96, 8, 16
    oracle, runner = _build(cfg, "awq", seed=11, num_blocks=NB, max_batch=MAXB, max_blocks=MAXBLK,
                            max_prefill_tokens=1024)
    with om.exact_mode():
        exact = om.LlamaOracle(cfg, oracle.w, NB)
    rng = np.random.default_rng(4)
    prompts = [rng.integers(0, cfg.vocab, size=L).tolist() for L in (37, 64, 65, 130, 300)]
    CH = 128                                              # small chunk => the 130/300 prompts are chunked
    osch = osched.Scheduler(NB, 64, MAXB, chunk_size=CH)
    eng = engine.LLMEngine(runner, NB, MAXB, chunk_size=CH)
    for p in prompts:
        osch.add(osched.Sequence(p, 64, max_tokens=6))
        eng.add_request(p, 6)
    # ---- oracle loop, recording tokens and top-2 margins
    o_out = {i: [] for i in range(len(prompts))}
    o_margin = {i: [] for i in range(len(prompts))}
    first_prefill_checked = False
    for _ in range(200):
        if not (osch.running or osch.waiting):
            break
        ids, pf = osch.schedule()
        if not ids:
            break
        seqs = [osch.running[i] for i in ids]
        if pf:
            meta = osched.prepare_prefill(seqs, 64, CH)
            logits = oracle.prefill(meta)
            if not first_prefill_checked:
                # same step on the GPU from the same metadata (fresh prompts; cache empty on both)
                with om.exact_mode():
                    truth = exact.prefill(meta)
                m2 = dict(meta)
                m2["block_tables"] = osched.prepare_block_tables(seqs)[0]
                m2 = {k: (np.asarray(v) if isinstance(v, list) else v) for k, v in m2.items()}
                got, _ = runner.prefill_device(m2)
                noise = (logits - truth).abs().max().item()
                err = (got.cpu() - truth).abs().max().item()
                assert err <= 2 * noise + 2e-3, f"prefill logits {err} vs noise floor {noise}"
                for l in range(cfg.num_layers):          # cache written at the mapped slots
                    d = (runner.k_cache[l].cpu().float() - oracle.k_cache[l].float()).abs().max()
                    assert d < 2e-2, d
                first_prefill_checked = True
            pos, fin = osch.filter_prefill_finished(ids)
            toks = logits.argmax(-1)[pos].tolist() if pos else []
            srt = logits.sort(-1, descending=True).values
            for j, i in zip(pos, fin):
                sid = osch.running[i].id
                o_out[sid].append(int(logits[j].argmax()))
                o_margin[sid].append(float(srt[j, 0] - srt[j, 1]))
            osch.postprocess(fin, toks)
        else:
            logits = oracle.decode(osched.prepare_decode(seqs, 64))
            srt = logits.sort(-1, descending=True).values
            toks = logits.argmax(-1).tolist()
            for j, s in enumerate(seqs):
                o_out[s.id].append(toks[j])
                o_margin[s.id].append(float(srt[j, 0] - srt[j, 1]))
            osch.postprocess(ids, toks)
        osch.clear_finished()
    assert first_prefill_checked
    # ---- GPU engine loop from scratch (fresh caches)
    for l in range(cfg.num_layers):
        runner.k_cache[l].zero_()
        runner.v_cache[l].zero_()
    g_out = eng.run_to_completion()
    assert eng.sched.num_free_blocks == NB
    checked = 0
    for sid in range(len(prompts)):
        for t, (a, b) in enumerate(zip(o_out[sid], g_out[sid])):
            if o_margin[sid][t] < 3e-2:       # inside fp16 noise: either token is legitimate
                break
            assert a == b, f"seq {sid} step {t}: oracle {a} vs gpu {b} (margin {o_margin[sid][t]})"
            checked += 1
        assert len(g_out[sid]) == len(o_out[sid]) == 7    # prefill token + 6 decode tokens
    assert checked >= 15
    # the loop above ran its decode steps through b200_engine_step (one native call per step); the same loop
    # spelled out over the individual C symbols must produce identical tokens
    for l in range(cfg.num_layers):
        runner.k_cache[l].zero_()
        runner.v_cache[l].zero_()
    eng2 = engine.LLMEngine(runner, NB, MAXB, chunk_size=CH, native_step=False)
    for p in prompts:
        eng2.add_request(p, 6)
    g2 = eng2.run_to_completion()
    assert {k: v for k, v in g2.items()} == {k: v for k, v in g_out.items()}


@pytest.mark.parametrize("fmt", ["awq", "gptq8"])
def test_large_prefill_step_dense_route(lib, fmt):
    """A prefill step of >= 1024 tokens takes the other launch plan: per projection a W4/W8 -> fp16
    dequant pass + the dense tcgen05 GEMM (CTA pairs), SiLU * mul inside the gate|up GEMM epilogue,
    tcgen05 attention (head_dim 128).  Checked against the oracle at the logits (noise-floor bar as in
    the other runner tests) and against the oracle's KV cache."""
    cfg = om.ModelConfig(hidden=512, num_layers=2, num_heads=4, num_kv_heads=2, head_dim=128, inter=1024,
                         vocab=1024, rope_theta=10000.0, rms_eps=1e-5, max_pos=4096)
    NB, MAXB, MAXBLK = 48, 4, 16
    oracle, runner = _build(cfg, fmt, seed=5, num_blocks=NB, max_batch=MAXB, max_blocks=MAXBLK,
                            max_prefill_tokens=2048)
    with om.exact_mode():
        exact = om.LlamaOracle(cfg, oracle.w, NB)
    rng = np.random.default_rng(9)
    osch = osched.Scheduler(NB, 64, MAXB, chunk_size=8192)
    for L in (700, 500):
        osch.add(osched.Sequence(rng.integers(0, cfg.vocab, size=L).tolist(), 64, max_tokens=4))
    ids, pf = osch.schedule()
    assert pf and len(ids) == 2
    seqs = [osch.running[i] for i in ids]
    meta = osched.prepare_prefill(seqs, 64, 8192)
    assert len(meta["input_ids"]) == 1200
    ref = oracle.prefill(meta)
    with om.exact_mode():
        truth = exact.prefill(meta)
    m2 = dict(meta)
    m2["block_tables"] = osched.prepare_block_tables(seqs)[0]
    m2 = {k: (np.asarray(v) if isinstance(v, list) else v) for k, v in m2.items()}
    got, ids_gpu = runner.prefill_device(m2)
    launches_big = runner.launches_per_step
    noise = (ref - truth).abs().max().item()
    err = (got.cpu() - truth).abs().max().item()
    print(f"[{fmt}] large prefill: noise {noise:.2e} err_truth {err:.2e} launches {launches_big}")
    assert err <= 3 * noise + 2e-3, f"prefill logits {err} vs noise floor {noise}"
    for l in range(cfg.num_layers):
        d = (runner.k_cache[l].cpu().float() - oracle.k_cache[l].float()).abs().max()
        assert d < 2e-2, d
    for b in range(2):
        if ref[b].sort(descending=True).values[:2].diff().abs() > 3 * noise + 2e-3:
            assert int(ids_gpu[b]) == int(ref[b].argmax())


def _oracle_loop(osch, oracle, chunk, out, margin):
    """The reference's step loop on the oracle; appends greedy tokens and top-2 margins per seq id."""
    for _ in range(400):
        if not (osch.running or osch.waiting):
            break
        ids, pf = osch.schedule()
        if not ids:
            break
        seqs = [osch.running[i] for i in ids]
        if pf:
            logits = oracle.prefill(osched.prepare_prefill(seqs, 64, chunk))
            pos, fin = osch.filter_prefill_finished(ids)
            srt = logits.sort(-1, descending=True).values
            for j, i in zip(pos, fin):
                out.setdefault(osch.running[i].id, []).append(int(logits[j].argmax()))
                margin.setdefault(osch.running[i].id, []).append(float(srt[j, 0] - srt[j, 1]))
            osch.postprocess(fin, logits.argmax(-1)[pos].tolist() if pos else [])
        else:
            logits = oracle.decode(osched.prepare_decode(seqs, 64))
            srt = logits.sort(-1, descending=True).values
            toks = logits.argmax(-1).tolist()
            for j, s in enumerate(seqs):
                out.setdefault(s.id, []).append(toks[j])
                margin.setdefault(s.id, []).append(float(srt[j, 0] - srt[j, 1]))
            osch.postprocess(ids, toks)
        osch.clear_finished()


def test_engine_loop_with_prefix_cache(lib):
    """Prefix-cache hits end to end: a finished request leaves its full blocks in the
    cache; later requests with the same prefix share those pages by reference count, prefill only the
    uncached tail (attention over the cached pages + the new tokens) and must produce the oracle's
    tokens.  Block sharing / cached-token counts are compared bit-exact on the way."""
    from vllm_rs_b200 import engine
    cfg = _toy_cfg()
    NB, MAXB, MAXBLK = 64, 8, 16
    oracle, runner = _build(cfg, "awq", seed=23, num_blocks=NB, max_batch=MAXB, max_blocks=MAXBLK,
                            max_prefill_tokens=1024)
    rng = np.random.default_rng(9)
    base = rng.integers(0, cfg.vocab, size=200).tolist()
    waves = [[base[:150]],                                                   # cold: 2 full blocks get cached
             [base[:150] + rng.integers(0, cfg.vocab, size=40).tolist(),     # hit: 2 blocks (+ the first's outputs differ)
              base[:128],                                                    # entirely cached blocks: keeps 1 block out
              base[:70] + rng.integers(0, cfg.vocab, size=100).tolist()],    # hit: 1 block
             [base[:200]]]                                                   # hit after the second wave extended the chain
    cap = osched.prefix_cache_max_blocks(NB, 64, None)
    osch = osched.Scheduler(NB, 64, MAXB, chunk_size=8192, prefix_cache_blocks=cap)
    eng = engine.LLMEngine(runner, NB, MAXB, prefix_cache=True)
    o_out, o_margin, hits = {}, {}, []
    for l in range(cfg.num_layers):
        runner.k_cache[l].zero_()
        runner.v_cache[l].zero_()
    for wave in waves:
        for p in wave:
            osch.add(osched.Sequence(p, 64, max_tokens=5))
            eng.add_request(p, 5)
        # peek at the allocation the first schedule() of this wave makes, on both sides
        _oracle_loop(osch, oracle, 8192, o_out, o_margin)
        eng.step()
        hits.append([eng.sched.seq_info(i)["num_cached_tokens"] for i in range(eng.sched.num_running)])
        eng.run_to_completion()
        assert eng.sched.free_block_ids() == list(osch.block_manager.free_block_ids)
        assert eng.sched.prefix_cache_blocks == osch.block_manager.prefix_cache_blocks()
    assert hits[0] == [0] and hits[1] == [128, 64, 64] and hits[2][0] >= 128, hits
    g_out = eng.outputs
    checked = 0
    for sid in sorted(o_out):
        assert len(g_out[sid]) == len(o_out[sid]) == 6
        for t, (a, b) in enumerate(zip(o_out[sid], g_out[sid])):
            if o_margin[sid][t] < 3e-2:
                break
            assert a == b, f"seq {sid} step {t}: oracle {a} vs gpu {b} (margin {o_margin[sid][t]})"
            checked += 1
    assert checked >= 12


@pytest.mark.parametrize("nseq", [1, 2, 4])
@pytest.mark.parametrize("plan", ["fused", "classic"])
def test_decode_small_batches(lib, nseq, plan):
    """Batches of 1-4 sequences (BASELINE: "batch 1-64") run the tcgen05 decode GEMM with a 16-token tile since
    the GEMV lost its M <= 4 slot (B200_GEMV_MAX_M): same parity bars as the larger batches, two steps."""
    from vllm_rs_b200 import engine
    cfg = _toy_cfg()
    NB, MAXB, MAXBLK = 32, 4, 8
    oracle, runner = _build(cfg, "awq", seed=40 + nseq, num_blocks=NB, max_batch=MAXB, max_blocks=MAXBLK)
    if plan == "classic":
        runner.set_allreduce(lambda buf, n, st: None)
    with om.exact_mode():
        exact = om.LlamaOracle(cfg, oracle.w, NB)
    osch, csch = osched.Scheduler(NB, 64, MAXB), engine.Scheduler(NB, 64, MAXB)
    rng = np.random.default_rng(nseq)
    for L in (37, 64, 130, 65)[:nseq]:
        toks = rng.integers(0, cfg.vocab, size=L).tolist()
        osch.add(osched.Sequence(toks, 64, max_tokens=8))
        csch.add(toks, 8)
    ids, _ = osch.schedule()
    csch.schedule()
    pmeta = osched.prepare_prefill([osch.running[i] for i in ids], 64)
    first = oracle.prefill(pmeta).argmax(-1).tolist()
    with om.exact_mode():
        exact.prefill(pmeta)
    for s in (osch, csch):
        _, fin = s.filter_prefill_finished(ids)
        s.postprocess(fin, first)
    _sync_cache_to_gpu(oracle, runner)
    for step in range(2):
        ids, _ = osch.schedule()
        csch.schedule()
        mo = osched.prepare_decode([osch.running[i] for i in ids], 64)
        mc = csch.prepare_decode(ids, max_blocks_cap=MAXBLK)
        ref = oracle.decode(mo)
        with om.exact_mode():
            truth = exact.decode(mo)
        B = len(ids)
        assert B == nseq
        bt = torch.zeros(B, MAXBLK, dtype=torch.int32)
        bt[:, :mc["max_blocks"]] = torch.from_numpy(mc["block_tables"].astype(np.int32))
        got, _ = runner.decode_device(
            torch.from_numpy(mc["input_ids"].astype(np.int32)).cuda(),
            torch.from_numpy(mc["positions"]).cuda(), torch.from_numpy(mc["slot_mapping"]).cuda(),
            bt.cuda(), torch.from_numpy(mc["context_lens"].astype(np.int32)).cuda())
        got = got.cpu()
        noise = (ref - truth).abs().max().item()
        assert (got - truth).abs().max().item() <= 3 * noise + 2e-3
        assert (got - ref).abs().max().item() <= 3 * noise + 2e-3
        toks = ref.argmax(-1).tolist()
        osch.postprocess(ids, toks)
        csch.postprocess(ids, toks)
        _sync_cache_to_gpu(oracle, runner)


def test_runner_parity_with_the_streamk_gemm_forced():
    """The automatic dispatch gives the stream-K decode GEMM (wna16_sk.cu) only big-N projections of small batches,
    which the toy / TP-rank shapes of this file never are.  B200_DEC_SK=1 routes EVERY decode GEMM with M <= 32
    through it -- the fused RoPE + KV-write, residual + norm and SiLU.mul finalisers included -- and the same
    oracle comparisons must hold.  The switch is read once per process, hence the child process."""
    import os
    import subprocess
    import sys
    env = dict(os.environ, B200_DEC_SK="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-p", "no:cacheprovider",
                        "-k", "test_decode_steps_match_oracle or test_decode_small_batches or "
                              "test_run_decode_host_path_and_graphs"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert " passed" in r.stdout
