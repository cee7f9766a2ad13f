"""Shared parity helpers (TEST INFRASTRUCTURE: drives the CUDA path and checks it against oracle/).

  ulp16_err       logits error in units of one fp16 ulp at the row's logit scale
  Checkpoint      a seeded synthetic quantised / dense checkpoint in the reference's on-disk formats, the same
                  on every rank, with the dequantised weights the unsharded oracle uses
  tp_parity       world-size-N check used by tests/test_tp_gpu.py AND by `bench.py --gpus N` before timing:
                  every rank shards the checkpoint as the reference does (wna16.rs:35-40,127-132,
                  distributed.rs:438-454,498-538), prefills and decodes through the CUDA runner with the real
                  exchange step; rank 0 replays the same steps on the unsharded CPU oracle.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def ulp16(x: torch.Tensor) -> torch.Tensor:
    """One fp16 unit in the last place at magnitude |x| (normal range; subnormal spacing below 2^-14)."""
    a = x.abs().to(torch.float64).clamp_min(2.0 ** -14)
    return torch.exp2(torch.floor(torch.log2(a)) - 10)


def ulp16_err(got: torch.Tensor, truth: torch.Tensor) -> float:
    """max |got - truth| in fp16 ulps of the ROW's largest |logit| (the scale a fp16 logits tensor of the
    reference is rounded at, llama.rs:315-319): 1.0 = one rounding step of the largest logit."""
    scale = ulp16(truth.abs().amax(dim=-1, keepdim=True))
    return float(((got.to(torch.float64) - truth.to(torch.float64)).abs() / scale).max())


def load_hf_fixture(name: str):
    """tests/golden/<name>.{json,npz} (HuggingFace transformers, fp32 math on fp16-valued weights; generator:
    tests/golden/make_golden.py) -> (meta, oracle ModelConfig, oracle ModelWeights, npz)."""
    import json
    from pathlib import Path

    from oracle import model as om
    gold = Path(__file__).parent / "golden"
    meta = json.loads((gold / f"{name}.json").read_text())
    z = np.load(gold / f"{name}.npz")
    cfg = om.ModelConfig(**meta["cfg"])
    w = om.ModelWeights(embed=torch.from_numpy(z["embed"]), final_norm=torch.from_numpy(z["final_norm"]),
                        lm_head=torch.from_numpy(z["lm_head"]))
    for l in range(cfg.num_layers):
        def t(k):
            return torch.from_numpy(z[f"l{l}.{k}"])
        lw = om.LayerWeights(t("input_norm"), t("post_norm"), t("wq"), t("wk"), t("wv"), t("wo"),
                             t("wgate"), t("wup"), t("wdown"))
        if cfg.qkv_bias:
            lw.bq, lw.bk, lw.bv = t("bq"), t("bk"), t("bv")
        w.layers.append(lw)
    return meta, cfg, w, z


class Checkpoint:
    """Synthetic checkpoint of a LLaMA / Qwen2 decoder: fmt in {"awq", "gptq", "gptq_sym", "fp16"}."""

    @classmethod
    def from_dense(cls, cfg, weights):
        """Wrap existing dense fp16 oracle weights (an HF fixture) as an fp16 checkpoint."""
        self = cls.__new__(cls)
        self.cfg, self.fmt, self.group, self.weights = cfg, "fp16", 128, weights
        self.packed = [dict(q=lw.wq, k=lw.wk, v=lw.wv, o=lw.wo, g=lw.wgate, u=lw.wup, d=lw.wdown)
                       for lw in weights.layers]
        return self

    def __init__(self, cfg, fmt: str, seed: int, group: int = 128, scale_boost: float = 1.5):
        from oracle import model as om
        from oracle import quant as oq
        self.cfg, self.fmt, self.group = cfg, fmt, group
        rng = np.random.default_rng(seed)
        g = torch.Generator().manual_seed(seed)
        H, Hkv, D, h, I = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim, cfg.hidden, cfg.inter

        def qlayer(K, N):
            if fmt == "fp16":
                w = (torch.randn(K, N, generator=g) * (1.0 / math.sqrt(K))).to(torch.float16)
                return w, w
            if fmt == "awq":
                qw, qz, s = oq.random_awq(rng, K, N, group)
            else:
                qw, qz, s = oq.random_gptq(rng, K, N, group, 4, sym=(fmt == "gptq_sym"))
            s = (s.astype(np.float32) * scale_boost).astype(np.float16)     # keeps activations O(1) deep in the net
            deq = oq.dequant_awq(qw, qz, s, group) if fmt == "awq" else oq.dequant_gptq(qw, qz, s, group, 4)
            return (qw, qz, s), torch.from_numpy(deq)

        def norm_w():
            return (1 + 0.05 * torch.randn(h, generator=g)).to(torch.float16)

        self.weights = om.ModelWeights(embed=torch.randn(cfg.vocab, h, generator=g).to(torch.float16),
                                       final_norm=norm_w(),
                                       lm_head=(0.05 * torch.randn(cfg.vocab, h, generator=g)).to(torch.float16))
        self.packed = []
        for _ in range(cfg.num_layers):
            parts = {n: qlayer(*kn) for n, kn in dict(q=(h, H * D), k=(h, Hkv * D), v=(h, Hkv * D), o=(H * D, h),
                                                       g=(h, I), u=(h, I), d=(I, h)).items()}
            lw = om.LayerWeights(norm_w(), norm_w(), *(parts[n][1] for n in "qkvogud"))
            if cfg.qkv_bias:
                lw.bq, lw.bk, lw.bv = [(0.1 * torch.randn(n, generator=g)).to(torch.float16)
                                       for n in (H * D, Hkv * D, Hkv * D)]
            self.weights.layers.append(lw)
            self.packed.append({n: parts[n][0] for n in parts})

    def build_rank_runner(self, rank: int, world: int, dev, *, num_blocks, max_batch, max_blocks,
                          max_prefill_tokens=0):
        """This rank's shard as a CUDA runner: q/k/v/gate/up column-parallel, o/down row-parallel, kv heads
        replicated when Hkv < world (attention.rs:445-530, mlp.rs:367-442, distributed.rs:526-537)."""
        from oracle import model as om
        from vllm_rs_b200 import engine, ops, tp
        cfg, fmt = self.cfg, self.fmt
        H, Hkv, D = cfg.num_heads, cfg.num_kv_heads, cfg.head_dim
        assert H % world == 0 and cfg.inter % world == 0
        kv_idx, kv_n = tp.kv_head_shard(rank, world, Hkv)
        bits = 16 if fmt == "fp16" else 4
        runner = engine.ModelRunner(
            hidden=cfg.hidden, num_layers=cfg.num_layers, num_heads=H // world, num_kv_heads=Hkv // kv_n,
            head_dim=D, inter=cfg.inter // world, vocab=cfg.vocab, num_blocks=num_blocks,
            max_blocks_per_seq=max_blocks, max_batch=max_batch, bits=bits, group_size=self.group,
            qfmt=ops.QFMT_AWQ if fmt == "awq" else ops.QFMT_GPTQ, qkv_bias=cfg.qkv_bias, max_pos=cfg.max_pos,
            rms_eps=cfg.rms_eps, max_prefill_tokens=max_prefill_tokens, device=dev)
        for l, (lw, pk) in enumerate(zip(self.weights.layers, self.packed)):
            if fmt == "fp16":
                def col(n, idx=rank, cnt=world):
                    w = pk[n]
                    step = w.shape[1] // cnt
                    return w[:, idx * step:(idx + 1) * step]

                def row(n):
                    w = pk[n]
                    step = w.shape[0] // world
                    return w[rank * step:(rank + 1) * step]

                qkv = engine.fuse_dense([col("q"), col("k", kv_idx, kv_n), col("v", kv_idx, kv_n)], dev)
                o = engine.fuse_dense([row("o")], dev)
                gu = engine.fuse_dense([col("g"), col("u")], dev)
                dn = engine.fuse_dense([row("d")], dev)
            else:
                fuse = engine.fuse_awq if fmt == "awq" else engine.fuse_gptq
                col = lambda n, idx=rank, cnt=world: tp.shard_column(*pk[n], idx, cnt)        # noqa: E731
                row = lambda n: tp.shard_row(*pk[n], rank, world, self.group)                 # noqa: E731
                qkv = fuse([col("q"), col("k", kv_idx, kv_n), col("v", kv_idx, kv_n)], dev)
                o = fuse([row("o")], dev)
                gu = fuse([col("g"), col("u")], dev)
                dn = fuse([row("d")], dev)
            bias = None
            if cfg.qkv_bias:
                hq, hk = H * D // world, Hkv * D // kv_n
                bias = torch.cat([lw.bq[rank * hq:(rank + 1) * hq], lw.bk[kv_idx * hk:(kv_idx + 1) * hk],
                                  lw.bv[kv_idx * hk:(kv_idx + 1) * hk]]).to(dev)
            runner.set_layer(l, input_norm=lw.input_norm.to(dev), post_norm=lw.post_norm.to(dev), qkv=qkv, o=o,
                             gate_up=gu, down=dn, qkv_bias=bias)
        cos, sin = om.rope_tables(cfg)
        w = self.weights
        runner.set_globals(w.embed.to(dev), w.final_norm.to(dev), w.lm_head.to(dev), cos.to(dev), sin.to(dev))
        return runner


def _dev_meta(meta, max_blocks, dev):
    B = len(meta["input_ids"])
    bt = torch.zeros(B, max_blocks, dtype=torch.int32)
    bt[:, :meta["max_blocks"]] = torch.tensor(meta["block_tables"], dtype=torch.int32)
    return (torch.tensor(meta["input_ids"], dtype=torch.int32).to(dev),
            torch.tensor(meta["positions"], dtype=torch.int64).to(dev),
            torch.tensor(meta["slot_mapping"], dtype=torch.int64).to(dev), bt.to(dev),
            torch.tensor(meta["context_lens"], dtype=torch.int32).to(dev))


def tp_parity(cfg, fmt: str, *, rank: int, world: int, dev, steps: int = 4, seed: int = 21,
              prompt_lens=(20, 70, 129, 5, 64, 33), with_oracle: bool | None = None, tp_fused=None) -> dict:
    """Prefill + `steps` decode steps of a `world`-way sharded model through the CUDA runner (real exchange
    steps), compared on rank 0 with the UNSHARDED oracle fed the same tokens.  Collective calls: every rank
    must call this with the same arguments.  Returns (rank 0; others get launches only):
      max_err / max_err_ulp16     vs the fp16 oracle;  truth_err_ulp16 vs the fp32 truth of the same weights,
      noise_ulp16                 the fp16 oracle's own distance from that truth,
      ids_equal                   greedy ids equal to the oracle's wherever its top-2 margin exceeds the noise,
      ranks_identical             every rank produced bit-identical logits (replicated lm_head)."""
    import torch.distributed as dist
    from oracle import model as om
    from oracle import sched as osched
    from vllm_rs_b200 import tp
    if with_oracle is None:
        with_oracle = rank == 0
    ck = Checkpoint(cfg, fmt, seed)
    NB, MAXB = 48, max(8, len(prompt_lens))
    MAXBLK = max(4, (max(prompt_lens) + steps + 64) // 64 + 1)
    T = sum(prompt_lens)
    runner = ck.build_rank_runner(rank, world, dev, num_blocks=NB, max_batch=MAXB, max_blocks=MAXBLK,
                                  max_prefill_tokens=max(T, 64))
    comm = None
    if world > 1:
        comm = tp.attach_allreduce(runner)
        if tp_fused is not None:
            runner._l.b200_runner_set_tp_fused(runner._h, int(tp_fused))
    osch = osched.Scheduler(NB, 64, MAXB)
    for L in prompt_lens:
        osch.add(osched.Sequence(np.random.default_rng(1000 + L).integers(0, cfg.vocab, size=L).tolist(), 64,
                                 max_tokens=steps + 4))
    oracle = exact = None
    if with_oracle:
        oracle = om.LlamaOracle(cfg, ck.weights, NB)
        with om.exact_mode():
            exact = om.LlamaOracle(cfg, ck.weights, NB)
    out = dict(max_err=0.0, max_err_ulp16=0.0, truth_err_ulp16=0.0, noise_ulp16=0.0, ids_equal=True,
               ranks_identical=True, steps=steps, world=world, oneshot=bool(comm.oneshot) if comm else None)

    def compare(got, got_ids, ref, truth):
        got = got.float().cpu()
        noise = (ref - truth).abs().max().item()
        out["max_err"] = max(out["max_err"], (got - ref).abs().max().item())
        out["max_err_ulp16"] = max(out["max_err_ulp16"], ulp16_err(got, ref))
        out["truth_err_ulp16"] = max(out["truth_err_ulp16"], ulp16_err(got, truth))
        out["noise_ulp16"] = max(out["noise_ulp16"], ulp16_err(ref, truth))
        ref_ids = ref.argmax(-1)
        for b in range(got.shape[0]):
            if int(got_ids[b]) != int(ref_ids[b]) and float(ref[b].max() - ref[b, int(got_ids[b])]) > 3 * noise + 2e-3:
                out["ids_equal"] = False

    def ranks_same(logits):
        if world == 1:
            return
        allr = [torch.zeros_like(logits) for _ in range(world)]
        dist.all_gather(allr, logits)
        if not all(torch.equal(a, allr[0]) for a in allr):
            out["ranks_identical"] = False

    # ---- prefill step through the CUDA path (row-parallel outputs all-reduced at prefill size)
    ids, pf = osch.schedule()
    assert pf
    pmeta = osched.prepare_prefill([osch.running[i] for i in ids], 64)
    gmeta = dict(pmeta)
    if gmeta.get("block_tables") is None:                       # the CUDA path always reads the paged cache
        tabs = [osch.running[i].block_table for i in ids]
        w = max(len(t) for t in tabs)
        bt = np.zeros((len(tabs), w), dtype=np.uint32)
        for r, t in enumerate(tabs):
            bt[r, :len(t)] = t
        gmeta["block_tables"] = bt
    gmeta = {k: (np.asarray(v) if isinstance(v, list) else v) for k, v in gmeta.items()}
    logits, toks = runner.prefill_device(gmeta)
    torch.cuda.synchronize()
    ranks_same(logits)
    toks = toks.cpu().tolist()
    if with_oracle:
        ref = oracle.prefill(pmeta)
        with om.exact_mode():
            truth = exact.prefill(pmeta)
        compare(logits, toks, ref, truth)
    _, fin = osch.filter_prefill_finished(ids)
    osch.postprocess(fin, toks)                                  # every rank follows the GPU's tokens
    # ---- decode steps
    for _ in range(steps):
        ids, pf = osch.schedule()
        assert not pf
        mo = osched.prepare_decode([osch.running[i] for i in ids], 64)
        logits, got_ids = runner.decode_device(*_dev_meta(mo, MAXBLK, dev))
        torch.cuda.synchronize()
        ranks_same(logits)
        got_ids = got_ids.cpu().tolist()
        if with_oracle:
            ref = oracle.decode(mo)
            with om.exact_mode():
                truth = exact.decode(mo)
            compare(logits, got_ids, ref, truth)
        osch.postprocess(ids, got_ids)
    out["launches_per_step"] = int(runner.launches_per_step)
    if world > 1:
        flag = torch.tensor([1 if out["ranks_identical"] else 0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        out["ranks_identical"] = bool(flag.item())
        dist.barrier()
        comm.close()
    del runner
    return out
