#!/usr/bin/env python
"""bench.py -- decode tokens/s of Llama-3-8B AWQ-int4, batch 32, context 4096 on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one decode step of the whole model (32 layers + lm_head + greedy sampling) for
`batch` sequences at `ctx` tokens of paged KV each: BASELINE.json configs[1].

  value  device-resident: step inputs already in HBM, K CUDA-graph replays, CUDA events.
  e2e    the reference-facing call: C++ scheduler.schedule() -> prepare_decode (host, pinned)
         -> b200_runner_decode_host (H2D of the step inputs, graph replay, D2H of the token
         ids, stream sync) -> scheduler.postprocess(), i.e. ModelRunner::run as the engine
         loop drives it (engine.rs:1693-1757).
  roofline       dominant kernel (paged attention decode), timed alone with CUDA events across
                 all 32 layer caches (17 GB >> L2), algorithmic bytes / duration vs measured HBM.
  fp8_kv         the same decode step with the e4m3 KV cache (`--fp8-kvcache`); reported beside the headline,
                 which stays the reference's default fp16 cache.
  cpu_baseline   oracle/c/decode_ref.c (C + OpenMP port of the same step; the reference has no
                 CPU path for AWQ) on a bounded sample: 2 of 32 layers + lm_head.

Multi-GPU (torchrun, --gpus N): tensor-parallel shards of the same model (strong scaling).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

MODEL = "llama3-8b"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ctx", type=int, default=4096)
    ap.add_argument("--layers", type=int, default=0, help="debug only: fewer layers (invalid as a bench)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp8-kv", action="store_true", help="skip the extra FP8-KV-cache decode leg")
    ap.add_argument("--py-step", action="store_true",
                    help="e2e: drive the step from Python over the individual C symbols instead of b200_engine_step")
    ap.add_argument("--prefill-tokens", type=int, default=8192,
                    help="also time one prefill chunk of this many tokens (0 = skip)")
    return ap.parse_args()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.stop, self.index = [], threading.Event(), index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                self.rows.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        sm = sorted(int(float(r[0])) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        reasons = []
        for i, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


# ---------------------------------------------------------------------- CPU baseline
def cpu_baseline(batch, ctx, steps=2, layers_sample=2):
    """C + OpenMP port on the host cores: `layers_sample` of 32 layers + lm_head per step."""
    import numpy as np
    from oracle import cref
    from vllm_rs_b200.synthetic import SHAPES
    hidden, L, H, Hkv, D, inter, vocab, _ = SHAPES[MODEL]
    G = 128
    rng = np.random.default_rng(0)

    def awq(K, N):
        qw = rng.integers(-2**31, 2**31 - 1, size=(K, N // 8), dtype=np.int32)
        qz = rng.integers(-2**31, 2**31 - 1, size=(K // G, N // 8), dtype=np.int32)
        s = (rng.uniform(0.5, 1.5, size=(K // G, N)) * (2 / 15) / np.sqrt(K)).astype(np.float16)
        return qw, qz, s

    blocks_per_seq = (ctx + 64) // 64 + 1
    nb = batch * blocks_per_seq
    chunk = (rng.standard_normal(1 << 22) * 1.0).astype(np.float16)

    def cache():
        n = nb * 64 * Hkv * D
        return np.resize(chunk, n).reshape(nb, 64, Hkv, D)

    layers = []
    for _ in range(layers_sample):
        lw = {"input_norm": np.ones(hidden, np.float16), "post_norm": np.ones(hidden, np.float16)}
        for nm, (K, N) in dict(q=(hidden, H * D), k=(hidden, Hkv * D), v=(hidden, Hkv * D),
                               o=(H * D, hidden), g=(hidden, inter), u=(hidden, inter),
                               d=(inter, hidden)).items():
            lw[f"{nm}_w"], lw[f"{nm}_z"], lw[f"{nm}_s"] = awq(K, N)
        lw["k_cache"], lw["v_cache"] = cache(), cache()
        layers.append(lw)
    lm_head = (rng.standard_normal((vocab, hidden)) * 0.02).astype(np.float16)
    inv = 1.0 / (500000.0 ** (np.arange(0, D, 2) / D))
    f = np.arange(8192)[:, None] * inv[None, :]
    cos, sin = np.cos(f).astype(np.float16), np.sin(f).astype(np.float16)
    bt = rng.permutation(nb).reshape(batch, blocks_per_seq).astype(np.uint32)
    pos = np.full(batch, ctx - 1, np.int64)
    slots = bt[:, (ctx - 1) // 64].astype(np.int64) * 64 + (ctx - 1) % 64
    cl = np.full(batch, ctx, np.uint32)
    # every host core the process may run on (torchrun exports OMP_NUM_THREADS=1 to its workers; under
    # torchrun only rank 0 runs this, the other ranks have exited or are idle)
    try:
        ncores = len(os.sched_getaffinity(0))
    except AttributeError:
        ncores = os.cpu_count() or 1
    cref.lib().ref_set_num_threads(ncores)
    threads = cref.lib().ref_num_threads()
    t_layers, t_head = [], []
    for _ in range(steps):
        x = rng.standard_normal((batch, hidden)).astype(np.float16)
        t0 = time.perf_counter()
        cref.decode_layers(x, layers, cos, sin, pos, slots, bt, cl, H=H, Hkv=Hkv, D=D, inter=inter, group=G)
        t1 = time.perf_counter()
        xn = cref.rmsnorm(x, np.ones(hidden, np.float16), 1e-5)
        logits = cref.dense_gemm(xn, lm_head)
        _ = logits.argmax(-1)
        t2 = time.perf_counter()
        t_layers.append(t1 - t0)
        t_head.append(t2 - t1)
    tl, th = min(t_layers), min(t_head)
    step_s = tl * (L / layers_sample) + th
    return {"value": batch / step_s, "unit": "tokens/s", "cores": threads, "kind": "port",
            "sample": f"{steps} steps of {layers_sample}/{L} layers ({tl:.2f} s) + lm_head ({th:.2f} s) "
                      f"at batch {batch} ctx {ctx}, layers scaled x{L // layers_sample}; "
                      f"oracle/c/decode_ref.c, {os.cpu_count()} host cpus"}, step_s


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = []
    base = None
    for _ in range(max(1, min(args.steps, 3))):
        base, step_s = cpu_baseline(args.batch, args.ctx, steps=1)
        times.append(step_s)
    step_s = min(times)
    v = args.batch / step_s
    base["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": f"decode tokens/s (Llama-3-8B AWQ-int4, b={args.batch}, ctx={args.ctx})",
        "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"{MODEL} AWQ-int4 g128 decode batch={args.batch} ctx={args.ctx} layers=32 "
                               f"block_size=64 tp=1 (BASELINE.json configs[1])",
                   "sample": "each step = 2 of 32 layers + lm_head on the host cores, layers scaled x16"},
        "cpu_baseline": base,
        "e2e": {"value": v, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def _trace(msg):
    """B200_BENCH_TRACE=1: progress lines on stderr (which phase a multi-GPU run is in)."""
    if os.environ.get("B200_BENCH_TRACE"):
        print(f"[bench rank {os.environ.get('RANK', '0')} +{time.time() - _T0:.1f}s] {msg}", file=sys.stderr, flush=True)


_T0 = time.time()


def _prof(which, start):
    """B200_PROFILE=decode|prefill brackets that region with cudaProfilerStart/Stop so that
    `ncu --profile-from-start off` lists exactly its launches (a run under ncu is never a bench value)."""
    if os.environ.get("B200_PROFILE") == which:
        import torch
        rt = torch.cuda.cudart()
        torch.cuda.synchronize()
        (rt.cudaProfilerStart if start else rt.cudaProfilerStop)()


# ---------------------------------------------------------------------- B200 arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from vllm_rs_b200 import _lib, engine, ops, synthetic

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, ctx, K, W = args.batch, args.ctx, args.steps, max(args.warmup, 3)
    blocks_per_seq = (ctx + K + W + 64) // 64 + 1
    # + 2: the reference's can_append rule wants two free blocks while a sequence is inside a block
    # (block_manager.rs:233-239); with one sequence the per-sequence slack alone is a single block
    num_blocks = B * blocks_per_seq + 2
    hidden, L, H, Hkv, D, inter, vocab, _ = synthetic.SHAPES[MODEL]
    if args.layers:
        L = args.layers
    tp = world
    pf_tokens = args.prefill_tokens if world == 1 else 0
    runner = synthetic.build_awq_runner(MODEL, num_blocks=num_blocks, max_blocks_per_seq=blocks_per_seq,
                                        max_batch=B, num_layers=L, tp_size=tp, device=dev)
    synthetic.fill_kv_random(runner)
    _trace("runner built")
    if tp > 1:
        from vllm_rs_b200 import tp as tpmod
        tpmod.attach_allreduce(runner)
    shape = runner.shape
    torch.cuda.synchronize()
    _trace("comm attached")

    # ---------------- device-resident value: graph replays of one step at ctx
    g = torch.Generator(device=dev).manual_seed(1)
    perm = torch.randperm(num_blocks, device=dev, generator=g).to(torch.int32)
    bt = perm[:B * blocks_per_seq].reshape(B, blocks_per_seq).contiguous()
    ids = torch.randint(0, vocab, (B,), device=dev, generator=g, dtype=torch.int32)
    pos = torch.full((B,), ctx - 1, dtype=torch.int64, device=dev)
    cl = torch.full((B,), ctx, dtype=torch.int32, device=dev)
    slots = bt[:, (ctx - 1) // 64].to(torch.int64) * 64 + (ctx - 1) % 64
    stream = torch.cuda.Stream(device=dev)
    l = _lib.lib()

    def step_direct():
        runner.decode_device(ids, pos, slots, bt, cl, want_logits=False)

    with torch.cuda.stream(stream):
        for _ in range(2):
            step_direct()
        stream.synchronize()
        _trace("2 eager steps done")
        launches_per_step = runner.launches_per_step
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            step_direct()
        _trace("graph captured")
        for _ in range(W):
            graph.replay()
        stream.synchronize()
        _trace("warm-up replays done")

        def barrier():
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(local) as clk:
            _prof("decode", True)
            e0.record(stream)
            for _ in range(K):
                graph.replay()
            e1.record(stream)
            stream.synchronize()
            _prof("decode", False)
            barrier()
            ms = e0.elapsed_time(e1)
            # keep the GPU under load ~1 s more so short runs still get a few clock samples.  The number
            # of extra replays must be IDENTICAL on every rank (each replay contains the TP exchange: a
            # rank that replays once more than its peers waits for them forever), so it is derived from
            # the rank-agreed step time, never from a local wall clock.
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            extra = max(1, min(2000, int(1000.0 / max(ms / K, 1e-3))))
            for _ in range(extra):
                graph.replay()
            stream.synchronize()
        clocks = clk.summary()
    ms_per_step = ms / K
    value = B * 1e3 / ms_per_step
    _trace(f"timed region done: {ms_per_step:.3f} ms/step")

    # ---------------- roofline of the dominant kernel (paged attention decode), timed alone
    peak, peak_src = measured_peaks()
    q = torch.randn(B, shape["H"], D, device=dev, dtype=torch.float16)
    ws = ops.paged_attn_workspace(B, shape["H"], D, blocks_per_seq, dev)
    out = torch.empty_like(q)
    scale = 1.0 / float(np.sqrt(D))
    with torch.cuda.stream(stream):
        for rep in range(2):      # first pass warms up, second is timed
            e0.record(stream)
            for li in range(L):
                ops.paged_attention_decode(q, runner.k_cache[li], runner.v_cache[li], bt, cl, scale, ws, out=out)
            e1.record(stream)
            stream.synchronize()
        attn_ms = e0.elapsed_time(e1) / L
    attn_bytes = B * ctx * shape["Hkv"] * D * 2 * 2 + 2 * B * shape["H"] * D * 2
    achieved = attn_bytes / (attn_ms * 1e-3) / 1e9
    alg = synthetic.decode_algorithmic_bytes(shape, B, ctx)
    step_gbs = alg["total"] / (ms_per_step * 1e-3) / 1e9

    # ---------------- e2e: scheduler -> prepare_decode (host) -> decode_host -> postprocess
    sched = engine.Scheduler(num_blocks, 64, max_num_seqs=max(B, 5), chunk_size=8192)
    rng = np.random.default_rng(0)
    for _ in range(B):
        sched.add(rng.integers(0, vocab, size=ctx - 1), max_tokens=K + W + 8)
    idx, is_prefill = sched.schedule()
    assert is_prefill and len(idx) == B, (is_prefill, len(idx))
    _, fin = sched.filter_prefill_finished(idx)     # KV of the prompts is the random fill above
    sched.postprocess(fin, rng.integers(0, vocab, size=B))
    _trace("attention roofline done; capturing the runner's own graph")
    runner.capture([B])
    _trace("runner graph captured")

    # the step loop itself is one native call (b200_engine_step: schedule -> prepare_decode into pinned staging
    # -> b200_runner_decode_host -> postprocess -> clear_finished); --py-step spells it out over the single symbols
    import ctypes as C
    eh = None
    if not args.py_step:
        eh = l.b200_engine_new(sched._h, runner._h, B, blocks_per_seq)
        _lib.check()
        e_idx = np.zeros(max(B, 5), np.int32)
        e_tok = np.zeros(max(B, 5), np.uint32)
        e_npf = C.c_int32(0)
        e_stream = runner._graph_stream.cuda_stream

    def engine_step():
        if eh is not None:
            n = l.b200_engine_step(eh, e_idx.ctypes.data, None, e_tok.ctypes.data, C.byref(e_npf), e_stream)
            if n != B:
                _lib.check()
                raise RuntimeError(f"engine_step returned {n}")
            return
        ii, pf = sched.schedule()
        meta = sched.prepare_decode(ii, max_blocks_cap=blocks_per_seq, pinned=runner.pinned)
        toks = runner.run_decode(meta)
        sched.postprocess(ii, toks)

    for _ in range(W):
        engine_step()
    h2d = B * (4 + 8 + 8 + 4) + B * len(sched.block_table(0)) * 4
    d2h = B * 4
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        engine_step()
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / K
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_value = B * 1e3 / e2e_ms
    if eh is not None:
        l.b200_engine_free(eh)
        eh = None
    _trace("e2e done")

    # ---------------- prefill chunk (second half of BASELINE's metric: "prefill TFLOPS")
    prefill = None
    if pf_tokens > 0:
        T = pf_tokens
        nblk = T // 64
        del runner, graph, sched                      # free 23 GB before building the prefill runner
        torch.cuda.empty_cache()
        runner = synthetic.build_awq_runner(MODEL, num_blocks=nblk + 2, max_blocks_per_seq=nblk + 1, max_batch=1,
                                            num_layers=L, device=dev, max_prefill_tokens=T)
        meta = dict(input_ids=rng.integers(0, vocab, size=T).astype(np.uint32),
                    positions=np.arange(T, dtype=np.int64),
                    slot_mapping=(np.repeat(np.arange(nblk, dtype=np.int64) * 64, 64) + np.tile(np.arange(64), nblk))[:T],
                    cu_seqlens_q=np.array([0, T], dtype=np.uint32), context_lens=None,
                    block_tables=np.arange(nblk, dtype=np.uint32).reshape(1, nblk), max_seqlen_q=T)
        with torch.cuda.stream(stream):
            runner.prefill_device(meta, want_logits=False)          # warm-up
            stream.synchronize()
            _prof("prefill", True)
            e0.record(stream)
            for _ in range(2):
                runner.prefill_device(meta, want_logits=False)
            e1.record(stream)
            stream.synchronize()
            _prof("prefill", False)
        pf_ms = e0.elapsed_time(e1) / 2
        p_lin = shape["hidden"] * (shape["H"] + 2 * shape["Hkv"]) * D + shape["H"] * D * shape["hidden"] + 3 * shape["hidden"] * shape["inter"]
        flops = 2.0 * p_lin * L * T + L * 4.0 * shape["H"] * D * (T * (T + 1) / 2)
        prefill = {"tokens": T, "ms": pf_ms, "tokens_per_s": T * 1e3 / pf_ms, "tflops": flops / (pf_ms * 1e-3) / 1e12,
                   "flops": flops, "launches": int(runner.launches_per_step),
                   "note": "one 8192-token chunk: per projection a W4->fp16 dequant pass + dense tcgen05 GEMM (CTA pairs, "
                           "256x256 tiles) when T >= B200_PREFILL_DENSE_MIN_T (default 1024), else the in-kernel-dequant "
                           "W4A16 tcgen05 GEMM (NT=256); tcgen05 varlen causal attention over the paged cache "
                           "(attn_prefill_tc.cu; B200_PREFILL_ATTN=mma selects the mma.sync kernel)"}

    # ---------------- the same decode step with the FP8 KV cache (`--fp8-kvcache`): NOT the
    # headline configuration (BASELINE.json's metric is the default fp16 cache) -- reported beside it.  Last,
    # and guarded, so nothing here can disturb the numbers above.
    fp8_kv = None
    if world == 1 and not args.no_fp8_kv and D == 128:
        try:
            runner = graph = None
            torch.cuda.empty_cache()
            r8 = synthetic.build_awq_runner(MODEL, num_blocks=num_blocks, max_blocks_per_seq=blocks_per_seq,
                                            max_batch=B, num_layers=L, device=dev, kv_fp8=True)
            synthetic.fill_kv_random(r8)
            with torch.cuda.stream(stream):
                for _ in range(2):
                    r8.decode_device(ids, pos, slots, bt, cl, want_logits=False)
                stream.synchronize()
                g8 = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g8, stream=stream):
                    r8.decode_device(ids, pos, slots, bt, cl, want_logits=False)
                for _ in range(W):
                    g8.replay()
                stream.synchronize()
                e0.record(stream)
                for _ in range(K):
                    g8.replay()
                e1.record(stream)
                stream.synchronize()
                ms8 = e0.elapsed_time(e1) / K
                for rep in range(2):
                    e0.record(stream)
                    for li in range(L):
                        ops.paged_attention_decode(q, r8.k_cache[li], r8.v_cache[li], bt, cl, scale, ws, out=out)
                    e1.record(stream)
                    stream.synchronize()
                a8_ms = e0.elapsed_time(e1) / L
            a8_bytes = B * ctx * shape["Hkv"] * D * 2 * 1 + 2 * B * shape["H"] * D * 2
            alg8 = synthetic.decode_algorithmic_bytes(shape, B, ctx, kv_bytes=1)
            fp8_kv = {"value": B * 1e3 / ms8, "unit": "tokens/s", "ms_per_step": ms8,
                      "kv_format": "e4m3, scale 1 (library-defined; parity unpinned)",
                      "attention": {"kernel": "paged_attn_decode_kernel<128, fp8>", "ms_per_launch": a8_ms,
                                    "bytes_per_launch": a8_bytes, "achieved": a8_bytes / (a8_ms * 1e-3) / 1e9,
                                    "peak": peak, "unit": "GB/s", "frac": a8_bytes / (a8_ms * 1e-3) / 1e9 / peak,
                                    # dram read + write of one launch, profiles/r1_v7_attn_decode_fp8_ncu_full.txt
                                    "traffic": 274544384 if (B, ctx, L) == (32, 4096, 32) else None},
                      "step_roofline": {"algorithmic_bytes": alg8["total"],
                                        "achieved": alg8["total"] / (ms8 * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                        "frac": alg8["total"] / (ms8 * 1e-3) / 1e9 / peak}}
            del r8, g8
        except Exception as e:      # noqa: BLE001 -- the headline numbers stand on their own
            fp8_kv = {"error": repr(e)}
        _trace("fp8 KV leg done")

    out_line = {
        "metric": f"decode tokens/s (Llama-3-8B AWQ-int4, b={B}, ctx={ctx})",
        "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": f"{MODEL} AWQ-int4 g128 decode batch={B} ctx={ctx} layers={L} "
                               f"block_size=64 tp={tp} (BASELINE.json configs[1])",
                   "l2": "inputs >> L2: 17.2 GB KV + 3.6 GB weights streamed per step (126 MB L2)",
                   "block_tables": "random permutation of pages (value) / scheduler FIFO (e2e)",
                   "tokens_per_sec_per_gpu": value / world},
        "e2e": {"value": e2e_value, "unit": "tokens/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches_per_step) * K,
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": "paged_attn_decode_kernel<128> (+ split merge)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch at this exact shape, from the
                     # committed `ncu --set full` capture profiles/r1_attn_decode_ncu_full.txt (537.18 + 8.35 MB)
                     "traffic": 545532160 if (B, ctx, tp, L) == (32, 4096, 1, 32) else None,
                     "peak_source": peak_src, "bytes_per_launch": attn_bytes,
                     "ms_per_launch": attn_ms},
        "step_roofline": {"algorithmic_bytes": alg["total"], "achieved": step_gbs, "peak": peak,
                          "unit": "GB/s", "frac": step_gbs / peak},
        "prefill": prefill,
        "fp8_kv": fp8_kv,
    }
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            try:
                out_line["cpu_baseline"], _ = cpu_baseline(B, ctx, steps=2)
            except Exception as e:       # the GPU numbers stand even if the host port cannot run
                out_line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": 0, "kind": "port",
                                            "sample": f"failed: {e!r}"}
        print(json.dumps(out_line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
