#!/usr/bin/env python
"""bench.py -- the forward hot path of vllm.rs on B200: decode tokens/s and prefill TFLOP/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config NAME]

Configs (BASELINE.json `configs`):
  decode8b        [1] Llama-3-8B AWQ-int4 decode, batch 32, context 4096 (the headline; TP = N under torchrun,
                      strong scaling).  One "step" = one decode step of the whole model (32 layers + lm_head +
                      greedy sampling).  The N = 1 line also carries `prefill` ([2] as written: fp16 weights,
                      8 x 8192 tokens in one step), `prefill_awq` (one 8192-token AWQ chunk), `sweep` (batch 1 / 8
                      / 64) and `fp8_kv`; the N = 4 / N = 8 lines also carry `configs.q72_tp4` / `configs.l70_tp8`.
  prefill8b_fp16  [2] Llama-3-8B fp16 chunked prefill, 8 sequences x 8192 tokens, as the headline line.
  q72_tp4         [3] Qwen2-72B GPTQ-int4, TP = 4, continuous batching of 16 running sequences (Poisson arrivals).
  l70_tp8         [4] Llama-3-70B AWQ-int4, TP = 8, 64 concurrent sequences, prefill and decode steps alternating
                      exactly as Scheduler::schedule does (scheduler.rs:200-380).

Keys of the JSON line:
  value  device-resident: step inputs already in HBM, K CUDA-graph replays of the SAME step (the same KV slot is
         rewritten every replay: bandwidth-identical to consecutive steps), CUDA events, max over ranks.
  e2e    the reference-facing call: b200_engine_step = scheduler.schedule() -> prepare_decode (pinned host) ->
         b200_runner_decode_host (H2D of the step inputs, graph replay, D2H of the token ids, stream sync) ->
         scheduler.postprocess(), i.e. ModelRunner::run as the engine loop drives it (engine.rs:1693-1757).
  roofline       dominant kernel (paged attention decode), timed alone with CUDA events across all layer caches
                 (>> L2), algorithmic bytes / duration vs the measured HBM peak.
  step_roofline  algorithmic bytes of the whole step (per rank) / step time.
  tp_parity      (N > 1) before timing: TP prefill + 4 decode steps of a 2-layer model of the same per-rank layer
                 shapes vs the unsharded CPU oracle on rank 0 (tests/parity_util.py).
  cpu_baseline   oracle/c/decode_ref.c (C + OpenMP port of the same step; the reference has no CPU path for AWQ,
                 SURVEY F3) -- a whole step: 32 layer passes (over 2 layers' worth of distinct tensors) + lm_head.
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

CONFIGS = {
    "decode8b": dict(model="llama3-8b", fmt="awq", kind="decode", baseline="configs[1]", label="Llama-3-8B AWQ-int4"),
    "prefill8b_fp16": dict(model="llama3-8b", fmt="fp16", kind="prefill", seqs=8, ctx=8192, baseline="configs[2]",
                           label="Llama-3-8B fp16"),
    "q72_tp4": dict(model="qwen2-72b", fmt="gptq_sym", kind="serve", gpus=4, max_num_seqs=16, requests=32,
                    prompt=(1536, 3584), max_tokens=128, mean_gap_steps=6.0, baseline="configs[3]",
                    label="Qwen2-72B GPTQ-int4"),
    "l70_tp8": dict(model="llama3-70b", fmt="awq", kind="serve", gpus=8, max_num_seqs=64, requests=96,
                    prompt=(512, 2048), max_tokens=64, mean_gap_steps=1.0, baseline="configs[4]",
                    label="Llama-3-70B AWQ-int4"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="decode8b", choices=sorted(CONFIGS))
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--ctx", type=int, default=4096)
    ap.add_argument("--layers", type=int, default=0, help="debug only: fewer layers (invalid as a bench)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp8-kv", action="store_true", help="skip the extra FP8-KV-cache decode leg")
    ap.add_argument("--no-sweep", action="store_true", help="skip the batch 1 / 8 / 64 sweep")
    ap.add_argument("--no-parity", action="store_true", help="skip the TP parity check before timing (N > 1)")
    ap.add_argument("--no-extra-configs", action="store_true", help="N = 4 / 8: skip the q72_tp4 / l70_tp8 legs")
    ap.add_argument("--py-step", action="store_true",
                    help="e2e: drive the step from Python over the individual C symbols instead of b200_engine_step")
    ap.add_argument("--prefill-tokens", type=int, default=8192,
                    help="N = 1: also time the prefill legs (0 = skip); the AWQ chunk has this many tokens")
    return ap.parse_args()


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1600.0, "bf16_tflops_sustained": 1350.0}, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel: str, key: str):
    """dram bytes (read + write) of one launch from THIS round's committed `ncu --set full` capture
    (profiles/r2_traffic.json: {kernel: {shape key: bytes, "source": file}}); None when no capture matches."""
    p = ROOT / "profiles" / "r2_traffic.json"
    if not p.exists():
        return None, None
    d = json.loads(p.read_text()).get(kernel, {})
    return d.get(key), d.get("source")


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
def cpu_baseline(batch, ctx, steps=1, model="llama3-8b", distinct_layers=2):
    """C + OpenMP port on the host cores: a WHOLE decode step per sample -- all L layer passes (cycling over
    `distinct_layers` layers' worth of distinct weights and KV so the host footprint stays ~1.3 GB; each pass
    still streams ~650 MB, far beyond any CPU cache) + final norm + lm_head + argmax.  No extrapolation."""
    import numpy as np
    from oracle import cref
    from vllm_rs_b200.synthetic import SHAPES
    hidden, L, H, Hkv, D, inter, vocab, _ = SHAPES[model]
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
    for _ in range(distinct_layers):
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
    all_layers = [layers[i % distinct_layers] for i in range(L)]
    times = []
    for _ in range(steps):
        x = rng.standard_normal((batch, hidden)).astype(np.float16)
        t0 = time.perf_counter()
        cref.decode_layers(x, all_layers, cos, sin, pos, slots, bt, cl, H=H, Hkv=Hkv, D=D, inter=inter, group=G)
        xn = cref.rmsnorm(x, np.ones(hidden, np.float16), 1e-5)
        logits = cref.dense_gemm(xn, lm_head)
        _ = logits.argmax(-1)
        times.append(time.perf_counter() - t0)
    step_s = min(times)
    return {"value": batch / step_s, "unit": "tokens/s", "cores": threads, "kind": "port",
            "sample": f"{steps} whole step(s) at batch {batch} ctx {ctx}: {L} layer passes over {distinct_layers} "
                      f"layers' distinct weights + KV, + lm_head; {sum(times):.1f} s of CPU work, no extrapolation; "
                      f"oracle/c/decode_ref.c, {os.cpu_count()} host cpus"}, step_s, times


def run_reference(args):
    """The reference arm for this tier: the reference's own CPU implementation of the path cannot be built
    here (Rust, un-vendored kernels; and it has no CPU path for AWQ at all), so this times the
    oracle PORT (kind "port") on every host core, on the b200 arm's metric / config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config != "decode8b":
        print(json.dumps({"impl": "reference", "unavailable": f"the CPU port only restates the decode step "
                                                              f"(config decode8b), not {args.config}"}))
        return
    sampled = max(1, min(args.steps, 2))
    base, step_s, times = cpu_baseline(args.batch, args.ctx, steps=sampled)
    v = args.batch / step_s
    print(json.dumps({
        "impl": "reference", "metric": f"decode tokens/s (Llama-3-8B AWQ-int4, b={args.batch}, ctx={args.ctx})",
        "value": v, "unit": "tokens/s", "n_gpus": args.gpus, "steps": sampled, "steps_requested": args.steps,
        "steps_sampled": sampled, "sampled_seconds": [round(t, 3) for t in times], "warmup": 0,
        "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"llama3-8b AWQ-int4 g128 decode batch={args.batch} ctx={args.ctx} layers=32 "
                               f"block_size=64 tp=1 (BASELINE.json configs[1])",
                   "sample": "every step is a whole decode step on the host cores (32 layer passes + lm_head); "
                             "ms_per_step is the fastest of the sampled steps, nothing is extrapolated",
                   "gpus_used": 0},
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


class Ctx:
    """Rank / device / distributed context of one bench process."""

    def __init__(self):
        import torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)

    def barrier(self):
        import torch
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v: float) -> float:
        if self.world == 1:
            return v
        import torch
        import torch.distributed as dist
        t = torch.tensor([v], device=self.dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


def graph_time(cx: Ctx, fn, K, W, prof=None, sample_clocks=False):
    """Capture `fn` (one step) into a CUDA graph on cx.stream, W warm-up replays, K timed replays bracketed
    by barrier + synchronize, CUDA events, max over ranks.  Returns (ms_per_step, clocks | None)."""
    import torch
    stream = cx.stream
    clocks = None
    # inputs and caches were produced on the default stream (decode_inputs, fill_kv_random): the step stream must
    # not start before they exist (the FP8-KV leg read unwritten positions -> wild cos / sin rows -> illegal address)
    stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(stream):
        for _ in range(2):
            fn()
        stream.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=stream):
            fn()
        for _ in range(W):
            graph.replay()
        stream.synchronize()
        cx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def timed():
            if prof:
                _prof(prof, True)
            e0.record(stream)
            for _ in range(K):
                graph.replay()
            e1.record(stream)
            stream.synchronize()
            if prof:
                _prof(prof, False)
            cx.barrier()
            return cx.max_over_ranks(e0.elapsed_time(e1))

        if sample_clocks:
            with ClockSampler(cx.local) as clk:
                ms = timed()
                # keep the GPU under load ~1 s more so short runs still get a few clock samples.  The number of
                # extra replays must be IDENTICAL on every rank (each replay contains the TP exchange: a rank that
                # replays once more than its peers waits for them forever), so it derives from the rank-agreed time.
                extra = max(1, min(2000, int(1000.0 / max(ms / K, 1e-3))))
                for _ in range(extra):
                    graph.replay()
                stream.synchronize()
            clocks = clk.summary()
        else:
            ms = timed()
    del graph
    return ms / K, clocks


def decode_inputs(cx: Ctx, runner, B, ctx, blocks_per_seq, num_blocks, vocab, seed=1):
    import torch
    dev = cx.dev
    g = torch.Generator(device=dev).manual_seed(seed)
    perm = torch.randperm(num_blocks, device=dev, generator=g).to(torch.int32)
    bt = perm[:B * blocks_per_seq].reshape(B, blocks_per_seq).contiguous()
    ids = torch.randint(0, vocab, (B,), device=dev, generator=g, dtype=torch.int32)
    pos = torch.full((B,), ctx - 1, dtype=torch.int64, device=dev)
    cl = torch.full((B,), ctx, dtype=torch.int32, device=dev)
    slots = bt[:, (ctx - 1) // 64].to(torch.int64) * 64 + (ctx - 1) % 64
    return ids, pos, slots, bt, cl


def attention_roofline(cx: Ctx, runner, B, ctx, blocks_per_seq, bt, cl, peak, peak_src, kv_bytes=2):
    """The dominant kernel timed alone over every layer's cache (>> L2), CUDA events on the launching stream."""
    import numpy as np
    import torch
    from vllm_rs_b200 import ops
    shape = runner.shape
    D, L = shape["D"], shape["L"]
    q = torch.randn(B, shape["H"], D, device=cx.dev, dtype=torch.float16)
    ws = ops.paged_attn_workspace(B, shape["H"], D, blocks_per_seq, cx.dev)
    out = torch.empty_like(q)
    scale = 1.0 / float(np.sqrt(D))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cx.stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cx.stream):
        for rep in range(2):      # first pass warms up, second is timed
            e0.record(cx.stream)
            for li in range(L):
                ops.paged_attention_decode(q, runner.k_cache[li], runner.v_cache[li], bt, cl, scale, ws, out=out)
            e1.record(cx.stream)
            cx.stream.synchronize()
        attn_ms = e0.elapsed_time(e1) / L
    attn_bytes = B * ctx * shape["Hkv"] * D * 2 * kv_bytes + 2 * B * shape["H"] * D * 2
    achieved = attn_bytes / (attn_ms * 1e-3) / 1e9
    name = "paged_attn_decode_kernel<128" + (", fp8>" if kv_bytes == 1 else ">")
    key = f"b{B}_ctx{ctx}_hkv{shape['Hkv']}_kv{kv_bytes}"
    traffic, src = measured_traffic(name, key)
    return {"bound": "hbm", "kernel": name + " (+ split merge)", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "traffic_source": src, "peak_source": peak_src,
            "bytes_per_launch": attn_bytes, "ms_per_launch": attn_ms}


# ---------------------------------------------------------------------- decode config (the headline)
def run_decode(args, cx: Ctx):
    import numpy as np
    import torch
    from vllm_rs_b200 import _lib, engine, synthetic
    c = CONFIGS["decode8b"]
    world, rank, dev = cx.world, cx.rank, cx.dev
    B, ctx, K, W = args.batch, args.ctx, args.steps, max(args.warmup, 3)
    model = c["model"]
    hidden, L, H, Hkv, D, inter, vocab, _ = synthetic.SHAPES[model]
    if args.layers:
        L = args.layers
    tp = world
    peaks, peak_src = measured_peaks()
    peak = float(peaks["hbm_gbs"])

    # ---------------- TP parity before timing (N > 1): same per-rank layer shapes, 2 layers, real exchange
    tp_par = None
    if world > 1 and not args.no_parity:
        try:
            sys.path.insert(0, str(ROOT / "tests"))
            from oracle import model as om
            from parity_util import tp_parity
            pcfg = om.ModelConfig(hidden=hidden, num_layers=2, num_heads=H, num_kv_heads=Hkv, head_dim=D, inter=inter,
                                  vocab=4096, rope_theta=500000.0, rms_eps=1e-5, max_pos=2048)
            t0 = time.time()
            tp_par = tp_parity(pcfg, "awq", rank=rank, world=world, dev=str(dev), steps=4)
            tp_par["seconds"] = round(time.time() - t0, 1)
            tp_par["model"] = f"{model} layer shapes, 2 layers, vocab 4096, TP prefill of 321 tokens + 4 decode steps"
        except Exception as e:      # noqa: BLE001 -- reported, never hidden: the line then carries the failure
            tp_par = {"error": repr(e)}
        cx.barrier()
        _trace(f"tp parity: {tp_par}")

    sweep_b = [] if (args.no_sweep or args.layers) else [1, 8, 64]
    maxB = max([B] + sweep_b)
    blocks_per_seq = (ctx + K + W + 64) // 64 + 1
    # + 2: the reference's can_append rule wants two free blocks while a sequence is inside a block
    # (block_manager.rs:233-239); with one sequence the per-sequence slack alone is a single block
    num_blocks = maxB * blocks_per_seq + 2
    runner = synthetic.build_awq_runner(model, num_blocks=num_blocks, max_blocks_per_seq=blocks_per_seq,
                                        max_batch=maxB, num_layers=L, tp_size=tp, tp_rank=rank, device=dev)
    synthetic.fill_kv_random(runner)
    _trace("runner built")
    if tp > 1:
        from vllm_rs_b200 import tp as tpmod
        tpmod.attach_allreduce(runner)
    shape = runner.shape
    torch.cuda.synchronize()
    _trace("comm attached")
    l = _lib.lib()

    # ---------------- device-resident value: graph replays of one step at ctx
    ids, pos, slots, bt, cl = decode_inputs(cx, runner, B, ctx, blocks_per_seq, num_blocks, vocab)

    def step_direct():
        runner.decode_device(ids, pos, slots, bt, cl, want_logits=False)

    ms_per_step, clocks = graph_time(cx, step_direct, K, W, prof="decode", sample_clocks=True)
    launches_per_step = runner.launches_per_step
    value = B * 1e3 / ms_per_step
    _trace(f"timed region done: {ms_per_step:.3f} ms/step")

    # ---------------- roofline of the dominant kernel (paged attention decode), timed alone
    roof = attention_roofline(cx, runner, B, ctx, blocks_per_seq, bt, cl, peak, peak_src)
    alg = synthetic.decode_algorithmic_bytes(shape, B, ctx)
    step_gbs = alg["total"] / (ms_per_step * 1e-3) / 1e9

    # ---------------- batch sweep (north star: batch 1-64): graph replays only
    sweep = None
    if sweep_b:
        sweep = {}
        for b in sweep_b:
            si = decode_inputs(cx, runner, b, ctx, blocks_per_seq, num_blocks, vocab, seed=2)
            ms_b, _ = graph_time(cx, lambda: runner.decode_device(*si, want_logits=False), max(10, K // 2), 3)
            a = synthetic.decode_algorithmic_bytes(shape, b, ctx)
            sweep[str(b)] = {"tokens_per_s": b * 1e3 / ms_b, "ms_per_step": ms_b,
                             "step_frac": a["total"] / (ms_b * 1e-3) / 1e9 / peak}
        _trace(f"sweep: {sweep}")

    # ---------------- e2e: scheduler -> prepare_decode (host) -> decode_host -> postprocess
    sched = engine.Scheduler(num_blocks, 64, max_num_seqs=max(B, 5), chunk_size=8192)
    rng = np.random.default_rng(0)
    for _ in range(B):
        sched.add(rng.integers(0, vocab, size=ctx - 1), max_tokens=K + W + 8)
    idx, is_prefill = sched.schedule()
    assert is_prefill and len(idx) == B, (is_prefill, len(idx))
    _, fin = sched.filter_prefill_finished(idx)     # KV of the prompts is the random fill above
    sched.postprocess(fin, rng.integers(0, vocab, size=B))
    runner.capture([B])
    _trace("runner graph captured")

    # the step loop itself is one native call (b200_engine_step: schedule -> prepare_decode into pinned staging
    # -> b200_runner_decode_host -> postprocess -> clear_finished); --py-step spells it out over the single symbols
    import ctypes as C
    eh = None
    if not args.py_step:
        eh = l.b200_engine_new(sched._h, runner._h, maxB, blocks_per_seq)
        _lib.check()
        e_idx = np.zeros(max(maxB, 5), np.int32)
        e_tok = np.zeros(max(maxB, 5), np.uint32)
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
    cx.barrier()
    t0 = time.perf_counter()
    for _ in range(K):
        engine_step()
    torch.cuda.synchronize()
    e2e_ms = cx.max_over_ranks((time.perf_counter() - t0) * 1e3 / K)
    e2e_value = B * 1e3 / e2e_ms
    if eh is not None:
        l.b200_engine_free(eh)
        eh = None
    _trace("e2e done")
    if world > 1 and getattr(runner, "_comm", None):
        cx.barrier()
    del runner, sched
    torch.cuda.empty_cache()

    out_line = {
        "metric": f"decode tokens/s (Llama-3-8B AWQ-int4, b={B}, ctx={ctx})",
        "value": value, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": f"{model} AWQ-int4 g128 decode batch={B} ctx={ctx} layers={L} "
                               f"block_size=64 tp={tp} (BASELINE.json configs[1])",
                   "l2": "inputs >> L2: 17.2 GB KV + 3.6 GB weights streamed per step (126 MB L2)",
                   "timed_step": "K graph replays of the SAME decode step (the same KV slots are rewritten); "
                                 "e2e advances the sequences step by step",
                   "block_tables": "random permutation of pages (value) / scheduler FIFO (e2e)",
                   "tokens_per_sec_per_gpu": value / world},
        "e2e": {"value": e2e_value, "unit": "tokens/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches_per_step) * K,
        "clocks": clocks,
        "roofline": roof,
        "step_roofline": {"algorithmic_bytes": alg["total"], "achieved": step_gbs, "peak": peak,
                          "unit": "GB/s", "frac": step_gbs / peak, "per": "rank"},
    }
    if tp_par is not None:
        out_line["tp_parity"] = tp_par
    if sweep is not None:
        out_line["sweep"] = sweep
    return out_line


# ---------------------------------------------------------------------- prefill legs
def prefill_flops(shape, L, seq_lens):
    D = shape["D"]
    p_lin = (shape["hidden"] * (shape["H"] + 2 * shape["Hkv"]) * D + shape["H"] * D * shape["hidden"]
             + 3 * shape["hidden"] * shape["inter"])
    T = sum(seq_lens)
    attn = sum(L * 4.0 * shape["H"] * D * (t * (t + 1) / 2) for t in seq_lens)
    return 2.0 * p_lin * L * T + attn + 2.0 * shape["vocab"] * shape["hidden"] * len(seq_lens)


def run_prefill_leg(args, cx: Ctx, fmt, seqs, tokens_per_seq, reps=2):
    """One prefill step of `seqs` sequences x `tokens_per_seq` tokens (BASELINE configs[2]: fp16 weights, 8 x 8192
    in one step = 65 536 tokens, causal varlen attention over the paged cache, last-token lm_head)."""
    import numpy as np
    import torch
    from vllm_rs_b200 import synthetic
    model = "llama3-8b"
    L = args.layers or synthetic.SHAPES[model][1]
    vocab = synthetic.SHAPES[model][6]
    T = seqs * tokens_per_seq
    nblk = tokens_per_seq // 64
    runner = synthetic.build_runner(model, fmt, num_blocks=seqs * nblk + 2, max_blocks_per_seq=nblk + 1,
                                    max_batch=max(seqs, 1), num_layers=L, device=cx.dev, max_prefill_tokens=T)
    rng = np.random.default_rng(0)
    bt = np.arange(seqs * nblk, dtype=np.uint32).reshape(seqs, nblk)
    slot = (np.repeat(np.arange(seqs * nblk, dtype=np.int64) * 64, 64) + np.tile(np.arange(64), seqs * nblk))[:T]
    meta = dict(input_ids=rng.integers(0, vocab, size=T).astype(np.uint32),
                positions=np.tile(np.arange(tokens_per_seq, dtype=np.int64), seqs), slot_mapping=slot,
                cu_seqlens_q=(np.arange(seqs + 1) * tokens_per_seq).astype(np.uint32), context_lens=None,
                block_tables=bt, max_seqlen_q=tokens_per_seq)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cx.stream.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(cx.stream):
        runner.prefill_device(meta, want_logits=False)          # warm-up
        cx.stream.synchronize()
        _prof("prefill", True)
        e0.record(cx.stream)
        for _ in range(reps):
            runner.prefill_device(meta, want_logits=False)
        e1.record(cx.stream)
        cx.stream.synchronize()
        _prof("prefill", False)
    pf_ms = e0.elapsed_time(e1) / reps
    flops = prefill_flops(runner.shape, L, [tokens_per_seq] * seqs)
    peaks, _ = measured_peaks()
    tf = flops / (pf_ms * 1e-3) / 1e12
    res = {"weights": fmt, "sequences": seqs, "tokens": T, "ms": pf_ms, "tokens_per_s": T * 1e3 / pf_ms, "tflops": tf,
           "flops": flops, "launches": int(runner.launches_per_step),
           "frac_of_sustained_tensor_peak": tf / float(peaks["bf16_tflops_sustained"]),
           "frac_of_burst_tensor_peak": tf / float(peaks["bf16_tflops"])}
    del runner
    torch.cuda.empty_cache()
    return res


def run_fp8_leg(args, cx: Ctx, K, W, peak):
    """The same decode step with the FP8 KV cache (`--fp8-kvcache`): NOT the headline
    configuration (BASELINE.json's metric is the default fp16 cache) -- reported beside it."""
    import torch
    from vllm_rs_b200 import synthetic
    model = "llama3-8b"
    B, ctx = args.batch, args.ctx
    L = args.layers or synthetic.SHAPES[model][1]
    vocab = synthetic.SHAPES[model][6]
    blocks_per_seq = (ctx + K + W + 64) // 64 + 1
    num_blocks = B * blocks_per_seq + 2
    r8 = synthetic.build_awq_runner(model, num_blocks=num_blocks, max_blocks_per_seq=blocks_per_seq,
                                    max_batch=B, num_layers=L, device=cx.dev, kv_fp8=True)
    synthetic.fill_kv_random(r8)
    si = decode_inputs(cx, r8, B, ctx, blocks_per_seq, num_blocks, vocab)
    ms8, _ = graph_time(cx, lambda: r8.decode_device(*si, want_logits=False), K, W)
    _, peak_src = measured_peaks()
    roof8 = attention_roofline(cx, r8, B, ctx, blocks_per_seq, si[3], si[4], peak, peak_src, kv_bytes=1)
    alg8 = synthetic.decode_algorithmic_bytes(r8.shape, B, ctx, kv_bytes=1)
    res = {"value": B * 1e3 / ms8, "unit": "tokens/s", "ms_per_step": ms8,
           "kv_format": "e4m3, scale 1 (library-defined; parity unpinned)", "attention": roof8,
           "step_roofline": {"algorithmic_bytes": alg8["total"], "achieved": alg8["total"] / (ms8 * 1e-3) / 1e9,
                             "peak": peak, "unit": "GB/s", "frac": alg8["total"] / (ms8 * 1e-3) / 1e9 / peak}}
    del r8
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------- serving configs (configs[3], configs[4])
def run_serve(args, cx: Ctx, name):
    """Continuous batching through the engine loop: Poisson arrivals (seed 2, in units of engine steps so that
    every TP rank makes the same decisions), Scheduler::schedule alternating prefill (<= 8192-token chunks) and
    decode steps (scheduler.rs:200-380), greedy sampling, every request runs to max_tokens."""
    import numpy as np
    import torch
    from vllm_rs_b200 import engine, synthetic
    c = CONFIGS[name]
    world, rank, dev = cx.world, cx.rank, cx.dev
    model, fmt = c["model"], c["fmt"]
    hidden, L, H, Hkv, D, inter, vocab, _ = synthetic.SHAPES[model]
    if args.layers:
        L = args.layers
    S, R = c["max_num_seqs"], c["requests"]
    lo, hi = c["prompt"]
    max_tokens = c["max_tokens"]
    peaks, peak_src = measured_peaks()
    peak = float(peaks["hbm_gbs"])
    blocks_per_seq = (hi + max_tokens + 64) // 64 + 1
    num_blocks = S * blocks_per_seq + 2
    runner = synthetic.build_runner(model, fmt, num_blocks=num_blocks, max_blocks_per_seq=blocks_per_seq,
                                    max_batch=max(S, 5), num_layers=L, tp_size=world, tp_rank=rank, device=dev,
                                    max_prefill_tokens=8192)
    if world > 1:
        from vllm_rs_b200 import tp as tpmod
        tpmod.attach_allreduce(runner)
    torch.cuda.synchronize()
    _trace(f"{name}: runner built")
    shape = runner.shape
    rng = np.random.default_rng(2)
    prompts = [rng.integers(0, vocab, size=int(n)) for n in rng.integers(lo, hi + 1, size=R)]
    arrive = np.floor(np.cumsum(rng.exponential(c["mean_gap_steps"], size=R))).astype(np.int64)
    arrive -= arrive[0]
    runner.capture([b for b in engine.GRAPH_BUCKETS if b <= max(S, 5)])
    eng = engine.LLMEngine(runner, num_blocks, S, chunk_size=8192)
    _trace(f"{name}: graphs captured")

    # warm-up: one short request end to end (prefill kernels, graphs, NCCL channels), then a clean scheduler
    eng.add_request(rng.integers(0, vocab, size=256), 4)
    eng.run_to_completion()
    cx.barrier()
    nxt, it, gen_tokens = 0, 0, 0
    t_pref = t_dec = 0.0
    n_pref = n_dec = 0
    pref_tokens = 0
    t0 = time.perf_counter()
    while nxt < R or eng.sched.num_running or eng.sched.num_waiting:
        while nxt < R and arrive[nxt] <= it:
            eng.add_request(prompts[nxt], max_tokens)
            nxt += 1
        if not (eng.sched.num_running or eng.sched.num_waiting):
            it = int(arrive[nxt])                      # idle: jump to the next arrival
            continue
        ts = time.perf_counter()
        before = sum(len(v) for v in eng.outputs.values())
        if not eng.step():
            raise RuntimeError(f"{name}: engine stalled")
        dt = time.perf_counter() - ts
        made = sum(len(v) for v in eng.outputs.values()) - before
        gen_tokens += made
        if eng.last_kind == "prefill":
            t_pref += dt
            n_pref += 1
            pref_tokens += eng.last_prefill_tokens
        else:
            t_dec += dt
            n_dec += 1
        it += 1
    torch.cuda.synchronize()
    wall = cx.max_over_ranks(time.perf_counter() - t0)
    _trace(f"{name}: served {R} requests, {gen_tokens} tokens in {wall:.2f} s")

    # device-resident decode step at the full running batch and the mean served context
    ctx_mean = int((lo + hi) // 2 + max_tokens // 2)
    si = decode_inputs(cx, runner, S, ctx_mean, blocks_per_seq, num_blocks, vocab)
    K, W = max(10, args.steps), max(args.warmup, 3)
    ms_step, clocks = graph_time(cx, lambda: runner.decode_device(*si, want_logits=False), K, W, sample_clocks=True)
    roof = attention_roofline(cx, runner, S, ctx_mean, blocks_per_seq, si[3], si[4], peak, peak_src)
    alg = synthetic.decode_algorithmic_bytes(shape, S, ctx_mean)
    launches = int(runner.launches_per_step)
    h2d = S * (4 + 8 + 8 + 4) + S * blocks_per_seq * 4
    res = {
        "metric": f"serving tokens/s ({c['label']}, TP={world}, {S} running sequences)",
        "value": S * 1e3 / ms_step, "unit": "tokens/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": f"{model} {fmt} g128 tp={world} max_num_seqs={S} requests={R} prompt U[{lo},{hi}] "
                               f"max_tokens={max_tokens} Poisson arrivals (mean gap {c['mean_gap_steps']} steps, seed 2) "
                               f"layers={L} block_size=64 chunk=8192 (BASELINE.json {c['baseline']})",
                   "value_is": f"graph replays of one decode step, {S} sequences at context {ctx_mean}",
                   "e2e_is": "generated tokens / wall time of the whole serving run through the engine loop "
                             "(prefill + decode steps, host scheduler, H2D / D2H every step)"},
        "e2e": {"value": gen_tokens / wall, "unit": "tokens/s", "seconds": wall, "generated_tokens": gen_tokens,
                "prefill_steps": n_pref, "prefill_tokens": pref_tokens, "prefill_seconds": t_pref,
                "decode_steps": n_dec, "decode_seconds": t_dec,
                "decode_tokens_per_s": (gen_tokens - n_pref) / t_dec if t_dec > 0 else None,
                "prefill_tokens_per_s": pref_tokens / t_pref if t_pref > 0 else None,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": S * 4},
        "gpu_launches": launches * K, "clocks": clocks, "roofline": roof,
        "step_roofline": {"algorithmic_bytes": alg["total"], "achieved": alg["total"] / (ms_step * 1e-3) / 1e9,
                          "peak": peak, "unit": "GB/s", "frac": alg["total"] / (ms_step * 1e-3) / 1e9 / peak,
                          "per": "rank"},
    }
    if world > 1:
        cx.barrier()
    del eng, runner
    torch.cuda.empty_cache()
    return res


# ---------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch.distributed as dist
    cx = Ctx()
    cfgname = args.config
    c = CONFIGS[cfgname]
    if c["kind"] == "serve":
        if cx.world != c["gpus"] and not args.layers:
            raise SystemExit(f"--config {cfgname} is defined on {c['gpus']} GPUs (torchrun --nproc-per-node {c['gpus']})")
        line = run_serve(args, cx, cfgname)
    elif c["kind"] == "prefill":
        if cx.world != 1:
            raise SystemExit("--config prefill8b_fp16 is a single-GPU configuration")
        pf = run_prefill_leg(args, cx, "fp16", c["seqs"], c["ctx"])
        peaks, peak_src = measured_peaks()
        line = {"metric": "prefill TFLOP/s (Llama-3-8B fp16, 8 x 8192 tokens)", "value": pf["tflops"],
                "unit": "TFLOP/s", "n_gpus": 1, "steps": 2, "warmup": 1, "ms_per_step": pf["ms"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": {"workload": "llama3-8b fp16 prefill 8 sequences x 8192 tokens, one step "
                                       "(BASELINE.json configs[2])"},
                "roofline": {"bound": "tensor", "achieved": pf["tflops"], "peak": float(peaks["bf16_tflops_sustained"]),
                             "unit": "TFLOP/s", "frac": pf["frac_of_sustained_tensor_peak"], "traffic": None,
                             "peak_source": peak_src + " sustained"},
                "gpu_launches": pf["launches"] * 2, "prefill": pf}
    else:
        line = run_decode(args, cx)
        if cx.world == 1 and args.prefill_tokens > 0 and not args.layers:
            for key, fn in (("prefill", lambda: run_prefill_leg(args, cx, "fp16", 8, 8192)),
                            ("prefill_awq", lambda: run_prefill_leg(args, cx, "awq", 1, args.prefill_tokens))):
                try:
                    line[key] = fn()
                except Exception as e:      # noqa: BLE001 -- the headline numbers stand on their own
                    line[key] = {"error": repr(e)}
                _trace(f"{key}: {line[key]}")
        if cx.world == 1 and not args.no_fp8_kv:
            try:
                line["fp8_kv"] = run_fp8_leg(args, cx, args.steps, max(args.warmup, 3), line["roofline"]["peak"])
            except Exception as e:          # noqa: BLE001
                line["fp8_kv"] = {"error": repr(e)}
        extra = {4: "q72_tp4", 8: "l70_tp8"}.get(cx.world)
        if extra and not args.no_extra_configs and not args.layers:
            # a failure here must be the same on every rank (same code path, same shapes): reported, not fatal
            try:
                line.setdefault("configs", {})[extra] = run_serve(args, cx, extra)
            except Exception as e:          # noqa: BLE001
                line.setdefault("configs", {})[extra] = {"error": repr(e)}
    if cx.rank == 0:
        if cx.world == 1 and cfgname == "decode8b" and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], _, _ = cpu_baseline(args.batch, args.ctx, steps=1)
            except Exception as e:       # the GPU numbers stand even if the host port cannot run
                line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": 0, "kind": "port",
                                        "sample": f"failed: {e!r}"}
        print(json.dumps(line))
    if cx.world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
