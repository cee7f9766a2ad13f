"""A whole serving run through the engine <-> runner protocol (runner/mod.rs:168-295) on CPU: the native scheduler on
the engine side, frames over a socket pair with the reference's acknowledgment handshake, a runner loop that rebuilds
every step's metadata from the frames alone.  The "model" is a deterministic function of that metadata, so any field a
frame loses or distorts (a slot at a block boundary, a chunked prompt's cached tokens, a block table after
reallocation) changes the generated tokens.  Oracle of the comparison: the same run with the metadata taken
directly from the engine's own prepare_prefill / prepare_decode."""
import socket
import threading

import numpy as np

VOCAB = 50021


def _mix(*vals):
    h = 1469598103934665603
    for v in vals:
        h = ((h ^ (int(v) & 0xFFFFFFFFFFFFFFFF)) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h % VOCAB


def fake_model(kind, meta):
    """One token per sequence from everything a real step would read: the step's last input id, its position, the
    slot its KV goes to, the context length, and (when the step reads cached context) the block table row."""
    out = []
    if kind == "prefill":
        cu = np.asarray(meta["cu_seqlens_q"]).tolist()
        for b in range(len(cu) - 1):
            a, e = cu[b], cu[b + 1]
            extra = []
            if meta.get("block_tables") is not None:
                extra = [int(meta["context_lens"][b])] + np.asarray(meta["block_tables"][b]).tolist()
            out.append(_mix(1, *np.asarray(meta["input_ids"][a:e]).tolist(), *np.asarray(meta["positions"][a:e]).tolist(),
                            *np.asarray(meta["slot_mapping"][a:e]).tolist(), *extra))
    else:
        for b in range(len(meta["input_ids"])):
            ctx = int(meta["context_lens"][b])
            nb = -(-ctx // 64)
            out.append(_mix(2, meta["input_ids"][b], meta["positions"][b], meta["slot_mapping"][b], ctx,
                            *np.asarray(meta["block_tables"][b][:nb]).tolist()))
    return out


def _serve_run(sched, run_prefill, run_decode, on_finished, max_steps=2000):
    """LLMEngine::step (engine.rs:865-1010) around the two step calls; returns {seq_id: generated tokens}."""
    gen = {}
    for _ in range(max_steps):
        ids, pf = sched.schedule()
        if not ids:
            break
        infos = {i: sched.seq_info(i) for i in ids}
        if pf:
            toks = run_prefill(ids)
            pos, fin = sched.filter_prefill_finished(ids)
            sched.postprocess(fin, [toks[p] for p in pos])
            done_rows = fin
            for p, i in zip(pos, fin):
                gen.setdefault(infos[i]["id"], []).append(toks[p])
        else:
            toks = run_decode(ids)
            sched.postprocess(ids, toks)
            done_rows = ids
            for i, t in zip(ids, toks):
                gen.setdefault(infos[i]["id"], []).append(t)
        for i in done_rows:
            if sched.seq_info(i)["status"] == 2:           # Finished
                on_finished(infos[i]["id"])
        sched.clear_finished()
    assert sched.num_running == 0 and sched.num_waiting == 0
    return gen


def _requests(rng):
    lens = [1, 37, 63, 64, 65, 128, 130, 200, 9000]          # the last one is prefilled in two chunks (8192 + 808)
    return [(rng.integers(0, VOCAB, size=L).tolist(), int(rng.integers(3, 70))) for L in lens]


def test_whole_run_over_the_wire_equals_the_direct_run(lib):
    from vllm_rs_b200.engine import Scheduler
    from vllm_rs_b200 import runner_ipc as ipc
    reqs = _requests(np.random.default_rng(11))

    def fresh():
        s = Scheduler(400, 64, 6, 0, 8192)               # 6 running sequences at most: later requests wait
        for toks, mt in reqs:
            s.add(toks, max_tokens=mt)
        return s

    direct = fresh()
    want = _serve_run(direct, lambda ids: fake_model("prefill", direct.prepare_prefill(ids)),
                      lambda ids: fake_model("decode", direct.prepare_decode(ids)), lambda sid: None)
    # (the finish test runs after the append, scheduler.rs:600-610: a sequence ends with max_tokens + 1 outputs)
    assert len(want) == len(reqs) and all(len(want[k]) == mt + 1 for k, (_, mt) in zip(sorted(want), reqs))

    a, b = socket.socketpair()
    finished, served = [], {}
    th = threading.Thread(target=lambda: served.setdefault("steps", ipc.serve(b, fake_model, on_finish=finished.append)))
    th.start()
    try:
        eng_sched = fresh()
        eng = ipc.WireEngine(eng_sched, a)
        got = _serve_run(eng_sched, eng.run_prefill, eng.run_decode, eng.finish)
        eng.shutdown()
    finally:
        th.join(timeout=60)
        a.close()
        b.close()
    assert not th.is_alive()
    assert got == want                                           # every token of every sequence, in order
    assert sorted(finished) == sorted(want)                      # one FinishDecode per sequence reached the runner
    assert served["steps"] > 60
    lib.b200_clear_error()


def test_handshake_rejects_a_wrong_acknowledgment():
    import pytest
    from vllm_rs_b200 import _lib, runner_ipc as ipc
    a, b = socket.socketpair()
    try:
        t = threading.Thread(target=lambda: (b.recv(64), b.sendall(b"\x02")))
        t.start()
        with pytest.raises(_lib.B200Error, match="acknowledgment"):
            ipc.send_local(a, ipc._simple(ipc.MT_HEARTBEAT))
        t.join()
    finally:
        a.close()
        b.close()


# ---------------------------------------------------------------- with a paged "KV cache" on the runner side
class PagedTokenStore:
    """Stand-in for the KV cache of a runner process: every step writes its input ids at the step's slots; a
    sequence's next token is a hash of what its block table READS BACK for the whole context.  KVCacheSwap frames
    move blocks between the GPU array and the CPU swap array exactly as swap_kvcache does (dst[d] = src[s])."""

    def __init__(self, num_blocks, num_cpu_blocks):
        self.gpu = np.full((num_blocks, 64), -1, np.int64)
        self.cpu = np.full((max(num_cpu_blocks, 1), 64), -1, np.int64)
        self.swaps = 0

    def swap(self, pairs, swap_in):
        self.swaps += 1
        for s, d in pairs:
            if swap_in:
                self.gpu[d] = self.cpu[s]
            else:
                self.cpu[d] = self.gpu[s]

    def read(self, table, ctx):
        return self.gpu[np.asarray(table[:-(-ctx // 64)], np.int64)].reshape(-1)[:ctx]

    def step(self, kind, meta):
        flat = self.gpu.reshape(-1)
        out = []
        if kind == "prefill":
            cu = np.asarray(meta["cu_seqlens_q"]).tolist()
            flat[np.asarray(meta["slot_mapping"], np.int64)] = np.asarray(meta["input_ids"], np.int64)
            for b in range(len(cu) - 1):
                ctx = meta["num_cached_tokens"][b] + cu[b + 1] - cu[b]
                out.append(_mix(*self.read(meta["seq_block_tables"][b], ctx).tolist()))
        else:
            flat[np.asarray(meta["slot_mapping"], np.int64)] = np.asarray(meta["input_ids"], np.int64)
            for b in range(len(meta["input_ids"])):
                out.append(_mix(*self.read(meta["block_tables"][b], int(meta["context_lens"][b])).tolist()))
        return out


def test_whole_run_with_cpu_swap_over_the_wire_reads_back_every_history(lib):
    """Few GPU blocks, CPU swap on: sequences are swapped out and back while others reuse their blocks.  The runner's
    paged store only ever sees frames (RunPrefill, RunDecode, KVCacheSwap); the direct run hashes each sequence's true
    token history.  Equal tokens <=> every slot, block table and swap copy list that crossed the wire was right."""
    from vllm_rs_b200.engine import Scheduler
    from vllm_rs_b200 import runner_ipc as ipc
    rng = np.random.default_rng(5)
    reqs = [(rng.integers(0, VOCAB, size=int(L)).tolist(), int(mt))
            for L, mt in zip(rng.integers(150, 300, size=10), rng.integers(150, 260, size=10))]
    # a swap-in wants 1000 tokens of head-room beyond the sequence (scheduler.rs:50): 30 blocks give that to a lone
    # sequence, while six running ones of up to 9 blocks each do not fit
    NB, NCPU = 30, 48

    def fresh():
        s = Scheduler(NB, 64, 6)
        s.enable_cpu_swap(NCPU)
        for toks, mt in reqs:
            s.add(toks, max_tokens=mt)
        return s

    def run(sched, run_prefill, run_decode, on_swap, on_finished):
        gen, now, steps = {}, 0, 0
        for steps in range(4000):
            now += (100, 700, 2600)[steps % 3]
            sched.set_time_ms(now)
            ids, pf = sched.schedule()
            for swap_in, pairs in sched.pop_swap_ops():              # the host runs these before the forward
                on_swap(pairs, swap_in)
            if not ids:
                if sched.num_swapped == 0 and sched.num_waiting == 0:
                    break
                continue
            infos = {i: sched.seq_info(i) for i in ids}
            assert all(v["status"] != 2 for v in infos.values())     # sizes chosen so no sequence is stranded
            if pf:
                toks = run_prefill(ids)
                pos, fin = sched.filter_prefill_finished(ids)
                sched.postprocess(fin, [toks[p] for p in pos])
                rows = list(zip(fin, [toks[p] for p in pos]))
            else:
                toks = run_decode(ids)
                sched.postprocess(ids, toks)
                rows = list(zip(ids, toks))
            for i, t in rows:
                gen.setdefault(infos[i]["id"], []).append(t)
                if sched.seq_info(i)["status"] == 2:
                    on_finished(infos[i]["id"])
            sched.clear_finished()
        assert sched.num_running == 0 and sched.num_waiting == 0 and sched.num_swapped == 0
        return gen

    # direct run: the token after a step is the hash of the sequence's true history up to and including the step's input
    d = fresh()
    swaps = []
    want = run(d, lambda ids: [_mix(*d.tokens(i)[:d.seq_info(i)["num_cached_tokens"] + min(8192, d.seq_info(i)["len"] - d.seq_info(i)["num_cached_tokens"])]) for i in ids],
               lambda ids: [_mix(*d.tokens(i)) for i in ids], lambda p, f: swaps.append(f), lambda sid: None)
    assert len(want) == len(reqs) and True in swaps and False in swaps      # swapped out AND back in

    store = PagedTokenStore(NB, NCPU)
    a, b = socket.socketpair()
    th = threading.Thread(target=lambda: ipc.serve(b, store.step, on_swap=store.swap))
    th.start()
    try:
        es = fresh()
        eng = ipc.WireEngine(es, a)
        got = run(es, eng.run_prefill, eng.run_decode, lambda p, f: eng.kv_swap(p, f), eng.finish)
        eng.shutdown()
    finally:
        th.join(timeout=60)
        a.close()
        b.close()
    assert not th.is_alive()
    assert got == want
    assert store.swaps == len(swaps)
    lib.b200_clear_error()
