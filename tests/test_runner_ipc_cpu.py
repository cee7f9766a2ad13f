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
