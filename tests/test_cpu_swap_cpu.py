"""CPU swap bookkeeping (host side; data path = b200_swap_blocks, tests/test_kernels_gpu.py).
C++ mirror (through the C ABI) vs the Python oracle vs hand-derived expectations, BIT-EXACT, no GPU.

Reference rules exercised (scheduler.rs:303-337,830-954; block_manager.rs:255-272,878-1000):
  * under pressure with no prefix cache to evict, the OLDEST sequence that cannot append is swapped out, only while
    another sequence keeps running; its GPU blocks go back to the free list in reverse table order, CPU blocks are
    taken from the front of the CPU free list in table order (the (gpu, cpu) copy list);
  * a swapped sequence comes back at most one per schedule(), only after 5 s (host clock), when nothing is
    preempted, usage < 0.855 and free blocks > its block count with >= 1000 spare tokens; it gets FRESH blocks from
    the front of the free list and the copy list maps its old CPU blocks onto them in order.
"""
import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import sched as osched

BS = 64


@pytest.fixture(scope="module")
def eng(lib):
    from vllm_rs_b200 import engine
    return engine


def _mk(eng, num_blocks, num_cpu, max_num_seqs=8, prefix_blocks=0):
    o = osched.Scheduler(num_blocks, BS, max_num_seqs, prefix_cache_blocks=prefix_blocks)
    o.block_manager.enable_cpu_swap(num_cpu)
    c = eng.Scheduler(num_blocks, BS, max_num_seqs, 0, 8192, ())
    if prefix_blocks:
        assert c.enable_prefix_cache(prefix_blocks * BS) == prefix_blocks
    c.enable_cpu_swap(num_cpu)
    return o, c


def _same(o, c):
    bm = o.block_manager
    assert len(o.running) == c.num_running and len(o.waiting) == c.num_waiting
    assert [s.id for s in o.cached] == c.swapped_seq_ids()
    assert list(bm.free_block_ids) == c.free_block_ids()
    assert list(bm.free_cpu_block_ids) == c.free_cpu_block_ids()
    assert [c.block_ref_count(b) for b in range(bm.num_blocks)] == bm.ref_count
    for i, s in enumerate(o.running):
        assert c.block_table(i) == s.block_table
        info = c.seq_info(i)
        assert (info["id"], info["status"], info["len"]) == (s.id, s.status, len(s))
    ops_o = [(d, [tuple(p) for p in pairs]) for d, pairs in bm.swap_ops]
    bm.swap_ops.clear()
    assert c.pop_swap_ops() == ops_o
    return ops_o


def _step(o, c, now, tok=7):
    o.now_ms = now
    c.set_time_ms(now)
    ids_o, pf_o = o.schedule()
    ids_c, pf_c = c.schedule()
    assert (ids_o, pf_o) == (ids_c, pf_c)
    ops = _same(o, c)
    if ids_o:
        if pf_o:
            pos_o, fin_o = o.filter_prefill_finished(ids_o)
            pos_c, fin_c = c.filter_prefill_finished(ids_c)
            assert (pos_o, fin_o) == (pos_c, fin_c)
            o.postprocess(fin_o, [tok] * len(fin_o))
            c.postprocess(fin_c, [tok] * len(fin_c))
        else:
            seqs = [o.running[i] for i in ids_o]
            mo = osched.prepare_decode(seqs, BS)
            mc = c.prepare_decode(ids_c)
            assert list(mc["slot_mapping"]) == mo["slot_mapping"] and mc["block_tables"].tolist() == mo["block_tables"]
            o.postprocess(ids_o, [tok] * len(ids_o))
            c.postprocess(ids_c, [tok] * len(ids_c))
    o.clear_finished()
    c.clear_finished()
    _same(o, c)
    return ids_o, pf_o, ops


def test_hand_trace_swap_out_and_back(eng):
    """8 GPU blocks, two 190-token prompts (3 blocks each, 2 free).  Both reach 193 tokens: sequence 0 takes block 6,
    sequence 1 cannot append (one free block, two wanted) and is skipped.  Next step BOTH cannot append (one free
    block): nothing to evict, so the oldest -- sequence 0 -- is swapped out: copy list GPU [0,1,2,6] -> CPU [0,1,2,3],
    its blocks return in reverse order.  Sequence 1 then runs alone; 5 s later sequence 0 still waits, because
    coming back needs MORE than 4 free blocks with 1000 spare tokens."""
    o, c = _mk(eng, 8, 8)
    for _ in range(2):
        toks = list(range(190))
        o.add(osched.Sequence(toks, BS, max_tokens=60))
        c.add(toks, 60)
    now = 1_000
    ids, pf, _ = _step(o, c, now)
    assert pf and [s.block_table for s in o.running] == [[0, 1, 2], [3, 4, 5]]
    for _ in range(2):                                   # len 191 -> 192 -> 193
        now += 100
        ids, pf, ops = _step(o, c, now)
        assert ids == [0, 1] and not ops
    now += 100
    ids, pf, ops = _step(o, c, now)                      # len 193: seq 0 appends block 6, seq 1 is skipped
    assert ids == [0] and o.running[0].block_table == [0, 1, 2, 6] and not ops
    now += 100
    ids, pf, ops = _step(o, c, now)
    assert ops == [(False, [(0, 0), (1, 1), (2, 2), (6, 3)])]
    assert [s.id for s in o.cached] == [0] and o.cached[0].status == osched.SWAPPED
    assert ids == [0] and o.running[0].id == 1 and o.running[0].block_table == [3, 4, 5, 7]
    assert list(o.block_manager.free_block_ids) == [6, 2, 1, 0]
    assert list(o.block_manager.free_cpu_block_ids) == [4, 5, 6, 7]
    for _ in range(3):
        now += 3_000                                     # well past the cooling period
        ids, pf, ops = _step(o, c, now)
        assert not ops and [s.id for s in o.cached] == [0]


def test_swap_in_gets_fresh_blocks(eng):
    """With room (40 blocks) the swapped sequence returns after the cooling period: fresh blocks from the FRONT of the
    free list, copy list CPU block i -> new GPU block i, CPU blocks back to the END of the CPU free list."""
    o, c = _mk(eng, 40, 12)
    for L in (190, 190):
        toks = list(range(L))
        o.add(osched.Sequence(toks, BS, max_tokens=400))
        c.add(toks, 400)
    now = 0
    _step(o, c, now)
    # exhaust the cache artificially: a long third prompt takes all but one block
    big = list(range(33 * BS - 5))
    o.add(osched.Sequence(big, BS, max_tokens=1))
    c.add(big, 1)
    swapped_at = None
    for i in range(40):
        now += 200
        ids, pf, ops = _step(o, c, now)
        if ops and not ops[0][0]:
            swapped_at = now
            out_pairs = ops[0][1]
            break
    assert swapped_at is not None, "pressure never triggered a swap-out"
    victim = o.cached[0].id
    back = None
    for i in range(80):
        now += 500
        ids, pf, ops = _step(o, c, now)
        if ops and ops[0][0]:
            back = ops[0][1]
            break
    assert back is not None, "the sequence never came back"
    assert now - swapped_at >= osched.SWAP_COOLING_PERIOD
    assert [p[0] for p in back] == [p[1] for p in out_pairs]          # its CPU blocks, in table order
    seq = [s for s in o.running if s.id == victim][0]
    assert [p[1] for p in back] == seq.block_table[:len(back)]         # onto its NEW table
    assert not o.cached and c.num_swapped == 0


@settings(max_examples=40, deadline=None)
@given(st.lists(st.tuples(st.integers(1, 400), st.integers(1, 80), st.integers(0, 6)), min_size=2, max_size=10),
       st.integers(6, 30), st.integers(0, 16), st.integers(1, 6), st.integers(0, 6),
       st.lists(st.integers(50, 4000), min_size=1, max_size=8))
def test_property_swap_traces(lib, reqs, num_blocks, num_cpu, max_seqs, prefix_blocks, dts):
    """Random arrivals under a tight block budget with CPU swap space (and optionally a prefix cache), random clock
    steps around the 5 s cooling period: C++ == oracle after every step (free lists in order, CPU free list, copy
    lists, swapped set); blocks are conserved on both devices."""
    from vllm_rs_b200 import engine
    o, c = _mk(engine, num_blocks, num_cpu, max_num_seqs=max_seqs, prefix_blocks=min(prefix_blocks, num_blocks))
    pending = list(reqs)
    now = 0
    for step in range(300):
        for L, mx, w in [r for r in pending if r[2] == 0]:
            toks = [(L * 31 + j) % 5000 for j in range(L)]
            o.add(osched.Sequence(toks, BS, max_tokens=mx))
            c.add(toks, mx)
        pending = [(L, mx, w - 1) for L, mx, w in pending if w > 0]
        now += dts[step % len(dts)]
        ids, pf, ops = _step(o, c, now)
        bm = o.block_manager
        in_cpu = sum(len(v) for v in bm.swapped_map.values())
        assert in_cpu + len(bm.free_cpu_block_ids) == num_cpu
        assert len(bm.free_block_ids) == sum(1 for x in bm.ref_count if x == 0)
        if not ids and not pending and not o.cached:
            break


def test_swap_copy_lists_preserve_every_sequences_kv():
    """Semantics of the copy lists, with numpy arrays standing in for the GPU cache and the CPU swap space: every token
    writes a marker (seq id, position) at its slot; the copy lists of schedule() are applied exactly as the host
    would (dst[d] = src[s] per pair, runner.rs:1626-1645); at every decode step the pages of each scheduled sequence
    must hold exactly its own history -- across swap-out, reuse of its old blocks by others, and swap-in."""
    rng = np.random.default_rng(0)
    for trial in range(6):
        NB, NCPU = int(rng.integers(10, 22)), int(rng.integers(8, 20))
        o = osched.Scheduler(NB, BS, 6)
        o.block_manager.enable_cpu_swap(NCPU)
        gpu = np.full((NB, BS, 2), -1, dtype=np.int64)
        cpu = np.full((NCPU, BS, 2), -1, dtype=np.int64)
        for _ in range(int(rng.integers(3, 7))):
            L = int(rng.integers(20, 300))
            o.add(osched.Sequence(rng.integers(0, 1000, size=L).tolist(), BS, max_tokens=int(rng.integers(20, 120))))
        now, swaps = 0, 0
        for step in range(600):
            now += int(rng.choice([100, 700, 2600]))
            o.now_ms = now
            ids, pf = o.schedule()
            for swap_in, pairs in o.block_manager.swap_ops:           # the host runs these before the forward
                swaps += 1
                for s_, d_ in pairs:
                    if swap_in:
                        gpu[d_] = cpu[s_]
                    else:
                        cpu[d_] = gpu[s_]
            o.block_manager.swap_ops.clear()
            if not ids:
                if not o.cached and not o.waiting:
                    break
                continue
            seqs = [o.running[i] for i in ids]
            flat = gpu.reshape(NB * BS, 2)
            if pf:
                m = osched.prepare_prefill(seqs, BS, o.chunk_size)
                k = 0
                for s in seqs:
                    n = m["cu_seqlens_q"][seqs.index(s) + 1] - m["cu_seqlens_q"][seqs.index(s)]
                    for j in range(n):
                        flat[m["slot_mapping"][k]] = (s.id, m["positions"][k])
                        k += 1
                pos, fin = o.filter_prefill_finished(ids)
                o.postprocess(fin, [5] * len(fin))
            else:
                m = osched.prepare_decode(seqs, BS)
                for b, s in enumerate(seqs):
                    if s.status == osched.FINISHED:
                        # reference quirk, preserved by the mirror (scheduler.rs:872-876): a sequence that can
                        # never be swapped back in is marked Finished and pushed to `running` INSIDE schedule(),
                        # so it is scheduled once more with its stale block table.  A host must drop such rows
                        # (their slot now belongs to someone else); this simulation does.
                        continue
                    flat[m["slot_mapping"][b]] = (s.id, m["positions"][b])
                    ctx = m["context_lens"][b]
                    pages = gpu[np.asarray(m["block_tables"][b][:-(-ctx // BS)])].reshape(-1, 2)[:ctx]
                    assert (pages[:, 0] == s.id).all() and (pages[:, 1] == np.arange(ctx)).all(), (trial, step, s.id)
                o.postprocess(ids, [5] * len(ids))
            o.clear_finished()
        assert not o.running or step == 599
    assert swaps > 0
