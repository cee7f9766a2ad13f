"""Prefix cache (host side), CPU only and BIT-EXACT.

Three layers of evidence:
  * the chain hash: Rust's `DefaultHasher::new()` (prefix_cache.rs:337-349) is SipHash-1-3 with the
    all-zero key.  `std` is not part of /root/reference, so the oracle restates the published algorithm
    and is pinned against (a) the SipHash paper's 2-4 test vector and (b) CPython's own SipHash-1-3
    (`hash(bytes)` under PYTHONHASHSEED=0 uses the zero key) -- an independent implementation;
  * the reference's own unit tests (prefix_cache.rs:362-403) replayed on the oracle's PrefixCache;
  * C++ (through the C ABI) == oracle on hand-derived and random traces: block tables, cached-token
    counts, reference counts, the ORDER of the free list, cache size, and the prefill metadata
    (context block tables, cu_seqlens_k, slot mapping offset) of prefix hits.
"""
import os
import subprocess
import sys

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


# ------------------------------------------------------------------------------------ hash
def test_siphash24_paper_vector():
    """Aumasson & Bernstein, "SipHash: a fast short-input PRF", appendix A: key 00..0f, input 00..0e."""
    k0 = int.from_bytes(bytes(range(8)), "little")
    k1 = int.from_bytes(bytes(range(8, 16)), "little")
    assert osched.siphash(bytes(range(15)), k0, k1, c_rounds=2, d_rounds=4) == 0xA129CA6149BE45E5


def test_siphash13_zero_key_matches_cpython():
    if sys.hash_info.algorithm != "siphash13":
        pytest.skip(f"this CPython hashes with {sys.hash_info.algorithm}")
    msgs = [bytes(i % 251 for i in range(n)) for n in (1, 7, 8, 9, 15, 16, 31, 64, 272)]
    code = "import sys\nfor h in sys.argv[1:]:\n    print(hash(bytes.fromhex(h)))\n"
    out = subprocess.run([sys.executable, "-c", code] + [m.hex() for m in msgs], capture_output=True, text=True,
                         env={**os.environ, "PYTHONHASHSEED": "0"}, check=True).stdout.split()
    for m, h in zip(msgs, out):
        want = int(h) & osched.M64
        got = osched.siphash(m)                  # zero key, 1-3
        # CPython maps the value -1 to -2; no other adjustment
        assert got == want or (got == osched.M64 and want == osched.M64 - 1), (m.hex(), hex(got), hex(want))


def test_hash_block_cpp_matches_oracle(lib):
    rng = np.random.default_rng(3)
    for n in (1, 4, 63, 64, 65):
        for parent in (0, 1, 0xDEADBEEFCAFEF00D, osched.M64):
            toks = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
            got = lib.b200_prefix_hash_block(parent, toks.ctypes.data, n)
            assert got == osched.hash_block(parent, toks.tolist())


# ------------------------------------------------------- the reference's own unit tests
def test_reference_prefix_cache_matches_full_blocks():
    """prefix_cache.rs:366-384."""
    cache = osched.PrefixCache(4, 8)
    inserted, evicted = cache.insert_prefix([1, 2, 3, 4, 5, 6, 7, 8], [10, 11])
    assert evicted == [] and len(inserted) == 2
    matched, last = cache.match_prefix([1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12])
    assert matched == 2
    assert cache.blocks_for_match(last) == [10, 11]


def test_reference_prefix_cache_evicts_leaf_blocks():
    """prefix_cache.rs:386-403."""
    cache = osched.PrefixCache(4, 1)
    tokens = [1, 2, 3, 4, 5, 6, 7, 8]
    inserted, evicted = cache.insert_prefix(tokens, [21, 22])
    assert evicted == [22]
    assert cache.match_prefix(tokens)[0] == 1


# ------------------------------------------------------------- scheduler traces, C++ == oracle
def _mk(eng, num_blocks, max_tokens=None, **kw):
    kw.setdefault("max_num_seqs", 32)
    cap = osched.prefix_cache_max_blocks(num_blocks, BS, max_tokens)
    o = osched.Scheduler(num_blocks, BS, kw["max_num_seqs"], None, kw.get("chunk_size", 8192),
                         kw.get("eos_token_ids", ()), prefix_cache_blocks=cap)
    c = eng.Scheduler(num_blocks, BS, kw["max_num_seqs"], 0, kw.get("chunk_size", 8192),
                      kw.get("eos_token_ids", ()))
    assert c.enable_prefix_cache(max_tokens) == cap
    return o, c


def _same_state(o, c):
    bm = o.block_manager
    assert len(o.running) == c.num_running and len(o.waiting) == c.num_waiting
    assert list(bm.free_block_ids) == c.free_block_ids()              # same ids in the same ORDER
    assert bm.prefix_cache_blocks() == c.prefix_cache_blocks
    assert [c.block_ref_count(b) for b in range(bm.num_blocks)] == bm.ref_count
    for i, s in enumerate(o.running):
        assert c.block_table(i) == s.block_table
        info = c.seq_info(i)
        assert (info["id"], info["num_cached_tokens"], info["status"]) == (s.id, s.num_cached_tokens, s.status)


def _step(o, c, token_fn):
    ids_o, pf_o = o.schedule()
    ids_c, pf_c = c.schedule()
    assert (ids_o, pf_o) == (ids_c, pf_c)
    if not ids_o:
        _same_state(o, c)
        return None
    seqs = [o.running[i] for i in ids_o]
    if pf_o:
        mo = osched.prepare_prefill(seqs, BS, o.chunk_size)
        mc = c.prepare_prefill(ids_c)
        for k in ("input_ids", "positions", "slot_mapping", "cu_seqlens_q", "cu_seqlens_k"):
            assert list(mc[k]) == mo[k], k
        for k in ("max_seqlen_q", "max_seqlen_k", "max_context_len"):
            assert mc[k] == mo[k], k
        if mo["block_tables"] is None:
            assert mc["block_tables"] is None
        else:
            assert mc["block_tables"].tolist() == mo["block_tables"]
            assert list(mc["context_lens"]) == mo["context_lens"]
        pos_o, fin_o = o.filter_prefill_finished(ids_o)
        pos_c, fin_c = c.filter_prefill_finished(ids_c)
        assert (pos_o, fin_o) == (pos_c, fin_c)
        toks = [token_fn(o.running[i]) for i in fin_o]
        o.postprocess(fin_o, toks)
        c.postprocess(fin_c, toks)
    else:
        mo = osched.prepare_decode(seqs, BS)
        mc = c.prepare_decode(ids_c)
        for k in ("input_ids", "positions", "slot_mapping", "context_lens"):
            assert list(mc[k]) == mo[k], k
        assert mc["block_tables"].tolist() == mo["block_tables"]
        toks = [token_fn(s) for s in seqs]
        o.postprocess(ids_o, toks)
        c.postprocess(ids_c, toks)
    o.clear_finished()
    c.clear_finished()
    _same_state(o, c)
    return pf_o, mo


def _add(o, c, toks, max_tokens):
    o.add(osched.Sequence(list(toks), BS, max_tokens=max_tokens))
    c.add(list(toks), max_tokens)


def _drain(o, c, token_fn=lambda s: 11, limit=2000):
    n = 0
    while (o.running or o.waiting) and n < limit:
        if _step(o, c, token_fn) is None:
            break
        n += 1
    return n


def test_hand_trace_hit_shares_blocks(eng):
    """A 130-token prompt finishes; its two FULL blocks stay in the cache with one reference each
    (cache_sequence before deallocate, scheduler.rs:617-620); the partial third block returns to the
    BACK of the free list.  A second prompt with the same first 130 tokens reuses blocks [0, 1]
    (allocate_with_prefix, block_manager.rs:346-442) and prefills only from token 128 on."""
    o, c = _mk(eng, 16, max_tokens=16 * BS)
    a = list(range(1000, 1130))
    _add(o, c, a, 1)
    pf, m = _step(o, c, lambda s: 5)                       # prefill: nothing cached yet
    assert pf and m["block_tables"] is None and o.running[0].block_table == [0, 1, 2]
    pf, _ = _step(o, c, lambda s: 5)                       # decode; output_len 1 >= max_tokens -> finished
    assert not pf and not o.running
    bm = o.block_manager
    assert bm.ref_count[:3] == [1, 1, 0] and bm.prefix_cache_blocks() == 2
    assert list(bm.free_block_ids) == list(range(3, 16)) + [2]

    b = a + list(range(7000, 7100))                        # 230 tokens: 3 full blocks + 38
    _add(o, c, b, 2)
    pf, m = _step(o, c, lambda s: 6)
    s = None
    assert pf and m["block_tables"] == [[0, 1, 3, 4]]      # shared prefix blocks, then fresh from the front
    assert m["context_lens"] == [230] and m["cu_seqlens_q"] == [0, 102] and m["cu_seqlens_k"] == [0, 230]
    assert m["positions"] == list(range(128, 230)) and m["input_ids"] == b[128:]
    assert m["slot_mapping"] == list(range(3 * 64, 3 * 64 + 64)) + list(range(4 * 64, 4 * 64 + 38))
    assert o.running[0].num_cached_tokens == 128 and bm.ref_count[:2] == [2, 2]
    _drain(o, c, lambda s: 6)
    # b finished at 232 tokens: block 3 (tokens 128..191) joins the chain; blocks 0, 1 are already cached
    assert bm.prefix_cache_blocks() == 3 and bm.ref_count[:5] == [1, 1, 0, 1, 0]
    assert c.prefix_cache_blocks == 3


def test_never_reuses_whole_prompt(eng):
    """block_manager.rs:291-299: a prompt that is entirely cached full blocks keeps its last block
    out of the match, so at least one token is prefilled."""
    o, c = _mk(eng, 16, max_tokens=16 * BS)
    a = list(range(128))
    _add(o, c, a + [9], 1)
    _drain(o, c)
    assert o.block_manager.prefix_cache_blocks() == 2
    _add(o, c, a, 1)
    pf, m = _step(o, c, lambda s: 3)
    assert pf and o.running[0].num_cached_tokens == 64
    assert m["cu_seqlens_q"] == [0, 64] and m["cu_seqlens_k"] == [0, 128]
    assert m["block_tables"][0][0] == 0 and m["block_tables"][0][1] != 1


def test_capacity_eviction_is_leaf_lru(eng):
    """Capacity 2 blocks: inserting a 3-block chain evicts its own newest leaf (prefix_cache.rs:248-257);
    a different prompt then evicts the older chain leaf-first."""
    o, c = _mk(eng, 16, max_tokens=2 * BS)
    _add(o, c, list(range(200)), 1)                        # 3 full blocks at finish (201 tokens)
    _drain(o, c)
    bm = o.block_manager
    assert bm.prefix_cache_blocks() == 2 and bm.ref_count[:4] == [1, 1, 0, 0]
    _add(o, c, list(range(5000, 5070)), 1)                 # 1 full block: evicts leaf (block 1)
    _drain(o, c)
    assert bm.prefix_cache_blocks() == 2 and bm.ref_count[0] == 1 and bm.ref_count[1] == 0


def test_eviction_under_pressure(eng):
    """scheduler.rs:288-337,1167-1175: a running sequence that cannot append makes schedule() drop
    ceil(10 %) of the cached blocks (at least one) BEFORE the decode batch is picked, so the freed
    block is usable in the same step."""
    o, c = _mk(eng, 6, max_tokens=6 * BS)
    bm = o.block_manager
    _add(o, c, list(range(100, 100 + 129)), 1)             # blocks 0,1,2; finishes at 130 tokens, 2 cached
    _drain(o, c)
    assert bm.prefix_cache_blocks() == 2 and list(bm.free_block_ids) == [3, 4, 5, 2]
    _add(o, c, list(range(9000, 9000 + 255)), 4)           # 4 fresh blocks -> nothing free
    pf, _ = _step(o, c, lambda s: 8)
    assert pf and o.running[0].block_table == [3, 4, 5, 2] and not bm.free_block_ids
    # len 256: appending needs one free block -> pressure evicts the LRU leaf (block 1); decode goes ahead
    pf, m = _step(o, c, lambda s: 8)
    assert not pf and m["context_lens"] == [256]
    assert bm.prefix_cache_blocks() == 1 and list(bm.free_block_ids) == [1]
    # len 257: needs two free blocks -> evicts block 0 as well; may_append then pops block 1
    pf, m = _step(o, c, lambda s: 8)
    assert not pf and bm.prefix_cache_blocks() == 0 and m["block_tables"] == [[3, 4, 5, 2, 1]]
    assert list(bm.free_block_ids) == [0]
    # len 258 wants two free blocks for the reference's can_append rule (block_manager.rs:233-239) and
    # the cache is empty: the sequence is filtered out (the reference would now try CPU swap: out of scope)
    assert _step(o, c, lambda s: 8) is None and len(o.running) == 1


def test_explicit_eviction_and_disabled(eng):
    o, c = _mk(eng, 16, max_tokens=16 * BS)
    _add(o, c, list(range(300)), 1)
    _drain(o, c)
    assert c.prefix_cache_blocks == 4
    assert c.evict_prefix_cache_blocks(3) == o.block_manager.evict_prefix_cache_blocks(3) == 3
    _same_state(o, c)
    # capacity 0 -> disabled exactly as build_prefix_cache_config does (scheduler.rs:96-102)
    c2 = eng.Scheduler(16, BS, 4, 0, 8192, ())
    assert c2.enable_prefix_cache(10) == 0
    c2.add(list(range(200)), 1)
    c2.schedule()
    assert c2.prefix_cache_blocks == 0


@settings(max_examples=40, deadline=None)
@given(st.lists(st.tuples(st.integers(0, 3), st.integers(1, 330), st.integers(1, 6), st.integers(0, 3)),
                min_size=2, max_size=14),
       st.integers(10, 40), st.integers(1, 8), st.integers(0, 12))
def test_property_shared_prefix_traces(lib, reqs, num_blocks, max_seqs, cap_blocks):
    """Random arrivals drawn from 4 prompt families (shared prefixes of varying length), arriving in
    waves so hits, capacity evictions and pressure evictions all occur; C++ == oracle after every
    step, and the cache's references are consistent: ref_count = live tables + cache membership."""
    from vllm_rs_b200 import engine
    o, c = _mk(engine, num_blocks, max_tokens=cap_blocks * BS, max_num_seqs=max_seqs)
    fam = [[(f * 7919 + j * 13) % 30000 for j in range(330)] for f in range(4)]
    pending = list(reqs)
    for _ in range(600):
        for f, L, mx, wave in [r for r in pending if r[3] == 0]:
            _add(o, c, fam[f][:L], mx)
        pending = [(f, L, mx, w - 1) for f, L, mx, w in pending if w > 0]
        r = _step(o, c, lambda s: 11)
        bm = o.block_manager
        live = {}
        for s in o.running + list(o.waiting):
            for b in s.block_table:
                live[b] = live.get(b, 0) + 1
        cached = {e[1] for e in bm.prefix_cache.entries.values()} if bm.prefix_cache else set()
        for b in range(num_blocks):
            assert bm.ref_count[b] == live.get(b, 0) + (1 if b in cached else 0)
        assert len(bm.free_block_ids) == sum(1 for x in bm.ref_count if x == 0)
        if r is None and not pending:
            break


# ------------------------------------------------------------------------------------ committed golden vectors
def test_golden_hash_chain_and_trace(eng, lib):
    """tests/golden/prefix_cache.json (made by tests/golden/make_prefix_golden.py from the pinned oracle): the chain
    hashes and a three-wave scheduler trace with prefix hits must be reproduced by the oracle AND by the C++
    mirror -- a drift of both at once would otherwise go unnoticed."""
    import json
    from pathlib import Path
    doc = json.loads((Path(__file__).parent / "golden" / "prefix_cache.json").read_text())
    toks = doc["tokens"]
    for rec in doc["hash_chain"]:
        parent, n, want = int(rec["parent"]), rec["n"], int(rec["hash"])
        assert osched.hash_block(parent, toks[:n]) == want
        arr = np.asarray(toks[:n], dtype=np.uint32)
        assert lib.b200_prefix_hash_block(parent, arr.ctypes.data, n) == want
    # replay the trace on both sides
    o, c = _mk(eng, 24, max_num_seqs=8)
    base = [(i * 7919 + 13) % 32000 for i in range(400)]
    waves = [[(base[:150], 3)], [(base[:150] + [5, 6, 7], 2), (base[:128], 2), (base[:70] + base[200:300], 4)],
             [(base[:330], 2), (base[:64], 1)]]
    it = iter(doc["trace"])
    steps = 0
    for wave in waves:
        for t, mx in wave:
            _add(o, c, t, mx)
        for _ in range(64):
            if not (o.running or o.waiting):
                break
            ids, pf = o.schedule()
            if not ids:
                c.schedule()
                break
            seqs_before = [(o.running[i].id, list(o.running[i].block_table), o.running[i].num_cached_tokens) for i in ids]
            cids, cpf = c.schedule()
            rec = next(it)
            assert (ids, pf) == (cids, cpf) == (rec["ids"], rec["prefill"])
            assert [s[0] for s in seqs_before] == rec["seq_ids"]
            assert [s[1] for s in seqs_before] == rec["block_tables"] == [c.block_table(i) for i in cids]
            assert [s[2] for s in seqs_before] == rec["num_cached_tokens"]
            seqs = [o.running[i] for i in ids]
            if pf:
                m = osched.prepare_prefill(seqs, BS, o.chunk_size)
                mc = c.prepare_prefill(cids)
                assert m["cu_seqlens_q"] == rec["cu_seqlens_q"] == list(mc["cu_seqlens_q"])
                assert m["cu_seqlens_k"] == rec["cu_seqlens_k"] == list(mc["cu_seqlens_k"])
                assert m["slot_mapping"][:4] == rec["slot_head"] == list(mc["slot_mapping"][:4])
                pos, fin = o.filter_prefill_finished(ids)
                c.filter_prefill_finished(cids)
                o.postprocess(fin, [11] * len(fin))
                c.postprocess(fin, [11] * len(fin))
            else:
                m = osched.prepare_decode(seqs, BS)
                mc = c.prepare_decode(cids)
                assert m["slot_mapping"] == rec["slot_mapping"] == list(mc["slot_mapping"])
                o.postprocess(ids, [11] * len(ids))
                c.postprocess(cids, [11] * len(ids))
            o.clear_finished()
            c.clear_finished()
            assert list(o.block_manager.free_block_ids) == rec["free_block_ids"] == c.free_block_ids()
            assert o.block_manager.ref_count == rec["ref_count"]
            assert o.block_manager.prefix_cache_blocks() == rec["cached_blocks"] == c.prefix_cache_blocks
            steps += 1
    assert steps == len(doc["trace"]) and steps >= 10
