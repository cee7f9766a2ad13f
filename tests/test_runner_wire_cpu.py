"""Engine <-> runner frames of the decode step (runner/mod.rs:168-295): native code vs the oracle restatement of
bincode 1.3.1 and vs a RunDecode frame written out by hand; the decode metadata a runner derives from it
(core/runner.rs:1243-1264) vs the scheduler mirror's own prepare_decode.  CPU only."""
import ctypes as C
import struct

import numpy as np

from oracle import sched as osched
from oracle import transfer as ot


def _hand_run_decode():
    """RunDecode(([DecodeSequence{id 3, last_token 77, len 130, last_block_tokens 2, block_table_last 9,
    block_tables [4, 5, 9], SamplingParams{temperature Some(0.5), max_tokens Some(16), ignore_eos false, top_k Some(-1),
    everything else None}}], false)) byte by byte."""
    b = bytes([4, 0, 0, 0])                                          # MessageType variant 4
    b += bytes([1, 0, 0, 0, 0, 0, 0, 0])                             # Vec<DecodeSequence> len 1
    b += bytes([3, 0, 0, 0, 0, 0, 0, 0])                             # id: usize
    b += bytes([77, 0, 0, 0])                                        # last_token: u32
    b += bytes([130, 0, 0, 0, 0, 0, 0, 0])                           # len
    b += bytes([2, 0, 0, 0, 0, 0, 0, 0])                             # last_block_tokens
    b += bytes([9, 0, 0, 0])                                         # block_table_last
    b += bytes([3, 0, 0, 0, 0, 0, 0, 0]) + bytes([4, 0, 0, 0, 5, 0, 0, 0, 9, 0, 0, 0])   # block_tables
    b += bytes([1]) + struct.pack("<f", 0.5)                         # temperature Some(0.5)
    b += bytes([1, 16, 0, 0, 0, 0, 0, 0, 0])                         # max_tokens Some(16)
    b += bytes([0])                                                  # ignore_eos
    b += bytes([1]) + bytes([0xFF] * 8)                              # top_k Some(-1): isize
    b += bytes([0, 0, 0, 0, 0])                                      # top_p, session_id, frequency, presence, stop_sequences
    b += bytes([0, 0, 0, 0, 0])                                      # thinking, mcp_mode, grammar, grammar_json, reasoning_effort
    b += bytes([0])                                                  # is_prefill
    return struct.pack("<I", len(b)) + b


def _seq(i, ltok, ln, bt, **sp):
    return dict(id=i, last_token=ltok, len=ln, last_block_tokens=(ln - 1) % 64 + 1, block_table_last=bt[-1],
                block_tables=bt, sampling=ot.sampling_params(**sp))


def test_oracle_matches_the_hand_written_run_decode():
    q = dict(id=3, last_token=77, len=130, last_block_tokens=2, block_table_last=9, block_tables=[4, 5, 9],
             sampling=ot.sampling_params(temperature=0.5, max_tokens=16, top_k=-1))
    assert ot.run_decode_frame([q], False) == _hand_run_decode()
    assert ot.run_response_frame([7, 9]) == bytes([20, 0, 0, 0, 5, 0, 0, 0, 2, 0, 0, 0, 0, 0, 0, 0, 7, 0, 0, 0, 9, 0, 0, 0])
    assert ot.simple_frame(ot.MT_SHUTDOWN) == bytes([4, 0, 0, 0, 34, 0, 0, 0])


def test_native_run_decode_round_trip(lib):
    from vllm_rs_b200 import _lib
    rng = np.random.default_rng(1)
    seqs = [_seq(11, 5, 130, [4, 5, 9], temperature=0.5, max_tokens=16, top_k=-1),
            _seq(12, 31999, 64, [7], session_id=b"abc", stop_sequences=[b"</s>", b"\n\n"], thinking=1,
                 grammar=b'{"grammars":[]}', reasoning_effort=2, top_p=0.9, frequency_penalty=0.25, presence_penalty=-0.5),
            _seq(13, 1, 1, [0], ignore_eos=True)]
    for _ in range(5):
        nb = int(rng.integers(1, 40))
        seqs.append(_seq(int(rng.integers(0, 1 << 40)), int(rng.integers(0, 128256)), nb * 64 - int(rng.integers(0, 64)),
                         rng.integers(0, 5000, nb).tolist()))
    fr = np.frombuffer(ot.run_decode_frame(seqs, False), np.uint8).copy()
    MS, MB = 16, 64
    ids, lt, ln, lbt = np.zeros(MS, np.uint64), np.zeros(MS, np.uint32), np.zeros(MS, np.uint64), np.zeros(MS, np.uint64)
    btl, bt, nblk = np.zeros(MS, np.uint32), np.full((MS, MB), 99, np.uint32), np.zeros(MS, np.int32)
    samp = (_lib.WireSampling * MS)()
    pf = C.c_int32(7)
    n = lib.b200_wire_parse_run_decode(fr.ctypes.data, len(fr), MS, MB, ids.ctypes.data, lt.ctypes.data, ln.ctypes.data,
                                       lbt.ctypes.data, btl.ctypes.data, bt.ctypes.data, nblk.ctypes.data, samp,
                                       C.byref(pf))
    assert n == len(seqs) and pf.value == 0
    assert lib.b200_wire_kind(fr.ctypes.data, len(fr)) == ot.MT_RUN_DECODE
    for i, q in enumerate(seqs):
        assert (ids[i], lt[i], ln[i], lbt[i], btl[i]) == (q["id"], q["last_token"], q["len"], q["last_block_tokens"],
                                                          q["block_table_last"])
        assert bt[i, :nblk[i]].tolist() == q["block_tables"] and not bt[i, nblk[i]:].any()
    assert (samp[0].has_temperature, samp[0].temperature, samp[0].has_max_tokens, samp[0].max_tokens,
            samp[0].has_top_k, samp[0].top_k, samp[0].ignore_eos) == (1, 0.5, 1, 16, 1, -1, 0)
    assert (samp[1].has_top_p, round(samp[1].top_p, 6), samp[1].frequency_penalty, samp[1].presence_penalty,
            samp[1].has_temperature) == (1, 0.9, 0.25, -0.5, 0)
    assert samp[2].ignore_eos == 1
    # the engine side: the native writer reproduces the oracle frame when the string fields are None
    plain = [dict(q, sampling=ot.sampling_params(temperature=0.7, top_k=40, top_p=0.95, max_tokens=128)) for q in seqs]
    want = ot.run_decode_frame(plain, True)
    ws = (_lib.WireSampling * MS)()
    for i in range(len(plain)):
        ws[i].has_temperature, ws[i].temperature = 1, 0.7
        ws[i].has_top_k, ws[i].top_k, ws[i].has_top_p, ws[i].top_p = 1, 40, 1, 0.95
        ws[i].has_max_tokens, ws[i].max_tokens = 1, 128
    out = np.zeros(len(want) + 8, np.uint8)
    got = lib.b200_wire_frame_run_decode(out.ctypes.data, len(out), len(plain), ids.ctypes.data, lt.ctypes.data,
                                         ln.ctypes.data, lbt.ctypes.data, btl.ctypes.data, bt.ctypes.data, MB,
                                         nblk.ctypes.data, ws, 1)
    assert got == len(want) and out[:got].tobytes() == want
    # malformed input: truncated, too many sequences, a block table longer than the caller's row
    assert lib.b200_wire_parse_run_decode(fr.ctypes.data, len(fr) - 1, MS, MB, ids.ctypes.data, lt.ctypes.data,
                                          ln.ctypes.data, lbt.ctypes.data, btl.ctypes.data, bt.ctypes.data,
                                          nblk.ctypes.data, samp, C.byref(pf)) == -1
    assert lib.b200_wire_parse_run_decode(fr.ctypes.data, len(fr), 3, MB, ids.ctypes.data, lt.ctypes.data,
                                          ln.ctypes.data, lbt.ctypes.data, btl.ctypes.data, bt.ctypes.data,
                                          nblk.ctypes.data, samp, C.byref(pf)) == -2
    assert lib.b200_wire_parse_run_decode(fr.ctypes.data, len(fr), MS, 2, ids.ctypes.data, lt.ctypes.data,
                                          ln.ctypes.data, lbt.ctypes.data, btl.ctypes.data, bt.ctypes.data,
                                          nblk.ctypes.data, samp, C.byref(pf)) == -3
    hand = np.frombuffer(_hand_run_decode(), np.uint8).copy()
    assert lib.b200_wire_parse_run_decode(hand.ctypes.data, len(hand), MS, MB, ids.ctypes.data, lt.ctypes.data,
                                          ln.ctypes.data, lbt.ctypes.data, btl.ctypes.data, bt.ctypes.data,
                                          nblk.ctypes.data, samp, C.byref(pf)) == 1
    assert (ids[0], lt[0], ln[0], lbt[0], btl[0], samp[0].top_k) == (3, 77, 130, 2, 9, -1)
    lib.b200_clear_error()


def test_native_small_frames_and_kv_swap(lib):
    ids = np.array([5, 128255, 0, 42], np.uint32)
    want = ot.run_response_frame(ids.tolist())
    out = np.zeros(64, np.uint8)
    assert lib.b200_wire_frame_run_response(out.ctypes.data, 64, ids.ctypes.data, 4) == len(want)
    assert out[:len(want)].tobytes() == want
    back = np.zeros(8, np.uint32)
    assert lib.b200_wire_parse_run_response(out.ctypes.data, len(want), back.ctypes.data, 8) == 4
    assert back[:4].tolist() == ids.tolist()
    pairs = np.array([[3, 0], [17, 1], [5, 2]], np.int64)
    want = ot.kv_swap_frame(pairs.tolist(), True)
    out = np.zeros(128, np.uint8)
    assert lib.b200_wire_frame_kv_swap(out.ctypes.data, 128, pairs.ctypes.data, 3, 1) == len(want)
    assert out[:len(want)].tobytes() == want
    got, flag = np.zeros((4, 2), np.int64), C.c_int32(0)
    assert lib.b200_wire_parse_kv_swap(out.ctypes.data, len(want), got.ctypes.data, 4, C.byref(flag)) == 3
    assert got[:3].tolist() == pairs.tolist() and flag.value == 1
    for kind, nb, val, payload in ((ot.MT_FINISH_DECODE, 8, 123456789, struct.pack("<Q", 123456789)),
                                   (ot.MT_KV_SWAP_RESPONSE, 1, 1, b"\x01"), (ot.MT_HEARTBEAT, 0, 0, b""),
                                   (ot.MT_SHUTDOWN, 0, 0, b"")):
        w = ot.simple_frame(kind, payload)
        assert lib.b200_wire_frame_simple(out.ctypes.data, 128, kind, nb, val) == len(w)
        assert out[:len(w)].tobytes() == w and lib.b200_wire_kind(out.ctypes.data, len(w)) == kind
    lib.b200_clear_error()


def test_decode_metadata_from_the_wire_equals_prepare_decode(lib):
    """A runner that only sees DecodeSequence must derive the same step metadata as Scheduler-side prepare_decode."""
    sch = osched.Scheduler(64, 64, 8)
    rng = np.random.default_rng(2)
    for L in (37, 64, 65, 1, 130):
        sch.add(osched.Sequence(rng.integers(0, 1000, size=L).tolist(), 64, max_tokens=8))
    ids, pf = sch.schedule()
    assert pf
    _, fin = sch.filter_prefill_finished(ids)
    sch.postprocess(fin, rng.integers(0, 1000, size=len(fin)).tolist())
    ids, pf = sch.schedule()
    assert not pf
    seqs = [sch.running[i] for i in ids]
    meta = osched.prepare_decode(seqs, 64)
    # DecodeSequence::new (sequence.rs:64-78)
    wire = [dict(id=q.id, last_token=q.last_token, len=len(q), last_block_tokens=q.last_block_num_tokens(),
                 block_table_last=q.block_table[-1], block_tables=list(q.block_table)) for q in seqs]
    assert ot.decode_metadata(wire) == dict(input_ids=meta["input_ids"], positions=meta["positions"],
                                            context_lens=meta["context_lens"], slot_mapping=meta["slot_mapping"])
    n = len(wire)
    lt = np.array([q["last_token"] for q in wire], np.uint32)
    ln = np.array([q["len"] for q in wire], np.uint64)
    lbt = np.array([q["last_block_tokens"] for q in wire], np.uint64)
    btl = np.array([q["block_table_last"] for q in wire], np.uint32)
    ii, pos, slot, cl = np.zeros(n, np.uint32), np.zeros(n, np.int64), np.zeros(n, np.int64), np.zeros(n, np.uint32)
    lib.b200_wire_decode_metadata(n, lt.ctypes.data, ln.ctypes.data, lbt.ctypes.data, btl.ctypes.data, 64,
                                  ii.ctypes.data, pos.ctypes.data, slot.ctypes.data, cl.ctypes.data)
    assert (ii.tolist(), pos.tolist(), slot.tolist(), cl.tolist()) == (meta["input_ids"], meta["positions"],
                                                                       meta["slot_mapping"], meta["context_lens"])


def test_run_decode_frames_property(lib):
    """Random batches: the oracle's bincode bytes parse back to the same fields, and the native writer reproduces the
    oracle's bytes (string-valued sampling fields None, as the engine sends them for plain requests)."""
    from hypothesis import given, settings, strategies as st
    from vllm_rs_b200 import _lib

    seq = st.tuples(st.integers(0, (1 << 63) - 1), st.integers(0, (1 << 32) - 1), st.integers(1, 1 << 20),
                    st.lists(st.integers(0, (1 << 32) - 1), min_size=1, max_size=24),
                    st.one_of(st.none(), st.floats(0, 2, width=32)), st.one_of(st.none(), st.integers(-1, 1 << 20)),
                    st.one_of(st.none(), st.floats(0, 1, width=32)), st.one_of(st.none(), st.integers(0, 1 << 40)),
                    st.booleans())

    @settings(max_examples=60, deadline=None)
    @given(st.lists(seq, min_size=0, max_size=12), st.booleans())
    def check(batch, is_prefill):
        seqs = [dict(id=i, last_token=tok, len=ln, last_block_tokens=(ln - 1) % 64 + 1, block_table_last=bt[-1],
                     block_tables=bt, sampling=ot.sampling_params(temperature=t, top_k=k, top_p=p, max_tokens=mt,
                                                                  ignore_eos=ie))
                for i, tok, ln, bt, t, k, p, mt, ie in batch]
        want = ot.run_decode_frame(seqs, is_prefill)
        fr = np.frombuffer(want, np.uint8).copy()
        MS, MB = 12, 24
        ids, lt, ln_, lbt = np.zeros(MS, np.uint64), np.zeros(MS, np.uint32), np.zeros(MS, np.uint64), np.zeros(MS, np.uint64)
        btl, bt_, nblk = np.zeros(MS, np.uint32), np.zeros((MS, MB), np.uint32), np.zeros(MS, np.int32)
        samp = (_lib.WireSampling * MS)()
        pf = C.c_int32(-1)
        n = lib.b200_wire_parse_run_decode(fr.ctypes.data, len(fr), MS, MB, ids.ctypes.data, lt.ctypes.data,
                                           ln_.ctypes.data, lbt.ctypes.data, btl.ctypes.data, bt_.ctypes.data,
                                           nblk.ctypes.data, samp, C.byref(pf))
        assert n == len(seqs) and pf.value == int(is_prefill)
        for i, q in enumerate(seqs):
            assert (int(ids[i]), int(lt[i]), int(ln_[i]), int(btl[i])) == (q["id"], q["last_token"], q["len"],
                                                                          q["block_table_last"])
            assert bt_[i, :nblk[i]].tolist() == q["block_tables"]
        out = np.zeros(len(want) + 16, np.uint8)
        got = lib.b200_wire_frame_run_decode(out.ctypes.data, len(out), n, ids.ctypes.data, lt.ctypes.data,
                                             ln_.ctypes.data, lbt.ctypes.data, btl.ctypes.data, bt_.ctypes.data, MB,
                                             nblk.ctypes.data, samp, int(is_prefill))
        assert got == len(want) and out[:got].tobytes() == want       # parse -> frame is the identity on these frames

    check()
    lib.b200_clear_error()


# ---------------------------------------------------------------- RunPrefill((Vec<Sequence>, bool))
def _hand_run_prefill():
    """RunPrefill(([Sequence{id 7, created_time 0, swapped_time None, status Running, token_ids [11, 12, 13],
    output_ids [], block_table [5], num_cached_tokens 0, mamba_prefix_hash None, last_token 13, block_size 64,
    SamplingParams{max_tokens Some(9), rest None / false}, pd_first_token None, images None, false, false, None}],
    true)) byte by byte."""
    b = bytes([3, 0, 0, 0]) + bytes([1, 0, 0, 0, 0, 0, 0, 0])        # variant 3, one sequence
    b += bytes([7, 0, 0, 0, 0, 0, 0, 0]) + bytes(8) + bytes([0])     # id, created_time, swapped_time None
    b += bytes([1, 0, 0, 0])                                         # status: Running
    b += bytes([3, 0, 0, 0, 0, 0, 0, 0]) + bytes([11, 0, 0, 0, 12, 0, 0, 0, 13, 0, 0, 0])   # token_ids
    b += bytes(8)                                                    # output_ids: empty
    b += bytes([1, 0, 0, 0, 0, 0, 0, 0]) + bytes([5, 0, 0, 0])       # block_table
    b += bytes(8) + bytes([0])                                       # num_cached_tokens, mamba_prefix_hash None
    b += bytes([13, 0, 0, 0]) + bytes([64, 0, 0, 0, 0, 0, 0, 0])     # last_token, block_size
    b += bytes([0]) + bytes([1, 9, 0, 0, 0, 0, 0, 0, 0]) + bytes([0])   # temperature None, max_tokens Some(9), ignore_eos
    b += bytes(11)                                                   # top_k .. reasoning_effort: eleven None tags
    b += bytes([0, 0, 0, 0, 0])                                      # pd_first_token, images, two flags, stop_sequence
    b += bytes([1])                                                  # is_prefill
    return struct.pack("<I", len(b)) + b


def test_run_prefill_frame_golden_and_round_trip(lib):
    from vllm_rs_b200.engine import Scheduler
    q = dict(id=7, status=1, token_ids=[11, 12, 13], output_ids=[], block_table=[5], num_cached_tokens=0,
             last_token=13, block_size=64, sampling=ot.sampling_params(max_tokens=9))
    assert ot.run_prefill_frame([q], True) == _hand_run_prefill()
    # engine side: three prompts scheduled for prefill, framed by the native scheduler == the oracle's bytes
    rng = np.random.default_rng(3)
    eng = Scheduler(64, 64, 8)
    prompts = [rng.integers(0, 1000, size=L).tolist() for L in (37, 64, 130)]
    for t in prompts:
        eng.add(t, max_tokens=12)
    ids, pf = eng.schedule()
    assert pf and len(ids) == 3
    fr = eng.frame_run_prefill(ids, True)
    want = ot.run_prefill_frame(
        [dict(id=eng.seq_info(i)["id"], status=1, token_ids=eng.tokens(i), output_ids=[], block_table=eng.block_table(i),
              num_cached_tokens=0, last_token=eng.tokens(i)[-1], block_size=64,
              sampling=ot.sampling_params(max_tokens=12)) for i in ids], True)
    assert fr == want
    assert lib.b200_wire_kind(np.frombuffer(fr, np.uint8).ctypes.data, len(fr)) == ot.MT_RUN_PREFILL
    # runner side: a fresh container loads the frame and derives the SAME step metadata
    run = Scheduler(64, 64, 8)
    rids, rpf, sp = run.load_run_prefill(fr)
    assert rpf and rids == [0, 1, 2] and (sp.has_max_tokens, sp.max_tokens, sp.has_temperature) == (1, 12, 0)
    a, b = eng.prepare_prefill(ids), run.prepare_prefill(rids)
    for k in a:
        if a[k] is None:
            assert b[k] is None
        else:
            assert np.array_equal(np.asarray(a[k]), np.asarray(b[k])), k
    # the oracle's frame of a sequence with fields the mirror ignores (session id, stop sequences, swapped_time, a
    # second chunk with cached tokens) still loads, and the metadata is the oracle's prepare_prefill
    sch = osched.Scheduler(64, 64, 8)
    sch.add(osched.Sequence(rng.integers(0, 1000, size=200).tolist(), 64, max_tokens=4))
    oids, _ = sch.schedule()
    s0 = sch.running[oids[0]]
    s0.num_cached_tokens = 128                                       # as after a 128-token chunk / prefix hit
    ofr = ot.run_prefill_frame([dict(
        id=s0.id, created_time=1700000000, swapped_time=55, status=1, token_ids=list(s0.token_ids), output_ids=[],
        block_table=list(s0.block_table), num_cached_tokens=128, mamba_prefix_hash=99, last_token=s0.last_token,
        block_size=64, pd_first_token=4, stop_sequence=b"END", hit_stop_sequence=True,
        sampling=ot.sampling_params(temperature=0.5, session_id=b"s", stop_sequences=[b"a", b"bc"], top_k=5))], True)
    rids, _, sp = run.load_run_prefill(ofr)
    meta = run.prepare_prefill(rids)
    want_meta = osched.prepare_prefill([s0], 64)
    for k in ("input_ids", "positions", "slot_mapping", "cu_seqlens_q", "cu_seqlens_k", "context_lens"):
        assert np.asarray(meta[k]).tolist() == list(want_meta[k]), k
    assert np.asarray(meta["block_tables"]).tolist() == [list(r) for r in want_meta["block_tables"]]
    assert (sp.has_temperature, sp.temperature, sp.top_k) == (1, 0.5, 5)
    # refused: truncated frame, a foreign block size, a block table too short for the tokens
    import pytest
    from vllm_rs_b200 import _lib
    with pytest.raises(_lib.B200Error, match="-1"):
        run.load_run_prefill(fr[:-3])
    bad = dict(q, block_size=16)
    with pytest.raises(_lib.B200Error, match="-3"):
        run.load_run_prefill(ot.run_prefill_frame([bad], True))
    bad = dict(q, token_ids=list(range(70)))
    with pytest.raises(_lib.B200Error, match="-3"):
        run.load_run_prefill(ot.run_prefill_frame([bad], True))
    lib.b200_clear_error()


def test_pd_sequence_messages(lib):
    """TransferPrefill (runner IPC and client -> PD server TCP), ReceivePrefillResponse, KvCacheSend, KvCacheReceive:
    one Sequence each (pinned by the hand-written RunPrefill frame above), plus KvCacheReceiveResponse."""
    import pytest
    from vllm_rs_b200 import _lib
    from vllm_rs_b200.engine import Scheduler
    rng = np.random.default_rng(9)
    eng = Scheduler(64, 64, 8)
    eng.add(rng.integers(0, 1000, size=150).tolist(), max_tokens=21)
    ids, pf = eng.schedule()
    i = ids[0]
    q = dict(id=eng.seq_info(i)["id"], status=1, token_ids=eng.tokens(i), output_ids=[], block_table=eng.block_table(i),
             num_cached_tokens=0, last_token=eng.tokens(i)[-1], block_size=64, sampling=ot.sampling_params(max_tokens=21))
    cases = [((15, 0, 0), ot.transfer_prefill_frame(q)), ((0, 1, 0), ot.transfer_prefill_frame(q, tcp=True)),
             ((18, 0, 1), ot.receive_prefill_response_frame(True, q)), ((23, 0, 777), ot.kv_send_frame(q, 777)),
             ((25, 0, 0), ot.kv_receive_frame(q))]
    box = Scheduler(64, 64, 8)
    for (kind, family, extra), want in cases:
        assert eng.frame_sequence(i, kind, family, extra) == want, kind
        assert box.load_sequence(want, family) == (True, kind, extra)
        assert (box.seq_info(0)["id"], box.tokens(0), box.block_table(0)) == (q["id"], q["token_ids"], q["block_table"])
        assert box.frame_sequence(0, kind, family, extra) == want          # load -> frame is the identity
    none = ot.receive_prefill_response_frame(True, None)
    assert eng.frame_sequence(-1, 18, 0, 1) == none and box.load_sequence(none) == (False, 18, 1)
    assert box.num_running == 0
    with pytest.raises(_lib.B200Error):
        box.load_sequence(ot.kv_send_frame(q, 5), family=1)                 # not a TransferMessage
    with pytest.raises(_lib.B200Error):
        box.load_sequence(cases[0][1][:-1])
    want = ot.kv_receive_response_frame(True, 4242, 1760000000123, 128)
    out = np.zeros(64, np.uint8)
    assert lib.b200_wire_frame_kv_receive_response(out.ctypes.data, 64, 1, 4242, 1760000000123, 128) == len(want)
    assert out[:len(want)].tobytes() == want
    ok, tok, t, c = C.c_int32(), C.c_uint32(), C.c_uint64(), C.c_uint64()
    assert lib.b200_wire_parse_kv_receive_response(out.ctypes.data, len(want), C.byref(ok), C.byref(tok), C.byref(t),
                                                   C.byref(c)) == 0
    assert (ok.value, tok.value, t.value, c.value) == (1, 4242, 1760000000123, 128)
    lib.b200_clear_error()
