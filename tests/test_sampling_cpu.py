"""CPU: the sampling oracle against published vectors, and the C-ABI host mirror (StdRng stream,
strategy selection) bit-exact against the oracle.  No GPU."""
import ctypes as C

import numpy as np
import pytest

from oracle import sampling as S


def test_chacha_core_matches_rfc7539_block_vector():
    """RFC 7539 section 2.3.2: key 00..1f, counter 1, nonce 00:00:00:09:00:00:00:4a:00:00:00:00."""
    key = [int.from_bytes(bytes(range(4 * i, 4 * i + 4)), "little") for i in range(8)]
    out = S.chacha_block(key, 1, 0x09000000, 0x4A000000, 0, 20)
    assert out == [0xE4E7F110, 0x15593BD1, 0x1FDD0F50, 0xC47120A3, 0xC7F4D1C7, 0x0368C033, 0x9AAA2204, 0x4E6CD4C3,
                   0x466482D2, 0x09AA9F07, 0x05D7C214, 0xA2028BD9, 0xD19C12B5, 0xB94E16DE, 0xE883D0CB, 0x4E3C50A2]


def test_pcg32_seed_expansion_is_xsh_rr():
    """seed_from_u64 is PCG32 XSH-RR with the rand_core constants; state 0 -> first outputs by hand:
    state1 = INC, x = rotr32(((s >> 18) ^ s) >> 27, s >> 59)."""
    inc = 11634580027462260723
    xs = (((inc >> 18) ^ inc) >> 27) & 0xFFFFFFFF
    rot = inc >> 59
    want0 = ((xs >> rot) | (xs << (32 - rot))) & 0xFFFFFFFF
    assert S.seed_from_u64(0)[0] == want0


def test_c_rng_stream_matches_oracle(lib):
    for seed in (0, 1, 42, 299792458, 2**64 - 1):
        o = S.StdRng(seed)
        h = lib.b200_rng_new(seed)
        try:
            for i in range(70):                    # crosses several 16-word blocks, mixes u32 / u64 draws
                if i % 5 == 3:
                    assert lib.b200_rng_next_u32(h) == o.next_u32()
                else:
                    assert lib.b200_rng_next_u64(h) == o.next_u64()
        finally:
            lib.b200_rng_free(h)


CASES = [(None, None, None), (0.0, 32, 0.9), (1e-9, None, None), (0.7, None, None), (0.7, 0, None), (0.7, -1, 0.5),
         (0.7, 32, None), (0.7, None, 0.95), (0.7, 32, 0.95), (1.3, 256, 1.0), (-1.0, 5, 0.5)]


@pytest.mark.parametrize("t,k,p", CASES)
def test_strategy_selection_matches_oracle(lib, t, k, p):
    kinds = {"argmax": 0, "all": 1, "topk": 2, "topp": 3, "topk_topp": 4}
    want = S.get_strategy(t, k, p)
    ko, po, to = C.c_int32(), C.c_float(), C.c_float()
    kind = lib.b200_sampling_strategy(int(t is not None), float(t or 0.0), int(k is not None), int(k or 0),
                                      int(p is not None), float(p or 0.0), C.byref(ko), C.byref(po), C.byref(to))
    assert kind == kinds[want[0]]
    kp = S.kernel_params(want)
    if kp is None:
        assert ko.value == 0
    else:
        assert (ko.value, np.float32(po.value), np.float32(to.value)) == (kp[0], np.float32(kp[1]), np.float32(kp[2]))


def test_default_strategy_is_top32_p095_t07():
    assert S.kernel_params(S.DEFAULT_STRATEGY) == (32, 0.95, 0.7)      # core/runner.rs:1483-1487


def test_oracle_draw_semantics():
    logits = np.log(np.array([0.5, 0.25, 0.125, 0.0625, 0.0625], np.float64)).astype(np.float32)
    ids, prs = S.candidates(logits, 3, 1.0)
    assert ids.tolist() == [0, 1, 2] and np.allclose(prs, [0.5, 0.25, 0.125])
    # the cut walks the UN-renormalised top-k mass: with p = 0.6 the second token is still kept
    # (cumsum 0.5 < 0.6 when it is visited), the third is zeroed
    assert S.topp_clamp(prs, 0.6).tolist() == [0.5, 0.25, 0.0]
    assert S.topp_clamp(prs, 0.9).tolist() == prs.tolist()             # p >= sum of the top-k mass: no cut
    assert S.weighted_index(np.array([0.5, 0.25, 0.0], np.float32), 0.0) == 0
    assert S.weighted_index(np.array([0.5, 0.25, 0.0], np.float32), 0.70) == 1
    # penalties: window rule and the count / presence formula
    assert S.penalty_context(list(range(128))) == [] and len(S.penalty_context(list(range(129)))) == 128
    out = S.apply_penalties(np.zeros(8, np.float32), [1, 1, 3], 0.5, 0.25)
    assert out.tolist() == [0, -1.25, 0, -0.75, 0, 0, 0, 0]
    assert S.apply_penalties(np.zeros(8, np.float32), [1], 0.5, 0.25).tolist() == [0] * 8      # len <= 1: no-op
