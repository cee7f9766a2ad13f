"""GPU parity of the device sampler and the penalty kernel against oracle/sampling.py.

Integer outputs (token ids): identical to the oracle's draw for the same (seed, row) uniform, except
where the uniform lands within float round-off of a CDF boundary or the top-p cut lands within
round-off of p (the kernel uses __expf; the oracle float64 exp) -- those rows must still pick an
ADJACENT candidate.  Penalised logits: bit-exact.  Distribution: chi-square against the exact law.
"""
import numpy as np
import pytest
import torch

from oracle import sampling as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(lib):
    from vllm_rs_b200 import ops as _ops
    return _ops


def _rows_agree(logits, got, k, p, t, seed):
    exact = near = 0
    for b in range(logits.shape[0]):
        ids, prs = S.candidates(logits[b], k, t)
        kept = S.topp_clamp(prs, p)
        u = S.row_uniform(seed, b)
        want = int(ids[S.weighted_index(kept, u)])
        if int(got[b]) == want:
            exact += 1
            continue
        # tolerate a neighbour when u * total is within round-off of a boundary, or the cut is
        pos = {int(x): i for i, x in enumerate(ids)}
        assert int(got[b]) in pos, f"row {b}: token {int(got[b])} is not among the top-{k}"
        cum = np.cumsum(kept.astype(np.float64))
        x = float(u) * cum[-1]
        gi, wi = pos[int(got[b])], pos[want]
        boundary = min(abs(cum[min(gi, wi)] - x), abs(np.cumsum(prs.astype(np.float64))[min(gi, wi)] - p))
        assert abs(gi - wi) <= 1 and boundary < 5e-5 * max(cum[-1], 1e-6) + 1e-6, \
            f"row {b}: got candidate {gi}, oracle {wi}, boundary distance {boundary}"
        near += 1
    return exact, near


@pytest.mark.parametrize("V,k,p,t", [(1024, 32, 0.95, 0.7), (128256, 32, 0.95, 0.7), (32000, 256, 1.0, 1.0),
                                     (128256, 256, 0.5, 1.3), (5000, 1, 0.9, 0.5), (300, 256, 0.3, 2.0),
                                     (100, 256, 1.0, 1.0)])
def test_sampler_matches_oracle(ops, V, k, p, t):
    rng = np.random.default_rng(V + k)
    B = 32
    logits = (rng.standard_normal((B, V)) * 3.0).astype(np.float32)
    logits[3, :50] = logits[3, 0]                      # exact ties at the top: lowest indices win
    logits[4] = 0.0                                    # a flat row
    ld = torch.from_numpy(logits).cuda()
    tot_exact = 0
    for seed in (0, 12345678901234567, 2**64 - 3):
        got = ops.sample_topk_topp(ld, k, p, t, seed).cpu().numpy()
        assert np.array_equal(got, ops.sample_topk_topp(ld, k, p, t, seed).cpu().numpy())    # deterministic
        exact, near = _rows_agree(logits, got, min(k, V), p, t, seed)
        tot_exact += exact
        assert near <= 2
    assert tot_exact >= 3 * B - 3


def test_sampler_top1_is_argmax_and_seed_changes_draws(ops):
    rng = np.random.default_rng(1)
    logits = torch.from_numpy(rng.standard_normal((16, 4096)).astype(np.float32)).cuda()
    assert torch.equal(ops.sample_topk_topp(logits, 1, 1.0, 1.0, 7).long(), logits.argmax(-1))
    a = ops.sample_topk_topp(logits, 64, 1.0, 5.0, 1)
    b = ops.sample_topk_topp(logits, 64, 1.0, 5.0, 2)
    assert not torch.equal(a, b)


def test_sampler_distribution_chi_square(ops):
    """4096 identical rows = 4096 draws of one distribution (row index varies the uniform)."""
    probs = np.array([0.4, 0.3, 0.15, 0.1, 0.05])
    V, B = 64, 4096
    row = np.full(V, -30.0, np.float32)
    row[[7, 3, 40, 41, 63]] = np.log(probs).astype(np.float32)
    logits = torch.from_numpy(np.tile(row, (B, 1))).cuda()
    got = ops.sample_topk_topp(logits, 5, 1.0, 1.0, 99).cpu().numpy()
    counts = np.array([(got == i).sum() for i in (7, 3, 40, 41, 63)])
    assert counts.sum() == B
    chi2 = (((counts - probs * B) ** 2) / (probs * B)).sum()
    assert chi2 < 23.5, (counts, chi2)                 # 4 dof, p ~ 1e-4
    # top-p 0.8 on the same law keeps {0.4, 0.3, 0.15} (mass before the 4th is 0.85 >= 0.8)
    got = ops.sample_topk_topp(logits, 5, 0.8, 1.0, 99).cpu().numpy()
    assert set(np.unique(got).tolist()) == {7, 3, 40}


def test_penalties_bit_exact(ops):
    rng = np.random.default_rng(2)
    B, V, cap = 6, 1000, 128
    logits = rng.standard_normal((B, V)).astype(np.float32)
    lens = np.array([128, 128, 1, 0, 77, 128], np.int32)
    ctx = rng.integers(0, 40, size=(B, cap)).astype(np.int32)           # many repeats
    ctx[5, :4] = [999, 999, 1000, 5000]                                 # out-of-vocabulary ids are skipped
    want = np.stack([S.apply_penalties(logits[b], ctx[b, :lens[b]].tolist(), 0.5, 0.25) for b in range(B)])
    got = ops.apply_penalties(torch.from_numpy(logits).cuda(), torch.from_numpy(ctx).cuda(),
                              torch.from_numpy(lens).cuda(), 0.5, 0.25).cpu().numpy()
    assert np.array_equal(got, want)
    # the reference's guards: penalty values 0 and 1 switch the pass off entirely
    same = ops.apply_penalties(torch.from_numpy(logits).cuda(), torch.from_numpy(ctx).cuda(),
                               torch.from_numpy(lens).cuda(), 1.0, 0.0).cpu().numpy()
    assert np.array_equal(same, logits)


def test_runner_samples_on_device_under_graph(lib):
    """The decode step with the device sampler: same logits as the greedy step, ids from the oracle's
    draw for the step seed, and a captured graph picks up a new seed without re-capture."""
    from oracle import sched as osched
    from oracle.toy import build_pair, toy_cfg
    cfg = toy_cfg()
    oracle, runner = build_pair(cfg, "awq", seed=11, num_blocks=16, max_batch=4, max_blocks=4)
    B = 4
    g = torch.Generator().manual_seed(0)
    ids = torch.randint(0, cfg.vocab, (B,), generator=g, dtype=torch.int32).cuda()
    pos = torch.zeros(B, dtype=torch.int64).cuda()
    bt = torch.arange(B, dtype=torch.int32).reshape(B, 1).repeat(1, 4).cuda()
    slots = (bt[:, 0].long() * 64)
    cl = torch.ones(B, dtype=torch.int32).cuda()
    logits, greedy = runner.decode_device(ids, pos, slots, bt, cl)
    assert torch.equal(greedy.long().cpu(), logits.argmax(-1).cpu())
    runner.set_sampling(32, 0.95, 0.7)
    rng = S.StdRng(299792458)
    for _ in range(3):
        seed = rng.next_u64()
        runner.set_step_seed(seed)
        lg, tok = runner.decode_device(ids, pos, slots, bt, cl)
        _rows_agree(lg.cpu().numpy(), tok.cpu().numpy(), 32, 0.95, 0.7, seed)
    # host path with CUDA graphs
    runner.capture([B])
    meta = dict(input_ids=ids.cpu().numpy().astype(np.uint32), positions=pos.cpu().numpy(),
                slot_mapping=slots.cpu().numpy(), block_tables=bt.cpu().numpy().astype(np.uint32),
                context_lens=cl.cpu().numpy().astype(np.uint32), max_blocks=4)
    seen = []
    for _ in range(4):
        seed = rng.next_u64()
        runner.set_step_seed(seed, stream=runner._graph_stream.cuda_stream)
        toks = runner.run_decode(meta)
        _rows_agree(lg.cpu().numpy(), np.asarray(toks), 32, 0.95, 0.7, seed)
        seen.append(tuple(int(x) for x in toks))
    assert len(set(seen)) > 1
