"""GPU parity of the tcgen05 prefill attention (csrc/attn_prefill_tc.cu, head_dim 128) against the
CPU oracle (oracle/model.py::attention_prefill), with the mma.sync kernel as a third arm.

Floating point: fp32 scores / softmax on both sides, P rounded to fp16 before the PV product on the
GPU (as flash attention does) => tolerance 2e-3 + 2e-3 |x| on outputs of magnitude <= 1 (the same bar
as the mma.sync kernel's test).  Stale cache rows past the context are poisoned with NaN.
"""
import math

import pytest
import torch

from oracle import model as om

pytestmark = pytest.mark.gpu
F16 = torch.float16


@pytest.fixture(scope="module")
def ops(lib):
    from vllm_rs_b200 import ops as _ops
    return _ops


def _case(H, Hkv, cases, seed, big_scores=False):
    D = 128
    g = torch.Generator().manual_seed(seed)
    B = len(cases)
    ctxs = [n + c for n, c in cases]
    nblk = [-(-c // 64) for c in ctxs]
    nb = sum(nblk) + 3
    perm = torch.randperm(nb, generator=g).tolist()
    kc = torch.randn(nb, 64, Hkv, D, generator=g).to(F16)
    vc = torch.randn(nb, 64, Hkv, D, generator=g).to(F16)
    if big_scores:                       # widely varying row maxima: exercises the lazy-rescale branch
        kc = (kc.float() * torch.linspace(0.2, 6.0, nb).view(nb, 1, 1, 1)).to(F16)
    bt = torch.zeros(B, max(nblk), dtype=torch.int32)
    off = 0
    for b, n in enumerate(nblk):
        bt[b, :n] = torch.tensor(perm[off:off + n], dtype=torch.int32)
        off += n
    cu = [0]
    for n, _ in cases:
        cu.append(cu[-1] + n)
    T = cu[-1]
    q = torch.randn(T, H, D, generator=g).to(F16)
    scale = 1.0 / math.sqrt(D)
    dummy = torch.zeros(T, Hkv, D, dtype=F16)
    ref = om.attention_prefill(q, dummy, dummy, cu, kc, vc, bt.tolist(), ctxs, scale)
    for b, n in enumerate(nblk):         # poison the tail of every last page
        last = ctxs[b] % 64
        if last:
            kc[bt[b, n - 1], last:] = float("nan")
            vc[bt[b, n - 1], last:] = float("nan")
    args = (q.cuda(), kc.cuda(), vc.cuda(), torch.tensor(cu, dtype=torch.int32).cuda(),
            torch.tensor(ctxs, dtype=torch.int32).cuda(), bt.cuda(), max(n for n, _ in cases), scale)
    return args, ref


CASES = [
    (32, 8, [(37, 0), (64, 0), (65, 0), (130, 64), (1, 0), (200, 128)]),     # Llama-3-8B heads, ragged + cached
    (8, 8, [(100, 0), (33, 64)]),                                            # MHA (G = 1: 128 tokens per CTA)
    (8, 1, [(129, 0), (17, 300)]),                                           # TP rank shape, G = 8
    (4, 2, [(700, 0)]),                                                      # G = 2, many pages
]


@pytest.mark.parametrize("H,Hkv,cases", CASES)
def test_attn_prefill_tc_matches_oracle(ops, H, Hkv, cases):
    args, ref = _case(H, Hkv, cases, seed=H * 131 + len(cases))
    errs = {}
    for name, mode in (("mma", ops.ATTN_MMA), ("tc", ops.ATTN_TC), ("tc_swapdesc", ops.ATTN_TC_SWAPDESC)):
        out = ops.attention_prefill(*args, mode=mode).float().cpu()
        nan = int(torch.isnan(out).sum())
        err = (out - ref.float()).abs()
        tol = 2e-3 + 2e-3 * ref.float().abs()
        errs[name] = (nan, float(err[~torch.isnan(err)].max()) if nan < out.numel() else float("nan"),
                      int((err > tol).sum()))
    print(f"H={H} Hkv={Hkv} {cases}: (nan, max err, #bad) = {errs}")
    assert errs["mma"][0] == 0 and errs["mma"][2] == 0
    assert errs["tc"][0] == 0 and errs["tc"][2] == 0, errs


def test_attn_prefill_tc_lazy_rescale_and_long_rows(ops):
    """Row maxima that keep growing along the sequence (scaled key pages) force O rescales; 3000-token
    sequences walk 47 pages through both ring stages and both barrier parities."""
    args, ref = _case(32, 8, [(1000, 0), (3000, 64)], seed=7, big_scores=True)
    out = ops.attention_prefill(*args, mode=ops.ATTN_TC).float().cpu()
    assert not torch.isnan(out).any()
    err = (out - ref.float()).abs()
    assert bool((err <= 3e-3 + 3e-3 * ref.float().abs()).all()), float(err.max())
    # the default dispatch is the tcgen05 kernel: identical bits
    assert torch.equal(out, ops.attention_prefill(*args).float().cpu())
