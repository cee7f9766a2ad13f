"""GPU parity of the hand-written dense fp16 GEMM (csrc/dense_gemm_tc.cu) and of the
dequant-to-fp16 pass that feeds it on large quantised prefill steps.

Oracle: torch fp32 matmul of the same fp16 operands on the CPU (a floating-point kernel, so the
bar is a tolerance: fp32 accumulation on both sides, ONE fp16 rounding of the output => 1 fp16 ulp
of |y| plus accumulation-order noise; fp32 outputs are held to 2e-4 relative).  The dequant pass is
BIT-EXACT against oracle/quant.py (one rounding, on the scale multiply).  cuBLAS (mode 0, the call the
reference makes, linear.rs:75-123) runs as a third arm on the same inputs.
"""
import numpy as np
import pytest
import torch

from oracle import quant as oq

pytestmark = pytest.mark.gpu
F16 = torch.float16


@pytest.fixture(scope="module")
def ops(lib):
    from vllm_rs_b200 import ops as _ops
    return _ops


def _case(M, N, K, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(M, K, generator=g).to(F16)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(F16)
    return x, w, x.float() @ w.float().t()


def _check(y, ref, f32):
    err = (y.float().cpu() - ref).abs()
    tol = (2e-4 + 2e-4 * ref.abs()) if f32 else (1e-3 + 1.5e-3 * ref.abs())
    assert bool((err <= tol).all()), f"max err {err.max().item():.3e} at {int(err.argmax())}"


# decode shapes take the swapped weight-streaming form (BN = 32 / 64 / 128), M > 128 the 128 x 256
# tiles (mode 1) or the CTA-pair 256 x 256 tiles (mode 2); ragged M / N exercise TMA zero fill and
# the store predicates, K = 64 the single-k-block pipeline, K = 14336 many ring wrap-arounds
SHAPES = [(1, 512, 64), (32, 4096, 1024), (7, 1000, 256), (64, 2048, 512), (100, 1536, 4096),
          (128, 256, 128), (129, 256, 64), (300, 1000, 512), (1024, 6144, 4096), (777, 4096, 14336),
          (2048, 7392, 1024), (512, 128256, 256)]


@pytest.mark.parametrize("mode", [1, 2], ids=["tc1", "tc2"])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_dense_gemm_tc_matches_fp32_oracle(ops, M, N, K, mode):
    if mode == 1 and M <= 128 and (M, N, K) != (32, 4096, 1024):
        pytest.skip("decode shapes do not depend on the mode")
    x, w, ref = _case(M, N, K, seed=M + N + K)
    xd, wd = x.cuda(), w.cuda()
    _check(ops.dense_gemm(xd, wd, mode=mode), ref, False)
    _check(ops.dense_gemm(xd, wd, out_f32=True, mode=mode), ref, True)
    # the library GEMM on the same inputs (what the reference runs)
    _check(ops.dense_gemm(xd, wd, mode=ops.DENSE_CUBLAS), ref, False)


@pytest.mark.parametrize("mode", [1, 2], ids=["tc1", "tc2"])
def test_dense_gemm_tc_bias_and_determinism(ops, mode):
    x, w, ref = _case(300, 1024, 512, seed=5)
    b = (torch.randn(1024) * 0.5).to(F16)
    y = ops.dense_gemm(x.cuda(), w.cuda(), bias=b.cuda(), mode=mode)
    _check(y, ref + b.float(), False)
    assert torch.equal(y, ops.dense_gemm(x.cuda(), w.cuda(), bias=b.cuda(), mode=mode))
    xs, ws_, refs = _case(16, 1024, 512, seed=6)
    _check(ops.dense_gemm(xs.cuda(), ws_.cuda(), bias=b.cuda(), mode=mode), refs + b.float(), False)


def test_dense_gemm_tc_persistent_many_tiles_identity(ops):
    """Size-independent property at a prefill size: W = I (padded) must reproduce the input exactly,
    whichever CTA / accumulator buffer / ring slot a tile lands in."""
    M, K = 8192, 1024
    g = torch.Generator().manual_seed(0)
    x = torch.randn(M, K, generator=g).to(F16).cuda()
    w = torch.zeros(2048, K, dtype=F16, device="cuda")
    w[:K] = torch.eye(K, dtype=F16, device="cuda")
    for mode in (1, 2):
        y = ops.dense_gemm(x, w, mode=mode)
        assert torch.equal(y[:, :K], x) and int(y[:, K:].abs().sum()) == 0


@pytest.mark.parametrize("fmt,bits", [("awq", 4), ("gptq", 4), ("gptq_sym", 4), ("gptq", 8)])
def test_wna16_dequant_bit_exact(ops, fmt, bits):
    K, N, G = 1024, 512, 128
    rng = np.random.default_rng(11)
    if fmt == "awq":
        qw, qz, s = oq.random_awq(rng, K, N, G)
        W = oq.dequant_awq(qw, qz, s, G)
    else:
        qw, qz, s = oq.random_gptq(rng, K, N, G, bits, sym=(fmt == "gptq_sym"))
        if fmt == "gptq_sym":
            qz = None
        W = oq.dequant_gptq(qw, qz, s, G, bits)
    qwd = torch.from_numpy(qw).cuda()
    packed = ops.awq_prepack(qwd) if fmt == "awq" else ops.gptq_prepack(qwd, bits)
    out = ops.wna16_dequant(packed, torch.from_numpy(s).cuda(), None if qz is None else torch.from_numpy(qz).cuda(),
                            k=K, n=N, bits=bits, group_size=G, qfmt=ops.QFMT_AWQ if fmt == "awq" else ops.QFMT_GPTQ)
    want = torch.from_numpy(np.ascontiguousarray(W.T))                 # oracle W is [K, N]
    assert torch.equal(out.cpu(), want)
