import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch
from oracle import quant as oq, model as om
from vllm_rs_b200 import ops
F16 = torch.float16
def case(K, N, M, fmt, group, bias, seed=0):
    rng = np.random.default_rng(seed)
    if fmt == "awq":
        qw, qz, s = oq.random_awq(rng, K, N, group); W = oq.dequant_awq(qw, qz, s, group)
    else:
        qw, qz, s = oq.random_gptq(rng, K, N, group, 4); W = oq.dequant_gptq(qw, qz, s, group, 4)
    x = torch.from_numpy(rng.standard_normal((M, K)).astype(np.float16))
    b = torch.randn(N).to(F16) if bias else None
    qwd = torch.from_numpy(qw).cuda()
    packed = ops.awq_prepack(qwd) if fmt == "awq" else ops.gptq_prepack(qwd, 4)
    ws = ops.wna16_workspace(max(M, 16), N)
    kw = dict(k=K, n=N, bits=4, group_size=group, qfmt=ops.QFMT_AWQ if fmt == "awq" else ops.QFMT_GPTQ,
              bias=None if b is None else b.cuda())
    ref = om.linear(x, torch.from_numpy(W), b)
    out = {}
    for v in (2, 5):
        errs = []
        for rep in range(20):
            y = ops.wna16_gemm(x.cuda(), packed, torch.from_numpy(s).cuda(), torch.from_numpy(qz).cuda(), ws, variant=v, **kw)
            errs.append((y.cpu().float() - ref.float()).abs().max().item())
        out[v] = (max(errs), min(errs))
    print(f"K={K} N={N} M={M} {fmt} g={group} bias={bias}: old {out[2]}  sk {out[5]}")
for M in (6, 1, 16, 32):
    for bias in (False, True):
        case(512, 768, M, "gptq", 128, bias)
        case(512, 768, M, "awq", 128, bias)
case(512, 512, 6, "gptq", 128, False); case(1024, 512, 6, "gptq", 128, False); case(512, 2048, 6, "gptq", 128, False)
