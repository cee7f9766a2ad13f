"""Does tcgen05.mma kind::f16 treat fp16 SUBNORMAL operands exactly?  (design probe for the decode GEMM:
a 4-bit code q placed in the low mantissa bits of an fp16 word is the subnormal q * 2^-24 -- one AND
per two weights instead of lop3 + sub + mul).  Uses the dense tcgen05 GEMM (weights = operand)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vllm_rs_b200 import ops

torch.manual_seed(0)
dev = "cuda"
for (M, N, K) in [(32, 256, 128), (32, 4096, 4096), (256, 512, 256)]:
    q = torch.randint(0, 16, (N, K), device=dev, dtype=torch.int16)
    w_sub = q.view(torch.float16)                       # bit pattern 0x000q = q * 2^-24
    w_hi = (q << 4).view(torch.float16)                 # 0x00q0 = 16 q * 2^-24
    x = torch.randn(M, K, device=dev, dtype=torch.float16)
    ref = x.double() @ q.double().t()
    for name, w, sc in (("lo", w_sub, 2.0 ** 24), ("hi", w_hi, 2.0 ** 20)):
        assert (w.double() * sc == q.double()).all()
        for mode in (ops.DENSE_TC1, ops.DENSE_TC2, ops.DENSE_CUBLAS):
            try:
                y = ops.dense_gemm(x, w.contiguous(), out_f32=True, mode=mode).double() * sc
            except Exception as e:                      # noqa: BLE001
                print(f"M{M} N{N} K{K} {name} mode{mode}: {e}")
                continue
            err = (y - ref).abs().max().item()
            print(f"M{M} N{N} K{K} {name} mode{mode}: max|err| = {err:.3e}  (|ref| max {ref.abs().max().item():.1f}, "
                  f"zero-output = {(y == 0).all().item()})")
