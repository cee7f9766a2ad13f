"""W4A16 GEMM at prefill token counts: the in-kernel dequant tcgen05 kernel (wna16_tc_kernel<4, 256>, variant 2)
against dequantise-once + dense tcgen05 GEMM (what the runner does for steps >= 1024 tokens), per Llama-3-8B
projection.  CUDA events, 5 timed iterations after 2 warm-ups, weights > L2 for gate|up.
    python tools/prefill_gemm_ab.py [M]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vllm_rs_b200 import ops, synthetic

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
gen = torch.Generator(device="cuda").manual_seed(0)
st = torch.cuda.current_stream()


def timed(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for name, K, N in (("qkv", 4096, 6144), ("o", 4096, 4096), ("gate_up", 4096, 28672), ("down", 14336, 4096)):
    l = synthetic._fused_awq_device(gen, K, [N], 128, "cuda")
    x = torch.randn(M, K, device="cuda", dtype=torch.float16)
    y = torch.empty(M, N, device="cuda", dtype=torch.float16)
    ws = ops.wna16_workspace(M, N)
    wd = torch.empty(N, K, device="cuda", dtype=torch.float16)
    flop = 2.0 * M * N * K
    t_in = timed(lambda: ops.wna16_gemm(x, l.w, l.s, l.z, ws, k=K, n=N, variant=2, out=y))
    y_in = y.clone()
    def two_step():
        ops.wna16_dequant(l.w, l.s, l.z, k=K, n=N, out=wd)
        ops.dense_gemm(x, wd, out=y)
    t_dd = timed(two_step)
    t_dq = timed(lambda: ops.wna16_dequant(l.w, l.s, l.z, k=K, n=N, out=wd))
    err = (y.float() - y_in.float()).abs().max().item()
    print(f"{name:8s} M={M} K={K} N={N}: in-kernel dequant {t_in*1e3:8.1f} us = {flop/t_in/1e9:6.0f} TF/s | "
          f"dequant + dense {t_dd*1e3:8.1f} us = {flop/t_dd/1e9:6.0f} TF/s (dequant alone {t_dq*1e3:6.1f} us) | "
          f"max |diff| {err:.2e}")
