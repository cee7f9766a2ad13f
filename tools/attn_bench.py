"""Prefill attention micro-benchmark: one Llama-3-8B layer (32 q heads, 8 kv heads, D = 128), causal,
one sequence of T tokens over the paged cache.  tcgen05 kernel vs the mma.sync kernel, CUDA events.
python tools/attn_bench.py [T]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vllm_rs_b200 import ops

T = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
H, Hkv, D = 32, 8, 128
nb = T // 64 + 1
dev = "cuda"
q = torch.randn(T, H, D, device=dev, dtype=torch.float16)
kc = torch.randn(nb, 64, Hkv, D, device=dev, dtype=torch.float16)
vc = torch.randn(nb, 64, Hkv, D, device=dev, dtype=torch.float16)
cu = torch.tensor([0, T], dtype=torch.int32, device=dev)
ctx = torch.tensor([T], dtype=torch.int32, device=dev)
bt = torch.randperm(nb, device=dev).to(torch.int32).reshape(1, nb)
out = torch.empty_like(q)
flops = 4.0 * H * D * (T * (T + 1) / 2)
res = {}
for name, mode in (("mma.sync", ops.ATTN_MMA), ("tcgen05", ops.ATTN_TC)):
    try:
        for _ in range(2):
            ops.attention_prefill(q, kc, vc, cu, ctx, bt, T, 1 / math.sqrt(D), out=out, mode=mode)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.attention_prefill(q, kc, vc, cu, ctx, bt, T, 1 / math.sqrt(D), out=out, mode=mode)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        res[name] = out.clone()
        print(f"T={T} {name}: {ms:.3f} ms  {flops / ms / 1e9:.0f} TF/s (causal flops)", flush=True)
    except Exception as e:
        print(f"T={T} {name}: FAILED {e!r}", flush=True)
        break
if len(res) == 2:
    print("max |tc - mma| =", float((res["tcgen05"].float() - res["mma.sync"].float()).abs().max()))
