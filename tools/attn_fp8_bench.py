"""Decode attention micro-benchmark, fp16 vs e4m3 KV cache (b=32, ctx=4096, Llama-3-8B heads), CUDA events over
L distinct layer caches (>> L2).  Env: B200_ATTN_F8 (kernel variant), B200_ATTN_PPS (pages per split),
B200_ATTN_DEBUG=1 prints the resident CTAs per SM.   python tools/attn_fp8_bench.py [batch] [ctx]"""
import math, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vllm_rs_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
ctx = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
H, Hkv, D, L = 32, 8, 128, 8
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
bps = ctx // 64 + 2
nb = B * bps
bt = torch.randperm(nb, device=dev, generator=g).to(torch.int32).reshape(B, bps).contiguous()
cl = torch.full((B,), ctx, dtype=torch.int32, device=dev)
q = torch.randn(B, H, D, device=dev, dtype=torch.float16, generator=g)
ws = ops.paged_attn_workspace(B, H, D, bps, dev)
out = torch.empty_like(q)
scale = 1 / math.sqrt(D)
res = {}
for name in ([os.environ["B200_ATTN_ONLY"]] if os.environ.get("B200_ATTN_ONLY") else ["f16", "fp8"]):
    if name == "f16":
        caches = [(torch.randn(nb, 64, Hkv, D, device=dev, dtype=torch.float16, generator=g),
                   torch.randn(nb, 64, Hkv, D, device=dev, dtype=torch.float16, generator=g)) for _ in range(L)]
        nbytes = B * ctx * Hkv * D * 2 * 2
    else:
        caches = [tuple(torch.randn(nb, 64, Hkv, D, device=dev, dtype=torch.float16, generator=g)
                        .to(torch.float8_e4m3fn).view(torch.uint8) for _ in range(2)) for _ in range(L)]
        nbytes = B * ctx * Hkv * D * 2 * 1
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(3):
            e0.record(st)
            for k, v in caches:
                ops.paged_attention_decode(q, k, v, bt, cl, scale, ws, out=out)
            e1.record(st)
            st.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / L
    res[name] = (us, nbytes / us / 1e3)
    del caches
    torch.cuda.empty_cache()
print(f"b={B} ctx={ctx} F8={os.environ.get('B200_ATTN_F8', 'dflt')} PPS={os.environ.get('B200_ATTN_PPS', 'plan')}: " +
      "  ".join(f"{k}: {v[0]:.1f} us {v[1]:.0f} GB/s" for k, v in res.items()))
