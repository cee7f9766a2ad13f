"""Device sampler timing: sample_topk_topp / penalties / argmax at the Llama-3 vocabulary,
CUDA events over graph replays; logits rotated over copies larger than L2.  python tools/sampler_bench.py [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vllm_rs_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
V = 128256
copies = max(2, int(300e6 // (B * V * 4)) + 1)
g = torch.Generator(device="cuda").manual_seed(0)
logits = [torch.randn(B, V, generator=g, device="cuda") * 3 for _ in range(copies)]
seed = torch.tensor([12345], dtype=torch.int64, device="cuda")
out = torch.empty(B, dtype=torch.int32, device="cuda")
ctx = torch.randint(0, V, (B, 4096), generator=g, device="cuda", dtype=torch.int32)
cl = torch.full((B,), 4096, dtype=torch.int32, device="cuda")


def timed(fn, label, bytes_per_call):
    st = torch.cuda.Stream()
    st.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(st):
        for i in range(copies):
            fn(i)
        st.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for i in range(copies):
                fn(i)
        graph.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5):
            graph.replay()
        e1.record(st); st.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * copies)
    print(f"B={B} V={V} {label}: {us:.1f} us/call, {bytes_per_call / us / 1e3:.0f} GB/s of logits")


nb = B * V * 4
timed(lambda i: ops.argmax(logits[i]), "argmax", nb)
for k, p in ((50, 0.9), (256, 0.95), (1, 1.0)):
    timed(lambda i: ops.sample_topk_topp(logits[i], k, p, 0.8, seed, out=out), f"top_k={k} top_p={p}", nb)
timed(lambda i: ops.sample_topk_topp(logits[i], 0, 1.0, 0.8, seed, out=out), "Sampling::All (k=0)", nb)
timed(lambda i: ops.apply_penalties(logits[i], ctx, cl, 0.1, 0.1), "penalties over 4096-token contexts", B * 4096 * 4 * 3)
