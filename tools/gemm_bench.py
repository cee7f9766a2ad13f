"""Micro-benchmark of the W4A16 GEMM at the Llama-3-8B decode shapes (CUDA events, weights
rotated over copies larger than L2).  python tools/gemm_bench.py [M] [variant]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vllm_rs_b200 import ops, synthetic

M = int(sys.argv[1]) if len(sys.argv) > 1 else 32
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(0)
shapes = {"qkv": (4096, 6144, 0), "o": (4096, 4096, 0), "gate_up": (4096, 28672, 1), "down": (14336, 4096, 0)}
ws = ops.wna16_workspace(max(M, 16), 28672)
if os.environ.get("B200_GEMM_ONLY"):
    shapes = {k: v for k, v in shapes.items() if k == os.environ["B200_GEMM_ONLY"]}
out = {}
for name, (K, N, epi) in shapes.items():
    copies = max(2, int(300e6 // (K * N // 2)) + 1)
    layers = [synthetic._fused_awq_device(gen, K, [N], 128, dev) for _ in range(copies)]
    x = torch.randn(M, K, device=dev, dtype=torch.float16)
    y = torch.empty(M, N // 2 if epi else N, device=dev, dtype=torch.float16)
    def run(i):
        l = layers[i % copies]
        ops.wna16_gemm(x, l.w, l.s, l.z, ws, k=K, n=N, epilogue=epi, variant=variant, out=y)
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(copies * 2): run(i)
        st.synchronize()
        graph = torch.cuda.CUDAGraph()          # graph replay: no CPU launch overhead in the timing
        with torch.cuda.graph(graph, stream=st):
            for i in range(copies): run(i)
        graph.replay(); st.synchronize()
        reps = 6
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps): graph.replay()
        e1.record(st); st.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * copies)
    out[name] = (us, K * N / 2 / us / 1e3)
print(f"M={M} variant={variant} dbg={os.environ.get('B200_TC_DBG','0')} pdl={os.environ.get('B200_PDL','1')}: " +
      "  ".join(f"{k}: {v[0]:.1f}us {v[1]:.0f}GB/s" for k, v in out.items()) +
      f"  | sum {sum(v[0] for v in out.values()):.1f}us")

# ---- chain mode: the four GEMMs of a layer back to back (what the decode step launches), over
# several layers' worth of distinct weights -- cold instruction / L2 state between shapes.
if os.environ.get("B200_GEMM_CHAIN", "1") != "0" and len(shapes) == 4:
    nl = 4
    Ls = [{n: synthetic._fused_awq_device(gen, K, [N], 128, dev) for n, (K, N, _) in shapes.items()}
          for _ in range(nl)]
    x4 = torch.randn(M, 4096, device=dev, dtype=torch.float16)
    x14 = torch.randn(M, 14336, device=dev, dtype=torch.float16)
    ys = {n: torch.empty(M, (N // 2 if e else N), device=dev, dtype=torch.float16) for n, (K, N, e) in shapes.items()}
    def layer(i):
        for n, (K, N, e) in shapes.items():
            l = Ls[i % nl][n]
            ops.wna16_gemm(x14 if K == 14336 else x4, l.w, l.s, l.z, ws, k=K, n=N, epilogue=e,
                           variant=variant, out=ys[n])
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for i in range(nl * 2): layer(i)
        st.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=st):
            for i in range(nl * 4): layer(i)
        graph.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(6): graph.replay()
        e1.record(st); st.synchronize()
    print(f"chain: {e0.elapsed_time(e1) * 1e3 / (6 * nl * 4):.1f} us per layer (4 GEMMs interleaved)")
