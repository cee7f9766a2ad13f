"""W4A16 decode GEMM at an arbitrary shape (CUDA events over graph replays, weights rotated over > L2 worth of copies).
    python tools/gemm_shape_bench.py K N [M] [epi] [variant]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vllm_rs_b200 import ops, synthetic

K, N = int(sys.argv[1]), int(sys.argv[2])
M = int(sys.argv[3]) if len(sys.argv) > 3 else 32
epi = int(sys.argv[4]) if len(sys.argv) > 4 else 0
variant = int(sys.argv[5]) if len(sys.argv) > 5 else 0
gen = torch.Generator(device="cuda").manual_seed(0)
copies = max(2, int(300e6 // (K * N // 2)) + 1)
copies = min(copies, 64)
layers = [synthetic._fused_awq_device(gen, K, [N], 128, "cuda") for _ in range(copies)]
x = torch.randn(M, K, device="cuda", dtype=torch.float16)
y = torch.empty(M, N // 2 if epi else N, device="cuda", dtype=torch.float16)
ws = ops.wna16_workspace(max(M, 16), max(N, 28672))
st = torch.cuda.Stream()
st.wait_stream(torch.cuda.current_stream())
def run(i):
    l = layers[i % copies]
    ops.wna16_gemm(x, l.w, l.s, l.z, ws, k=K, n=N, epilogue=epi, variant=variant, out=y)
with torch.cuda.stream(st):
    for i in range(copies * 2): run(i)
    st.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=st):
        for i in range(copies): run(i)
    g.replay(); st.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(6): g.replay()
    e1.record(st); st.synchronize()
us = e0.elapsed_time(e1) * 1e3 / (6 * copies)
print(f"K={K} N={N} M={M} epi={epi} variant={variant} max_splits={os.environ.get('B200_DEC_MAX_SPLITS','-')}: {us:.2f} us  {K * N / 2 / us / 1e3:.0f} GB/s")
