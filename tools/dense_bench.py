"""Micro-benchmark of the dense fp16 GEMM (csrc/dense_gemm_tc.cu) against cuBLAS on the Llama-3-8B
prefill shapes (tensor roofline) and the decode lm_head (HBM roofline).  CUDA events, weights rotated
over copies larger than L2 for the HBM-bound case.   python tools/dense_bench.py [M]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vllm_rs_b200 import ops

M = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
dev = "cuda"
peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
names = {0: "cublas", 1: "tc1", 2: "tc2"}


def timeit(fn, reps):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        for _ in range(3):
            fn(0)
        st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for i in range(reps):
            fn(i)
        e1.record(st)
        st.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"peaks: {peaks.get('bf16_tflops')} TF/s burst, {peaks.get('bf16_tflops_sustained')} sustained, {peaks.get('hbm_gbs')} GB/s")
for name, (N, K) in {"qkv": (6144, 4096), "o": (4096, 4096), "gate_up": (28672, 4096), "down": (4096, 14336)}.items():
    x = torch.randn(M, K, device=dev, dtype=torch.float16)
    w = torch.randn(N, K, device=dev, dtype=torch.float16) * 0.02
    y = torch.empty(M, N, device=dev, dtype=torch.float16)
    row = []
    for mode in (0, 1, 2):
        try:
            ms = timeit(lambda i: ops.dense_gemm(x, w, out=y, mode=mode), 10)
            row.append(f"{names[mode]} {ms:.3f} ms {2.0 * M * N * K / ms / 1e9:.0f} TF/s")
        except Exception as e:          # a failing mode must not hide the others
            row.append(f"{names[mode]} FAILED {e!r}")
            break
    print(f"M={M} {name} [N={N} K={K}]: " + " | ".join(row), flush=True)

# decode lm_head: M = 32, vocab 128256 x 4096 fp16 = 1.05 GB per call (>> L2), fp32 logits
N, K = 128256, 4096
ws = [torch.randn(N, K, device=dev, dtype=torch.float16) * 0.02 for _ in range(2)]
for Md in (1, 32, 64, 128):
    x = torch.randn(Md, K, device=dev, dtype=torch.float16)
    y = torch.empty(Md, N, device=dev, dtype=torch.float32)
    row = []
    for mode in (0, 2):
        ms = timeit(lambda i: ops.dense_gemm(x, ws[i & 1], out_f32=True, out=y, mode=mode), 10)
        row.append(f"{names[mode]} {ms * 1e3:.0f} us {N * K * 2 / ms / 1e6:.0f} GB/s")
    print(f"lm_head M={Md}: " + " | ".join(row), flush=True)
