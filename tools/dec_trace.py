"""Pipeline trace of the decode GEMM (needs a build with B200_NVCC_EXTRA=-DB200_DEC_TRACE):
per-chunk SM-clock timestamps of CTA (0,0): weight copy issued, dequant group waits for / gets the
chunk, TMEM store, MMA issue, activation copy issued.  python tools/dec_trace.py [shape] [M]"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from vllm_rs_b200 import ops, synthetic, _lib

name = sys.argv[1] if len(sys.argv) > 1 else "gate_up"
M = int(sys.argv[2]) if len(sys.argv) > 2 else 32
shapes = {"qkv": (4096, 6144, 0), "o": (4096, 4096, 0), "gate_up": (4096, 28672, 1), "down": (14336, 4096, 0)}
K, N, epi = shapes[name]
dev = "cuda"
gen = torch.Generator(device=dev).manual_seed(0)
lib = _lib.lib()
lib.b200_debug_set_dec_trace.argtypes = [ctypes.c_void_p]
copies = max(2, int(300e6 // (K * N // 2)) + 1)
layers = [synthetic._fused_awq_device(gen, K, [N], 128, dev) for _ in range(copies)]
x = torch.randn(M, K, device=dev, dtype=torch.float16)
y = torch.empty(M, N // 2 if epi else N, device=dev, dtype=torch.float16)
ws = ops.wna16_workspace(max(M, 16), 28672)
tr = torch.zeros(16 * 256, device=dev, dtype=torch.int32)
def run(i):
    l = layers[i % copies]
    ops.wna16_gemm(x, l.w, l.s, l.z, ws, k=K, n=N, epilogue=epi, out=y)
for i in range(copies): run(i)
torch.cuda.synchronize()
lib.b200_debug_set_dec_trace(ctypes.c_void_p(tr.data_ptr()))
for i in range(copies): run(i)          # back to back: the traced launch follows a warm pipeline
torch.cuda.synchronize()
t = tr.cpu().view(16, 256).numpy().astype("int64") & 0xFFFFFFFF
nz = t[:8][t[:8] > 0]
t0 = nz.min()
names = ["w_issue", "dq_wait", "dq_got", "dq_store", "dq_arrive", "mma_issue", "x_issue", "mma_wait"]
nch = int((t[2] > 0).sum())
print(f"{name} M={M}: {nch} chunks in CTA (0,0); times in cycles since first event")
print("chunk " + " ".join(f"{n:>10s}" for n in names))
for i in range(nch):
    print(f"{i:5d} " + " ".join(f"{(t[e][i] - t0) if t[e][i] else -1:10d}" for e in range(8)))
print("epilogue of tile 0 per split: start, accumulators complete, partials stored, arrived, reduced (last only), exit")
for sp in range(8):
    if t[8 + sp][0]:
        print(f"split {sp}: " + " ".join(f"{(t[8 + sp][k] - t0) if t[8 + sp][k] else -1:9d}" for k in range(6)))
