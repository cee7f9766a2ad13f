"""ctypes wrapper + build recipe for oracle/c/decode_ref.c (CPU baseline / C restatement).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Output goes to oracle/_build/ (git-ignored,
travels to the GPU box with gpurun like the product .so).
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
SRC = HERE / "c" / "decode_ref.c"
OUT = HERE / "_build" / "libdecode_ref.so"
_lib = None


def build(force: bool = False) -> Path:
    OUT.parent.mkdir(exist_ok=True)
    if force or not OUT.exists() or OUT.stat().st_mtime < SRC.stat().st_mtime:
        # x86-64-v3 (AVX2 + F16C + FMA): portable between this container and the GPU host
        cmd = ["gcc", "-O3", "-march=x86-64-v3", "-fopenmp", "-shared", "-fPIC", str(SRC), "-o",
               str(OUT), "-lm"]
        subprocess.run(cmd, check=True)
    return OUT


class RefLayer(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "input_norm", "post_norm",
        "q_w", "q_z", "k_w", "k_z", "v_w", "v_z", "o_w", "o_z", "g_w", "g_z", "u_w", "u_z",
        "d_w", "d_z", "q_s", "k_s", "v_s", "o_s", "g_s", "u_s", "d_s", "k_cache", "v_cache")]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(OUT))
        _lib.ref_num_threads.restype = C.c_int
        _lib.ref_set_num_threads.argtypes = [C.c_int]
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def awq_gemm(x, qweight, qzeros, scales, group, bias=None):
    M, K = x.shape
    N = scales.shape[1]
    y = np.empty((M, N), dtype=np.float16)
    lib().ref_awq_gemm(_p(x), _p(qweight), _p(qzeros), _p(scales), _p(bias), _p(y), M, K, N, group)
    return y


def paged_attn_decode(q, k_cache, v_cache, block_tables, context_lens, scale):
    B, H, D = q.shape
    nb, bs, Hkv, _ = k_cache.shape
    out = np.empty_like(q)
    bt = np.ascontiguousarray(block_tables, dtype=np.uint32)
    cl = np.ascontiguousarray(context_lens, dtype=np.uint32)
    lib().ref_paged_attn_decode(_p(q), _p(k_cache), _p(v_cache), _p(bt), _p(cl), _p(out), B, H, Hkv,
                                D, bs, bt.shape[1], C.c_float(scale))
    return out


def dense_gemm(x, w):
    M, K = x.shape
    N = w.shape[0]
    y = np.empty((M, N), dtype=np.float32)
    lib().ref_dense_gemm(_p(x), _p(w), _p(y), M, N, K)
    return y


def rmsnorm(x, w, eps):
    out = np.empty_like(x)
    lib().ref_rmsnorm(_p(x), _p(w), _p(out), x.shape[0], x.shape[1], C.c_float(eps))
    return out


def decode_layers(x, layers, cos, sin, positions, slots, block_tables, context_lens, *, H, Hkv, D,
                  inter, group, eps=1e-5):
    """layers: list of dicts with the RefLayer field names -> numpy arrays.  x updated in place."""
    arr = (RefLayer * len(layers))()
    for i, lw in enumerate(layers):
        for name, _ in RefLayer._fields_:
            setattr(arr[i], name, lw[name].ctypes.data)
    B, hidden = x.shape
    bt = np.ascontiguousarray(block_tables, dtype=np.uint32)
    cl = np.ascontiguousarray(context_lens, dtype=np.uint32)
    pos = np.ascontiguousarray(positions, dtype=np.int64)
    sl = np.ascontiguousarray(slots, dtype=np.int64)
    bs = layers[0]["k_cache"].shape[1]
    lib().ref_decode_layers(_p(x), arr, len(layers), _p(cos), _p(sin), _p(pos), _p(sl), _p(bt),
                            _p(cl), B, hidden, H, Hkv, D, inter, bs, bt.shape[1], group,
                            C.c_float(eps))
    return x
