"""Build the C-ABI shared library (sm_100a only) in-tree with nvcc.

    python -m vllm_rs_b200.build          # incremental
    python -m vllm_rs_b200.build --force

Output: vllm_rs_b200/libb200_vllm.so (git-ignored; travels to the GPU box with gpurun).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
# B200_BUILD_TAG=<name>: a side build (objects under build/obj-<name>, library build/<name>/libb200_vllm.so)
_TAG = os.environ.get("B200_BUILD_TAG", "")
BUILD = ROOT / "build" / ("obj-" + _TAG if _TAG else "obj")
LIB = (ROOT / "build" / _TAG / "libb200_vllm.so") if _TAG else PKG / "libb200_vllm.so"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-I", str(ROOT / "include"),
] + os.environ.get("B200_NVCC_EXTRA", "").split()   # e.g. -DB200_DEC_TRACE (tools/dec_trace.py)


def sources() -> list[Path]:
    return sorted(list(CSRC.glob("*.cu")) + list((CSRC / "host").glob("*.cpp")))


def _deps_hash(src: Path) -> str:
    h = hashlib.sha1()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in [src] + sorted(CSRC.glob("*.cuh")) + sorted((CSRC / "host").glob("*.hpp")) + [
        ROOT / "include" / "b200_vllm.h"
    ]:
        h.update(p.read_bytes())
    return h.hexdigest()


def _compile(src: Path, force: bool) -> Path:
    obj = BUILD / (src.stem + ".o")
    stamp = BUILD / (src.stem + ".hash")
    want = _deps_hash(src)
    if not force and obj.exists() and stamp.exists() and stamp.read_text() == want:
        return obj
    cmd = [NVCC, *NVCC_FLAGS, "-x", "cu", "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
    if r.stderr.strip():
        sys.stderr.write(r.stderr)
    stamp.write_text(want)
    return obj


def build(force: bool = False, verbose: bool = True) -> Path:
    BUILD.mkdir(parents=True, exist_ok=True)
    LIB.parent.mkdir(parents=True, exist_ok=True)
    srcs = sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile(s, force), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if force or not LIB.exists() or LIB.stat().st_mtime < newest:
        cmd = [NVCC, "-shared", "-o", str(LIB), *map(str, objs),
               "-gencode", "arch=compute_100a,code=sm_100a",
               "-L/usr/local/cuda/lib64", "-lcublas", "-lcudart",
               "-Xlinker", "-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if verbose:
        print(f"built {LIB} ({LIB.stat().st_size >> 10} KiB) from {len(srcs)} sources")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
