"""AWQ / GPTQ checkpoint formats: pack, unpack, dequantise (numpy).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Formats as implied by the reference loader
/root/reference/src/models/layers/wna16.rs:56-83,127-152 (tensor names / shapes / shard dims)
and the public AutoGPTQ / AutoAWQ packing:

  GPTQ  qweight i32 [K/pack, N]  element (k,n) = (w[k//pack, n] >> bits*(k%pack)) & mask
        qzeros  i32 [K/G, N/pack] zero(g,n)    = ((z[g, n//pack] >> bits*(n%pack)) & mask) + 1
        scales  f16 [K/G, N]
  AWQ   qweight i32 [K, N/8]     nibble j of a word holds output 8c + AWQ_ORDER[j]
        qzeros  i32 [K/G, N/8]   same nibble order, no +1
        scales  f16 [K/G, N]
  w[k, n] = fp16((q - z) * s)   -- dequantised weight is materialised in fp16 (one rounding),
  which is what a register-dequant tensor-core kernel (Marlin in the reference) feeds the MMA.
"""
from __future__ import annotations

import numpy as np

AWQ_ORDER = [0, 2, 4, 6, 1, 3, 5, 7]


def pack_gptq(q: np.ndarray, bits: int) -> np.ndarray:
    """q uint [K, N] -> i32 [K/pack, N] (pack along K)."""
    pack = 32 // bits
    K, N = q.shape
    assert K % pack == 0
    out = np.zeros((K // pack, N), dtype=np.uint32)
    for j in range(pack):
        out |= q[j::pack].astype(np.uint32) << np.uint32(bits * j)
    return out.view(np.int32)


def unpack_gptq(qw: np.ndarray, bits: int) -> np.ndarray:
    pack = 32 // bits
    w = qw.view(np.uint32)
    Kp, N = w.shape
    out = np.zeros((Kp * pack, N), dtype=np.uint8)
    mask = np.uint32((1 << bits) - 1)
    for j in range(pack):
        out[j::pack] = ((w >> np.uint32(bits * j)) & mask).astype(np.uint8)
    return out


def pack_gptq_zeros(z: np.ndarray, bits: int) -> np.ndarray:
    """z int [K/G, N] (true zero point, 1..2^bits) -> i32 [K/G, N/pack] storing z - 1."""
    pack = 32 // bits
    G, N = z.shape
    zm = (z.astype(np.int64) - 1).astype(np.uint32) & np.uint32((1 << bits) - 1)
    out = np.zeros((G, N // pack), dtype=np.uint32)
    for j in range(pack):
        out |= zm[:, j::pack] << np.uint32(bits * j)
    return out.view(np.int32)


def unpack_gptq_zeros(qz: np.ndarray, bits: int) -> np.ndarray:
    pack = 32 // bits
    w = qz.view(np.uint32)
    G, Np = w.shape
    out = np.zeros((G, Np * pack), dtype=np.int32)
    mask = np.uint32((1 << bits) - 1)
    for j in range(pack):
        out[:, j::pack] = ((w >> np.uint32(bits * j)) & mask).astype(np.int32) + 1
    return out


def pack_awq(q: np.ndarray) -> np.ndarray:
    """q uint4 [R, N] -> i32 [R, N/8] in AWQ nibble order (used for qweight and qzeros)."""
    R, N = q.shape
    assert N % 8 == 0
    out = np.zeros((R, N // 8), dtype=np.uint32)
    for j, o in enumerate(AWQ_ORDER):
        out |= q[:, o::8].astype(np.uint32) << np.uint32(4 * j)
    return out.view(np.int32)


def unpack_awq(qw: np.ndarray) -> np.ndarray:
    w = qw.view(np.uint32)
    R, Np = w.shape
    out = np.zeros((R, Np * 8), dtype=np.uint8)
    for j, o in enumerate(AWQ_ORDER):
        out[:, o::8] = ((w >> np.uint32(4 * j)) & np.uint32(0xF)).astype(np.uint8)
    return out


def dequant(q: np.ndarray, z: np.ndarray, scales: np.ndarray, group_size: int) -> np.ndarray:
    """q [K, N] ints, z [K/G, N] ints, scales [K/G, N] f16 -> W [K, N] fp16."""
    K, N = q.shape
    g = np.arange(K) // group_size
    d = (q.astype(np.int32) - z[g].astype(np.int32)).astype(np.float32)
    return (d * scales[g].astype(np.float32)).astype(np.float16)


def dequant_awq(qweight, qzeros, scales, group_size):
    return dequant(unpack_awq(qweight), unpack_awq(qzeros).astype(np.int32), scales, group_size)


def dequant_gptq(qweight, qzeros, scales, group_size, bits):
    q = unpack_gptq(qweight, bits)
    if qzeros is None:
        z = np.full((q.shape[0] // group_size, q.shape[1]), 1 << (bits - 1), dtype=np.int32)
    else:
        z = unpack_gptq_zeros(qzeros, bits)
    return dequant(q, z, scales, group_size)


def random_awq(rng: np.random.Generator, K: int, N: int, group_size: int = 128):
    """Synthetic AWQ layer per SURVEY 8(d) config 2: uniform nibbles, scales
    U(0.5,1.5) * (2/15) / sqrt(K) so activations stay O(1)."""
    q = rng.integers(0, 16, size=(K, N), dtype=np.uint8)
    z = rng.integers(0, 16, size=(K // group_size, N), dtype=np.uint8)
    s = (rng.uniform(0.5, 1.5, size=(K // group_size, N)) * (2.0 / 15.0) / np.sqrt(K)).astype(
        np.float16)
    return pack_awq(q), pack_awq(z), s


def random_gptq(rng: np.random.Generator, K: int, N: int, group_size: int = 128, bits: int = 4,
                sym: bool = False):
    top = 1 << bits
    q = rng.integers(0, top, size=(K, N), dtype=np.uint8 if bits <= 8 else np.uint16)
    if sym:
        z = np.full((K // group_size, N), top // 2, dtype=np.int32)
    else:
        z = rng.integers(1, top + 1, size=(K // group_size, N)).astype(np.int32)
    s = (rng.uniform(0.5, 1.5, size=(K // group_size, N)) * (2.0 / (top - 1)) / np.sqrt(K)).astype(
        np.float16)
    return pack_gptq(q, bits), pack_gptq_zeros(z, bits), s
