"""Sampling oracle.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, in numpy / pure Python:
  * LogitsProcessor::get_strategy                 (/root/reference/src/utils/logits_processor.rs:48-70)
  * the (k, p, t) the CUDA sampler is called with (logits_processor.rs:199-231)
  * sample_topk_topp / sample_topk / sample_topp  (logits_processor.rs:84-196; the CPU branch, which is
    the only source-visible definition of "top-k then top-p then multinomial")
  * apply_penalties / apply_batch_repeat_penalty  (logits_processor.rs:286-343) and the "last 128
    tokens, only once more than 128 exist" window of ModelRunner::sample (core/runner.rs:1522-1536)
  * StdRng::seed_from_u64 + next_u64 (`LogitsProcessor::from_sampling`, logits_processor.rs:29-30; the
    per-step kernel seed is `rng.next_u64()`, :222-225).  StdRng lives in the un-vendored `rand` crate
    (0.9: ChaCha12, seed expanded from the u64 with PCG32) -- PUBLISHED algorithm restated here; its
    ChaCha core is pinned against the RFC 7539 section 2.3.2 block-function vector in tests/.

PARITY UNPINNED for the device draw: `attention_rs::sampler::Sampler::sample_cuda` (how a row's
uniform is derived from the step seed) is not source-visible, so `row_uniform` below is THIS repo's
definition; the reference's CPU branch draws rows under rayon in a nondeterministic order anyway.
"""
from __future__ import annotations

import numpy as np

M64 = (1 << 64) - 1


# ------------------------------------------------------------------ StdRng (rand 0.9 = ChaCha12)
def _rotl(x, n):
    return ((x << n) | (x >> (32 - n))) & 0xFFFFFFFF


def chacha_block(key_words, w12, w13, w14, w15, rounds):
    """One 64-byte ChaCha block as 16 little-endian u32 words (state words 12..15 given explicitly)."""
    st = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + list(key_words) + [w12, w13, w14, w15]
    x = list(st)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = _rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = _rotl(x[b] ^ x[c], 7)

    for _ in range(rounds // 2):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return [(a + b) & 0xFFFFFFFF for a, b in zip(x, st)]


def seed_from_u64(state):
    """rand_core SeedableRng::seed_from_u64: PCG32 (XSH-RR) stream -> 32-byte seed as 8 LE words."""
    MUL, INC = 6364136223846793005, 11634580027462260723
    words = []
    for _ in range(8):
        state = (state * MUL + INC) & M64
        xorshifted = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        words.append(((xorshifted >> rot) | (xorshifted << ((32 - rot) & 31))) & 0xFFFFFFFF)
    return words


class StdRng:
    """rand::rngs::StdRng::seed_from_u64(seed): ChaCha12, 64-bit block counter from 0, stream 0."""

    def __init__(self, seed):
        self.key = seed_from_u64(seed & M64)
        self.counter = 0
        self.buf, self.idx = [], 0

    def next_u32(self):
        if self.idx >= len(self.buf):
            self.buf = chacha_block(self.key, self.counter & 0xFFFFFFFF, self.counter >> 32, 0, 0, 12)
            self.counter += 1
            self.idx = 0
        v = self.buf[self.idx]
        self.idx += 1
        return v

    def next_u64(self):
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)


# ------------------------------------------------------------------ strategy selection
def get_strategy(temperature, top_k, top_p):
    """logits_processor.rs:48-70 -> ("argmax",) | ("all", t) | ("topk", k, t) | ("topp", p, t) | ("topk_topp", k, p, t)"""
    t = None if temperature is None or temperature < 1e-7 else temperature
    k = top_k if (top_k is not None and top_k > 0) else None
    if t is None or not t > 0.0:
        return ("argmax",)
    if k is None and top_p is None:
        return ("all", t)
    if top_p is None:
        return ("topk", k, t)
    if k is None:
        return ("topp", top_p, t)
    return ("topk_topp", k, top_p, t)


DEFAULT_STRATEGY = ("topk_topp", 32, 0.95, 0.7)          # core/runner.rs:1483-1487


def kernel_params(strategy):
    """(k, p, t) handed to the device sampler (logits_processor.rs:208-216); None = not a device strategy."""
    kind = strategy[0]
    if kind == "topk_topp":
        return strategy[1], strategy[2], strategy[3]
    if kind == "topk":
        return strategy[1], 1.0, strategy[2]
    if kind == "topp":
        return 256, strategy[1], strategy[2]
    return None


# ------------------------------------------------------------------ per-row uniform (this repo's definition)
def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def row_uniform(seed, row):
    """u in [0, 1) with 24 random bits for batch row `row` of the step whose seed is `seed`."""
    return np.float32((splitmix64((seed ^ ((row + 1) * 0xD1342543DE82EF95)) & M64) >> 40)) * np.float32(1.0 / (1 << 24))


# ------------------------------------------------------------------ the draw
def candidates(logits, k, t):
    """softmax(logits / t) over the whole vocabulary, top-k by probability (ties: lower index first).
    Returns (token ids, probabilities as float32) in descending order."""
    z = logits.astype(np.float32) / np.float32(t)
    z = z - z.max()
    e = np.exp(z.astype(np.float64))
    prs = (e / e.sum()).astype(np.float32)
    k = min(k, logits.shape[0])
    order = np.lexsort((np.arange(logits.shape[0]), -logits.astype(np.float64)))[:k]
    return order.astype(np.int64), prs[order]


def topp_clamp(prs, top_p):
    """sample_topk_topp's cut (logits_processor.rs:171-186): zero everything after the cumulative
    mass first reaches top_p -- on the UN-renormalised top-k probabilities."""
    prs = prs.copy()
    sum_p = np.float32(prs.sum(dtype=np.float32))
    if top_p <= 0.0 or top_p >= sum_p:
        return prs
    cumsum = np.float32(0.0)
    for i in range(len(prs)):
        if cumsum >= np.float32(top_p):
            prs[i] = 0.0
        else:
            cumsum = np.float32(cumsum + prs[i])
    return prs


def weighted_index(prs, u):
    """WeightedIndex::sample with x = u * total: first i whose cumulative weight exceeds x."""
    cum = np.cumsum(prs.astype(np.float64))
    x = float(u) * cum[-1]
    return int(min(np.searchsorted(cum, x, side="right"), len(prs) - 1))


def sample_row(logits, k, p, t, u):
    ids, prs = candidates(logits, k, t)
    prs = topp_clamp(prs, p)
    return int(ids[weighted_index(prs, u)])


def sample_batch(logits, k, p, t, seed):
    return [sample_row(logits[b], k, p, t, row_uniform(seed, b)) for b in range(logits.shape[0])]


# ------------------------------------------------------------------ penalties
def penalty_context(tokens):
    """core/runner.rs:1522-1533: the last 128 generated tokens, and only once MORE than 128 exist."""
    return list(tokens[-128:]) if len(tokens) > 128 else []


def apply_penalties(logits, context, frequency_penalty, presence_penalty):
    """logits_processor.rs:286-306 behind the guards of :325-329."""
    out = logits.astype(np.float32).copy()
    active = len(context) > 1 and ((frequency_penalty not in (1.0, 0.0)) or (presence_penalty not in (1.0, 0.0)))
    if not active:
        return out
    counts = np.zeros(out.shape[0], np.float32)
    for tok in context:
        if tok < out.shape[0]:
            counts[tok] += 1.0
    out = out - counts * np.float32(frequency_penalty) - (counts > 0).astype(np.float32) * np.float32(presence_penalty)
    return out.astype(np.float32)
