"""Integer-state oracle: Sequence, BlockManager, Scheduler, prepare_decode / prepare_prefill.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Line-for-line Python restatement of

  /root/reference/src/core/sequence.rs:150-238          Sequence
  /root/reference/src/core/block_manager.rs:51-299,346-442,552-604,645-697   BlockManager (prefix cache, no swap)
  /root/reference/src/core/prefix_cache.rs              PrefixCache (hash chain of full blocks, leaf LRU)
  /root/reference/src/core/scheduler.rs:186-380         add, schedule
  /root/reference/src/core/scheduler.rs:500-654         postprocess, clear_finished
  /root/reference/src/core/scheduler.rs:718-785         filter_prefill_finished
  /root/reference/src/core/runner.rs:952-1388           prepare_block_tables / prefill / decode

Written independently of the C++ mirror (vllm_rs_b200/csrc/host/scheduler.cpp) -- plain
Python containers, same control flow as the Rust -- so that the two can be compared.
"""
from __future__ import annotations

from collections import deque
from dataclasses import dataclass, field

WAITING, RUNNING, FINISHED, CACHED, SWAPPED, FINISH_SWAPPED = 0, 1, 2, 3, 4, 5    # sequence.rs:8-15
SWAP_COOLING_PERIOD = 5000               # ms, scheduler.rs:49
MIN_KVCACHE_TOKENS_LEFT_FOR_SWAP = 1000  # scheduler.rs:50
MIN_NUM_SCHEDULED_REQS = 5        # scheduler.rs:44
KVCACHE_SWAP_THRESHOLD = 0.95     # scheduler.rs:48
PREFIX_CACHE_RATIO_NORMAL = 0.65  # scheduler.rs:55
PREFIX_CACHE_PRESSURE_EVICT_PERCENT = 0.1   # scheduler.rs:58


def prefix_cache_max_blocks(num_blocks, block_size, max_tokens=None):
    """build_prefix_cache_config (scheduler.rs:60-104) for a plain (non-PD) engine with the cache enabled."""
    import numpy as _np
    n = max_tokens // block_size if max_tokens is not None else int(_np.float32(num_blocks) * _np.float32(PREFIX_CACHE_RATIO_NORMAL))
    return min(n, num_blocks)
CHUNK_SIZE_CUDA = 8192            # scheduler.rs:203 / runner.rs:989


@dataclass
class Sequence:
    """sequence.rs:32-61, 150-238."""

    token_ids: list
    block_size: int
    max_tokens: int = 16384          # sampling_params.max_tokens.unwrap_or(16384)
    id: int = 0
    status: int = WAITING
    output_ids: list = field(default_factory=list)
    block_table: list = field(default_factory=list)
    num_cached_tokens: int = 0
    last_token: int = 0
    swapped_time: object = None      # ms since the epoch of the last swap, sequence.rs
    stop_token_ids: object = None    # sampling_params.stop_token_ids: list of token sequences (or None)

    def __post_init__(self):
        self.token_ids = list(self.token_ids)
        self.last_token = self.token_ids[-1] if self.token_ids else 0   # sequence.rs:180

    def __len__(self):                       # sequence.rs:189-195
        return len(self.token_ids)

    def num_blocks(self):                    # sequence.rs:207-209  len.div_ceil(block_size)
        return -(-len(self.token_ids) // self.block_size)

    def last_block_num_tokens(self):         # sequence.rs:211-213
        return len(self.token_ids) - (self.num_blocks() - 1) * self.block_size

    def num_cached_blocks(self):             # sequence.rs:215-217
        return self.num_cached_tokens // self.block_size

    def append_token(self, token):           # sequence.rs:219-223
        self.token_ids.append(token)
        self.output_ids.append(token)
        self.last_token = token

    def clone(self):
        c = Sequence(self.token_ids, self.block_size, self.max_tokens)
        c.id, c.status = self.id, self.status
        c.output_ids = list(self.output_ids)
        c.block_table = list(self.block_table)
        c.num_cached_tokens = self.num_cached_tokens
        c.last_token = self.last_token
        c.swapped_time = self.swapped_time
        c.stop_token_ids = self.stop_token_ids
        return c



M64 = (1 << 64) - 1


def siphash(data: bytes, k0: int = 0, k1: int = 0, c_rounds: int = 1, d_rounds: int = 3) -> int:
    """SipHash-c-d (Aumasson & Bernstein).  Rust's `DefaultHasher::new()` -- what prefix_cache.rs:337-349
    hashes with -- is SipHash-1-3 with the all-zero key (std is not part of /root/reference; this is the
    published algorithm, pinned in tests/ against the paper's SipHash-2-4 vector)."""
    v0 = k0 ^ 0x736F6D6570736575
    v1 = k1 ^ 0x646F72616E646F6D
    v2 = k0 ^ 0x6C7967656E657261
    v3 = k1 ^ 0x7465646279746573

    def rotl(x, b):
        return ((x << b) | (x >> (64 - b))) & M64

    def rnd(v0, v1, v2, v3):
        v0 = (v0 + v1) & M64; v1 = rotl(v1, 13); v1 ^= v0; v0 = rotl(v0, 32)
        v2 = (v2 + v3) & M64; v3 = rotl(v3, 16); v3 ^= v2
        v0 = (v0 + v3) & M64; v3 = rotl(v3, 21); v3 ^= v0
        v2 = (v2 + v1) & M64; v1 = rotl(v1, 17); v1 ^= v2; v2 = rotl(v2, 32)
        return v0, v1, v2, v3

    n = len(data)
    for off in range(0, n - n % 8, 8):
        m = int.from_bytes(data[off:off + 8], "little")
        v3 ^= m
        for _ in range(c_rounds):
            v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
        v0 ^= m
    b = ((n & 0xFF) << 56) | int.from_bytes(data[n - n % 8:], "little")
    v3 ^= b
    for _ in range(c_rounds):
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    v0 ^= b
    v2 ^= 0xFF
    for _ in range(d_rounds):
        v0, v1, v2, v3 = rnd(v0, v1, v2, v3)
    return v0 ^ v1 ^ v2 ^ v3


def hash_block(parent_hash: int, tokens) -> int:
    """prefix_cache.rs:337-342: `parent_hash.hash(h); tokens.hash(h)` -- u64 LE, then the slice as its
    length (usize LE) followed by the raw little-endian u32s."""
    data = parent_hash.to_bytes(8, "little") + len(tokens).to_bytes(8, "little")
    data += b"".join(int(t).to_bytes(4, "little") for t in tokens)
    return siphash(data)


class PrefixCache:
    """prefix_cache.rs:38-350 (no image seeds): entries hash -> [parent, block_id, children, access_id],
    the set of leaves and a lazily-invalidated LRU queue of (leaf hash, access id)."""

    def __init__(self, block_size, max_cached_blocks):
        self.block_size = block_size
        self.max_cached_blocks = max_cached_blocks
        self.entries = {}
        self.leaf_set = set()
        self.leaf_lru = deque()
        self.access_counter = 0

    def enabled(self):                                          # :59-61
        return self.max_cached_blocks > 0

    def cached_blocks(self):
        return len(self.entries)

    def _next_access_id(self):                                  # :332-335
        self.access_counter = (self.access_counter + 1) & M64
        return self.access_counter

    def _compact_lru_if_needed(self):                           # :316-330
        threshold = max(len(self.entries), 64) * 4
        if len(self.leaf_lru) <= threshold:
            return
        self.leaf_lru = deque((h, a) for h, a in self.leaf_lru
                              if h in self.leaf_set and h in self.entries and self.entries[h][3] == a)

    def _touch_leaf(self, h):                                   # :307-314
        if h in self.leaf_set and h in self.entries:
            self.leaf_lru.append((h, self.entries[h][3]))
        self._compact_lru_if_needed()

    def _touch(self, h):                                        # :297-305
        if h in self.entries:
            self.entries[h][3] = self._next_access_id()
            self._touch_leaf(h)

    def match_prefix(self, tokens):                             # :67-116 -> (matched_blocks, last_hash)
        if not self.enabled():
            return 0, None
        full_blocks = len(tokens) // self.block_size
        matched, parent, last = 0, 0, None
        for i in range(full_blocks):
            h = hash_block(parent, tokens[i * self.block_size:(i + 1) * self.block_size])
            if h in self.entries:
                matched += 1
                parent = h
                last = h
                self._touch(h)
            else:
                break
        return matched, last

    def blocks_for_match(self, last_hash):                      # :118-132
        blocks, cur = [], last_hash
        while cur is not None and cur in self.entries:
            blocks.append(self.entries[cur][1])
            cur = self.entries[cur][0]
        return blocks[::-1]

    def insert_prefix(self, tokens, blocks):                    # :174-262 -> (inserted, evicted)
        if not self.enabled():
            return [], []
        max_blocks = min(len(tokens) // self.block_size, len(blocks))
        if max_blocks == 0:
            return [], []
        inserted, parent = [], None
        for i in range(max_blocks):
            base = parent if parent is not None else 0
            h = hash_block(base, tokens[i * self.block_size:(i + 1) * self.block_size])
            if h in self.entries:
                self.entries[h][3] = self._next_access_id()
                self._touch_leaf(h)
            else:
                if parent is not None and parent in self.entries:
                    pe = self.entries[parent]
                    if pe[2] == 0:
                        self.leaf_set.discard(parent)
                    pe[2] += 1
                a = self._next_access_id()
                self.entries[h] = [parent, blocks[i], 0, a]
                self.leaf_set.add(h)
                self.leaf_lru.append((h, a))
                inserted.append(blocks[i])
            parent = h
        excess = max(len(self.entries) - self.max_cached_blocks, 0)
        evicted = self.evict_blocks(excess) if excess > 0 else []
        return inserted, evicted

    def evict_blocks(self, num_blocks):                         # :264-295
        evicted = []
        while num_blocks > 0:
            if not self.leaf_lru:
                break
            h, a = self.leaf_lru.popleft()
            if h not in self.leaf_set or h not in self.entries:
                continue
            e = self.entries[h]
            if e[3] != a or e[2] > 0:
                continue
            del self.entries[h]
            self.leaf_set.discard(h)
            evicted.append(e[1])
            num_blocks -= 1
            if e[0] is not None and e[0] in self.entries:
                pe = self.entries[e[0]]
                if pe[2] > 0:
                    pe[2] -= 1
                if pe[2] == 0:
                    self.leaf_set.add(e[0])
                    self.leaf_lru.append((e[0], pe[3]))
        return evicted


class BlockManager:
    """block_manager.rs:51-299,346-442,552-697 without CPU swap or mamba state."""

    def __init__(self, num_blocks, block_size, prefix_cache_blocks=0):
        self.ref_count = [0] * num_blocks                      # block_manager.rs:62-68
        self.free_block_ids = deque(range(num_blocks))
        self.used_block_ids = set()
        self.block_size = block_size
        self.num_blocks = num_blocks
        # :80-84: the cache exists only when enabled with a positive capacity
        self.prefix_cache = PrefixCache(block_size, prefix_cache_blocks) if prefix_cache_blocks > 0 else None
        # CPU swap space (block_manager.rs:69-78): ids 0..n-1, FIFO; swapped_map: seq id -> its CPU block ids
        self.free_cpu_block_ids = deque()
        self.swapped_map = {}
        self.swap_ops = []          # (swap_in, [(src, dst), ...]) for the host to execute with swap_blocks

    def enable_cpu_swap(self, num_cpu_blocks):
        self.free_cpu_block_ids = deque(range(num_cpu_blocks))

    # ---- CPU swap (block_manager.rs:255-272, 878-1000)
    def ensure_allocate(self, seq):                            # :255-272
        for _ in range(len(seq.block_table), seq.num_blocks()):
            if not self.free_block_ids:
                raise RuntimeError("No free blocks available, retry later!")
            b = self.free_block_ids.popleft()
            self._allocate_block(b)
            seq.block_table.append(b)

    def can_swap_out(self, seq):                               # :878-894
        if any(self.ref_count[b] > 1 for b in seq.block_table):
            return False
        return len(self.free_cpu_block_ids) > seq.num_blocks()

    def can_swap_in(self, seq):                                # :898-906
        return len(self.free_block_ids) > seq.num_blocks()

    def swap_out(self, seq, now_ms):                           # :910-957
        n = len(seq.block_table)
        if len(self.free_cpu_block_ids) < n:
            raise RuntimeError("Not enough CPU swap blocks")
        cpu_ids = [self.free_cpu_block_ids.popleft() for _ in range(n)]
        self.swap_ops.append((False, list(zip(seq.block_table, cpu_ids))))
        seq.swapped_time = now_ms
        self.swapped_map[seq.id] = cpu_ids

    def free_cpu_swap_for_seq(self, seq_id):                   # :993-1000
        for b in self.swapped_map.pop(seq_id, []):
            self.free_cpu_block_ids.append(b)

    def swap_in(self, seq):                                    # :959-990
        if seq.id not in self.swapped_map:
            raise RuntimeError("No CPU-swap entry for seq")
        cpu_ids = self.swapped_map.pop(seq.id)
        if len(cpu_ids) > len(seq.block_table):
            self.swapped_map[seq.id] = cpu_ids
            self.free_cpu_swap_for_seq(seq.id)
            raise RuntimeError("Insufficient GPU blocks to swap in sequence")
        self.swap_ops.append((True, [(c, seq.block_table[i]) for i, c in enumerate(cpu_ids)]))
        for c in cpu_ids:
            self.free_cpu_block_ids.append(c)

    def evict_prefix_cache_until_free(self, min_free_blocks):  # :674-697
        total = 0
        while len(self.free_block_ids) < min_free_blocks:
            if self.prefix_cache is None:
                break
            evicted = self.prefix_cache.evict_blocks(1)
            if not evicted:
                break
            total += len(evicted)
            for b in evicted:
                self._decrement_block_ref(b)
        return total

    def _allocate_block(self, block_id):                       # block_manager.rs:113-120
        assert self.ref_count[block_id] == 0
        self.ref_count[block_id] = 1
        self.used_block_ids.add(block_id)

    def _adjusted_matched_blocks(self, tokens_len, matched):    # :291-299: never reuse the WHOLE prompt
        full_blocks = tokens_len // self.block_size
        if matched == full_blocks and tokens_len % self.block_size == 0 and matched > 0:
            return matched - 1
        return matched

    def required_blocks(self, seq):                            # :179-200 (the match touches the LRU)
        if self.prefix_cache is not None:
            matched, _ = self.prefix_cache.match_prefix(seq.token_ids)
            return max(seq.num_blocks() - self._adjusted_matched_blocks(len(seq), matched), 0)
        return seq.num_blocks()

    def can_allocate(self, seq):                               # block_manager.rs:202-204
        return len(self.free_block_ids) >= self.required_blocks(seq)

    def _increment_block_ref(self, b):                         # :275-282
        if self.ref_count[b] == 0:
            if b in self.free_block_ids:
                self.free_block_ids.remove(b)
            self.used_block_ids.add(b)
        self.ref_count[b] += 1

    def _decrement_block_ref(self, b):                         # :283-289, 139-144
        self.ref_count[b] = max(self.ref_count[b] - 1, 0)
        if self.ref_count[b] == 0 and b in self.used_block_ids:
            self.used_block_ids.remove(b)
            self.free_block_ids.append(b)

    def _allocate_with_prefix(self, seq):                      # :346-442
        pc = self.prefix_cache
        matched, last_hash = pc.match_prefix(seq.token_ids)
        matched = self._adjusted_matched_blocks(len(seq), matched)
        if matched > 0 and last_hash is not None:
            for b in pc.blocks_for_match(last_hash)[:matched]:
                self._increment_block_ref(b)
                seq.block_table.append(b)
        seq.num_cached_tokens = matched * self.block_size
        for _ in range(len(seq.block_table), seq.num_blocks()):
            if not self.free_block_ids:
                raise RuntimeError("No free blocks available, retry later!")
            b = self.free_block_ids.popleft()
            self._allocate_block(b)
            seq.block_table.append(b)

    def cache_sequence(self, seq):                             # :552-604
        pc = self.prefix_cache
        if pc is None or not pc.enabled():
            return
        full_blocks = len(seq) // self.block_size
        if full_blocks == 0 or len(seq.block_table) < full_blocks:
            return
        inserted, evicted = pc.insert_prefix(seq.token_ids, seq.block_table[:full_blocks])
        for b in inserted:
            self._increment_block_ref(b)
        for b in evicted:
            self._decrement_block_ref(b)

    def evict_prefix_cache_blocks(self, num_blocks):           # :657-672
        if self.prefix_cache is None or num_blocks == 0:
            return 0
        evicted = self.prefix_cache.evict_blocks(num_blocks)
        for b in evicted:
            self._decrement_block_ref(b)
        return len(evicted)

    def prefix_cache_blocks(self):                             # :612-616
        return self.prefix_cache.cached_blocks() if self.prefix_cache is not None else 0

    def allocate(self, seq):                                   # :210-221; allocate_fresh :122-137
        assert not seq.block_table
        if self.prefix_cache is not None:
            return self._allocate_with_prefix(seq)
        seq.num_cached_tokens = 0
        for _ in range(seq.num_blocks()):
            if not self.free_block_ids:
                raise RuntimeError("No free blocks available, retry later!")
            b = self.free_block_ids.popleft()
            self._allocate_block(b)
            seq.block_table.append(b)

    def deallocate(self, seq):                                 # :227-231
        for b in reversed(seq.block_table):
            self._decrement_block_ref(b)

    def can_append(self, seq):                                 # :233-239
        need = 1
        if len(seq) % self.block_size != 0:
            need += 1
        return len(self.free_block_ids) >= need

    def may_append(self, seq):                                 # :241-253
        if len(seq) % self.block_size == 1:
            if not self.free_block_ids:
                raise RuntimeError("No free blocks available, retry later!")
            b = self.free_block_ids.popleft()
            self._allocate_block(b)
            seq.block_table.append(b)


class Scheduler:
    """scheduler.rs (CUDA build, no PD / swap / tool-call handling)."""

    def __init__(self, num_blocks, block_size=64, max_num_seqs=32, max_num_batched_tokens=None,
                 chunk_size=CHUNK_SIZE_CUDA, eos_token_ids=(), prefix_cache_blocks=0):
        self.waiting = deque()
        self.running = []
        self.block_manager = BlockManager(num_blocks, block_size, prefix_cache_blocks)
        self.cached = []            # swapped-out sequences (scheduler.rs `cached`; session caching is out of scope)
        self.now_ms = 0             # the host's clock (SystemTime::now() in the reference), set before schedule()
        self.next_seq_id = 0
        self.eos_token_id = list(eos_token_ids)
        self.block_size = block_size
        self.max_num_seqs = max_num_seqs
        # kvcache_allocator.rs:668: max_num_batched_tokens = num_blocks * block_size
        self.max_num_batched_tokens = (max_num_batched_tokens if max_num_batched_tokens
                                       else num_blocks * block_size)
        self.chunk_size = chunk_size
        self.is_last_prefill = False

    def add(self, seq):                                        # scheduler.rs:186-192
        seq.id = self.next_seq_id
        self.next_seq_id += 1
        self.waiting.append(seq)
        return seq.id

    def schedule(self):                                        # scheduler.rs:200-380
        scheduled_ids = []
        num_tokens = 0
        pre_existing_running = len(self.running)
        max_seqs_limit = max(self.max_num_seqs, MIN_NUM_SCHEDULED_REQS)
        while self.waiting:
            seq = self.waiting.popleft()
            effective_tokens = min(self.chunk_size, len(seq) - seq.num_cached_tokens)
            if (len(self.running) >= max_seqs_limit
                    or len(scheduled_ids) >= max(self.max_num_seqs, MIN_NUM_SCHEDULED_REQS)
                    or num_tokens + effective_tokens >= self.max_num_batched_tokens - 1
                    or (not seq.block_table and not self.block_manager.can_allocate(seq))
                    or (self.is_last_prefill and pre_existing_running > 0)):
                self.waiting.appendleft(seq)
                break
            if not seq.block_table:
                self.block_manager.allocate(seq)
            seq.status = RUNNING
            num_tokens += effective_tokens
            self.running.append(seq)
            scheduled_ids.append(len(self.running) - 1)
        if scheduled_ids:
            self.is_last_prefill = True
            return scheduled_ids, True
        # :288-337: sequences that cannot append are "preempt" candidates.  With nothing to preempt and room
        # in the cache, ONE swapped-out sequence may come back (try_swap_in); otherwise, under pressure, 10 % of
        # the prefix cache (at least one block) is evicted, and if there is nothing to evict the oldest
        # preempt candidate is swapped out to CPU memory (only when another sequence keeps running).
        import numpy as _np
        f32 = _np.float32
        preempt = [i for i, q in enumerate(self.running)
                   if not self.block_manager.can_append(q) and q.status not in (SWAPPED, FINISH_SWAPPED)]
        bm = self.block_manager
        usage = f32(1.0) - (f32(len(bm.free_block_ids)) * f32(1.0) / f32(bm.num_blocks))        # :1255-1259
        swap_in = (not preempt) and (usage < f32(KVCACHE_SWAP_THRESHOLD) * f32(0.9)
                                     or (not self.running and usage <= f32(0.3)))
        if swap_in:
            self.try_swap_in()
        elif preempt or usage > f32(KVCACHE_SWAP_THRESHOLD):
            evicted = 0
            cached = bm.prefix_cache_blocks()                                                   # :1167-1175
            if cached > 0:
                blocks = max(int(_np.ceil(f32(cached) * f32(PREFIX_CACHE_PRESSURE_EVICT_PERCENT))), 1)
                evicted = bm.evict_prefix_cache_blocks(blocks)
            if evicted == 0 and preempt and len(self.running) > 1:
                idx = min(preempt, key=lambda i: self.running[i].id)                            # the oldest
                self.try_swap_out(idx)
        decode_ids = []
        decode_max_seqs = max(self.max_num_seqs, MIN_NUM_SCHEDULED_REQS)
        for idx, seq in enumerate(self.running):
            if len(decode_ids) >= decode_max_seqs:
                break
            if not self.block_manager.can_append(seq):
                continue
            self.block_manager.may_append(seq)
            decode_ids.append(idx)
        self.is_last_prefill = False
        return decode_ids, False

    def try_swap_out(self, idx):                               # scheduler.rs:904-954 (running sequences only)
        if idx >= len(self.running):
            return False
        seq = self.running[idx]
        bm = self.block_manager
        if seq.block_table and seq.status == RUNNING and bm.can_swap_out(seq):
            try:
                bm.ensure_allocate(seq)         # same number of blocks when it comes back
            except RuntimeError:
                return False
            try:
                bm.swap_out(seq, self.now_ms)
            except RuntimeError:
                return False
            seq = self.running.pop(idx)
            seq.status = SWAPPED
            bm.deallocate(seq)
            self.cached.append(seq.clone())     # the block table is rebuilt at swap-in
            return True
        return False

    def try_swap_in(self):                                     # scheduler.rs:830-901
        cur = self.now_ms
        bm = self.block_manager
        i = 0
        while i < len(self.cached):
            seq = self.cached[i]
            st = seq.swapped_time if seq.swapped_time is not None else cur
            if seq.status != SWAPPED or cur - st < SWAP_COOLING_PERIOD:
                i += 1
                continue
            available = len(bm.free_block_ids) * self.block_size
            if not bm.can_swap_in(seq) or available - len(seq) < MIN_KVCACHE_TOKENS_LEFT_FOR_SWAP:
                if self.running:
                    i += 1
                    continue                    # wait for a later step
                if bm.evict_prefix_cache_until_free(seq.num_blocks() + 1) > 0:
                    break
                seq = self.cached.pop(i)
                seq.status = FINISHED           # "No KvCache left for swap in"
                self.running.append(seq)
                break
            seq = self.cached.pop(i)
            seq.swapped_time = cur
            seq.block_table = []
            try:
                bm.ensure_allocate(seq)
            except RuntimeError:
                continue                        # (the reference drops the sequence here as well)
            try:
                bm.swap_in(seq)
                seq.status = RUNNING
            except RuntimeError:
                seq.status = FINISHED
            self.running.append(seq)
            break

    def filter_prefill_finished(self, scheduled_ids):          # scheduler.rs:718-785
        finished_seqs, remove_ids = [], []
        C = self.chunk_size
        for i, idx in enumerate(scheduled_ids):
            if idx < len(self.running):
                seq = self.running[idx]
                if len(seq) < C or seq.num_cached_tokens + C >= len(seq):
                    finished_seqs.append((i, seq.id))
                else:
                    remove_ids.append(seq.id)
                    c = seq.clone()
                    c.num_cached_tokens += C
                    c.status = WAITING
                    self.waiting.append(c)
        self.running = [s for s in self.running if s.id not in remove_ids]
        indices = [i for i, _ in finished_seqs]
        finished_indices = []
        for _, sid in finished_seqs:
            for pos, s in enumerate(self.running):
                if s.id == sid:
                    finished_indices.append(pos)
                    break
        return indices, finished_indices

    def postprocess(self, ids, output_ids):                    # scheduler.rs:500-629
        for i, idx in enumerate(ids):
            if idx >= len(self.running):
                continue
            seq = self.running[idx]
            token = output_ids[i]
            if (stop_sequence_match_index(token, seq) is not None
                    or token in self.eos_token_id
                    or len(seq.output_ids) >= seq.max_tokens
                    or len(seq) > self.max_num_batched_tokens):
                seq.status = FINISHED
                self.block_manager.cache_sequence(seq)                  # :617-620: cache, then drop the refs
                self.block_manager.deallocate(seq)
            else:
                seq.append_token(token)

    def clear_finished(self):                                  # scheduler.rs:631-654
        self.running = [s for s in self.running if s.status != FINISHED]
        self.waiting = deque(s for s in self.waiting if s.status != FINISHED)


def stop_sequence_match_index(token, seq):                     # scheduler.rs:1309-1341
    """Index of the first stop sequence completed by `token`: a one-token stop matches the token itself, a longer
    one needs its head to equal the tail of the tokens generated so far (output_ids, prompt excluded)."""
    stops = seq.stop_token_ids
    if not stops:
        return None
    for idx, stop in enumerate(stops):
        if not stop:
            continue
        if len(stop) == 1:
            if stop[0] == token:
                return idx
            continue
        prior = len(seq.output_ids)
        if len(stop) - 1 > prior:
            continue
        start = prior + 1 - len(stop)
        if seq.output_ids[start:] == list(stop[:-1]) and stop[-1] == token:
            return idx
    return None


def prepare_block_tables(seqs):                                # runner.rs:952-974
    max_len = max((len(s.block_table) for s in seqs), default=0)
    return [list(s.block_table) + [0] * (max_len - len(s.block_table)) for s in seqs], max_len


def prepare_decode(seqs, block_size):                          # runner.rs:1243-1280
    input_ids, positions, slot_mapping, context_lens = [], [], [], []
    for s in seqs:
        input_ids.append(s.last_token)
        positions.append(len(s) - 1)
        context_lens.append(len(s))
        slot = s.block_table[-1] * block_size + s.last_block_num_tokens() - 1
        slot_mapping.append(slot)
    block_tables, max_blocks = prepare_block_tables(seqs)
    return dict(input_ids=input_ids, positions=positions, slot_mapping=slot_mapping,
                context_lens=context_lens, block_tables=block_tables, max_blocks=max_blocks,
                max_context_len=max(context_lens))


def prepare_prefill(seqs, block_size, chunk_size=CHUNK_SIZE_CUDA):   # runner.rs:978-1082
    input_ids, positions, slot_mapping = [], [], []
    cu_q, cu_k = [0], [0]
    max_q = max_k = max_ctx = 0
    prefill_tokens = []
    for s in seqs:
        seqlen = len(s)
        num_tokens = min(chunk_size, seqlen - s.num_cached_tokens)
        input_ids += s.token_ids[s.num_cached_tokens:s.num_cached_tokens + num_tokens]
        positions += list(range(s.num_cached_tokens, s.num_cached_tokens + num_tokens))
        prefill_tokens.append(num_tokens)
        seqlen_q = num_tokens
        seqlen_k = s.num_cached_tokens + num_tokens if s.num_cached_tokens > 0 else num_tokens
        max_ctx = max(max_ctx, s.num_cached_tokens + num_tokens)
        cu_q.append(cu_q[-1] + seqlen_q)
        cu_k.append(cu_k[-1] + seqlen_k)
        max_q, max_k = max(max_q, seqlen_q), max(max_k, seqlen_k)
        emitted = 0
        for i in range(s.num_cached_blocks(), s.num_blocks()):
            start = s.block_table[i] * block_size
            if i == s.num_cached_blocks():
                start += s.num_cached_tokens % block_size
            end = start + min(num_tokens - emitted, block_size)
            slot_mapping += list(range(start, end))
            emitted += end - start
            if emitted >= num_tokens:
                break
    assert len(input_ids) == len(slot_mapping) == cu_q[-1]
    out = dict(input_ids=input_ids, positions=positions, slot_mapping=slot_mapping,
               cu_seqlens_q=cu_q, cu_seqlens_k=cu_k, max_seqlen_q=max_q, max_seqlen_k=max_k,
               max_context_len=max_ctx, block_tables=None, context_lens=None, max_blocks=0)
    if cu_k[-1] > cu_q[-1]:
        out["block_tables"], out["max_blocks"] = prepare_block_tables(seqs)
        out["context_lens"] = [s.num_cached_tokens + n for s, n in zip(seqs, prefill_tokens)]
    return out


def kvcache_plan(available, num_layers, num_kv_heads, head_dim, num_shards=1, block_size=64, dtype_size=2,
                 config_model_len=8192, user_max_num_seqs=None, user_max_model_len=None, cpu_mem_fold=0.5):
    """KVCacheAllocator::plan_allocation for dense models (kvcache_allocator.rs:126-141, 447-468,
    553-612, 616-707).  Raises ValueError('invalid') / MemoryError like the two KVCacheError arms."""
    shards = max(num_shards, 1)
    heads = num_kv_heads // shards if num_kv_heads >= shards else 1       # :126-137
    per_block = block_size * heads * head_dim * dtype_size * 2 * num_layers   # :460-466
    if per_block == 0:
        raise ValueError("invalid")
    total_blocks = available // per_block
    if user_max_num_seqs and user_max_model_len:                       # :636-648
        blocks_per_seq = -(-user_max_model_len // block_size)
        if user_max_num_seqs * blocks_per_seq * per_block > available:
            raise MemoryError
        max_seqs, max_len = user_max_num_seqs, user_max_model_len
    else:                                                              # auto_decide_params :553-612
        if total_blocks == 0:
            raise MemoryError
        total_tokens = total_blocks * block_size
        if total_tokens <= config_model_len:
            max_seqs, max_len = 1, total_tokens
        else:
            max_seqs, max_len = 1, total_tokens
            for cand in (config_model_len, config_model_len // 2, config_model_len // 4, config_model_len // 8,
                         16 * 1024, 8 * 1024, 4 * 1024, 1024):
                if cand == 0:
                    continue
                bps = -(-cand // block_size)
                if total_blocks >= bps:
                    max_seqs, max_len = min(total_blocks // bps, 8), cand
                    break
    if total_blocks == 0 or max_seqs == 0:
        raise MemoryError
    import numpy as _np
    return dict(per_block_bytes=per_block, num_gpu_blocks=total_blocks,
                num_cpu_blocks=int(_np.float32(total_blocks) * _np.float32(cpu_mem_fold)),   # `as f32 * fold` :681
                max_num_seqs=max_seqs, max_model_len=max_len, kvcache_memory_bytes=total_blocks * per_block,
                max_num_batched_tokens=total_blocks * block_size)
