"""Integer-state oracle: Sequence, BlockManager, Scheduler, prepare_decode / prepare_prefill.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Line-for-line Python restatement of

  /root/reference/src/core/sequence.rs:150-238          Sequence
  /root/reference/src/core/block_manager.rs:51-299      BlockManager (no prefix cache / swap)
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

WAITING, RUNNING, FINISHED = 0, 1, 2
MIN_NUM_SCHEDULED_REQS = 5        # scheduler.rs:44
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
        return c


class BlockManager:
    """block_manager.rs:51-299 without prefix cache, CPU swap or mamba state."""

    def __init__(self, num_blocks, block_size):
        self.ref_count = [0] * num_blocks                      # block_manager.rs:62-68
        self.free_block_ids = deque(range(num_blocks))
        self.used_block_ids = set()
        self.block_size = block_size

    def _allocate_block(self, block_id):                       # block_manager.rs:113-120
        assert self.ref_count[block_id] == 0
        self.ref_count[block_id] = 1
        self.used_block_ids.add(block_id)

    def can_allocate(self, seq):                               # block_manager.rs:202-208
        return len(self.free_block_ids) >= seq.num_blocks()

    def allocate(self, seq):                                   # allocate_fresh, :122-137
        assert not seq.block_table
        seq.num_cached_tokens = 0
        for _ in range(seq.num_blocks()):
            if not self.free_block_ids:
                raise RuntimeError("No free blocks available, retry later!")
            b = self.free_block_ids.popleft()
            self._allocate_block(b)
            seq.block_table.append(b)

    def deallocate(self, seq):                                 # :227-231, 283-289, 139-144
        for b in reversed(seq.block_table):
            self.ref_count[b] = max(self.ref_count[b] - 1, 0)
            if self.ref_count[b] == 0 and b in self.used_block_ids:
                self.used_block_ids.remove(b)
                self.free_block_ids.append(b)

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
                 chunk_size=CHUNK_SIZE_CUDA, eos_token_ids=()):
        self.waiting = deque()
        self.running = []
        self.block_manager = BlockManager(num_blocks, block_size)
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
            if (token in self.eos_token_id
                    or len(seq.output_ids) >= seq.max_tokens
                    or len(seq) > self.max_num_batched_tokens):
                seq.status = FINISHED
                self.block_manager.deallocate(seq)
            else:
                seq.append_token(token)

    def clear_finished(self):                                  # scheduler.rs:631-654
        self.running = [s for s in self.running if s.status != FINISHED]
        self.waiting = deque(s for s in self.waiting if s.status != FINISHED)


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
