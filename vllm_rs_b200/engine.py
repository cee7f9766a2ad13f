"""Host-side mirror of the reference's per-step interface over the C ABI.

  Scheduler    <->  src/core/scheduler.rs  (C++ state machine in csrc/host/scheduler.cpp)
  ModelRunner  <->  src/core/runner.rs     (decode branch; C++ launch plan in
                                            csrc/host/model_runner.cpp, CUDA graphs per bucket)

torch is used only to allocate device / pinned host memory and to name the CUDA stream.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, ops
from ._lib import LayerWeights, ModelConfig

GRAPH_BUCKETS = list(range(1, 16)) + [16, 32, 64]     # graph.rs:370-377 stops at 32; we add 64


class Scheduler:
    """Scheduler + BlockManager (scheduler.rs / block_manager.rs) behind the C ABI."""

    def __init__(self, num_blocks, block_size=64, max_num_seqs=32, max_num_batched_tokens=0,
                 chunk_size=8192, eos_token_ids=()):
        self._l = _lib.lib()
        eos = (C.c_uint32 * max(1, len(eos_token_ids)))(*eos_token_ids)
        self._h = self._l.b200_scheduler_new(num_blocks, block_size, max_num_seqs,
                                             max_num_batched_tokens, chunk_size, eos,
                                             len(eos_token_ids))
        _lib.check()
        self.block_size = block_size
        self.chunk_size = chunk_size
        self.cap = 4096

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.b200_scheduler_free(self._h)
            self._h = None

    def add(self, token_ids, max_tokens=0, stop_token_ids=None) -> int:
        arr = np.ascontiguousarray(token_ids, dtype=np.uint32)
        sid = self._l.b200_scheduler_add(self._h, arr.ctypes.data, len(arr), max_tokens)
        _lib.check()
        if stop_token_ids:                  # SamplingParams.stop_token_ids: a list of token sequences
            flat = np.ascontiguousarray([t for st_ in stop_token_ids for t in st_], dtype=np.uint32)
            lens = np.ascontiguousarray([len(st_) for st_ in stop_token_ids], dtype=np.int32)
            rc = self._l.b200_scheduler_set_stop_sequences(self._h, sid, flat.ctypes.data, lens.ctypes.data, len(lens))
            assert rc == 0
        return sid

    def schedule(self):
        idx = np.zeros(self.cap, dtype=np.int32)
        pf = C.c_int32(0)
        n = self._l.b200_scheduler_schedule(self._h, idx.ctypes.data, self.cap, C.byref(pf))
        _lib.check()
        if n < 0:
            raise _lib.B200Error("schedule failed")
        return idx[:n].tolist(), bool(pf.value)

    def filter_prefill_finished(self, ids):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        pos = np.zeros(max(1, len(a)), dtype=np.int32)
        idx = np.zeros(max(1, len(a)), dtype=np.int32)
        n = self._l.b200_scheduler_filter_prefill_finished(self._h, a.ctypes.data, len(a),
                                                           pos.ctypes.data, idx.ctypes.data)
        return pos[:n].tolist(), idx[:n].tolist()

    def postprocess(self, ids, tokens):
        a = np.ascontiguousarray(ids, dtype=np.int32)
        t = np.ascontiguousarray(tokens, dtype=np.uint32)
        self._l.b200_scheduler_postprocess(self._h, a.ctypes.data, len(a), t.ctypes.data)

    def clear_finished(self):
        self._l.b200_scheduler_clear_finished(self._h)

    @property
    def num_running(self):
        return self._l.b200_scheduler_num_running(self._h)

    @property
    def num_waiting(self):
        return self._l.b200_scheduler_num_waiting(self._h)

    @property
    def num_free_blocks(self):
        return self._l.b200_scheduler_num_free_blocks(self._h)

    def enable_prefix_cache(self, max_tokens=None) -> int:
        """build_prefix_cache_config (scheduler.rs:60-104); returns the capacity in blocks."""
        return self._l.b200_scheduler_enable_prefix_cache(self._h, -1 if max_tokens is None else int(max_tokens))

    @property
    def prefix_cache_blocks(self):
        return self._l.b200_scheduler_prefix_cache_blocks(self._h)

    def evict_prefix_cache_blocks(self, n):
        return self._l.b200_scheduler_evict_prefix_cache_blocks(self._h, int(n))

    def block_ref_count(self, block_id):
        return self._l.b200_scheduler_block_ref_count(self._h, int(block_id))

    def free_block_ids(self):
        n = self._l.b200_scheduler_free_block_ids(self._h, None, 0)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self._l.b200_scheduler_free_block_ids(self._h, out.ctypes.data, n)
        return out[:n].tolist()

    # ---- CPU swap bookkeeping (scheduler.rs:303-337,830-954)
    def enable_cpu_swap(self, num_cpu_blocks):
        self._l.b200_scheduler_enable_cpu_swap(self._h, int(num_cpu_blocks))

    def set_time_ms(self, now_ms):
        self._l.b200_scheduler_set_time_ms(self._h, int(now_ms))

    @property
    def num_swapped(self):
        return self._l.b200_scheduler_num_swapped(self._h)

    def swapped_seq_ids(self):
        return [self._l.b200_scheduler_swapped_seq_id(self._h, i) for i in range(self.num_swapped)]

    def free_cpu_block_ids(self):
        n = self._l.b200_scheduler_free_cpu_block_ids(self._h, None, 0)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self._l.b200_scheduler_free_cpu_block_ids(self._h, out.ctypes.data, n)
        return out[:n].tolist()

    def pop_swap_ops(self):
        """Pending block copies, oldest first: [(swap_in, [(src, dst), ...]), ...] for ops.swap_blocks."""
        ops_, buf, flag = [], np.zeros(2 * 4096, np.int64), C.c_int32(0)
        while True:
            n = self._l.b200_scheduler_pop_swap_op(self._h, buf.ctypes.data, 4096, C.byref(flag))
            if n <= 0:
                return ops_
            ops_.append((bool(flag.value), [tuple(p) for p in buf[:2 * n].reshape(n, 2).tolist()]))

    def seq_info(self, i):
        sid, cached, st, nout = C.c_int64(), C.c_int32(), C.c_int32(), C.c_int32()
        n = self._l.b200_scheduler_seq_info(self._h, i, C.byref(sid), C.byref(cached),
                                            C.byref(st), C.byref(nout))
        return dict(len=n, id=sid.value, num_cached_tokens=cached.value, status=st.value,
                    num_output_tokens=nout.value)

    def block_table(self, i):
        n = self._l.b200_scheduler_seq_block_table(self._h, i, None, 0)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self._l.b200_scheduler_seq_block_table(self._h, i, out.ctypes.data, n)
        return out[:n].tolist()

    def tokens(self, i):
        n = self._l.b200_scheduler_seq_tokens(self._h, i, None, 0)
        out = np.zeros(max(n, 1), dtype=np.uint32)
        self._l.b200_scheduler_seq_tokens(self._h, i, out.ctypes.data, n)
        return out[:n].tolist()

    def prepare_decode(self, ids, max_blocks_cap=1024, pinned=None):
        """prepare_decode (runner.rs:1243-1388) into (optionally pinned) host numpy buffers."""
        n = len(ids)
        a = np.ascontiguousarray(ids, dtype=np.int32)
        if pinned is None:
            bufs = dict(input_ids=np.zeros(n, np.uint32), positions=np.zeros(n, np.int64),
                        slot_mapping=np.zeros(n, np.int64), context_lens=np.zeros(n, np.uint32),
                        block_tables=np.zeros(n * max_blocks_cap, np.uint32))
        else:
            bufs = pinned
        mctx = C.c_int32(0)
        mb = self._l.b200_prepare_decode(self._h, a.ctypes.data, n, bufs["input_ids"].ctypes.data,
                                         bufs["positions"].ctypes.data,
                                         bufs["slot_mapping"].ctypes.data,
                                         bufs["context_lens"].ctypes.data,
                                         bufs["block_tables"].ctypes.data, max_blocks_cap,
                                         C.byref(mctx))
        if mb < 0:
            raise _lib.B200Error("prepare_decode: block table wider than cap")
        out = {k: v[:n] for k, v in bufs.items() if k != "block_tables"}
        out["block_tables"] = bufs["block_tables"][:n * mb].reshape(n, mb)
        out["max_blocks"] = mb
        out["max_context_len"] = mctx.value
        return out

    def prepare_prefill(self, ids, token_cap=1 << 17, max_blocks_cap=1024):
        n = len(ids)
        a = np.ascontiguousarray(ids, dtype=np.int32)
        input_ids = np.zeros(token_cap, np.uint32)
        positions = np.zeros(token_cap, np.int64)
        slots = np.zeros(token_cap, np.int64)
        cu_q = np.zeros(n + 1, np.uint32)
        cu_k = np.zeros(n + 1, np.uint32)
        ctx = np.zeros(n, np.uint32)
        bt = np.zeros(n * max_blocks_cap, np.uint32)
        mb, mq, mk, mc, hc = (C.c_int32() for _ in range(5))
        T = self._l.b200_prepare_prefill(self._h, a.ctypes.data, n, token_cap,
                                         input_ids.ctypes.data, positions.ctypes.data,
                                         slots.ctypes.data, cu_q.ctypes.data, cu_k.ctypes.data,
                                         ctx.ctypes.data, bt.ctypes.data, max_blocks_cap,
                                         C.byref(mb), C.byref(mq), C.byref(mk), C.byref(mc),
                                         C.byref(hc))
        _lib.check()
        if T < 0:
            raise _lib.B200Error("prepare_prefill: capacity exceeded")
        out = dict(input_ids=input_ids[:T], positions=positions[:T], slot_mapping=slots[:T],
                   cu_seqlens_q=cu_q, cu_seqlens_k=cu_k, max_seqlen_q=mq.value,
                   max_seqlen_k=mk.value, max_context_len=mc.value, block_tables=None,
                   context_lens=None, max_blocks=mb.value)
        if hc.value:
            out["block_tables"] = bt[:n * mb.value].reshape(n, mb.value)
            out["context_lens"] = ctx
        return out


    # ---- RunPrefill frames (runner/mod.rs:178): the engine side frames its scheduled sequences, a runner side loads them
    def frame_run_prefill(self, ids, is_prefill=True) -> bytes:
        a = np.ascontiguousarray(ids, dtype=np.int32)
        cap = 64 + sum(256 + 4 * (2 * len(self.tokens(i)) + len(self.block_table(i))) for i in ids)
        out = np.zeros(cap, np.uint8)
        n = self._l.b200_scheduler_frame_run_prefill(self._h, a.ctypes.data, len(a), out.ctypes.data, cap,
                                                     int(is_prefill))
        _lib.check()
        return out[:n].tobytes()

    def load_run_prefill(self, frame: bytes, cap=1024):
        """Replace this scheduler's running list with the sequences of a RunPrefill frame; returns (ids, is_prefill,
        sampling of the first sequence)."""
        buf = np.frombuffer(frame, np.uint8)
        idx = np.zeros(cap, np.int32)
        pf = C.c_int32()
        sp = _lib.WireSampling()
        n = self._l.b200_scheduler_load_run_prefill(self._h, buf.ctypes.data, len(buf), idx.ctypes.data, cap,
                                                    C.byref(pf), C.byref(sp))
        if n < 0:
            raise _lib.B200Error(f"load_run_prefill: {n}")
        return idx[:n].tolist(), bool(pf.value), sp


    # ---- single-Sequence messages of the PD hand-off (family 0: runner IPC, 1: client <-> PD server TCP)
    def frame_sequence(self, i, kind, family=0, extra=0) -> bytes:
        n_tok = 0 if i < 0 else 2 * len(self.tokens(i)) + len(self.block_table(i))
        out = np.zeros(512 + 4 * n_tok, np.uint8)
        n = self._l.b200_scheduler_frame_sequence(self._h, int(i), family, kind, int(extra), out.ctypes.data, len(out))
        _lib.check()
        return out[:n].tobytes()

    def load_sequence(self, frame: bytes, family=0):
        """Returns (has_sequence, kind, extra); the sequence becomes running[0] of this container."""
        buf = np.frombuffer(frame, np.uint8)
        kind, extra = C.c_int32(), C.c_uint32()
        n = self._l.b200_scheduler_load_sequence(self._h, buf.ctypes.data, len(buf), family, C.byref(kind),
                                                 C.byref(extra))
        if n < 0:
            raise _lib.B200Error(f"load_sequence: {n}")
        return bool(n), kind.value, extra.value


def plan_kvcache(available_bytes, *, num_layers, num_kv_heads, head_dim, num_shards=1, block_size=64,
                 dtype_size=2, config_model_len=8192, max_num_seqs=None, max_model_len=None,
                 cpu_mem_fold=0.5) -> dict:
    """KVCacheAllocator::plan_allocation (kvcache_allocator.rs:616-707) through the C ABI.
    Raises ValueError (InvalidConfiguration) / MemoryError (InsufficientGpuMemory)."""
    out = _lib.KvCachePlan()
    rc = _lib.lib().b200_kvcache_plan_allocation(int(available_bytes), num_layers, num_kv_heads, head_dim, num_shards,
                                          block_size, dtype_size, int(config_model_len),
                                          int(max_num_seqs or 0), int(max_model_len or 0),
                                          float(cpu_mem_fold), C.byref(out))
    if rc == -1:
        raise ValueError("invalid KV-cache configuration")
    if rc == -2:
        raise MemoryError("insufficient GPU memory for the KV cache")
    return {n: int(getattr(out, n)) for n, _ in out._fields_}


@dataclass
class QuantLinear:
    """One prepacked W4A16/W8A16 (or dense fp16) linear, possibly several projections fused
    along N (q|k|v, gate|up) -- wna16.rs:25-261 loads them separately."""
    w: torch.Tensor            # tile-major packed u32, or dense fp16 [N, K]
    s: torch.Tensor | None     # scales [K/G, N] fp16
    z: torch.Tensor | None     # qzeros, checkpoint packing
    k: int
    n: int


def fuse_awq(parts, device):
    """parts: list of (qweight i32 [K, N_i/8], qzeros i32 [K/G, N_i/8], scales f16 [K/G, N_i])
    numpy arrays (AWQ checkpoint tensors of projections sharing the same input).  They are
    concatenated along N in checkpoint format (word granularity = 8 channels) and prepacked once,
    so only the fused N has to be a multiple of 128."""
    qw = np.concatenate([p[0] for p in parts], axis=1)
    qz = np.concatenate([p[1] for p in parts], axis=1)
    sc = np.concatenate([p[2] for p in parts], axis=1)
    w = ops.awq_prepack(torch.from_numpy(np.ascontiguousarray(qw)).to(device)).reshape(-1)
    return QuantLinear(w, torch.from_numpy(np.ascontiguousarray(sc)).to(device),
                       torch.from_numpy(np.ascontiguousarray(qz)).to(device), qw.shape[0], sc.shape[1])


def fuse_gptq(parts, device, bits=4):
    qw = np.concatenate([p[0] for p in parts], axis=1)
    sc = np.concatenate([p[2] for p in parts], axis=1)
    z = None if parts[0][1] is None else torch.from_numpy(
        np.ascontiguousarray(np.concatenate([p[1] for p in parts], axis=1))).to(device)
    w = ops.gptq_prepack(torch.from_numpy(np.ascontiguousarray(qw)).to(device), bits).reshape(-1)
    return QuantLinear(w, torch.from_numpy(np.ascontiguousarray(sc)).to(device), z,
                       qw.shape[0] * (32 // bits), sc.shape[1])


def fuse_dense(parts_kn, device):
    """parts: list of fp16 torch [K, N_i] -> dense [sum N_i, K] (Linear weight layout)."""
    w = torch.cat([p.t().contiguous() for p in parts_kn], dim=0).to(device).contiguous()
    return QuantLinear(w, None, None, parts_kn[0].shape[0], w.shape[0])


class ModelRunner:
    """Decode-step runner over the C++ launch plan (b200_runner_*)."""

    def __init__(self, *, hidden, num_layers, num_heads, num_kv_heads, head_dim, inter, vocab,
                 num_blocks, max_blocks_per_seq, max_batch, bits=4, group_size=128,
                 qfmt=ops.QFMT_AWQ, rms_eps=1e-5, qkv_bias=False, max_pos=8192,
                 max_prefill_tokens=0, device="cuda", kv_fp8=False):
        self._l = _lib.lib()
        self.device = torch.device(device)
        self.cfg = ModelConfig(
            hidden=hidden, num_layers=num_layers, num_heads=num_heads, num_kv_heads=num_kv_heads,
            head_dim=head_dim, inter=inter, vocab=vocab, bits=bits, group_size=group_size,
            qfmt=qfmt, block_size=64, num_blocks=num_blocks, max_blocks_per_seq=max_blocks_per_seq,
            max_batch=max_batch, max_pos=max_pos, qkv_bias=int(qkv_bias), cos_sin_f32=0,
            rms_eps=rms_eps, attn_scale=1.0 / float(np.sqrt(head_dim)), tp_rank=0, tp_size=1,
            max_prefill_tokens=max_prefill_tokens, kv_fp8=int(kv_fp8))
        self._h = self._l.b200_runner_new(C.byref(self.cfg))
        _lib.check()
        self._keep = []            # tensors whose pointers the runner holds
        shape = (num_blocks, 64, num_kv_heads, head_dim)
        cache_dtype = torch.uint8 if kv_fp8 else torch.float16   # kvcache_allocator.rs:776
        self.k_cache = [torch.zeros(shape, dtype=cache_dtype, device=self.device)
                        for _ in range(num_layers)]
        self.v_cache = [torch.zeros(shape, dtype=cache_dtype, device=self.device)
                        for _ in range(num_layers)]
        nbytes = self._l.b200_runner_workspace_bytes(self._h)
        self.workspace = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        self._l.b200_runner_set_workspace(self._h, self.workspace.data_ptr(), nbytes)
        _lib.check()
        B, mb = max_batch, max_blocks_per_seq
        self.pinned = dict(
            input_ids=torch.zeros(B, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
            positions=torch.zeros(B, dtype=torch.int64).pin_memory().numpy(),
            slot_mapping=torch.zeros(B, dtype=torch.int64).pin_memory().numpy(),
            context_lens=torch.zeros(B, dtype=torch.int32).pin_memory().numpy().view(np.uint32),
            block_tables=torch.zeros(B * mb, dtype=torch.int32).pin_memory().numpy().view(np.uint32))
        self.out_ids = torch.zeros(B, dtype=torch.int32).pin_memory().numpy().view(np.uint32)
        self._ar_cb = None

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.b200_runner_free(self._h)
            self._h = None

    def set_layer(self, l, *, input_norm, post_norm, qkv: QuantLinear, o: QuantLinear,
                  gate_up: QuantLinear, down: QuantLinear, qkv_bias=None):
        ts = [input_norm, post_norm, qkv.w, qkv.s, qkv.z, qkv_bias, o.w, o.s, o.z, gate_up.w,
              gate_up.s, gate_up.z, down.w, down.s, down.z]
        self._keep.append(ts)
        p = [0 if t is None else t.data_ptr() for t in ts]
        lw = LayerWeights(*p, self.k_cache[l].data_ptr(), self.v_cache[l].data_ptr())
        self._l.b200_runner_set_layer(self._h, l, C.byref(lw))
        _lib.check()

    def set_globals(self, embed, final_norm, lm_head, cos, sin):
        self._keep.append([embed, final_norm, lm_head, cos, sin])
        self._l.b200_runner_set_globals(self._h, embed.data_ptr(), final_norm.data_ptr(),
                                        lm_head.data_ptr(), cos.data_ptr(), sin.data_ptr())

    def set_allreduce(self, fn):
        """fn(ptr, count_halfs, stream) is called after each row-parallel GEMM (TP)."""
        if fn is None:
            self._ar_cb = None
            self._l.b200_runner_set_allreduce(self._h, _lib.ALLREDUCE_FN(), None)
            return
        self._ar_cb = _lib.ALLREDUCE_FN(lambda buf, n, user, st: fn(buf, n, st))
        self._l.b200_runner_set_allreduce(self._h, self._ar_cb, None)

    def set_sampling(self, top_k=0, top_p=1.0, temperature=1.0):
        """Strategy cached at prefill (runner.rs:1402-1510): top_k = 0 is greedy, else the device sampler."""
        self._l.b200_runner_set_sampling(self._h, int(top_k), float(top_p), float(temperature))
        _lib.check()

    def set_step_seed(self, seed, stream=None):
        """The step's kernel seed = LogitsProcessor's rng.next_u64() (logits_processor.rs:222-225)."""
        st = stream if stream is not None else torch.cuda.current_stream().cuda_stream
        self._l.b200_runner_set_step_seed(self._h, int(seed) & ((1 << 64) - 1), st)
        _lib.check()

    def capture(self, buckets=None):
        """warmup_capture (graph.rs:448-690)."""
        b = [x for x in (buckets or GRAPH_BUCKETS) if x <= self.cfg.max_batch]
        arr = (C.c_int32 * len(b))(*b)
        # capture needs a non-default stream
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            rc = self._l.b200_runner_capture(self._h, arr, len(b), s.cuda_stream)
        s.synchronize()
        _lib.check()
        if rc != 0:
            raise _lib.B200Error("graph capture failed")
        self._graph_stream = s

    def decode_device(self, input_ids, positions, slot_mapping, block_tables, context_lens,
                      want_logits=True):
        """One decode step from device tensors; returns (logits f32 [B,V] | None, ids i32 [B])."""
        B = input_ids.shape[0]
        assert block_tables.shape[1] == self.cfg.max_blocks_per_seq
        logits = (torch.empty(B, self.cfg.vocab, dtype=torch.float32, device=self.device)
                  if want_logits else None)
        ids = torch.empty(B, dtype=torch.int32, device=self.device)
        self._l.b200_runner_decode_device(
            self._h, input_ids.data_ptr(), positions.data_ptr(), slot_mapping.data_ptr(),
            block_tables.data_ptr(), context_lens.data_ptr(), B,
            0 if logits is None else logits.data_ptr(), ids.data_ptr(),
            torch.cuda.current_stream().cuda_stream)
        _lib.check()
        return logits, ids

    def prefill_device(self, meta, want_logits=True):
        """One prefill step (prefill branch of ModelRunner::run) from prepare_prefill metadata
        (numpy, host).  Block tables are required for every sequence: keys are read back from
        the paged cache.  Returns (logits f32 [num_seqs, V] | None, ids i32 [num_seqs])."""
        dev = self.device
        B = len(meta["cu_seqlens_q"]) - 1
        T = len(meta["input_ids"])

        def d(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a).astype(dt)).to(dev)

        ids = d(meta["input_ids"], np.int32)
        pos = d(meta["positions"], np.int64)
        slots = d(meta["slot_mapping"], np.int64)
        cu = d(meta["cu_seqlens_q"], np.int32)
        bt = d(meta["block_tables"], np.int32)
        ctx = None if meta.get("context_lens") is None else d(meta["context_lens"], np.int32)
        last = d(np.asarray(meta["cu_seqlens_q"][1:], dtype=np.int64) - 1, np.int32)
        logits = (torch.empty(B, self.cfg.vocab, dtype=torch.float32, device=dev)
                  if want_logits else None)
        out = torch.empty(B, dtype=torch.int32, device=dev)
        self._l.b200_runner_prefill_device(
            self._h, ids.data_ptr(), pos.data_ptr(), slots.data_ptr(), cu.data_ptr(),
            0 if ctx is None else ctx.data_ptr(), bt.data_ptr(), last.data_ptr(), B, T,
            int(meta["max_seqlen_q"]), bt.shape[1], 0 if logits is None else logits.data_ptr(),
            out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        _lib.check()
        return logits, out

    def run_decode(self, meta, stream=None) -> np.ndarray:
        """ModelRunner::run for a decode batch: host metadata in, host token ids out
        (H2D of the step inputs, graph replay or direct launch, D2H of the ids, sync)."""
        B = len(meta["input_ids"])
        st = stream if stream is not None else getattr(self, "_graph_stream", None)
        sid = st.cuda_stream if st is not None else torch.cuda.current_stream().cuda_stream
        rc = self._l.b200_runner_decode_host(
            self._h, meta["input_ids"].ctypes.data, meta["positions"].ctypes.data,
            meta["slot_mapping"].ctypes.data, meta["block_tables"].ctypes.data,
            int(meta["max_blocks"]), meta["context_lens"].ctypes.data, B,
            self.out_ids.ctypes.data, sid)
        _lib.check()
        if rc != 0:
            raise _lib.B200Error("decode_host failed")
        return self.out_ids[:B].copy()

    @property
    def launches_per_step(self):
        return self._l.b200_runner_launches_per_step(self._h)


class LLMEngine:
    """The reference's 3-phase step loop (engine.rs:1693-1757) over the C++ scheduler and the
    C++/CUDA runner: prepare_step (schedule + prepare_*) -> run_forward -> finish_step
    (filter_prefill_finished / postprocess / clear_finished).  Greedy sampling."""

    def __init__(self, runner: ModelRunner, num_blocks, max_num_seqs, eos_token_ids=(),
                 chunk_size=8192, prefix_cache=False, prefix_cache_max_tokens=None, native_step=True,
                 max_num_batched_tokens=0, seed=None, num_cpu_blocks=0):
        self.runner = runner
        # the scheduler hands out up to max(max_num_seqs, 5) sequences per step (scheduler.rs:44); the runner
        # must be able to take such a decode batch, and one chunk of one sequence as a prefill step
        if max(max_num_seqs, 5) > runner.cfg.max_batch:
            raise ValueError(f"runner max_batch {runner.cfg.max_batch} < max(max_num_seqs, 5) = {max(max_num_seqs, 5)}")
        self.sched = Scheduler(num_blocks, 64, max_num_seqs, max_num_batched_tokens, chunk_size, eos_token_ids)
        if prefix_cache:                    # EngineConfig.prefix_cache / prefix_cache_max_tokens
            self.sched.enable_prefix_cache(prefix_cache_max_tokens)
        # native_step: schedule -> prepare_decode -> run -> postprocess of a decode batch is ONE C call
        # (b200_engine_step); False keeps the same loop spelled out over the individual C symbols
        self._l = _lib.lib()
        self._eh = None
        if native_step:
            self._eh = self._l.b200_engine_new(self.sched._h, runner._h, runner.cfg.max_batch,
                                               runner.cfg.max_blocks_per_seq)
            _lib.check()
            cap = max(runner.cfg.max_batch, 5)
            self._idx = np.zeros(cap, np.int32)
            self._sid = np.zeros(cap, np.int64)
            self._tok = np.zeros(cap, np.uint32)
            if seed is not None:            # SamplingParams.seed -> the engine's StdRng (logits_processor.rs:48-70)
                self._l.b200_engine_set_seed(self._eh, int(seed) & ((1 << 64) - 1))
        self._rng = None
        if seed is not None and not native_step:
            self._rng = self._l.b200_rng_new(int(seed) & ((1 << 64) - 1))
        self.max_blocks = runner.cfg.max_blocks_per_seq
        self.outputs = {}
        self.cpu_k = self.cpu_v = None
        if num_cpu_blocks > 0:              # CPU swap: pinned-host mirrors of the layer caches (kvcache plan)
            self.sched.enable_cpu_swap(num_cpu_blocks)
            shp = (num_cpu_blocks,) + tuple(runner.k_cache[0].shape[1:])
            dt = runner.k_cache[0].dtype
            self.cpu_k = [torch.zeros(shp, dtype=dt).pin_memory() for _ in runner.k_cache]
            self.cpu_v = [torch.zeros(shp, dtype=dt).pin_memory() for _ in runner.v_cache]
            if self._eh:
                L = len(runner.k_cache)
                arr = lambda ts: (C.c_void_p * L)(*[t.data_ptr() for t in ts])      # noqa: E731
                bb = runner.k_cache[0][0].numel() * runner.k_cache[0].element_size()
                rc = self._l.b200_engine_set_swap_caches(self._eh, L, arr(runner.k_cache), arr(runner.v_cache),
                                                         arr(self.cpu_k), arr(self.cpu_v), bb)
                _lib.check()
                assert rc == 0

    def __del__(self):
        if getattr(self, "_eh", None):
            self._l.b200_engine_free(self._eh)
            self._eh = None
        if getattr(self, "_rng", None):
            self._l.b200_rng_free(self._rng)
            self._rng = None

    def _drain_swaps(self):
        """Python-step twin of the native drain: every queued block copy runs before the forward."""
        for swap_in, pairs in self.sched.pop_swap_ops():
            if self.cpu_k is None:
                raise _lib.B200Error("the scheduler queued KV swap copies but the engine has no CPU caches")
            keep = []
            for gk, gv, ck, cv in zip(self.runner.k_cache, self.runner.v_cache, self.cpu_k, self.cpu_v):
                keep.append(ops.swap_blocks(ck if swap_in else gk, gk if swap_in else ck, pairs))
                keep.append(ops.swap_blocks(cv if swap_in else gv, gv if swap_in else cv, pairs))
            torch.cuda.synchronize()

    def _prefill_batches(self, ids):
        """Split a scheduled prefill batch at sequence boundaries so every runner call fits the runner's
        max_prefill_tokens / max_batch (the scheduler's budget is max_num_batched_tokens, scheduler.rs:258-260)."""
        cap_t, cap_b = self.runner.cfg.max_prefill_tokens, self.runner.cfg.max_batch
        out, cur, tok = [], [], 0
        for i in ids:
            info = self.sched.seq_info(i)
            eff = min(self.sched.chunk_size, info["len"] - info["num_cached_tokens"])
            if eff > cap_t:
                raise _lib.B200Error(f"a prefill chunk of {eff} tokens exceeds the runner's max_prefill_tokens {cap_t}")
            if cur and (tok + eff > cap_t or len(cur) >= cap_b):
                out.append(cur)
                cur, tok = [], 0
            cur.append(i)
            tok += eff
        if cur:
            out.append(cur)
        return out

    def add_request(self, token_ids, max_tokens):
        sid = self.sched.add(token_ids, max_tokens)
        self.outputs[sid] = []
        return sid

    def _all_block_tables(self, ids):
        tabs = [self.sched.block_table(i) for i in ids]
        w = max(len(t) for t in tabs)
        bt = np.zeros((len(tabs), w), dtype=np.uint32)
        for r, t in enumerate(tabs):
            bt[r, :len(t)] = t
        return bt

    def _stream(self):
        st = getattr(self.runner, "_graph_stream", None)
        return st.cuda_stream if st is not None else torch.cuda.current_stream().cuda_stream

    def step(self):
        self.last_kind, self.last_prefill_tokens = "decode", 0
        if self._eh:
            npf = C.c_int32(0)
            n = self._l.b200_engine_step(self._eh, self._idx.ctypes.data, self._sid.ctypes.data,
                                         self._tok.ctypes.data, C.byref(npf), self._stream())
            _lib.check()
            if n == -1:
                raise _lib.B200Error("engine_step failed")
            if n == 0:
                return False
            if n > 0:
                for sid, t in zip(self._sid[:n].tolist(), self._tok[:n].tolist()):
                    self.outputs[sid].append(int(t))
                return True
            ids, is_prefill = self._idx[:npf.value].tolist(), True       # n == -2: a prefill batch was scheduled
        else:
            ids, is_prefill = self.sched.schedule()
            self._drain_swaps()
        if not ids:
            return False
        if is_prefill:
            parts = []
            self.last_kind = "prefill"
            for sub in self._prefill_batches(ids):
                meta = self.sched.prepare_prefill(sub, token_cap=self.runner.cfg.max_prefill_tokens,
                                                  max_blocks_cap=self.max_blocks)
                self.last_prefill_tokens += len(meta["input_ids"])
                if meta["block_tables"] is None:            # the CUDA path always reads the paged cache
                    meta["block_tables"] = self._all_block_tables(sub)
                _, t = self.runner.prefill_device(meta, want_logits=False)
                parts.append(t)
            toks = torch.cat(parts).cpu().numpy().astype(np.uint32)
            pos, fin = self.sched.filter_prefill_finished(ids)
            sel = toks[pos] if len(pos) else toks[:0]
            for i, t in zip(fin, sel):
                self.outputs[self.sched.seq_info(i)["id"]].append(int(t))
            self.sched.postprocess(fin, sel)
        else:
            meta = self.sched.prepare_decode(ids, max_blocks_cap=self.max_blocks,
                                             pinned=self.runner.pinned)
            if self._rng:
                self.runner.set_step_seed(self._l.b200_rng_next_u64(self._rng), self._stream())
            toks = self.runner.run_decode(meta)
            for i, t in zip(ids, toks):
                self.outputs[self.sched.seq_info(i)["id"]].append(int(t))
            self.sched.postprocess(ids, toks)
        self.sched.clear_finished()
        return True

    def run_to_completion(self, max_steps=100000):
        n = 0
        while (self.sched.num_running or self.sched.num_waiting) and n < max_steps:
            if not self.step():
                raise _lib.B200Error(
                    f"engine stalled: nothing schedulable with {self.sched.num_running} running / "
                    f"{self.sched.num_waiting} waiting sequences and {self.sched.num_free_blocks} free blocks")
            n += 1
        return self.outputs
