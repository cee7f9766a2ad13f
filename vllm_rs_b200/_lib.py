"""ctypes binding of the C ABI declared in include/b200_vllm.h.

The product path has NO fallback: if the shared library is missing this raises, and every
op checks the latched error state after launching.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
# B200_LIB_PATH: load an instrumented build instead (tools/dec_trace.py builds one with -DB200_DEC_TRACE)
LIB_PATH = Path(os.environ["B200_LIB_PATH"]) if os.environ.get("B200_LIB_PATH") else PKG / "libb200_vllm.so"
HEADER = PKG.parent / "include" / "b200_vllm.h"

_lib = None

c_i32, c_i64, c_f32, c_vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


class ModelConfig(C.Structure):
    """b200_model_config (include/b200_vllm.h)."""

    _fields_ = [(n, c_i32) for n in (
        "hidden", "num_layers", "num_heads", "num_kv_heads", "head_dim", "inter", "vocab",
        "bits", "group_size", "qfmt",
        "block_size", "num_blocks", "max_blocks_per_seq", "max_batch", "max_pos",
        "qkv_bias", "cos_sin_f32")] + [("rms_eps", c_f32), ("attn_scale", c_f32),
                                       ("tp_rank", c_i32), ("tp_size", c_i32),
                                       ("max_prefill_tokens", c_i32), ("kv_fp8", c_i32)]


class LayerWeights(C.Structure):
    """b200_layer_weights (include/b200_vllm.h)."""

    _fields_ = [(n, c_vp) for n in (
        "input_norm", "post_norm",
        "qkv_w", "qkv_s", "qkv_z", "qkv_bias",
        "o_w", "o_s", "o_z",
        "gate_up_w", "gate_up_s", "gate_up_z",
        "down_w", "down_s", "down_z",
        "k_cache", "v_cache")]


class KvCachePlan(C.Structure):
    """b200_kvcache_plan (include/b200_vllm.h)."""

    _fields_ = [(n, c_i64) for n in ("per_block_bytes", "num_gpu_blocks", "num_cpu_blocks", "max_num_seqs",
                                     "max_model_len", "kvcache_memory_bytes", "max_num_batched_tokens")]


class PdKvHeader(C.Structure):
    """b200_pd_kv_header (include/b200_vllm.h)."""

    _fields_ = [("seq_id", C.c_uint64), ("first_token", C.c_uint32), ("num_layers", c_i32),
                ("layer_bytes", C.c_uint64), ("num_blocks", C.c_uint64), ("sending_time", C.c_uint64),
                ("num_cached_tokens", C.c_uint64)]


class WireSampling(C.Structure):
    """b200_wire_sampling (include/b200_vllm.h)."""

    _fields_ = [(n, c_i32) for n in ("has_temperature", "has_max_tokens", "ignore_eos", "has_top_k", "has_top_p",
                                     "has_frequency_penalty", "has_presence_penalty", "reserved")] + \
               [(n, c_f32) for n in ("temperature", "top_p", "frequency_penalty", "presence_penalty")] + \
               [("max_tokens", c_i64), ("top_k", c_i64)]


ALLREDUCE_FN = C.CFUNCTYPE(None, c_vp, c_i64, c_vp, c_i64)

# name -> (restype, argtypes)
_SIGS = {
    "b200_last_error": (C.c_int, []),
    "b200_last_error_string": (C.c_char_p, []),
    "b200_clear_error": (None, []),
    "b200_abi_version": (C.c_int, []),
    "b200_launch_count": (c_i64, []),
    "b200_awq_prepack": (None, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i64]),
    "b200_gptq_prepack": (None, [c_vp, c_vp, c_i32, c_i32, c_i32, c_i64]),
    "b200_marlin_awq_4bit_f16": (None, [c_vp] * 6 + [c_i32] * 3 + [c_vp, c_i32, c_i64]),
    "b200_marlin_4bit_f16": (None, [c_vp] * 6 + [c_i32] * 3 + [c_vp, c_i32, c_i64]),
    "b200_marlin_awq_4bit_bf16": (None, [c_vp] * 6 + [c_i32] * 3 + [c_vp, c_i32, c_i64]),
    "b200_marlin_4bit_bf16": (None, [c_vp] * 6 + [c_i32] * 3 + [c_vp, c_i32, c_i64]),
    "b200_gemm_half_q_half_8bit": (None, [c_vp] * 6 + [c_i32] * 4 + [c_vp, c_i64]),
    "b200_gemm_half_q_half_alt": (None, [c_vp] * 6 + [c_i32] * 4 + [c_i64]),
    "b200_gemm_half_q_half_alt_release": (None, []),
    "b200_wna16_gemm_ex": (None, [c_vp] * 6 + [c_i32] * 8 + [c_vp, c_i64]),
    "b200_wna16_workspace_bytes": (c_i64, [c_i32, c_i32]),
    "b200_wna16_dequant_f16": (None, [c_vp] * 4 + [c_i32] * 5 + [c_i64]),
    "b200_dense_gemm_f16": (None, [c_vp] * 3 + [c_i32] * 4 + [c_i64]),
    "b200_dense_gemm_ex": (None, [c_vp] * 4 + [c_i32] * 5 + [c_i64]),
    "b200_rope_kv_write_f16": (None, [c_vp] * 9 + [c_i32] * 8 + [c_i64]),
    "b200_rope_kv_write_bf16": (None, [c_vp] * 9 + [c_i32] * 8 + [c_i64]),
    "b200_rope_kv_write_fp8kv_f16": (None, [c_vp] * 9 + [c_i32] * 8 + [c_i64]),
    "b200_kv_write_f16": (None, [c_vp] * 5 + [c_i32] * 4 + [c_i64]),
    "b200_swap_blocks": (None, [c_vp, c_vp, c_vp, c_i32, c_i64, c_i64]),
    "b200_pd_kv_frame_bytes": (c_i64, [c_i32, c_i64]),
    "b200_pd_kv_frame_begin": (c_i64, [c_vp, c_i64, c_vp, c_vp, c_vp]),
    "b200_pd_kv_frame_parse": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i32]),
    "b200_pd_ipc_frame": (c_i64, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i32, c_i32, c_vp]),
    "b200_pd_ipc_frame_parse": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_i32]),
    "b200_pd_frame_small": (c_i64, [c_vp, c_i64, c_i32, C.c_uint64]),
    "b200_pd_frame_kind": (c_i32, [c_vp, c_i64, c_vp]),
    "b200_pd_export_blocks": (None, [c_vp, c_vp, c_i32, c_vp, c_i32, c_i64, c_vp, c_vp, c_vp, c_i64]),
    "b200_pd_import_blocks": (None, [c_vp, c_vp, c_i32, c_vp, c_i32, c_i64, c_vp, c_vp, c_vp, c_i64]),
    "b200_pd_copy_blocks_peer": (None, [c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_i32, c_i64, c_i64]),
    "b200_enable_peer_access": (c_i32, [c_i32, c_i32]),
    "b200_kv_write_l1": (None, [c_vp] * 5 + [c_i32] * 4 + [c_i64]),
    "b200_rope_kv_write_l1": (None, [c_vp] * 9 + [c_i32] * 9 + [c_i64]),
    "b200_paged_attn_decode_l1": (None, [c_vp] * 6 + [c_i32] * 7 + [c_f32, c_f32, c_i32, c_vp, c_i32, c_i64]),
    "b200_attn_prefill_l1": (None, [c_vp] * 7 + [c_i32] * 10 + [c_f32, c_i32, c_i64]),
    "b200_wire_kind": (c_i32, [c_vp, c_i64]),
    "b200_wire_parse_run_decode": (c_i32, [c_vp, c_i64, c_i32, c_i32] + [c_vp] * 9),
    "b200_wire_frame_run_decode": (c_i64, [c_vp, c_i64, c_i32] + [c_vp] * 6 + [c_i32, c_vp, c_vp, c_i32]),
    "b200_wire_frame_run_response": (c_i64, [c_vp, c_i64, c_vp, c_i32]),
    "b200_wire_parse_run_response": (c_i32, [c_vp, c_i64, c_vp, c_i32]),
    "b200_wire_parse_kv_swap": (c_i32, [c_vp, c_i64, c_vp, c_i32, c_vp]),
    "b200_wire_frame_kv_swap": (c_i64, [c_vp, c_i64, c_vp, c_i32, c_i32]),
    "b200_wire_frame_simple": (c_i64, [c_vp, c_i64, c_i32, c_i32, C.c_uint64]),
    "b200_scheduler_frame_run_prefill": (c_i64, [c_vp, c_vp, c_i32, c_vp, c_i64, c_i32]),
    "b200_scheduler_load_run_prefill": (c_i32, [c_vp, c_vp, c_i64, c_vp, c_i32, c_vp, c_vp]),
    "b200_scheduler_frame_sequence": (c_i64, [c_vp, c_i32, c_i32, c_i32, C.c_uint32, c_vp, c_i64]),
    "b200_scheduler_load_sequence": (c_i32, [c_vp, c_vp, c_i64, c_i32, c_vp, c_vp]),
    "b200_wire_frame_kv_receive_response": (c_i64, [c_vp, c_i64, c_i32, C.c_uint32, C.c_uint64, C.c_uint64]),
    "b200_wire_parse_kv_receive_response": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "b200_wire_decode_metadata": (None, [c_i32, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "b200_paged_attn_decode_f16": (None, [c_vp] * 6 + [c_i32] * 7 + [c_f32, c_f32, c_i32, c_vp, c_i64]),
    "b200_paged_attn_decode_bf16": (None, [c_vp] * 6 + [c_i32] * 7 + [c_f32, c_f32, c_i32, c_vp, c_i64]),
    "b200_paged_attn_decode_fp8kv_f16": (None, [c_vp] * 6 + [c_i32] * 7 + [c_f32, c_f32, c_i32, c_vp, c_i64]),
    "b200_paged_attn_workspace_bytes": (c_i64, [c_i32] * 4),
    "b200_attn_prefill_f16": (None, [c_vp] * 7 + [c_i32] * 8 + [c_f32, c_i64]),
    "b200_attn_prefill_ex_f16": (None, [c_vp] * 7 + [c_i32] * 10 + [c_f32, c_i64]),
    "b200_attn_prefill_ex_bf16": (None, [c_vp] * 7 + [c_i32] * 10 + [c_f32, c_i64]),
    "b200_rmsnorm_f16": (None, [c_vp] * 3 + [c_i32, c_i32, c_f32, c_i64]),
    "b200_rmsnorm_bf16": (None, [c_vp] * 3 + [c_i32, c_i32, c_f32, c_i64]),
    "b200_add_rmsnorm_f16": (None, [c_vp] * 4 + [c_i32, c_i32, c_f32, c_i64]),
    "b200_add_rmsnorm_bf16": (None, [c_vp] * 4 + [c_i32, c_i32, c_f32, c_i64]),
    "b200_silu_mul_f16": (None, [c_vp, c_vp, c_i32, c_i32, c_i64]),
    "b200_silu_mul_bf16": (None, [c_vp, c_vp, c_i32, c_i32, c_i64]),
    "b200_embedding_f16": (None, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i32, c_i64]),
    "b200_gather_rows_f16": (None, [c_vp, c_vp, c_vp, c_i32, c_i32, c_i64]),
    "b200_argmax_f32": (None, [c_vp, c_vp, c_i32, c_i32, c_i64]),
    "b200_add_bias_f16": (None, [c_vp, c_vp, c_i32, c_i32, c_i64]),
    "b200_sample_topk_topp_f32": (None, [c_vp, c_vp, c_i32, c_i32, c_i32, c_f32, c_f32, c_vp, c_i64]),
    "b200_apply_penalties_f32": (None, [c_vp, c_i32, c_i32, c_vp, c_vp, c_i32, c_f32, c_f32, c_i64]),
    "b200_sampling_strategy": (c_i32, [c_i32, c_f32, c_i32, c_i64, c_i32, c_f32, c_vp, c_vp, c_vp]),
    "b200_rng_new": (c_vp, [C.c_uint64]),
    "b200_rng_free": (None, [c_vp]),
    "b200_rng_next_u32": (C.c_uint32, [c_vp]),
    "b200_rng_next_u64": (C.c_uint64, [c_vp]),
    "b200_runner_set_sampling": (None, [c_vp, c_i32, c_f32, c_f32]),
    "b200_runner_set_step_seed": (None, [c_vp, C.c_uint64, c_i64]),
    "b200_runner_set_sampling_strategy": (c_i32, [c_vp, c_i32, c_i32, c_f32, c_f32]),
    "b200_engine_set_swap_caches": (c_i32, [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_i64]),
    "b200_engine_set_seed": (None, [c_vp, C.c_uint64]),
    "b200_scheduler_new": (c_vp, [c_i32] * 5 + [c_vp, c_i32]),
    "b200_scheduler_free": (None, [c_vp]),
    "b200_scheduler_add": (c_i64, [c_vp, c_vp, c_i32, c_i32]),
    "b200_scheduler_set_stop_sequences": (c_i32, [c_vp, c_i64, c_vp, c_vp, c_i32]),
    "b200_scheduler_schedule": (c_i32, [c_vp, c_vp, c_i32, c_vp]),
    "b200_scheduler_filter_prefill_finished": (c_i32, [c_vp, c_vp, c_i32, c_vp, c_vp]),
    "b200_scheduler_postprocess": (None, [c_vp, c_vp, c_i32, c_vp]),
    "b200_scheduler_clear_finished": (None, [c_vp]),
    "b200_scheduler_num_running": (c_i32, [c_vp]),
    "b200_scheduler_num_waiting": (c_i32, [c_vp]),
    "b200_scheduler_num_free_blocks": (c_i32, [c_vp]),
    "b200_scheduler_enable_prefix_cache": (c_i32, [c_vp, c_i64]),
    "b200_scheduler_prefix_cache_blocks": (c_i32, [c_vp]),
    "b200_scheduler_evict_prefix_cache_blocks": (c_i32, [c_vp, c_i32]),
    "b200_scheduler_block_ref_count": (c_i32, [c_vp, c_i32]),
    "b200_scheduler_free_block_ids": (c_i32, [c_vp, c_vp, c_i32]),
    "b200_prefix_hash_block": (C.c_uint64, [C.c_uint64, c_vp, c_i32]),
    "b200_scheduler_enable_cpu_swap": (None, [c_vp, c_i32]),
    "b200_scheduler_set_time_ms": (None, [c_vp, c_i64]),
    "b200_scheduler_num_swapped": (c_i32, [c_vp]),
    "b200_scheduler_swapped_seq_id": (c_i64, [c_vp, c_i32]),
    "b200_scheduler_free_cpu_block_ids": (c_i32, [c_vp, c_vp, c_i32]),
    "b200_scheduler_pop_swap_op": (c_i32, [c_vp, c_vp, c_i32, c_vp]),
    "b200_scheduler_seq_info": (c_i32, [c_vp, c_i32, c_vp, c_vp, c_vp, c_vp]),
    "b200_scheduler_seq_block_table": (c_i32, [c_vp, c_i32, c_vp, c_i32]),
    "b200_scheduler_seq_tokens": (c_i32, [c_vp, c_i32, c_vp, c_i32]),
    "b200_prepare_decode": (c_i32, [c_vp, c_vp, c_i32] + [c_vp] * 5 + [c_i32, c_vp]),
    "b200_prepare_prefill": (c_i32, [c_vp, c_vp, c_i32, c_i32] + [c_vp] * 7 + [c_i32] + [c_vp] * 5),
    "b200_kvcache_plan_allocation": (c_i32, [c_i64] + [c_i32] * 6 + [c_i64] * 3 + [c_f32, C.POINTER(KvCachePlan)]),
    "b200_runner_new": (c_vp, [C.POINTER(ModelConfig)]),
    "b200_runner_free": (None, [c_vp]),
    "b200_runner_set_layer": (None, [c_vp, c_i32, C.POINTER(LayerWeights)]),
    "b200_runner_set_globals": (None, [c_vp] * 6),
    "b200_runner_workspace_bytes": (c_i64, [c_vp]),
    "b200_runner_set_workspace": (None, [c_vp, c_vp, c_i64]),
    "b200_runner_set_allreduce": (None, [c_vp, ALLREDUCE_FN, c_vp]),
    "b200_runner_decode_device": (None, [c_vp] * 6 + [c_i32, c_vp, c_vp, c_i64]),
    "b200_runner_decode_host": (c_i32, [c_vp] * 5 + [c_i32, c_vp, c_i32, c_vp, c_i64]),
    "b200_runner_capture": (c_i32, [c_vp, c_vp, c_i32, c_i64]),
    "b200_runner_prefill_device": (None, [c_vp] * 8 + [c_i32] * 4 + [c_vp, c_vp, c_i64]),
    "b200_runner_launches_per_step": (c_i64, [c_vp]),
    "b200_engine_new": (c_vp, [c_vp, c_vp, c_i32, c_i32]),
    "b200_engine_free": (None, [c_vp]),
    "b200_engine_step": (c_i32, [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64]),
    "b200_runner_set_comm": (None, [c_vp, c_vp]),
    "b200_runner_set_tp_fused": (None, [c_vp, c_i32]),
    "b200_nccl_unique_id": (c_i32, [c_vp]),
    "b200_comm_init_rank": (c_vp, [c_vp, c_i32, c_i32]),
    "b200_comm_enable_oneshot": (c_i32, [c_vp, c_i64]),
    "b200_comm_set_force_nccl": (None, [c_vp, c_i32]),
    "b200_comm_free": (None, [c_vp]),
    "b200_allreduce_f16": (None, [c_vp, c_vp, c_i64, c_i64]),
    "b200_comm_rank": (c_i32, [c_vp]),
    "b200_comm_world": (c_i32, [c_vp]),
}


def declared_symbols() -> list[str]:
    """Every function name declared in include/b200_vllm.h."""
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)) - {"b200_allreduce_fn"})


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -m vllm_rs_b200.build` (there is no CPU fallback)")
        _lib = C.CDLL(str(LIB_PATH), mode=os.RTLD_GLOBAL if hasattr(os, "RTLD_GLOBAL") else 0)
        for name, (res, args) in _SIGS.items():
            fn = getattr(_lib, name)
            fn.restype = res
            fn.argtypes = args
    return _lib


class B200Error(RuntimeError):
    pass


def check() -> None:
    """Raise if any call since the last check latched an error."""
    l = lib()
    code = l.b200_last_error()
    if code:
        msg = l.b200_last_error_string().decode()
        l.b200_clear_error()
        raise B200Error(f"b200 error {code}: {msg}")
