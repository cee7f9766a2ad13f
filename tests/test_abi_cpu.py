"""CPU-only: the C-ABI library loads and exports every symbol include/b200_vllm.h declares;
no compute calls (no GPU here)."""
import ctypes


def test_every_declared_symbol_is_exported(lib):
    from vllm_rs_b200 import _lib
    names = _lib.declared_symbols()
    assert len(names) >= 40
    for n in names:
        assert hasattr(lib, n), f"{n} declared in b200_vllm.h but not exported"
    assert set(_lib._SIGS) <= set(names), set(_lib._SIGS) - set(names)
    assert lib.b200_abi_version() == 1
    assert lib.b200_last_error() == 0


def test_workspace_queries_and_arg_errors(lib):
    from vllm_rs_b200 import _lib
    assert lib.b200_wna16_workspace_bytes(64, 28672) >= 64 * 28672 * 4
    assert lib.b200_paged_attn_workspace_bytes(32, 32, 128, 66) > 0
    # argument validation happens before any CUDA call: latched error, readable message
    lib.b200_rmsnorm_f16(None, None, None, 2, 100, ctypes.c_float(1e-5), 0)
    assert lib.b200_last_error() == 1
    assert b"hidden" in lib.b200_last_error_string()
    lib.b200_clear_error()
    assert lib.b200_last_error() == 0
    cfg = _lib.ModelConfig()
    assert not lib.b200_runner_new(ctypes.byref(cfg))        # zeroed config is rejected
    lib.b200_clear_error()
