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


def _header_prototypes():
    """name -> (return type, [parameter type strings]) for every function declared in include/b200_vllm.h."""
    import re
    from vllm_rs_b200 import _lib
    text = re.sub(r"/\*.*?\*/", "", _lib.HEADER.read_text(), flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(b200_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", text, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        if name == "b200_allreduce_fn" or ret.startswith("typedef"):
            continue
        params = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
        protos[name] = (ret, params)
    return protos


def test_bindings_match_the_header_prototypes(lib):
    """Every ctypes signature in _lib._SIGS has the arity of its prototype in the header, and each parameter's
    class (pointer / 32-bit int / 64-bit int / float) agrees -- a binding that drifts from the header would
    otherwise only show up as memory corruption on the GPU box."""
    import ctypes as C
    from vllm_rs_b200 import _lib
    protos = _header_prototypes()
    assert len(protos) >= 80
    ptr_like = (C.c_void_p, C.c_char_p)

    def klass_c(t):
        if t in ptr_like or (isinstance(t, type) and issubclass(t, (C._Pointer, C._CFuncPtr))):
            return "ptr"
        return {C.c_int32: "i32", C.c_int64: "i64", C.c_uint32: "i32", C.c_uint64: "i64", C.c_float: "f32",
                C.c_int: "i32"}.get(t, str(t))

    def klass_h(p):
        p = p.strip()
        if "*" in p or p.startswith("b200_allreduce_fn"):
            return "ptr"
        base = p.replace("const ", "").split()[0]
        return {"int32_t": "i32", "uint32_t": "i32", "int": "i32", "int64_t": "i64", "uint64_t": "i64",
                "float": "f32"}.get(base, base)

    for name, (res, args) in _lib._SIGS.items():
        assert name in protos, f"{name} bound but not declared"
        ret, params = protos[name]
        assert len(args) == len(params), f"{name}: {len(args)} bound arguments, header declares {len(params)}: {params}"
        for i, (a, p) in enumerate(zip(args, params)):
            assert klass_c(a) == klass_h(p), f"{name} arg {i}: bound {a} vs header '{p}'"
        want = "void" if res is None else klass_c(res)
        got = "void" if ret == "void" else klass_h(ret + " x") if "*" not in ret else "ptr"
        assert want == got, f"{name}: return bound {res} vs header '{ret}'"


def test_product_path_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under vllm_rs_b200/ may import, load or execute it, and the library
    front door must fail loudly (never fall back) when the CUDA library is missing."""
    import re
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[/.]_build|oracle[/.]_ref|libdecode_ref", re.M)
    for p in sorted((root / "vllm_rs_b200").rglob("*")):
        if p.suffix in (".py", ".cu", ".cuh", ".cpp", ".hpp", ".h"):
            assert not pat.search(p.read_text()), p
    src = (root / "vllm_rs_b200" / "_lib.py").read_text()
    assert "B200Error" in src and "libb200_vllm.so" in src
