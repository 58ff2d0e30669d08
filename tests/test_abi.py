"""CPU checks of the drop-in boundary: libvbnn.so loads, exports every symbol include/vbnn.h
declares, reports config.lua's defaults, and fails loudly (no CPU fallback) without a GPU."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "vbnn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vbnn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vbnn_b200 import _lib
    lib = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 50
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vbnn.h but not exported"
    # and the ctypes table binds exactly the declared set
    assert sorted(_lib._PROTOS) == syms


def test_abi_version_and_defaults_match_config_lua():
    from vbnn_b200 import _lib
    lib = _lib.lib()
    assert lib.vbnn_abi_version() == 2
    o = _lib.VbnnOpts()
    lib.vbnn_opts_default(o)
    assert abs(o.var_init - 0.001) < 1e-9 and o.mu_init == 0 and o.S == 30       # config.lua:32,43-44
    assert o.B == 1e6                                                            # config.lua:30
    assert abs(o.lr_bias - 1e-3) < 1e-9 and abs(o.lr_mu - 1e-4) < 1e-9 and abs(o.lr_var - 0.05) < 1e-9
    assert o.strict_reference == 1 and o.reparam == 0


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from vbnn_b200 import _lib
    lib = _lib.lib()
    h = C.c_void_p()
    rc = lib.vbnn_ctx_create(0, None, 3, C.byref(h))
    assert rc == _lib.E_CUDA
    assert b"no CPU fallback" in lib.vbnn_last_error()
    import vbnn_b200
    with pytest.raises(vbnn_b200.VbnnError):
        vbnn_b200.Context(0)


def test_python_mirror_has_the_reference_interface():
    import vbnn_b200
    for name in ["sample", "compute_prior", "compute_mugrads", "compute_vargrads", "calc_lc", "clamp_to_map",
                 "accGradParameters", "resetAcc", "update", "updateOutput", "updateGradInput", "parameters",
                 "forward", "backward"]:                                         # VBLinear.lua:7-166
        assert callable(getattr(vbnn_b200.VBLinear, name)), name
    for name in ["buildModel", "resetGradients", "sample", "run", "test", "calc_lc", "update"]:   # mlp.lua
        assert callable(getattr(vbnn_b200.MLP, name)), name
    opt = vbnn_b200.default_opt()
    assert opt["S"] == 30 and opt["hidden"] == [10] and opt["B"] == 1e6


def test_lua_shim_binds_only_exported_symbols():
    """lua/vbnn_ffi.lua cannot be executed here (no LuaJIT); at least every function its ffi.cdef declares and
    every C.<name> call in lua/*.lua must be a symbol libvbnn.so exports."""
    import ctypes
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    lib = ctypes.CDLL(os.path.join(root, "vbnn_b200", "libvbnn.so"))
    cdef = open(os.path.join(root, "lua", "vbnn_ffi.lua")).read()
    declared = set(re.findall(r"\b(vbnn_[a-z0-9_]+)\s*\(", cdef))
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), name
    for fn in ("VBLinear.lua", "mlp.lua", "vbnn_ffi.lua"):
        src = open(os.path.join(root, "lua", fn)).read()
        for name in set(re.findall(r"\bC\.(vbnn_[a-z0-9_]+)", src)):
            assert name in declared, (fn, name)


def test_debug_knobs_and_ctx_create_ex_argument_checks():
    """(no GPU needed) vbnn_debug_knob round trip, unknown names, and the argument validation of
    vbnn_ctx_create_ex that precedes any CUDA call."""
    from vbnn_b200 import _lib
    lib = _lib.lib()
    old = _lib.knob("tc_bn", 256)
    assert _lib.knob("tc_bn", old) == 256                                # returns the previous value
    assert _lib.knob("tc_bn") == old and _lib.knob("tc_bn") == old       # KNOB_DEFAULT restores env / default
    with pytest.raises(_lib.VbnnError) as e:
        _lib.knob("no_such_knob", 1)
    assert e.value.code == _lib.E_INVALID
    h = C.c_void_p()
    assert lib.vbnn_ctx_create_ex(0, None, 7, 3, C.byref(h)) == _lib.E_INVALID          # bad stream mode
    fake = C.c_void_p(0x10)
    assert lib.vbnn_ctx_create_ex(0, fake, _lib.STREAM_LEGACY_DEFAULT, 3, C.byref(h)) == _lib.E_INVALID   # stream must be NULL
    assert b"stream must be NULL" in lib.vbnn_last_error()


def test_lua_shim_blocks_are_balanced():
    """No Lua interpreter exists here: at least every function / if / for / while / do has its `end` and the brackets match
    in the three shim files (a slip there would only surface on a user's machine)."""
    for fn in ("vbnn_ffi.lua", "VBLinear.lua", "mlp.lua"):
        src = open(os.path.join(ROOT, "lua", fn)).read()
        src = re.sub(r"\[\[.*?\]\]", "", src, flags=re.S)
        src = re.sub(r"--[^\n]*", "", src)
        src = re.sub(r"'[^'\n]*'|\"[^\"\n]*\"", "''", src)
        depth = pending_do = 0
        for t in re.findall(r"\b(function|if|for|while|do|end)\b", src):
            if t in ("function", "if"):
                depth += 1
            elif t in ("for", "while"):
                depth += 1; pending_do += 1
            elif t == "do":
                if pending_do:
                    pending_do -= 1
                else:
                    depth += 1
            else:
                depth -= 1
            assert depth >= 0, fn
        assert depth == 0, fn
        assert src.count("(") == src.count(")") and src.count("{") == src.count("}"), fn
