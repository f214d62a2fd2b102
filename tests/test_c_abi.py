"""The C-ABI library loads and exports every symbol include/vrr.h declares (no compute calls: CPU)."""
import ctypes
import os
import re

import pytest

from vit_rpe_rope_b200 import _lib


@pytest.fixture(scope="module")
def lib():
    if not os.path.isfile(_lib.LIB_PATH):
        import __graft_entry__ as g
        g.build()
    return _lib.load()


def _declared_functions():
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vrr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = _declared_functions()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vrr.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype in _lib.SIGNATURES"
    assert set(_lib.SIGNATURES) == set(names)


def _prototypes():
    """name -> list of parameter kinds ('p' pointer, 'i' int, 'f' float, 'z' size_t) parsed from the header."""
    src = open(_lib.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(vrr_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        params = [q.strip() for q in m.group(2).split(",") if q.strip() and q.strip() != "void"]
        kinds = []
        for q in params:
            if "*" in q:
                kinds.append("p")
            elif re.match(r"(const\s+)?float\b", q):
                kinds.append("f")
            elif re.match(r"(const\s+)?size_t\b", q):
                kinds.append("z")
            else:
                kinds.append("i")
        out[m.group(1)] = kinds
    return out


def test_ctypes_prototypes_match_the_header():
    """Every ctypes argtypes list has the header prototype's parameter count and pointer / int / float kinds: a
    drifted binding would pass garbage through the C ABI without any error."""
    kind_of = {ctypes.c_void_p: "p", ctypes.c_char_p: "p", ctypes.c_int: "i", ctypes.c_float: "f", ctypes.c_size_t: "z"}
    protos = _prototypes()
    assert set(protos) == set(_lib.SIGNATURES)
    for name, (_, argtypes) in _lib.SIGNATURES.items():
        got = ["p" if (isinstance(a, type) and issubclass(a, ctypes._Pointer)) else kind_of[a] for a in argtypes]
        assert got == protos[name], (name, got, protos[name])


def test_abi_version_and_enums(lib):
    assert lib.vrr_abi_version() == 1
    hdr = open(_lib.HEADER_PATH).read()
    for name, val in (("VRR_F32", _lib.VRR_F32), ("VRR_BF16", _lib.VRR_BF16), ("VRR_ROPE_AXIAL", _lib.ROPE_AXIAL),
                      ("VRR_ROPE_MIXED", _lib.ROPE_MIXED), ("VRR_BIAS_TABLE", _lib.BIAS_TABLE),
                      ("VRR_BIAS_POLY", _lib.BIAS_POLY), ("VRR_IMPL_TCGEN05", _lib.IMPL_TCGEN05)):
        assert re.search(rf"\b{name}\s*=\s*{val}\b", hdr), name
    assert ctypes.sizeof(_lib.BiasDesc) == 24


def test_argument_validation_without_device(lib):
    """Bad arguments are rejected before any CUDA call; a missing sm_100 device is an error, never a fallback."""
    rc = lib.vrr_attn_fwd(None, None, None, None, 1, 1, 1, 64, 0.125, 0, None)
    assert rc == -1 and "NULL" in _lib.last_error()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.vrr_attn_fwd(p, None, p, p, 1, 1, 4, 48, 0.125, 0, None)
    assert rc == -2 and "head dim" in _lib.last_error()
    d = _lib.BiasDesc(mode=_lib.BIAS_TABLE, heads=2, len=5, grid=0, param=p.value)
    rc = lib.vrr_attn_fwd(p, ctypes.byref(d), p, p, 1, 2, 4, 64, 0.125, 0, None)
    assert rc == -1 and "2N-1" in _lib.last_error()
    import torch
    if not torch.cuda.is_available():
        rc = lib.vrr_attn_fwd(p, None, p, p, 1, 1, 4, 64, 0.125, 0, None)
        assert rc == -3, "without a GPU the library must fail loudly"
        assert lib.vrr_device_ok() == 0


def test_workspace_query(lib):
    n = lib.vrr_attn_bwd_workspace_bytes(2, 3, 65, 32, None)
    assert n >= 2 * 3 * 65 * 4 and n % 256 == 0
