"""The C-ABI library builds, loads, exports every symbol include/mtsv_b200.h declares, and refuses to
compute without a GPU (no CPU fallback).  No compute calls here."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "mtsv_b200.h")).read()
    return sorted(set(re.findall(r"MTSVGPU_API [\w\s\*]*?(mtsvgpu_\w+)\(", hdr)))


def test_header_symbols_listed():
    from mtsv_tools_b200 import _lib
    assert _declared_symbols() == sorted(_lib.EXPORTS)


def test_library_loads_and_exports_everything():
    from mtsv_tools_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build_library()
    L = _lib.load_library()
    for sym in _declared_symbols():
        assert getattr(L, sym) is not None
    assert b"sm_100a" in L.mtsvgpu_version()


def test_struct_sizes_match_header():
    import ctypes as C
    from mtsv_tools_b200 import _lib
    assert C.sizeof(_lib.HitStruct) == 24
    assert C.sizeof(_lib.BinStruct) == 24
    assert C.sizeof(_lib.ParamsStruct) == 64
    assert C.sizeof(_lib.OptsStruct) == 24


def test_no_cpu_fallback():
    """Without a CUDA device every computing entry point must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mtsv_tools_b200 import LibraryError
    from mtsv_tools_b200.index import edit_distance
    with pytest.raises(LibraryError) as ei:
        edit_distance([b"ACGT"], [b"ACGA"])
    assert ei.value.code == -4


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "mtsv_tools_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "pyoracle" not in src and "liboracle" not in src and "mtsv_oracle" not in src, f
