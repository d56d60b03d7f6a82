"""The C-ABI library loads, exports every symbol include/poy5_b200.h declares, and refuses to run
without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "poy5_b200.h")).read()
    return sorted(set(re.findall(r"POY_API[^;(]*?\b(poy_\w+)\s*\(", text)))


def test_header_symbols_exported():
    from poy5_b200 import _lib
    L = ctypes.CDLL(_lib.lib_path())
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), n
    assert set(_lib.EXPORTS) <= set(names)


def test_no_cpu_fallback():
    import torch
    import poy5_b200 as pb
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pb.PoyError) as e:
        pb.Context(0)
    assert e.value.status == -2


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "poy5_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".c", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), os.path.join(dirpath, f)
