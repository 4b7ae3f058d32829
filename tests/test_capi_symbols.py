"""CPU: the C-ABI library loads, exports every symbol include/sbo_b200.h declares, and refuses to
create a context without a GPU (no CPU fallback).  No compute calls here."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "sbo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sbo_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    import sbo_b200
    from sbo_b200 import _capi
    names = declared_functions()
    assert len(names) >= 25
    lib = ctypes.CDLL(_capi.LIB_PATH)
    for nm in names:
        assert hasattr(lib, nm), f"{nm} declared in sbo_b200.h but not exported"
        assert nm in _capi.SIGNATURES, f"{nm} has no ctypes signature"
    assert sorted(_capi.SIGNATURES) == names
    assert _capi.load().sbo_version() == 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import sbo_b200
    with pytest.raises(sbo_b200.SboError, match="no CUDA device|no CPU fallback"):
        sbo_b200.GridEngine(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "safe-bayesian-optimization_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "gp_oracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f
