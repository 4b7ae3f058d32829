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


def test_header_is_plain_c_and_the_ctypes_structs_mirror_it(tmp_path):
    """include/sbo_b200.h compiles as C99 (no C++ / torch types on the boundary), and the ctypes mirrors of its result
    structs have the header's size and field offsets -- a field added on one side only would silently shift every later
    field the binding reads."""
    import subprocess
    from sbo_b200 import _capi
    mirrors = {"sbo_sets_result": _capi.SetsResult, "sbo_pair_result": _capi.PairResult,
               "sbo_step_result": _capi.StepResult, "sbo_pairs_info": _capi.PairsInfo}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "sbo_b200.h"', 'int main(void) {']
    for cname, cls in mirrors.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append(f'  printf("SBO_MAX_G %d\\n", SBO_MAX_G);')
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines) + "\n")
    exe = tmp_path / "layout"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout
    got = {}
    for ln in out.splitlines():
        parts = ln.split()
        got[tuple(parts[:-1])] = int(parts[-1])
    assert got[("SBO_MAX_G",)] == _capi.MAX_G
    for cname, cls in mirrors.items():
        assert got[(cname, "size")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
