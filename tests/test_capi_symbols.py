"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every
symbol include/lnaphylodyn_b200.h declares; compute calls fail loudly without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "lnaphylodyn_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lna_[A-Za-z0-9_]+)\s*\(", txt)))


def test_header_symbols_exported():
    from lnaphylodyn_b200 import _lib
    L = _lib.lib()
    decl = declared_symbols()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(L, name), f"{name} declared in the header but not exported by the .so"
    assert set(decl) == set(_lib.exported_symbols()), set(decl) ^ set(_lib.exported_symbols())
    assert L.lna_abi_version() == 1


def test_shard_helper():
    import ctypes as C
    from lnaphylodyn_b200 import _lib
    L = _lib.lib()
    covered = []
    for r in range(8):
        b, e = C.c_int64(), C.c_int64()
        assert L.lna_shard(8192 + 5, r, 8, C.byref(b), C.byref(e)) == 0
        covered.append((b.value, e.value))
    assert covered[0][0] == 0 and covered[-1][1] == 8197
    assert all(covered[i][1] == covered[i + 1][0] for i in range(7))
    assert max(e - b for b, e in covered) - min(e - b for b, e in covered) <= 1


def test_no_cpu_fallback(changepoint):
    """Without a CUDA device every compute entry point must fail loudly (never silently run on CPU)."""
    from lnaphylodyn_b200 import _lib, api
    if _lib.lib().lna_device_count() > 0:
        pytest.skip("a GPU is present")
    g = changepoint
    with pytest.raises(_lib.LnaError, match="no CUDA device"):
        api.problem_for(g.times, g.x_r, g.x_i, g.gridsize, g.t_correct, g.init)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under lnaphylodyn_b200/ may reference it."""
    pkg = os.path.join(ROOT, "lnaphylodyn_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "oracle" not in src.lower() or f == "__init__.py" and False, f"{f} mentions the oracle"
