"""CPU-side checks of the C-ABI library: it loads and exports every symbol the header declares.
No compute calls (no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "tes_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(tes_[a-z0-9_]+)\s*\(", txt)))


def test_library_loads_and_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from tes_b200 import _lib
    L = ctypes.CDLL(_lib.SO_PATH)
    syms = _header_symbols()
    assert len(syms) >= 6
    for s in syms:
        assert hasattr(L, s), "libtes_b200.so does not export %s" % s
    # the Python binding table covers exactly the header
    assert sorted(_lib.SIGNATURES) == syms
    assert _lib.lib().tes_abi_version() == 1


def test_argument_validation_without_gpu():
    from tes_b200 import _lib
    L = _lib.lib()
    assert L.tes_rff_topk_workspace_bytes(100, 0, 4, 8, 1) < 0      # d = 0
    assert L.tes_rff_topk_workspace_bytes(100, 2, 4, 8, 17) < 0     # k > 16
    assert L.tes_rff_topk_workspace_bytes(100, 2, 4, 8, 3) > 0
    # NULL pointers are rejected before any CUDA call
    assert L.tes_rff_topk_f64(None, 10, 2, None, None, None, 4, 8, 0, 1.0, 1, 0, None, None, None, 0, None) == -1
    assert b"argument" in L.tes_error_string(-4)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "trusted-maximizers-entropy-search-bo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("oracle/tes_oracle.c", "").replace("the oracle", "") \
                    .replace("oracle's", ""), "%s references the oracle" % f
