"""The C-ABI shared library loads and exports every symbol include/dktb200.h declares (no compute, no GPU)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    import __graft_entry__ as ge
    from deep_kernel_transfer_b200 import _lib
    path = ge.build()
    lib = _lib.DktbLib(path)
    header = open(os.path.join(ROOT, "include", "dktb200.h")).read()
    declared = set(re.findall(r"\b(dktb_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    missing = sorted(d for d in declared if not hasattr(lib._c, d))
    assert not missing, missing
    unbound = sorted(d for d in declared if d not in _lib.SIGNATURES)
    assert not unbound, unbound
    assert lib.dktb_version() >= 100


def test_product_has_no_oracle_import():
    """The product package must never import the oracle or the emulation library."""
    pkg = os.path.join(ROOT, "deep_kernel_transfer_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
                assert "emu" not in src or f == "_lib.py", f
