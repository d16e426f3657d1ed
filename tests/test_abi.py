"""CPU-side checks of the drop-in boundary: the shared library loads, exports
every symbol include/hifiasm_b200.h declares, and refuses to work without a GPU
(no compute call is made here)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "hifiasm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import hifiasm_b200
    p = hifiasm_b200.lib_path()
    assert os.path.exists(p), "build the extension first: python -c 'import __graft_entry__ as g; g.build()'"
    L = C.CDLL(p)
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "missing export: " + n


def test_no_cpu_fallback():
    import hifiasm_b200
    L = C.CDLL(hifiasm_b200.lib_path())
    L.hb_device_count.restype = C.c_int
    if L.hb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(hifiasm_b200.HBError):
        hifiasm_b200.Engine(0)


def test_product_does_not_reference_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py may touch oracle/"""
    bad = []
    for dp, dn, fn in os.walk(os.path.join(ROOT, "hifiasm_b200")):
        if "build" in dp:
            continue
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                t = open(os.path.join(dp, f), errors="ignore").read()
                if "ha_oracle" in t or "oracle/" in t.replace("tests/hostemu", ""):
                    bad.append(os.path.join(dp, f))
    assert not bad, bad
