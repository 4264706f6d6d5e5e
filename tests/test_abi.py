"""The C-ABI library loads and exports every symbol include/lidiff_b200.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "lidiff_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lb2_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    from lidiff_b200 import _lib
    if not os.path.exists(_lib._SO):
        import __graft_entry__ as g
        g.build()
    lib = _lib.get_lib()
    decl = _declared()
    assert len(decl) >= 18
    missing = [s for s in decl if not hasattr(lib.dll, s)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == decl, set(_lib.EXPORTS) ^ set(decl)
    assert lib.dll.lb2_version() >= 100


def test_no_cpu_fallback():
    import torch
    from lidiff_b200 import _lib, me
    with pytest.raises(RuntimeError):
        _lib.get_lib().handle("cpu")
    with pytest.raises(RuntimeError):
        me.TensorField(torch.zeros(4, 3), torch.zeros(4, 4))
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            _lib.get_lib().handle("cuda:0")


def test_product_never_imports_oracle():
    """the oracle is test infrastructure: nothing under lidiff_b200/ may import it"""
    bad = []
    for dp, _, fns in os.walk(os.path.join(ROOT, "lidiff_b200")):
        for fn in fns:
            if fn.endswith(".py"):
                txt = open(os.path.join(dp, fn)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M):
                    bad.append(os.path.join(dp, fn))
    assert not bad, bad
