"""CPU: the C-ABI library builds for sm_100a, loads, exports every symbol include/gdiet_cuda.h declares,
and fails loudly (no CPU fallback) when there is no GPU."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "gdiet_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\([^;{}]*\)\s*;", txt)
    return sorted(set(n for n in names if n not in ("defined",)))


def test_library_builds_and_exports_every_declared_symbol(gd):
    gd.build()
    L = gd.load()
    declared = _declared()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(L, name), "libgdiet_cuda.so does not export %s" % name
    assert sorted(gd.EXPORTS) == declared


def test_no_silent_fallback_without_gpu(gd):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gd.GdietError):
        gd.Context(0)


def test_product_does_not_reference_the_oracle():
    """nothing under the product tree may include / import / dlopen oracle code"""
    bad = []
    prod = os.path.join(ROOT, "genome-on-diet_b200")
    for d, _, files in os.walk(prod):
        for f in files:
            if f.endswith((".cu", ".cuh", ".h", ".py", ".c", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"gd_oracle|libgd_oracle|oraclelib|libgdref|oracle/_ref", txt):
                    bad.append(os.path.join(d, f))
    assert not bad, bad
