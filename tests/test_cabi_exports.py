"""The C-ABI boundary: the product library exports exactly the entry points include/*.h declares, and fails loudly without a GPU
(no CPU fallback).  No compute calls here: this tier runs where no device exists."""
import ctypes as C
import os
import re
import subprocess

import pytest

from calitas_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for h in ("calitas_b200.h", "calitas_b200_tools.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = re.sub(r"static inline[^\n]*\n", "\n", text)          # the record accessors are header-only (static inline), not exports
        names += re.findall(r"\b(calitas_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


@pytest.fixture(scope="module")
def product():
    if not os.path.exists(_capi.PRODUCT_LIB):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "calitas_b200", "csrc")])
    return _capi.Library()


def test_every_declared_symbol_is_exported(product):
    decl = declared_functions()
    assert len(decl) >= 20
    for name in decl:
        assert hasattr(product.L, name), name
    assert sorted(_capi.EXPORTS) == decl          # the Python binding lists exactly the header's functions


def test_exported_symbols_are_plain_c(product):
    out = subprocess.run(["nm", "-D", "--defined-only", product.path], capture_output=True, text=True).stdout
    syms = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for name in declared_functions():
        assert name in syms, name


def test_struct_layouts_match_the_header():
    assert C.sizeof(_capi.Hit) == 32 and C.sizeof(_capi.HitWide) == 64 and _capi.Hit.ops.offset == 20 and _capi.HitWide.ops.offset == 20
    assert C.sizeof(_capi.Limits) == 20 and C.sizeof(_capi.Costs) == 16
    assert C.sizeof(_capi.RegionTask) == 24 and C.sizeof(_capi.TargetTask) == 24


def test_no_cpu_fallback(product):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    with pytest.raises(_capi.CalitasError) as ei:
        _capi.Engine(0, lib=product)
    assert ei.value.code == 2 and "no CPU fallback" in ei.value.message       # CALITAS_ECUDA


def test_product_does_not_link_the_oracle_or_hostsim(product):
    out = subprocess.run(["ldd", product.path], capture_output=True, text=True).stdout
    assert "oracle" not in out and "hostsim" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "calitas_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(root, f), errors="replace").read()
                assert "pyoracle" not in text and "libcalitas_oracle" not in text, f
