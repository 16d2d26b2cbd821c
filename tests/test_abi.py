"""The C-ABI library loads without a GPU and exports every symbol include/katome_gpu.h declares."""
import os
import re

import pytest

from katome_b200 import _lib
from tests.helpers import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "katome_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ktg_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    L = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(_lib.SYMBOLS) == declared


def test_no_cpu_fallback_without_a_device():
    L = _lib.lib()
    if L.ktg_device_count() > 0:
        pytest.skip("a GPU is present")
    from katome_b200 import GpuGIR, KatomeError
    with pytest.raises(KatomeError) as e:
        GpuGIR(31)
    assert e.value.code == _lib.KTG_ERR_NO_DEVICE


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "katome_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r'#\s*include\s*[<"][^>"]*oracle', src), f
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libkatome_oracle" not in src and "dlopen" not in src, f
