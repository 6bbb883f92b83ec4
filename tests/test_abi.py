"""CPU tests: the C-ABI library loads and exports every symbol include/sst_b200.h declares
(no compute calls without a GPU), and fails loudly without a device."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sst_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sst_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(sst):
    L = sst.lib()
    names = declared_symbols()
    assert len(names) >= 35
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/sst_b200.h but not exported"


def test_binding_covers_header(sst):
    bound = set(re.findall(r'"(sst_[a-z0-9_]+)":', open(sst.__file__).read()))
    assert set(declared_symbols()) <= bound


def test_version_and_device_count(sst):
    L = sst.lib()
    assert b"sm_100a" in L.sst_version()
    assert sst.device_count() >= 0


def test_no_cpu_fallback(sst):
    """Without a usable sm_100 device every build must fail loudly instead of falling back."""
    if sst.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(sst.SstError):
        sst.STree16.new(np.array([1, 2, 3, sst.MAX], np.uint32))
    with pytest.raises(sst.SstError):
        sst.SaNaive.build(np.zeros(10, np.uint8))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "suffix-array-searching_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", ".rs")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "from oracle" not in text and "import oracle" not in text, f
