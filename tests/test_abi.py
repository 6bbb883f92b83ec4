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


def test_build_rs_compiles_every_source():
    """rust/build.rs must build the same set of CUDA sources as csrc/Makefile ($(wildcard *.cu)): it globs csrc/ and
    hard-codes no file list (round 1 listed six of seven files by hand and the crate could not link)."""
    import glob

    rs = open(os.path.join(ROOT, "suffix-array-searching_b200", "rust", "build.rs")).read()
    assert "read_dir" in rs and 'x == "cu"' in rs
    named = set(re.findall(r'"([a-z_0-9]+\.cu)"', rs))
    assert not named, f"build.rs names CUDA sources by hand: {named}"
    assert len(glob.glob(os.path.join(ROOT, "suffix-array-searching_b200", "csrc", "*.cu"))) >= 7
    mk = open(os.path.join(ROOT, "suffix-array-searching_b200", "csrc", "Makefile")).read()
    assert "$(wildcard *.cu)" in mk


def test_options_without_a_device(sst):
    """The option table is host-side state: it works (and validates) on a box without a GPU."""
    assert sst.get_option("BK_CHUNK2_LOG2") >= 14
    with pytest.raises(sst.SstError):
        sst.set_option("SA_LANES", 0)
    sst.set_option("SA_LANES", 8)
    assert sst.get_option("SST_SA_LANES") == 8
    sst.reset_options()
    assert sst.get_option("SA_LANES") == 1


def test_no_getenv_on_the_query_path():
    """Only runtime.cu (the load-time option table) may read the environment."""
    csrc = os.path.join(ROOT, "suffix-array-searching_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith((".cu", ".cuh")) and f != "runtime.cu":
            text = re.sub(r"//.*", "", open(os.path.join(csrc, f)).read())
            assert "getenv" not in text, f
