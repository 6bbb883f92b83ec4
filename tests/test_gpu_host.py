"""GPU test: the C++ host mirror (suffix-array-searching_b200/host/sst.hpp) running the reference's own
tests through the C ABI."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "suffix-array-searching_b200", "host")


@pytest.mark.gpu
def test_cpp_host_mirror(gpu):
    exe = os.path.join(HOST, "test_host")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", HOST, "-s"])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "host mirror OK" in out.stdout
