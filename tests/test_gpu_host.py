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


@pytest.mark.gpu
def test_bind_thread_to_device(gpu):
    """sst_bind_thread_to_device: CPU set local to the GPU (0 = topology not visible, nothing changed); the thread keeps a
    non-empty affinity inside the set it started with, and queries still work afterwards."""
    import os

    sst = gpu
    before = os.sched_getaffinity(0)
    try:
        n = sst.bind_thread_to_device(0)
        assert n >= 0
        after = os.sched_getaffinity(0)
        assert after and after <= before
        if n:
            assert len(after) == n
    finally:
        os.sched_setaffinity(0, before)
