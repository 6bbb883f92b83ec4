"""Golden fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py with the oracle):
the oracle must keep reproducing them (CPU) and the CUDA path must match them (GPU)."""
import hashlib
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _lb():
    return np.load(os.path.join(GOLD, "lower_bound.npz"))


def _sa():
    return np.load(os.path.join(GOLD, "sa_search.npz"))


KW = {"plain": {}, "left_max": {"left_max": True}, "reverse": {"reverse": True}, "full": {"left_max": True, "full": True}}


def test_oracle_reproduces_golden(oracle):
    g = _lb()
    ev, ei = oracle.lower_bound(g["vals"], g["qs"])
    assert np.array_equal(ev, g["values"]) and np.array_equal(ei, g["indices"])
    want = dict(zip(g["image_names"], g["image_sha256"]))
    for name, kw in KW.items():
        assert sha(oracle.Tree.stree(g["vals"], **kw).image()) == want["stree_" + name]
    for var in ("simple", "compact", "l1", "overlap", "map"):
        assert sha(oracle.Tree.pstree(g["vals"], 8, var).image()) == want["pstree_" + var]
    s = _sa()
    assert np.array_equal(oracle.sa_build(s["text"]), s["sa"])
    flat = np.concatenate([s["flat"], np.zeros(64, np.uint8)])
    lo, hi, pos, _ = oracle.sa_search(s["text"], s["sa"], flat, s["off"])
    assert np.array_equal(lo, s["lo"]) and np.array_equal(hi, s["hi"]) and np.array_equal(pos, s["pos"])


@pytest.mark.gpu
def test_gpu_matches_golden(gpu):
    sst = gpu
    g = _lb()
    want = dict(zip(g["image_names"], g["image_sha256"]))
    for name, kw in KW.items():
        t = sst.STree16.new_params(g["vals"], kw.get("left_max", False), kw.get("reverse", False), kw.get("full", False))
        assert sha(t.image()) == want["stree_" + name]
        v, i = t.query(g["qs"], want_index=True)
        assert np.array_equal(v, g["values"]) and np.array_equal(i, g["indices"])
    classes = {"simple": sst.PartitionedSTree16, "compact": sst.PartitionedSTree16C, "l1": sst.PartitionedSTree16L,
               "overlap": sst.PartitionedSTree16O, "map": sst.PartitionedSTree16M}
    for var, cls in classes.items():
        t = cls.new(g["vals"], 8)
        assert sha(t.image()) == want["pstree_" + var]
        v, i = t.query(g["qs"], want_index=True)
        assert np.array_equal(v, g["values"]) and np.array_equal(i, g["indices"])
    s = _sa()
    sa = sst.SaNaive.build(s["text"])
    assert np.array_equal(sa.sa, s["sa"])
    for mode in (sst.SA_BINARY, sst.SA_MLR):
        lo, hi, pos = sa.search(s["flat"], s["off"], mode)
        assert np.array_equal(lo, s["lo"]) and np.array_equal(hi, s["hi"]) and np.array_equal(pos, s["pos"])
