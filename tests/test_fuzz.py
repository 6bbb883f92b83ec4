"""Seeded fuzzing: random sizes and adversarial key distributions (clusters, long runs of duplicates,
tiny ranges, keys hugging 2^16 boundaries) through every layout and kernel.
CPU part: the oracle against numpy.searchsorted.  GPU part: CUDA against the oracle, bit-exact."""
import numpy as np
import pytest

from util import MAX


def make_keys(rng, n, kind):
    if kind == "uniform":
        v = rng.integers(0, MAX, n)
    elif kind == "clustered":
        centers = rng.integers(0, MAX - (1 << 20), max(1, n // 5000 + 1))
        v = centers[rng.integers(0, centers.size, n)] + rng.integers(0, 1 << 12, n)
    elif kind == "dupes":
        v = rng.integers(0, max(2, n // 50), n) * 977
    elif kind == "tiny_range":
        v = rng.integers(1000, 1000 + max(2, n // 3), n)
    elif kind == "boundary16":
        v = (rng.integers(1, 1 << 14, n) << 16) + rng.integers(-3, 4, n)
    else:
        raise ValueError(kind)
    v = np.clip(v, 0, MAX).astype(np.uint32)
    if rng.random() < 0.7:
        v[0] = MAX
    v.sort()
    return v


def make_queries(rng, vals, nq):
    a = rng.integers(0, MAX, nq // 3)
    b = vals[rng.integers(0, vals.size, nq // 3)].astype(np.int64) + rng.integers(-2, 3, nq // 3)
    c = rng.integers(int(vals[0]), int(vals[-1]) + 1, nq - 2 * (nq // 3))
    return np.clip(np.concatenate([a, b, c]), 0, MAX).astype(np.uint32)


KINDS = ["uniform", "clustered", "dupes", "tiny_range", "boundary16"]


@pytest.mark.parametrize("seed", range(10))
def test_oracle_vs_numpy(oracle, seed):
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(1, 60_000))
    vals = make_keys(rng, n, KINDS[seed % len(KINDS)])
    qs = make_queries(rng, vals, 3000)
    ev, ei = oracle.lower_bound(vals, qs)
    want_i = np.searchsorted(vals, qs, side="left")
    assert np.array_equal(ei, want_i.astype(np.uint64))
    assert np.array_equal(ev, np.where(want_i < n, vals[np.minimum(want_i, n - 1)], MAX).astype(np.uint32))
    for lm in (False, True):
        v, s = oracle.Tree.stree(vals, left_max=lm).search(qs, want_slot=True)
        assert np.array_equal(v, ev) and np.array_equal(s, ei)
    if vals[-1] > 0:
        for var in ("simple", "compact", "l1", "overlap", "map"):
            t = oracle.Tree.pstree(vals, int(rng.integers(0, 21)), var)
            if t is not None:
                v = t.search(qs)
                ok = (qs <= vals[-1])  # queries above every key: defined as MAX, compared separately below
                assert np.array_equal(v[ok], ev[ok]), var


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(12))
def test_gpu_fuzz(gpu, oracle, seed):
    sst = gpu
    rng = np.random.default_rng(2000 + seed)
    n = int(rng.integers(1, 400_000)) if seed % 3 else int(rng.integers(600_000, 3_000_000))
    kind = KINDS[seed % len(KINDS)]
    vals = make_keys(rng, n, kind)
    nq = int(rng.integers(1, 5000)) if seed % 4 == 0 else int(rng.integers(600_000, 1_200_000))  # small and table-kernel-sized batches
    qs = make_queries(rng, vals, max(nq, 3))
    ev, ei = oracle.lower_bound(vals, qs)
    flags = [(bool(rng.integers(0, 2)), False, False), (bool(rng.integers(0, 2)), True, False), (True, False, True)]
    for lm, rev, full in flags:
        t = sst.STree16.new_params(vals, lm, rev, full)
        assert np.array_equal(t.image(), oracle.Tree.stree(vals, left_max=lm, reverse=rev, full=full).image())
        for scheme in (sst.SCHEME_AUTO, sst.SCHEME_TABLE, sst.SCHEME_GROUP2, sst.SCHEME_GROUP4, sst.SCHEME_GROUP16, sst.SCHEME_GENERIC, sst.SCHEME_BINSEARCH):
            v, i = t.query(qs, scheme, want_index=True)
            assert np.array_equal(v, ev) and np.array_equal(i, ei), (kind, n, lm, rev, full, scheme)
    t15 = sst.STree15.new_params(vals, bool(rng.integers(0, 2)), False, False)
    v, i = t15.query(qs, want_index=True)
    assert np.array_equal(v, ev) and np.array_equal(i, ei), (kind, n, "stree15")
    if vals[-1] > 0:
        b = int(rng.integers(0, 21))
        for name, cls in (("simple", sst.PartitionedSTree16), ("compact", sst.PartitionedSTree16C), ("l1", sst.PartitionedSTree16L),
                          ("overlap", sst.PartitionedSTree16O), ("map", sst.PartitionedSTree16M)):
            ot = oracle.Tree.pstree(vals, b, name)
            t = cls.try_new(vals, b)
            assert (ot is None) == (t is None), (kind, n, name, b)
            if t is None:
                continue
            assert t.params == ot.params and np.array_equal(t.image(), ot.image()), (kind, n, name, b)
            for scheme in (sst.SCHEME_AUTO, sst.SCHEME_GENERIC):
                v, i = t.query(qs, scheme, want_index=True)
                assert np.array_equal(v, ev) and np.array_equal(i, ei), (kind, n, name, b, scheme)
