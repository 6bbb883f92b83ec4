"""Eytzinger baseline (static-search-tree/src/eytzinger.rs): the reference's own golden vectors
(:199-230) pin the oracle; the GPU builder/search must match the oracle bit for bit."""
import numpy as np
import pytest

from util import MAX, gen_queries, gen_vals

U32MAX = 0xFFFFFFFF


def test_reference_golden_vectors(oracle):
    # eytzinger_test_pow2_min_1 (:199-206)
    assert oracle.eytzinger_build(np.arange(1, 16)).tolist() == [U32MAX, 8, 4, 12, 2, 6, 10, 14, 1, 3, 5, 7, 9, 11, 13, 15]
    # eytzinger_test_non_pow2 (:208-214)
    e = oracle.eytzinger_build(np.arange(0, 10))
    assert e.tolist() == [U32MAX, 6, 3, 8, 1, 5, 7, 9, 0, 2, 4]
    # eyetzinger_search_test / eyetzinger_search_oob (:216-229)
    assert oracle.eytzinger_search(e, [3, 12]).tolist() == [3, U32MAX]


@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 15, 16, 17, 1000, 4097, 65535, 65536, 100_003])
def test_oracle_eytzinger_equals_binary_search(oracle, n):
    """test.rs:203 runs Eytzinger through the same differential test as every other index."""
    vals = gen_vals(n, seed=n)
    qs = gen_queries(2000, seed=n + 1, vals=vals)
    ev, _ = oracle.lower_bound(vals, qs)
    e = oracle.eytzinger_build(vals)
    assert sorted(e[1:].tolist()) == vals.tolist()
    assert np.array_equal(oracle.eytzinger_search(e, qs), ev)  # MAX is a key, so nothing is out of range


@pytest.mark.gpu
@pytest.mark.parametrize("n", [1, 2, 3, 7, 8, 15, 16, 17, 1000, 4097, 65535, 65536, 100_003, 3_000_001])
def test_gpu_eytzinger_matches_oracle(gpu, oracle, n):
    sst = gpu
    vals = gen_vals(n, seed=n + 5) if n > 16 else np.arange(n, dtype=np.uint32) * 3 + 1
    t = sst.Eytzinger.new(vals)
    e = oracle.eytzinger_build(vals)
    assert np.array_equal(t.image(), e)
    assert t.size() == 4 * (n + 1) and t.layers() == int(np.log2(n + 1)) + 1
    qs = np.concatenate([gen_queries(5000, seed=n, vals=vals), np.array([0, 1, MAX, U32MAX, 0x80000000], np.uint32)])
    v, i = t.query(qs, want_index=True)
    assert np.array_equal(v, oracle.eytzinger_search(e, qs))
    assert np.array_equal(i, np.searchsorted(vals, qs, side="left").astype(np.uint64))


@pytest.mark.gpu
def test_gpu_eytzinger_golden(gpu):
    sst = gpu
    assert sst.Eytzinger.new(np.arange(1, 16, dtype=np.uint32)).image().tolist() == [U32MAX, 8, 4, 12, 2, 6, 10, 14, 1, 3, 5, 7, 9, 11, 13, 15]
    t = sst.Eytzinger.new(np.arange(0, 10, dtype=np.uint32))
    assert t.image().tolist() == [U32MAX, 6, 3, 8, 1, 5, 7, 9, 0, 2, 4]
    assert t.query(np.array([3, 12], np.uint32)).tolist() == [3, U32MAX]
    with pytest.raises(sst.SstError):
        sst.Eytzinger.new(np.array([3, 1, 2], np.uint32))
