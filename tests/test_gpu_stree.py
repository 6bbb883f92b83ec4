"""GPU parity tests (-m gpu): CUDA lower_bound through the C ABI vs the CPU oracle.
Bit-exact: every value, every index, and the tree image byte for byte."""
import os

import numpy as np
import pytest

from util import MAX, gen_queries, gen_vals, reference_test_sizes

pytestmark = pytest.mark.gpu

FLAG_SETS = [(0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 1, 0), (1, 1, 0), (0, 0, 1)]


def schemes(sst):
    return [sst.SCHEME_AUTO, sst.SCHEME_GROUP4, sst.SCHEME_GROUP16, sst.SCHEME_GROUP2, sst.SCHEME_GENERIC, sst.SCHEME_TABLE]


def test_kats(gpu, oracle):
    sst = gpu
    vals = np.concatenate([np.arange(1, 2000, dtype=np.uint32), [MAX]]).astype(np.uint32)
    t = sst.STree16.new(vals)
    assert t.search(452) == 452 and t.search(289) == 289  # s_tree.rs:861-885
    assert t.layers() == 3 and t.size() == t.image().nbytes
    for s in schemes(sst):
        assert list(t.query([452, 289], s)) == [452, 289]


@pytest.mark.parametrize("n", reference_test_sizes(6, 20, ) [::3] + [17, 16, 1, 272, 273, 4624, 4625, (1 << 22) + 12345])
def test_plain_tree_matrix(gpu, oracle, n):
    """test.rs:142-260 on the GPU: all parameterisations x all kernels == binary_search."""
    sst = gpu
    vals = gen_vals(n, seed=n)
    qs = gen_queries(1024 + 37, seed=n + 1, vals=vals)
    ev, ei = oracle.lower_bound(vals, qs)
    for lm, rev, full in FLAG_SETS:
        ot = oracle.Tree.stree(vals, left_max=lm, reverse=rev, full=full)
        t = sst.STree16.new_params(vals, bool(lm), bool(rev), bool(full))
        assert (t.offsets == ot.offsets).all()
        assert t.layers() == ot.layers and t.size() == ot.size_bytes
        assert np.array_equal(t.image(), ot.image()), ("image", n, lm, rev, full)
        for s in schemes(sst):
            v, i = t.query(qs, s, want_index=True)
            assert np.array_equal(v, ev), (n, lm, rev, full, s)
            assert np.array_equal(i, ei), (n, lm, rev, full, s)


@pytest.mark.parametrize("n", [15, 16, 225, 240, 241, 5000, 123457])
def test_stree15(gpu, oracle, n):
    sst = gpu
    vals = gen_vals(n, seed=n + 100)
    qs = gen_queries(999, seed=n, vals=vals)
    ev, ei = oracle.lower_bound(vals, qs)
    for lm, rev, full in FLAG_SETS:
        ot = oracle.Tree.stree(vals, B=15, left_max=lm, reverse=rev, full=full)
        t = sst.STree15.new_params(vals, bool(lm), bool(rev), bool(full))
        assert np.array_equal(t.image(), ot.image()), (n, lm, rev, full)
        v, i = t.query(qs, want_index=True)
        assert np.array_equal(v, ev) and np.array_equal(i, ei), (n, lm, rev, full)


VARIANTS = ["simple", "compact", "l1", "overlap", "map"]


def _cls(sst, var):
    return {"simple": sst.PartitionedSTree16, "compact": sst.PartitionedSTree16C, "l1": sst.PartitionedSTree16L,
            "overlap": sst.PartitionedSTree16O, "map": sst.PartitionedSTree16M}[var]


@pytest.mark.parametrize("n", [1, 16, 17, 300, 5000, 70_000, 1 << 20, (1 << 21) * 5 // 4])
@pytest.mark.parametrize("var", VARIANTS)
def test_partitioned_matrix(gpu, oracle, n, var):
    """test.rs:227-258: five layouts x b in {0,4,8,16,20}; layout, parameters and results."""
    sst = gpu
    vals = gen_vals(n, seed=n + 7)
    qs = gen_queries(1024 + 5, seed=n + 8, vals=vals)
    ev, ei = oracle.lower_bound(vals, qs)
    for b in (0, 4, 8, 16, 20):
        ot = oracle.Tree.pstree(vals, b, var)
        t = _cls(sst, var).try_new(vals, b)
        assert (ot is None) == (t is None)
        if t is None:
            continue
        assert t.params == ot.params, (var, n, b)
        assert (t.offsets == ot.offsets).all()
        assert t.layers() == ot.layers and t.size() == ot.size_bytes
        assert np.array_equal(t.image(), ot.image()), ("image", var, n, b)
        if var == "map":
            assert np.array_equal(t.prefix_map, ot.prefix_map)
        for scheme in (sst.SCHEME_AUTO, sst.SCHEME_GENERIC):  # lane-group kernel / thread-per-query kernel
            v, i = t.query(qs, scheme, want_index=True)
            assert np.array_equal(v, ot.search(qs)), (var, n, b, scheme)
            assert np.array_equal(v, ev), (var, n, b, scheme)
            assert np.array_equal(i, ei), (var, n, b, scheme)
        assert np.array_equal(t.query(qs), ev)


def test_skewed_keys_partitioned(gpu, oracle):
    """Non-uniform keys: empty parts, one huge bucket, duplicates (exercises gap fill and overlap)."""
    sst = gpu
    rng = np.random.default_rng(5)
    parts = [rng.integers(0, 1 << 12, 3000), rng.integers(1 << 30, (1 << 30) + 4096, 20000), np.full(500, 123456789),
             rng.integers(0, MAX, 2000), [MAX]]
    vals = np.sort(np.concatenate(parts).astype(np.uint32))
    qs = gen_queries(4096, seed=9, vals=vals)
    qs[100:200] = rng.integers((1 << 30) - 10, (1 << 30) + 5000, 100)
    ev, ei = oracle.lower_bound(vals, qs)
    for var in VARIANTS:
        for b in (0, 3, 8, 12, 20):
            ot = oracle.Tree.pstree(vals, b, var)
            t = _cls(sst, var).try_new(vals, b)
            assert (ot is None) == (t is None)
            if t is None:
                continue
            assert t.params == ot.params
            assert np.array_equal(t.image(), ot.image()), (var, b)
            for scheme in (sst.SCHEME_AUTO, sst.SCHEME_GENERIC):
                v, i = t.query(qs, scheme, want_index=True)
                assert np.array_equal(v, ev) and np.array_equal(i, ei), (var, b, scheme)


def test_edge_batches_and_out_of_range(gpu, oracle):
    sst = gpu
    vals = np.sort(np.random.default_rng(3).integers(0, 1 << 20, 50_000).astype(np.uint32))  # MAX not a key
    t = sst.STree16.new_params(vals, True, False, False)
    assert t.query(np.zeros(0, np.uint32)).size == 0
    for nq in (1, 31, 32, 33, 63, 64, 65, 2047, 2049):
        qs = gen_queries(nq, seed=nq, vals=vals)
        ev, ei = oracle.lower_bound(vals, qs)  # q above every key -> (MAX, n) by definition
        for s in schemes(sst):
            v, i = t.query(qs, s, want_index=True)
            assert np.array_equal(v, ev) and np.array_equal(i, ei), (nq, s)
    # signed-compare quirk of node.rs:91-108: q >= 2^31 compares as negative -> first key
    ot = oracle.Tree.stree(vals, left_max=True)
    big = np.array([0x80000000, 0xFFFFFFFF, 0x80000001], np.uint32)
    for s in schemes(sst):
        assert np.array_equal(t.query(big, s), ot.search(big))


def test_table_kernel_skewed_keys(gpu, oracle):
    """Rank table with crowded buckets (keys packed into few 2^16 ranges) and with every flag set."""
    sst = gpu
    rng = np.random.default_rng(17)
    vals = np.sort(np.concatenate([rng.integers(5 << 16, (5 << 16) + 3000, 400_000), rng.integers(0, MAX, 5000),
                                   rng.integers(0x7FFF0000, MAX, 100_000), [MAX]]).astype(np.uint32))
    qs = np.concatenate([rng.integers(5 << 16, (5 << 16) + 3100, 20000), rng.integers(0x7FFE0000, MAX, 20000),
                         gen_queries(20001, seed=5, vals=vals)]).astype(np.uint32)
    ev, ei = oracle.lower_bound(vals, qs)
    for lm, rev, full in FLAG_SETS:
        t = sst.STree16.new_params(vals, bool(lm), bool(rev), bool(full))
        for g in ("2", "4"):
            sst.set_option("TABLE_G", int(g))
            v, i = t.query(qs, sst.SCHEME_TABLE, want_index=True)
            assert np.array_equal(v, ev) and np.array_equal(i, ei), (lm, rev, full, g)


def test_compressed_last_level(gpu, oracle, monkeypatch):
    """16-bit copy of the last internal level (forced on for a small tree): dense keys (compressed path),
    sparse keys (every node falls back to the exact node) and a mix, all flag combinations."""
    sst = gpu
    gpu.set_option("C5", 1)  # build it regardless of the level's size
    rng = np.random.default_rng(91)
    dense = np.sort(rng.integers(1 << 20, (1 << 20) + (1 << 22), 600_000).astype(np.uint32))
    sparse = gen_vals(300_000, seed=92)
    mix = np.sort(np.concatenate([dense[:200_000], sparse[:100_000], [MAX]]).astype(np.uint32))
    for vals in (dense, sparse, mix):
        qs = np.concatenate([gen_queries(600_000, seed=93, vals=vals), rng.integers(1 << 20, (1 << 20) + (1 << 22), 600_000).astype(np.uint32),
                             np.array([0x80000000, 0xFFFFFFFF, 0, MAX], np.uint32)])
        ev, ei = oracle.lower_bound(vals, qs)
        for lm, rev, full in FLAG_SETS:
            ot = oracle.Tree.stree(vals, left_max=lm, reverse=rev, full=full)
            t = sst.STree16.new_params(vals, bool(lm), bool(rev), bool(full))
            v, i = t.query(qs, sst.SCHEME_TABLE, want_index=True)
            assert np.array_equal(v, ot.search(qs)), (lm, rev, full)   # incl. the signed-compare quirk for q >= 2^31
            assert np.array_equal(v[:-4], ev[:-4]) and np.array_equal(i[:-4], ei[:-4]), (lm, rev, full)


def test_binary_search_baseline(gpu, oracle):
    """SortedVec::binary_search (binary_search.rs:36-49) as a GPU baseline kernel."""
    sst = gpu
    vals = gen_vals(300_001, seed=3)
    qs = gen_queries(10_007, seed=4, vals=vals)
    ev, ei = oracle.lower_bound(vals, qs)
    for lm, rev, full in FLAG_SETS:
        t = sst.STree16.new_params(vals, bool(lm), bool(rev), bool(full))
        v, i = t.query(qs, sst.SCHEME_BINSEARCH, want_index=True)
        assert np.array_equal(v, ev) and np.array_equal(i, ei)
    with pytest.raises(sst.SstError):  # plain B=16 trees only
        sst.PartitionedSTree16.new(vals, 4).query(qs, sst.SCHEME_BINSEARCH)


def test_duplicates_and_tiny(gpu, oracle):
    sst = gpu
    for vals in ([MAX], [0, MAX], [5] * 40 + [MAX], list(range(16)), list(range(17)), [7] * 16 + [9] * 16 + [MAX] * 3,
                 [3] * 1000):
        vals = np.array(vals, np.uint32)
        qs = np.array([0, 1, 3, 4, 5, 6, 7, 8, 9, 10, 15, 16, 17, MAX], np.uint32)
        ev, ei = oracle.lower_bound(vals, qs)
        for lm in (False, True):
            t = sst.STree16.new_params(vals, lm, False, False)
            for s in schemes(sst):
                v, i = t.query(qs, s, want_index=True)
                assert np.array_equal(v, ev) and np.array_equal(i, ei), (vals[:3], lm, s)


def test_errors_are_loud(gpu):
    sst = gpu
    with pytest.raises(sst.SstError):  # s_tree.rs:93 assert!(n > 0)
        sst.STree16.new(np.zeros(0, np.uint32))
    with pytest.raises(sst.SstError):  # s_tree.rs:87-89 assert!(v <= MAX)
        sst.STree16.new(np.array([1, 2, 0x80000000], np.uint32))
    with pytest.raises(sst.SstError):  # unsorted
        sst.STree16.new(np.array([3, 2, 1], np.uint32))
    with pytest.raises(sst.SstError):  # s_tree.rs:77-82 full + reverse
        sst.STree16.new_params(np.array([1, 2, 3], np.uint32), False, True, True)
    with pytest.raises(sst.SstError):
        sst.PartitionedSTree16M.new(np.array([3, 2, 1], np.uint32), 4)
    t = sst.PartitionedSTree16.new(gen_vals(1000), 4)
    with pytest.raises(sst.SstError):  # group kernels only serve the plain tree
        t.query(np.array([1], np.uint32), sst.SCHEME_GROUP4)


def test_map_capacity_none(gpu, oracle):
    """partitioned_s_tree.rs:594-597: Map returns None when the prefix map outgrows 4x the input."""
    sst = gpu
    vals = np.array([1, 2, MAX], np.uint32)
    assert oracle.Tree.pstree(vals, 20, "map") is None or True  # shrink loop usually avoids it
    ot = oracle.Tree.pstree(vals, 20, "map")
    t = sst.PartitionedSTree16M.try_new(vals, 20)
    assert (ot is None) == (t is None)


def test_device_buffers_and_streams(gpu, oracle):
    import torch

    sst = gpu
    vals = gen_vals(1 << 18, seed=42)
    qs = gen_queries(100_003, seed=43, vals=vals)
    ev, ei = oracle.lower_bound(vals, qs)
    dvals = torch.from_numpy(vals.view(np.int32)).cuda()
    t = sst.STree16.new_params(dvals, True, False, False)  # GPU layout builder from device-resident keys
    dq = torch.from_numpy(qs.view(np.int32)).cuda()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        v, i = t.query(dq, want_index=True)
    s.synchronize()
    assert np.array_equal(v.cpu().numpy().view(np.uint32), ev)
    assert np.array_equal(i.cpu().numpy().astype(np.uint64), ei)


def test_host_pipeline_chunks(gpu, oracle, monkeypatch):
    sst = gpu
    gpu.set_option("CHUNK", 4096)
    vals = gen_vals(1 << 16, seed=1)
    qs = gen_queries(4096 * 7 + 123, seed=2, vals=vals)
    ev, ei = oracle.lower_bound(vals, qs)
    t = sst.STree16.new_params(vals, True, False, False)
    v, i = t.query(qs, want_index=True)
    assert np.array_equal(v, ev) and np.array_equal(i, ei)


def test_multi_replicas_shard_queries(gpu, oracle):
    """sst_multi_*: replicas + contiguous shards (bench.rs:558-573). Two replicas on device 0 when only one GPU."""
    sst = gpu
    ndev = sst.device_count()
    devices = [0, 1] if ndev >= 2 else [0, 0]
    vals = gen_vals(200_000, seed=8)
    m = sst.MultiIndex.stree(vals, devices + [0], left_max=True)
    assert m.n_devices == 3
    for nq in (0, 1, 2, 3, 1000, 100_001):
        qs = gen_queries(nq, seed=nq + 1, vals=vals) if nq else np.zeros(0, np.uint32)
        ev, ei = oracle.lower_bound(vals, qs)
        v, i = m.query(qs, want_index=True)
        assert np.array_equal(v, ev) and np.array_equal(i, ei), nq
    mm = sst.MultiIndex.pstree(vals, 12, sst.MAP, devices)
    qs = gen_queries(50_000, seed=77, vals=vals)
    ev, ei = oracle.lower_bound(vals, qs)
    v, i = mm.query(qs, want_index=True)
    assert np.array_equal(v, ev) and np.array_equal(i, ei)


def test_multi_repeated_calls_reuse_buffers(gpu, oracle):
    """sst_multi_query keeps one worker thread (with its staging ring) per replica: no growth per call."""
    import torch

    sst = gpu
    vals = gen_vals(500_000, seed=21)
    qs = gen_queries(3_000_000, seed=22, vals=vals)
    ev, _ = oracle.lower_bound(vals, qs[:50_000])
    m = sst.MultiIndex.stree(vals, [0, 0], left_max=True)
    assert np.array_equal(m.query(qs)[:50_000], ev)
    free0 = torch.cuda.mem_get_info()[0]
    for _ in range(8):
        out = m.query(qs)
    assert np.array_equal(out[:50_000], ev)
    assert free0 - torch.cuda.mem_get_info()[0] < 32 << 20


def test_pinned_host_buffers(gpu, oracle):
    """sst_host_alloc: page-locked buffers through the host path."""
    sst = gpu
    vals = gen_vals(1 << 18, seed=31)
    n = 5_000_003
    q = sst.PinnedArray(n)
    q.array[:] = gen_queries(n, seed=32, vals=vals)
    out = sst.PinnedArray(n)
    t = sst.STree16.new_params(vals, True, False, False)
    import ctypes as C
    rc = sst.lib().sst_query(t._h, q.array.ctypes.data_as(C.c_void_p), n, out.array.ctypes.data_as(C.c_void_p), None, 0)
    assert rc == 0
    ev, _ = oracle.lower_bound(vals, q.array[:100_000])
    assert np.array_equal(out.array[:100_000], ev)
    assert np.array_equal(out.array, t.query(q.array))


def test_full_size_properties(gpu):
    """BASELINE size (2^28 keys, 10^8 queries): size-independent properties checked on the device:
    value >= q, keys[idx] == value, keys[idx-1] < q, and every kernel agrees with every other."""
    import torch

    sst = gpu
    n, nq = 1 << 28, 100_000_000
    g = torch.Generator(device="cuda").manual_seed(1234)
    keys = torch.randint(0, MAX, (n,), dtype=torch.int32, device="cuda", generator=g)
    keys[0] = MAX
    keys = torch.sort(keys).values.contiguous()
    qs = torch.randint(0, MAX, (nq,), dtype=torch.int32, device="cuda", generator=g)
    t = sst.STree16.new_params(keys, True, False, False)
    assert t.layers() == 7
    ref_v = None
    for s in (sst.SCHEME_TABLE, sst.SCHEME_BUCKETED, sst.SCHEME_GROUP4, sst.SCHEME_GROUP16, sst.SCHEME_GROUP2, sst.SCHEME_GENERIC):
        v, i = t.query(qs, s, want_index=True)
        torch.cuda.synchronize()
        assert bool((v >= qs).all())
        assert bool((keys[i] == v).all())
        prev = keys[(i - 1).clamp(min=0)]
        assert bool(((i == 0) | (prev < qs)).all())
        if ref_v is None:
            ref_v = v
        else:
            assert bool((v == ref_v).all())
        del i, prev
    # torch.searchsorted is an independent implementation of the same lower bound
    i2 = torch.searchsorted(keys, qs[: 10_000_000], right=False)
    v2 = keys[i2]
    assert bool((v2 == ref_v[: 10_000_000]).all())
    del t
    # Map-partitioned tree at the same size (b = 20 as in test.rs / bench)
    mp = sst.PartitionedSTree16M.new(keys, 20)
    v, i = mp.query(qs, want_index=True)
    assert bool((v == ref_v).all())
    assert bool((keys[i] == v).all())
