"""GPU parity tests (-m gpu) of the reordered-batch pipeline (SCHEME_BUCKETED, csrc/bucketed.cu): partition ->
search from shared memory -> un-permute must give exactly the oracle's values and indices.  Small separator
windows (SST_BK_R) force many buckets on small trees; the full-size run is in test_gpu_stree.py."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

from test_fuzz import KINDS, make_keys, make_queries
from util import MAX, gen_queries, gen_vals

pytestmark = pytest.mark.gpu

FLAG_SETS = [(0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 1, 0), (1, 1, 0)]


def _check(sst, oracle, vals, qs, flags=FLAG_SETS):
    ev, ei = oracle.lower_bound(vals, qs)
    for lm, rev, full in flags:
        t = sst.STree16.new_params(vals, bool(lm), bool(rev), bool(full))
        v, i = t.query(qs, sst.SCHEME_BUCKETED, want_index=True)
        assert np.array_equal(v, ev), (len(vals), len(qs), lm, rev, full, np.flatnonzero(v != ev)[:5])
        assert np.array_equal(i, ei), (len(vals), len(qs), lm, rev, full, np.flatnonzero(i != ei)[:5])
        assert np.array_equal(t.query(qs, sst.SCHEME_BUCKETED), ev)  # values only (no index pass)


@pytest.mark.parametrize("r,n,nq", [(64, 5000, 40_000), (64, 100_000, 100_001), (64, 1_000_000, 70_000), (256, 1 << 20, 300_000),
                                    (1024, 2_000_003, 50_000), (16384, (1 << 22) + 12345, 1_000_003), (32768, 3_000_000, 500_000), (64, 17, 33), (64, 600, 1)])
def test_bucketed_uniform(gpu, oracle, monkeypatch, r, n, nq):
    gpu.set_option("BK_MIN_N", 0)
    gpu.set_option("BK_R", int(str(r)))
    vals = gen_vals(n, seed=n + r)
    _check(gpu, oracle, vals, gen_queries(nq, seed=n + 1, vals=vals))


@pytest.mark.parametrize("r,n,nq", [(64, 5000, 40_000), (64, 1_000_000, 70_000), (256, (1 << 20) + 5, 300_000), (32768, 3_000_000, 500_000), (64, 16, 33),
                                    (64, 17, 5)])
def test_bucketed_node_granularity(gpu, oracle, monkeypatch, r, n, nq):
    """SST_BK_G=16: one separator per 16-key node (the layout used above 2^29 keys), two leaf sectors per query."""
    gpu.set_option("BK_MIN_N", 0)
    gpu.set_option("BK_R", int(str(r)))
    gpu.set_option("BK_G", 16)
    vals = gen_vals(n, seed=n + r + 1)
    _check(gpu, oracle, vals, gen_queries(nq, seed=n + 2, vals=vals))
    rng = np.random.default_rng(n)
    skew = make_keys(rng, n, "dupes")
    _check(gpu, oracle, skew, make_queries(rng, skew, max(nq // 4, 3)), flags=[(1, 0, 0), (0, 1, 0)])


@pytest.mark.parametrize("seed", range(10))
def test_bucketed_fuzz(gpu, oracle, monkeypatch, seed):
    """Adversarial key distributions (clusters, duplicate runs longer than a bucket, tiny ranges): crowded bucket-table
    cells, empty buckets, all queries in one bucket, jump-table cells with many separators."""
    rng = np.random.default_rng(5000 + seed)
    gpu.set_option("BK_MIN_N", 0)
    gpu.set_option("BK_R", int(str(int(rng.choice([64, 128, 1024])))))
    n = int(rng.integers(1, 900_000))
    vals = make_keys(rng, n, KINDS[seed % len(KINDS)])
    nq = int(rng.integers(1, 200_000))
    qs = make_queries(rng, vals, max(nq, 3))
    if seed % 2:  # every query in a narrow range -> one or two buckets, many chunks
        qs = np.clip(qs.astype(np.int64) % 5000 + int(vals[len(vals) // 2]), 0, MAX).astype(np.uint32)
    _check(gpu, oracle, vals, qs, flags=[(1, 0, 0), (0, 1, 0)])


def test_bucketed_edges(gpu, oracle, monkeypatch):
    sst = gpu
    gpu.set_option("BK_MIN_N", 0)
    gpu.set_option("BK_R", 64)
    vals = np.sort(np.random.default_rng(3).integers(0, 1 << 20, 50_000).astype(np.uint32))  # MAX is not a key
    for nq in (1, 31, 33, 16383, 16384, 16385, 32768 + 7):
        _check(sst, oracle, vals, gen_queries(nq, seed=nq, vals=vals), flags=[(1, 0, 0)])
    for keys in ([MAX], [0, MAX], [5] * 4000 + [MAX], [3] * 1000, list(range(16)), [7] * 16 + [9] * 16 + [MAX] * 3):
        keys = np.array(keys, np.uint32)
        _check(sst, oracle, keys, np.array([0, 1, 3, 4, 5, 6, 7, 8, 9, 10, 15, 16, 17, MAX], np.uint32), flags=[(0, 0, 0), (1, 0, 0)])
    # signed-compare quirk of node.rs:91-108: q >= 2^31 compares as negative -> first key, like every other kernel
    t = sst.STree16.new_params(vals, True, False, False)
    big = np.array([0x80000000, 0xFFFFFFFF, 0x80000001, 5, MAX], np.uint32)
    v, i = t.query(big, sst.SCHEME_BUCKETED, want_index=True)
    v2, i2 = t.query(big, sst.SCHEME_GROUP2, want_index=True)
    assert np.array_equal(v, v2) and np.array_equal(i, i2)
    assert t.query(np.zeros(0, np.uint32), sst.SCHEME_BUCKETED).size == 0


def test_bucketed_unsupported_is_loud(gpu, monkeypatch):
    sst = gpu
    vals = gen_vals(100_000, seed=9)
    qs = gen_queries(100, seed=1, vals=vals)
    t = sst.STree16.new_params(vals, True, False, False)  # default: trees below 2^22 keys have no bucket index
    with pytest.raises(sst.SstError):
        t.query(qs, sst.SCHEME_BUCKETED)
    gpu.set_option("BK_MIN_N", 0)
    gpu.set_option("BK_R", 64)
    big = gen_vals(64 * 8 * 2048 + 9, seed=10)  # one bucket too many for the 2048-bucket partition
    with pytest.raises(sst.SstError):
        sst.STree16.new_params(big, True, False, False).query(qs, sst.SCHEME_BUCKETED)
    with pytest.raises(sst.SstError):
        sst.STree15.new_params(vals, True, False, False).query(qs, sst.SCHEME_BUCKETED)
    gpu.set_option("BK_COMPACT", 0)  # Compact without its dense key copy: the image holds no flat leaf level
    with pytest.raises(sst.SstError):
        sst.PartitionedSTree16C.new(vals, 4).query(qs, sst.SCHEME_BUCKETED)


def test_bucketed_streams_and_repeats(gpu, oracle, monkeypatch):
    """Scratch buffers are per host thread: back-to-back calls on two torch streams must not trample each other."""
    import torch

    sst = gpu
    gpu.set_option("BK_MIN_N", 0)
    gpu.set_option("BK_R", 256)
    vals = gen_vals(1 << 19, seed=5)
    t = sst.STree16.new_params(vals, True, False, False)
    qa = gen_queries(400_000, seed=6, vals=vals)
    qb = gen_queries(123_457, seed=7, vals=vals)
    ea, _ = oracle.lower_bound(vals, qa)
    eb, _ = oracle.lower_bound(vals, qb)
    da, db = torch.from_numpy(qa.view(np.int32)).cuda(), torch.from_numpy(qb.view(np.int32)).cuda()
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    outs = []
    for _ in range(3):
        with torch.cuda.stream(s1):
            va = t.query(da, sst.SCHEME_BUCKETED)
        with torch.cuda.stream(s2):
            vb = t.query(db, sst.SCHEME_BUCKETED)
        outs.append((va, vb))
    torch.cuda.synchronize()
    for va, vb in outs:
        assert np.array_equal(va.cpu().numpy().view(np.uint32), ea)
        assert np.array_equal(vb.cpu().numpy().view(np.uint32), eb)


def test_bucketed_unaligned_device_buffers(gpu, oracle, monkeypatch):
    """The vector move kernels need 16-byte aligned buffers; a sliced tensor takes the checked scalar kernels."""
    import torch

    sst = gpu
    gpu.set_option("BK_MIN_N", 0)
    gpu.set_option("BK_R", 128)
    vals = gen_vals(300_000, seed=15)
    t = sst.STree16.new_params(vals, True, False, False)
    qs = gen_queries(100_000, seed=16, vals=vals)
    ev, ei = oracle.lower_bound(vals, qs)
    d = torch.from_numpy(qs.view(np.int32)).cuda()
    for k in (1, 2, 3, 4):
        v, i = t.query(d[k:], sst.SCHEME_BUCKETED, want_index=True)
        assert np.array_equal(v.cpu().numpy().view(np.uint32), ev[k:]) and np.array_equal(i.cpu().numpy().astype(np.uint64), ei[k:]), k


def test_auto_picks_the_pipeline_for_large_batches(gpu, oracle):
    """SCHEME_AUTO: reordered-batch pipeline for large batches over >= 2^25 keys (from 2^24 queries on, the crossover of
    profiles/r2_s3_size_batch_sweep.jsonl), the rank-table kernel below; sst_query_plan reports the choice and the launch count."""
    import ctypes as C

    import torch

    sst = gpu
    n = (1 << 25) + 77
    g = torch.Generator(device="cuda").manual_seed(7)
    keys = torch.randint(0, MAX, (n,), dtype=torch.int32, device="cuda", generator=g)
    keys[0] = MAX
    keys = torch.sort(keys).values.contiguous()
    t = sst.STree16.new_params(keys, True, False, False)
    sch, launches = C.c_int(0), C.c_int(0)
    L = sst.lib()
    assert L.sst_query_plan(t._h, 1 << 26, 0, 0, C.byref(sch), C.byref(launches)) == 0
    assert sch.value == sst.SCHEME_BUCKETED and launches.value == 4  # partition, work items, search, un-permute
    assert L.sst_query_plan(t._h, 1 << 26, 0, 1, C.byref(sch), C.byref(launches)) == 0 and launches.value == 5
    assert L.sst_query_plan(t._h, (1 << 28) + 1, 0, 0, C.byref(sch), C.byref(launches)) == 0 and launches.value == 4  # one run of up to 2^30 queries
    sst.set_option("BK_SUB_LOG2", 27)
    assert L.sst_query_plan(t._h, (1 << 28) + 1, 0, 0, C.byref(sch), C.byref(launches)) == 0 and launches.value == 12  # three runs of 2^27
    sst.set_option("BK_SUB_LOG2", 30)
    assert L.sst_query_plan(t._h, 1 << 24, 0, 0, C.byref(sch), C.byref(launches)) == 0 and sch.value == sst.SCHEME_BUCKETED
    assert L.sst_query_plan(t._h, (1 << 24) - 1, 0, 0, C.byref(sch), C.byref(launches)) == 0 and sch.value == sst.SCHEME_TABLE
    assert L.sst_query_plan(t._h, 1 << 20, 0, 0, C.byref(sch), C.byref(launches)) == 0
    assert sch.value == sst.SCHEME_TABLE and launches.value == 1
    qs = torch.randint(0, MAX, ((1 << 26) + 5,), dtype=torch.int32, device="cuda", generator=g)
    v, i = t.query(qs, sst.SCHEME_AUTO, want_index=True)
    v2, i2 = t.query(qs, sst.SCHEME_TABLE, want_index=True)
    assert bool((v == v2).all()) and bool((i == i2).all())
    hk, hq = keys.cpu().numpy().view(np.uint32), qs[:200_000].cpu().numpy().view(np.uint32)
    ev, ei = oracle.lower_bound(hk, hq)
    assert np.array_equal(v[:200_000].cpu().numpy().view(np.uint32), ev) and np.array_equal(i[:200_000].cpu().numpy().astype(np.uint64), ei)


def test_bucketed_degenerate_batches(gpu, oracle, monkeypatch):
    """Every query in one bucket / one key (all lanes of a warp collide in the ranking step), sorted and reverse-sorted batches."""
    sst = gpu
    gpu.set_option("BK_MIN_N", 0)
    gpu.set_option("BK_R", 256)
    vals = gen_vals(700_000, seed=33)
    for qs in (np.full(300_001, vals[123_456], np.uint32), np.full(50_000, 0, np.uint32), np.full(50_000, MAX, np.uint32),
               np.sort(gen_queries(200_000, seed=34, vals=vals)), np.sort(gen_queries(200_000, seed=35, vals=vals))[::-1].copy(),
               (vals[100_000] + (np.arange(150_000) % 7)).astype(np.uint32)):
        _check(sst, oracle, vals, qs, flags=[(1, 0, 0)])


@pytest.mark.parametrize("r,n,nq,b", [(64, 5000, 40_000, 4), (64, 1_000_000, 70_000, 8), (256, (1 << 20) + 5, 300_000, 16), (1024, 2_000_003, 150_000, 20),
                                      (64, 17, 33, 0), (16384, (1 << 22) + 999, 600_000, 20)])
@pytest.mark.parametrize("layout", ["PartitionedSTree16M", "PartitionedSTree16", "PartitionedSTree16L", "PartitionedSTree16O", "PartitionedSTree16C"])
def test_bucketed_map_partitioned(gpu, oracle, monkeypatch, r, n, nq, b, layout):
    """The pipeline over a Map-partitioned tree (its leaf level is the sorted array, partitioned_s_tree.rs:503) and over the
    flat leaf level of Simple / L1 / Overlapping (gaps hold the next part's first key, :502-515; positions converted to
    sorted-array indices), and for Compact (whose image interleaves the parts' levels) over a dense GPU-only copy of the keys:
    same values and indices as the oracle and as the layout's own lane-group kernel, including queries above MAX (no part:
    (MAX, n))."""
    sst = gpu
    gpu.set_option("BK_MIN_N", 0)
    gpu.set_option("BK_R", int(str(r)))
    vals = gen_vals(n, seed=n + r + 7)
    t = getattr(sst, layout).try_new(vals, b)
    if t is None:
        pytest.skip("layout exceeds the reference's memory cap for these keys")
    try:
        t.query(gen_queries(8, seed=1, vals=vals), sst.SCHEME_BUCKETED)
    except sst.SstError as e:  # the padded flat leaf level needs more than 2048 buckets at this separator window
        assert "reordered-batch" in str(e) or "group/table" in str(e), str(e)
        pytest.skip("flat leaf level too large for the bucket partition at this SST_BK_R")
    qs = gen_queries(nq, seed=n + 3, vals=vals)
    ev, ei = oracle.lower_bound(vals, qs)
    v, i = t.query(qs, sst.SCHEME_BUCKETED, want_index=True)
    assert np.array_equal(v, ev) and np.array_equal(i, ei)
    assert np.array_equal(t.query(qs, sst.SCHEME_BUCKETED), ev)
    # queries above MAX mixed in: must equal the lane-group kernel (AUTO at this batch size) and the generic kernel
    qb = qs.copy()
    qb[:: max(1, nq // 97)] = 0x80000000 + (qb[:: max(1, nq // 97)] >> 1)
    qb[-1] = 0xFFFFFFFF
    v1, i1 = t.query(qb, sst.SCHEME_BUCKETED, want_index=True)
    v2, i2 = t.query(qb, want_index=True)
    v3, i3 = t.query(qb, sst.SCHEME_GENERIC, want_index=True)
    assert np.array_equal(v1, v2) and np.array_equal(i1, i2) and np.array_equal(v1, v3) and np.array_equal(i1, i3)
    assert (v1[qb > MAX] == MAX).all() and (i1[qb > MAX] == n).all()
    assert np.array_equal(t.query(qb, sst.SCHEME_BUCKETED), v1)


def test_bucketed_map_skewed_keys(gpu, oracle, monkeypatch):
    """Map tree over keys that do not reach 31 bits / with long duplicate runs, MAX not a key."""
    sst = gpu
    gpu.set_option("BK_MIN_N", 0)
    gpu.set_option("BK_R", 256)
    rng = np.random.default_rng(77)
    served = 0
    for kind, layout in (("dupes", "PartitionedSTree16M"), ("clustered", "PartitionedSTree16M"), ("dupes", "PartitionedSTree16"),
                         ("clustered", "PartitionedSTree16L"), ("tiny_range", "PartitionedSTree16O"), ("boundary16", "PartitionedSTree16"),
                         ("dupes", "PartitionedSTree16C"), ("clustered", "PartitionedSTree16C")):
        vals = make_keys(rng, 300_000, kind)
        qs = make_queries(rng, vals, 100_000)
        t = getattr(sst, layout).try_new(vals, 12)
        if t is None:
            continue
        try:
            t.query(qs[:8], sst.SCHEME_BUCKETED)
        except sst.SstError as e:  # padded flat leaf level too large for 2048 buckets of 64 separators
            assert "reordered-batch" in str(e) or "group/table" in str(e), str(e)
            continue
        served += 1
        ev, ei = oracle.lower_bound(vals, qs)
        v, i = t.query(qs, sst.SCHEME_BUCKETED, want_index=True)
        v2, i2 = t.query(qs, want_index=True)
        assert np.array_equal(v, v2) and np.array_equal(i, i2)
        assert np.array_equal(v, ev) and np.array_equal(i, ei)
    assert served >= 3


def test_calibrate_sets_the_auto_crossover(gpu, oracle):
    """sst_query_calibrate times the direct kernel against the pipeline on the index's own device and SCHEME_AUTO follows the
    measured crossover: below it the direct kernel, from it on the pipeline; results do not change."""
    import ctypes as C

    sst = gpu
    sst.set_option("BK_MIN_N", 0)
    sst.set_option("BK_AUTO_MIN_N", 0)
    sst.set_option("BK_R", 1024)
    vals = gen_vals(1 << 21, seed=51)
    t = sst.STree16.new_params(vals, True, False, False)
    L = sst.lib()

    def auto_scheme(nq):
        sch, ln = C.c_int(0), C.c_int(0)
        L.sst_query_plan(t._h, nq, sst.SCHEME_AUTO, 0, C.byref(sch), C.byref(ln))
        return sch.value

    assert auto_scheme(1 << 22) != sst.SCHEME_BUCKETED and auto_scheme(1 << 24) == sst.SCHEME_BUCKETED  # the default rule
    cross = t.calibrate(1 << 22)
    assert cross == 2**64 - 1 or (1 << 20) <= cross <= (1 << 22)
    if cross != 2**64 - 1:
        assert auto_scheme(cross) == sst.SCHEME_BUCKETED and (cross == 1 << 20 or auto_scheme(cross // 2) != sst.SCHEME_BUCKETED)
    else:
        assert auto_scheme(1 << 24) != sst.SCHEME_BUCKETED
    qs = gen_queries(300_000, seed=52, vals=vals)
    ev, ei = oracle.lower_bound(vals, qs)
    v, i = t.query(qs, want_index=True)
    assert np.array_equal(v, ev) and np.array_equal(i, ei)
    small = sst.STree16.new_params(gen_vals(1000, seed=53), True, False, False)
    sst.reset_options()
    assert sst.STree16.new_params(gen_vals(1000, seed=53), True, False, False).calibrate() == 0  # nothing to calibrate


@pytest.mark.parametrize("g", [8, 16])
@pytest.mark.parametrize("kind,n,nq", [("uniform", 3_000_000, 400_000), ("uniform", 600_001, 100_000), ("clustered", 1_500_000, 200_000),
                                       ("dupes", 2_000_000, 200_000), ("tiny_range", 700_000, 100_000), ("uniform", 1000, 5000)])
def test_bucketed_sep16(gpu, oracle, kind, n, nq, g):
    """16-bit separator mode (what leaf levels above 2^28 slots take): 65536 separators per bucket stored as offsets inside their
    jump cell, both ends of the cell read, up to six offsets per query from three 32-bit loads, more by bisection.  Forced onto
    small trees: uniform keys (two separators per cell), clustered / duplicate keys (crowded cells: the bisection path, empty
    cells), plain and Map / Compact layouts, with the index output.  g = 16 keys per separator (two leaf sectors per query) is
    what a leaf level above 2^30 slots takes."""
    sst = gpu
    gpu.set_option("BK_MIN_N", 0)
    gpu.set_option("BK_SEP16", 1)
    if g == 16:
        gpu.set_option("BK_G", 16)
    rng = np.random.default_rng(n + nq)
    vals = gen_vals(n, seed=n) if kind == "uniform" else make_keys(rng, n, kind)
    qs = make_queries(rng, vals, nq)
    _check(sst, oracle, vals, qs, flags=[(1, 0, 0), (0, 1, 0)])
    ev, ei = oracle.lower_bound(vals, qs)
    for layout in ("PartitionedSTree16M", "PartitionedSTree16C"):
        t = getattr(sst, layout).try_new(vals, 8)
        if t is None:
            continue
        v, i = t.query(qs, sst.SCHEME_BUCKETED, want_index=True)
        assert np.array_equal(v, ev) and np.array_equal(i, ei), layout


def test_bucketed_several_runs(gpu, oracle):
    """A batch larger than the pipeline's run size (BK_SUB_LOG2; 2^30 by default, halved when the scratch does not fit) is
    answered run by run: full runs, a partial last run, the index output."""
    sst = gpu
    gpu.set_option("BK_MIN_N", 0)
    gpu.set_option("BK_R", 256)
    gpu.set_option("BK_SUB_LOG2", 20)
    vals = gen_vals(700_000, seed=61)
    qs = gen_queries((3 << 20) + 12_345, seed=62, vals=vals)
    _check(sst, oracle, vals, qs, flags=[(1, 0, 0)])


@pytest.mark.gpu
def test_bucketed_two_to_the_31_keys():
    """The largest tree a u32 index can hold (2^31 slots: 2048 buckets of 65536 16-bit separators, 16 keys each) takes the pipeline:
    a sample of values and indices equals torch.searchsorted, the whole batch equals the direct kernel (tools/big_tree.py)."""
    import json
    import subprocess
    env = dict(os.environ, N=str(1 << 31), NQ=str(1 << 24), TAG="test")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "big_tree.py")], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    r = json.loads(out.stdout.strip().splitlines()[-1])
    assert r["auto_scheme"] == 7 and r["values_ok"] and r["index_ok"] and r["pipeline_equals_direct"], r
