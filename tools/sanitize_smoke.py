#!/usr/bin/env python
"""Small all-kernel exercise without torch, for `compute-sanitizer --tool memcheck python tools/sanitize_smoke.py`."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "suffix-array-searching_b200"))
import sst_b200 as sst
from oracle import oracle as O

rng = np.random.default_rng(1)
MAX = sst.MAX
for n in (1, 17, 4625, 70_001, 300_000):
    vals = np.sort(rng.integers(0, MAX, n, dtype=np.uint32))
    vals[-1] = MAX
    qs = rng.integers(0, MAX, 5000 + 13, dtype=np.uint32)
    qs[:3] = [0, MAX, vals[0]]
    ev, ei = O.lower_bound(vals, qs)
    for flags in ((0, 0, 0), (1, 0, 0), (0, 1, 0), (1, 0, 1)):
        t = sst.STree16.new_params(vals, *map(bool, flags))
        for scheme in (1, 2, 3, 4, 5):
            v, i = t.query(qs, scheme, want_index=True)
            assert (v == ev).all() and (i == ei).all(), (n, flags, scheme)
    t15 = sst.STree15.new(vals)
    assert (t15.query(qs) == ev).all()
    for cls in (sst.PartitionedSTree16, sst.PartitionedSTree16C, sst.PartitionedSTree16L, sst.PartitionedSTree16O, sst.PartitionedSTree16M):
        for b in (0, 6, 20):
            t = cls.try_new(vals, b)
            if t is not None:
                v, i = t.query(qs, want_index=True)
                assert (v == ev).all() and (i == ei).all(), (n, cls.__name__, b)
# one big-batch run of the table kernel (staging via TMA)
vals = np.sort(rng.integers(0, MAX, 2_000_000, dtype=np.uint32)); vals[-1] = MAX
qs = rng.integers(0, MAX, 1_200_000, dtype=np.uint32)
ev, ei = O.lower_bound(vals, qs)
t = sst.STree16.new_params(vals, True, False, False)
v, i = t.query(qs, sst.SCHEME_TABLE, want_index=True)
assert (v == ev).all() and (i == ei).all()
# the reordered-batch pipeline (partition / plan / search / un-permute), forced onto a small tree; full and partial tiles,
# skewed batches (one bucket, one key) and the index output
sst.set_option("BK_MIN_N", 0)
sst.set_option("BK_R", 256)
vals = np.sort(rng.integers(0, MAX, 600_000, dtype=np.uint32)); vals[-1] = MAX
tb = sst.STree16.new_params(vals, True, False, False)
for qs in (rng.integers(0, MAX, 3 * 16384 + 77, dtype=np.uint32), np.full(40_000, vals[1234], np.uint32),
           (vals[5000] + rng.integers(0, 3000, 50_000)).astype(np.uint32)):
    ev, ei = O.lower_bound(vals, qs)
    v, i = tb.query(qs, sst.SCHEME_BUCKETED, want_index=True)
    assert (v == ev).all() and (i == ei).all()
    assert (tb.query(qs, sst.SCHEME_BUCKETED) == ev).all()
# the same over the partitioned layouts (Compact reads a dense copy of the keys), queries above MAX mixed in
for cls in (sst.PartitionedSTree16M, sst.PartitionedSTree16, sst.PartitionedSTree16C):
    tp = cls.new(vals, 8)
    qs = rng.integers(0, MAX, 2 * 16384 + 5, dtype=np.uint32)
    qs[::97] |= 0x80000000
    v1, i1 = tp.query(qs, sst.SCHEME_BUCKETED, want_index=True)
    v2, i2 = tp.query(qs, sst.SCHEME_GENERIC, want_index=True)
    assert (v1 == v2).all() and (i1 == i2).all(), cls.__name__
sst.reset_options()
# replicas on "several" devices (device 0 listed twice) and the probe counter
ms = sst.MultiSa.build(rng.integers(0, 4, 20_000, dtype=np.uint8), [0, 0])
flat, off = sst.pack_patterns([bytes(rng.integers(0, 4, 12, dtype=np.uint8)) for _ in range(100)])
assert ms.search(flat, off)[0].size == 100
# suffix arrays
for n, sigma in ((1, 4), (50, 2), (30_000, 4), (20_000, 256)):
    text = rng.integers(0, sigma, n, dtype=np.uint8)
    sa = sst.SaNaive.build(text)
    assert sa.check() == 0
    pats = [text[s : s + l].tobytes() for s, l in zip(rng.integers(0, max(1, n - 1), 300), rng.integers(0, 70, 300))] + [b"", bytes([255] * 40)]
    flat, off = sst.pack_patterns(pats)
    of, oo = O.pack_patterns(pats)
    elo, ehi, epos, _ = O.sa_search(text, sa.sa, of, oo)
    for lanes in ("1", "8", "32"):
        sst.set_option("SA_LANES", int(lanes))
        for mode in (0, 1):
            lo, hi, pos = sa.search(flat, off, mode)
            assert (lo == elo).all() and (hi == ehi).all() and (pos == epos).all(), (n, sigma, lanes, mode)
print("sanitize smoke OK")
