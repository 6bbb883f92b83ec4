"""CPU tests (-m "not gpu"): the oracle against the reference's own known answers and its
differential property (static-search-tree/src/test.rs:142-260, s_tree.rs:846-895)."""
import numpy as np
import pytest

from util import MAX, gen_queries, gen_vals, random_patterns, random_text, reference_test_sizes


def test_kat_bottom_layer_and_top_node(oracle):
    # s_tree.rs:861-885: vals = 1..2000 + MAX; search(452) == 452 and search(289) == 289
    vals = np.concatenate([np.arange(1, 2000, dtype=np.uint32), [MAX]]).astype(np.uint32)
    t = oracle.Tree.stree(vals)
    assert list(t.search([452, 289])) == [452, 289]
    ev, ei = oracle.lower_bound(vals, [452, 289])
    assert list(ev) == [452, 289] and list(ei) == [451, 288]


def test_kat_simd_cmp(oracle):
    # s_tree.rs:887-895: vals = 1..16 + MAX; tree[0].find(1) == 0
    vals = np.concatenate([np.arange(1, 16, dtype=np.uint32), [MAX]]).astype(np.uint32)
    t = oracle.Tree.stree(vals)
    assert oracle.node_find(t.image()[:16], 1) == 0
    # find == number of keys < q, signed (node.rs:93-109)
    node = np.array([1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 25, 27, 29, MAX], np.uint32)
    for q, want in [(0, 0), (1, 0), (2, 1), (29, 14), (30, 15), (MAX, 15)]:
        assert oracle.node_find(node, q) == want
    assert oracle.node_find(node, 0x80000000) == 0  # signed compare: q >= 2^31 is "negative"


def test_shape_table(oracle):
    # SURVEY 8(a) a3, computed from TreeBase (s_tree.rs:22-45)
    want = {
        1 << 10: [1, 4, 64],
        1 << 20: [1, 14, 227, 3856, 65536],
        1 << 28: [1, 12, 201, 3415, 58053, 986896, 16777216],
        1 << 30: [1, 3, 48, 804, 13660, 232211, 3947581, 67108864],
    }
    for n, nodes in want.items():
        H = oracle.height(n)
        assert [-(-oracle.layer_size(n, h, H) // 16) for h in range(H)] == nodes
    # 6 % space overhead (readme.org:12-13)
    assert abs(sum(want[1 << 20]) * 64 / (4 * (1 << 20)) - 1 - 0.063) < 0.002


@pytest.mark.parametrize("n", reference_test_sizes(6, 18) + [(1 << 22) * 5 // 16])
def test_differential_all_layouts(oracle, n):
    """test.rs:142-260: every index/scheme equals SortedVec::binary_search."""
    vals = gen_vals(n, seed=n)
    qs = gen_queries(1024, seed=n + 1, vals=vals)
    ev, ei = oracle.lower_bound(vals, qs)
    for B in (16, 15):
        for lm, rev, full in [(0, 0, 0), (1, 0, 0), (1, 0, 1), (0, 1, 0), (1, 1, 0), (0, 0, 1)]:
            t = oracle.Tree.stree(vals, B=B, left_max=lm, reverse=rev, full=full)
            v, s = t.search(qs, want_slot=True)
            assert (v == ev).all(), (n, B, lm, rev, full)
            assert (s == ei).all(), (n, B, lm, rev, full)
    for var in ("simple", "compact", "l1", "overlap", "map"):
        for b in (0, 4, 8, 16, 20):
            t = oracle.Tree.pstree(vals, b, var)
            assert t is not None
            assert (t.search(qs) == ev).all(), (n, var, b)


def test_batch_final_cpu_baseline(oracle):
    vals = gen_vals(100_000, seed=5)
    qs = gen_queries(128 * 50 + 17, seed=6, vals=vals)
    ev, _ = oracle.lower_bound(vals, qs)
    t = oracle.Tree.stree(vals, left_max=True)
    for threads in (1, 3):
        v, secs = t.batch_final(qs, threads)
        assert (v == ev).all() and secs >= 0


@pytest.mark.parametrize("n", [1, 15, 16, 17, 300, 5000, 100_000, 1_200_000])
def test_batch_interleave_cpu_baseline(oracle, n):
    """batch_interleave_all_128 (s_tree.rs:684-832), the reference's fastest CPU scheme, restated as a rotating pipeline:
    every tree height of the dispatch table that fits a test, batch sizes around P * L, several threads."""
    vals = gen_vals(n, seed=n)
    t = oracle.Tree.stree(vals, left_max=True)
    for nq in (0, 1, 127, 128 * 7 + 5, 20_011):
        qs = gen_queries(nq, seed=nq + 1, vals=vals)
        ev, _ = oracle.lower_bound(vals, qs)
        for threads in (1, 3):
            v, secs = t.batch_interleave(qs, threads)
            assert (v == ev).all() and secs >= 0, (n, nq, threads)


def test_duplicates_and_tiny(oracle):
    for vals in ([MAX], [0, MAX], [5] * 40 + [MAX], list(range(16)), list(range(17)), [7] * 16 + [9] * 16 + [MAX] * 3):
        vals = np.array(vals, np.uint32)
        qs = np.array([0, 1, 5, 6, 7, 8, 9, 10, 15, 16, 17, MAX], np.uint32)
        ev, ei = oracle.lower_bound(vals, qs)
        for lm in (0, 1):
            v, s = oracle.Tree.stree(vals, left_max=lm).search(qs, want_slot=True)
            assert (v == ev).all() and (s == ei).all()


def test_partition_params_match_survey(oracle):
    # SURVEY 8(a) a8 (uniform keys): n = 2^20, b = 20
    vals = gen_vals(1 << 20, seed=11)
    for var in ("simple", "compact", "l1", "overlap", "map"):
        t = oracle.Tree.pstree(vals, 20, var)
        p = t.params
        assert p["parts"] == 1 << (31 - p["shift"])
        assert t.layers == t.levels + (1 if var == "map" else 0)


# ---- suffix arrays (parity unpinned by the reference; property tests) ------------------------
def test_sa_oracle_properties(oracle):
    text = random_text(20_000, seed=21)
    sa = oracle.sa_build(text)
    assert sorted(sa.tolist()) == list(range(text.size))
    assert oracle.sa_check(text, sa) == 0  # sa_search.rs:36-38
    bad = sa.copy()
    bad[[10, 11]] = bad[[11, 10]]
    assert oracle.sa_check(text, bad) > 0
    pats = random_patterns(text, 500, seed=22) + [b"", bytes([3] * 50), bytes([0]), text[-5:].tobytes(), text[:40].tobytes()]
    flat, off = oracle.pack_patterns(pats)
    lo, hi, pos, cnt = oracle.sa_search(text, sa, flat, off)
    tb = text.tobytes()
    for i, p in enumerate(pats):
        # brute force: lo = #suffixes < p ; hi - lo = #occurrences
        occ = sum(1 for s in range(len(tb) - len(p) + 1) if tb.startswith(p, s)) if len(p) else len(tb)
        assert hi[i] - lo[i] == occ, i
        if lo[i] < text.size:
            assert pos[i] == sa[lo[i]]
            assert tb[sa[lo[i]] :] >= p
        if lo[i] > 0:
            assert tb[sa[lo[i] - 1] :] < p
    # LCP-accelerated search returns the same l
    mlo, _ = oracle.sa_search_mlr(text, sa, flat, off)
    assert (mlo == lo).all()
    # 16-byte cmp variant (sa_search.rs:346-374) agrees on substring queries
    padded = np.concatenate([text, np.zeros(200, np.uint8)])
    sub = random_patterns(text, 300, seed=23)
    f2, o2 = oracle.pack_patterns(sub)
    l2, _, p2, _ = oracle.sa_search(text, sa, f2, o2)
    cl, cp = oracle.sa_search_cmp(padded, text.size, sa, f2, o2)
    assert (cl == l2).all() and (cp == p2).all()
    bl, bp, _ = oracle.sa_search_batch32(text, sa, f2, o2, threads=2)
    assert (bl == l2).all() and (bp == p2).all()
