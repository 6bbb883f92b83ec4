"""bench.py's exact host-side parity samples (check_lower_bound_sample / check_sa_sample) must accept the oracle's answers
and flag every kind of wrong answer: they gate the bench's exit code for configs C3 / C4 / C5."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from util import MAX, gen_queries, gen_vals, random_patterns, random_text  # noqa: E402


def _lb_inputs(oracle, n=20_000, nq=3000):
    vals = gen_vals(n, seed=1)
    qs = gen_queries(nq, seed=2, vals=vals)
    v, i = oracle.lower_bound(vals, qs)
    i = i.astype(np.int64)
    return vals, qs, v, i, vals[np.maximum(i - 1, 0)], vals[np.minimum(i, n - 1)]


def test_lower_bound_checker_accepts_the_oracle(oracle):
    vals, qs, v, i, kp, ka = _lb_inputs(oracle)
    assert bench.check_lower_bound_sample(qs, v, i, kp, ka, vals.size) == 0


def test_lower_bound_checker_flags_wrong_answers(oracle):
    vals, qs, v, i, kp, ka = _lb_inputs(oracle)
    n = vals.size
    # an index one too high / too low is still "a key >= q" or "value matches" -- only the lower-bound property catches it
    j = int(np.argmax((i > 0) & (i < n - 1) & (vals[np.minimum(i + 1, n - 1)] > vals[np.minimum(i, n - 1)]) & (vals[np.minimum(i, n - 1)] >= qs)))
    up = i.copy(); up[j] += 1
    assert bench.check_lower_bound_sample(qs, vals[np.minimum(up, n - 1)], up, vals[np.maximum(up - 1, 0)], vals[np.minimum(up, n - 1)], n) >= 1
    k = int(np.argmax((i > 1) & (vals[np.maximum(i - 1, 0)] < qs)))
    dn = i.copy(); dn[k] -= 1
    assert bench.check_lower_bound_sample(qs, vals[np.minimum(dn, n - 1)], dn, vals[np.maximum(dn - 1, 0)], vals[np.minimum(dn, n - 1)], n) >= 1
    wrong_val = v.copy(); wrong_val[5] ^= 1
    assert bench.check_lower_bound_sample(qs, wrong_val, i, kp, ka, n) == 1
    # a query above MAX behaves like 0 (signed compare, node.rs:91-108)
    assert bench.check_lower_bound_sample([MAX + 5], [vals[0]], [0], [vals[0]], [vals[0]], n) == 0


def _sa_inputs(oracle):
    text = random_text(30_000, seed=3)
    sa = oracle.sa_build(text)
    pats = [bytes(p) for p in random_patterns(text, 400, seed=4, lo=1, hi=40)] + [bytes([3] * 30), bytes([0]), text[-5:].tobytes()]
    flat, off = oracle.pack_patterns(pats)
    lo, hi, pos, _ = oracle.sa_search(text, sa, flat, off)
    return text, sa, pats, lo.astype(np.int64), hi.astype(np.int64), pos


def _windows(text, sa, pats, lo, hi):
    n = text.size
    want = np.stack([lo - 1, lo, hi - 1, hi])
    valid = (want >= 0) & (want < n)
    sa_at = np.where(valid, sa[np.where(valid, want, 0)], 0).astype(np.int64)
    win = [[text[sa_at[j][i]: sa_at[j][i] + len(q)].tobytes() if valid[j][i] else b"" for i, q in enumerate(pats)] for j in range(4)]
    return sa_at, win


def test_sa_checker_accepts_the_oracle(oracle):
    text, sa, pats, lo, hi, pos = _sa_inputs(oracle)
    sa_at, win = _windows(text, sa, pats, lo, hi)
    assert bench.check_sa_sample(pats, lo, hi, pos, text.size, sa_at, win) == 0


def test_sa_checker_flags_wrong_bounds(oracle):
    text, sa, pats, lo, hi, pos = _sa_inputs(oracle)
    n = text.size
    multi = int(np.argmax(hi - lo >= 2))  # a pattern with at least two occurrences
    for name, dlo, dhi in (("lo too high", 1, 0), ("lo too low", -1, 0), ("hi too low", 0, -1), ("hi too high", 0, 1)):
        l2, h2 = lo.copy(), hi.copy()
        l2[multi] += dlo; h2[multi] += dhi
        p2 = pos.copy(); p2[multi] = sa[l2[multi]]
        sa_at, win = _windows(text, sa, pats, l2, h2)
        assert bench.check_sa_sample(pats, l2, h2, p2, n, sa_at, win) == 1, name
    p3 = pos.copy(); p3[multi] = (int(p3[multi]) + 1) % n
    sa_at, win = _windows(text, sa, pats, lo, hi)
    assert bench.check_sa_sample(pats, lo, hi, p3, n, sa_at, win) == 1
