"""Shared generators for the tests (fixed seeds; distributions of static-search-tree/src/util.rs:13-42
and suffix-array-searching/src/util.rs:9-26)."""
import numpy as np

MAX = 0x7FFFFFFF


def gen_vals(n, seed=0):
    """util.rs:31-42: n uniform keys in [0, MAX), vals[0] = MAX, sorted."""
    rng = np.random.default_rng(seed)
    v = rng.integers(0, MAX, n, dtype=np.uint32)
    v[0] = MAX
    v.sort()
    return v


def gen_queries(nq, seed=1, vals=None):
    """util.rs:16-21 uniform in [0, MAX) plus edge cases."""
    rng = np.random.default_rng(seed)
    q = rng.integers(0, MAX, nq, dtype=np.uint32)
    edge = [0, 1, MAX, MAX - 1]
    if vals is not None and len(vals):
        n = len(vals)
        edge += [int(vals[0]), int(vals[-1]), int(vals[n // 2]), min(int(vals[n // 2]) + 1, MAX), max(int(vals[n // 3]), 1) - 1]
        # hit node boundaries: last key of a leaf, first key of the next
        for j in (15, 16, 17 * 16 - 1, 17 * 16, 271, 272):
            if j < n:
                edge.append(int(vals[j]))
    edge = np.array(edge[:nq], dtype=np.uint32)
    q[: edge.size] = edge
    return q


def reference_test_sizes(lo_pow2=6, hi_pow2=20):
    """sizes in BYTES of static-search-tree/src/test.rs:146-153, converted to key counts."""
    out = []
    for p in range(lo_pow2, hi_pow2 + 1):
        x = 1 << p
        for s in (x, x * 5 // 4, x * 6 // 4, x * 7 // 4):
            out.append(s // 4)
    return out


def random_text(n, seed=2, sigma=4):
    """suffix-array-searching/src/util.rs:9-15: bytes uniform in 0..4."""
    return np.random.default_rng(seed).integers(0, sigma, n, dtype=np.uint8)


def random_patterns(text, npat, seed=3, lo=30, hi=100):
    """util.rs:18-26: substrings t[i..i+len], i in [0, n-200), len in [30,100)."""
    rng = np.random.default_rng(seed)
    n = len(text)
    starts = rng.integers(0, max(1, n - 200), npat)
    lens = rng.integers(lo, hi, npat)
    return [text[s : s + l].tobytes() for s, l in zip(starts, lens)]
