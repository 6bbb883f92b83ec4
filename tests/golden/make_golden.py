#!/usr/bin/env python
"""Generates tests/golden/*.npz with the CPU oracle (seeded inputs + expected outputs).

The reference's tests use an unseeded RNG and hold no golden vectors, and the Rust
reference cannot run here, so these vectors come from the oracle restatement after it passed the
reference's known-answer tests and differential property (tests/test_oracle.py).
Run:  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import oracle as O  # noqa: E402
from util import gen_queries, gen_vals, random_patterns, random_text  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    vals = gen_vals(5000, seed=101)
    qs = gen_queries(512, seed=102, vals=vals)
    ev, ei = O.lower_bound(vals, qs)
    images = {}
    for name, kw in (("plain", {}), ("left_max", {"left_max": True}), ("reverse", {"reverse": True}), ("full", {"left_max": True, "full": True})):
        t = O.Tree.stree(vals, **kw)
        images["stree_" + name] = sha(t.image())
    for var in ("simple", "compact", "l1", "overlap", "map"):
        t = O.Tree.pstree(vals, 8, var)
        images["pstree_" + var] = sha(t.image())
    np.savez_compressed(os.path.join(HERE, "lower_bound.npz"), vals=vals, qs=qs, values=ev, indices=ei,
                        image_names=np.array(list(images.keys())), image_sha256=np.array(list(images.values())))
    text = random_text(3000, seed=103)
    sa = O.sa_build(text)
    pats = random_patterns(text, 200, seed=104, lo=1, hi=60) + [b"", bytes([3] * 20), text[-9:].tobytes()]
    flat, off = O.pack_patterns(pats)
    lo, hi, pos, _ = O.sa_search(text, sa, flat, off)
    np.savez_compressed(os.path.join(HERE, "sa_search.npz"), text=text, sa=sa, flat=flat[: int(off[-1])], off=off, lo=lo, hi=hi, pos=pos)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
