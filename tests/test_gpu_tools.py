"""GPU smoke test of tools/reference_results.py: GPU rows in the reference's `Result` schema
(static-search-tree/src/bin/bench.rs:519-545) with its CLI switches (bench.rs:26-46)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "reference_results.py")
KEYS = {"params", "scheme", "size", "index_size", "queries", "threads", "run", "duration", "latency", "layers", "cycles", "freq"}  # bench.rs:519-533


def _run(tmp_path, *args):
    out = tmp_path / "gpu-results.json"
    p = subprocess.run([sys.executable, TOOL, "--out", str(out), *args], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    return json.load(open(out))


def test_result_schema_and_range(gpu, tmp_path):
    rows = _run(tmp_path, "--from", "14", "--to", "16", "--queries", "10000", "--runs", "2", "--range")
    assert rows and all(set(r) == KEYS for r in rows)
    assert {r["size"] for r in rows} == {1 << 14, 1 << 15, 1 << 16} and {r["run"] for r in rows} == {0, 1}
    nq = -(-10000 // 768) * 768  # next_multiple_of(256 * 3), bench.rs:78
    single = [r for r in rows if r["scheme"] == "gpu::single"]
    rng = [r for r in rows if r["scheme"] == "gpu::range"]
    assert single and rng and all(r["queries"] == nq for r in single) and all(r["queries"] == 2 * nq for r in rng)  # bench.rs:84: [q, q+1] pairs
    for r in rows:
        assert set(r["duration"]) == {"secs", "nanos"} and r["latency"] > 0 and r["layers"] >= 1 and r["index_size"] >= r["size"]
        assert abs(r["cycles"] - r["latency"] * 1e-9 * r["freq"]) < 1e-9


def test_dense_positive_and_human(gpu, tmp_path):
    rows = _run(tmp_path, "--from", "14", "--to", "15", "--queries", "5000", "--dense", "--positive")
    assert {r["size"] for r in rows} == {1 << 14, (1 << 14) * 5 // 4, (1 << 14) * 6 // 4, (1 << 14) * 7 // 4, 1 << 15}  # bench.rs:455-472
    assert {r["params"] for r in rows} == {"STree16 left_max", "PartitionedSTree16M b=20"}
    g = np.random.default_rng(3)
    fa = tmp_path / "genome.fa"
    seq = bytes(g.choice(list(b"ACGT"), 20_000).tolist())
    fa.write_bytes(b">chr1\n" + b"\n".join(seq[i:i + 70] for i in range(0, len(seq), 70)) + b"\n")
    rows = _run(tmp_path, "--from", "14", "--to", "14", "--queries", "2000", "--human", str(fa))
    assert rows and all(set(r) == KEYS and r["size"] == 1 << 14 for r in rows)
