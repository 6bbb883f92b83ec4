"""GPU test: bench.py's JSON contract on a small configuration (keys the driver reads)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_bench_line_contract(gpu):
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--n-keys", str(1 << 22), "--queries", str(1 << 22), "--steps", "3", "--warmup", "3",
         "--cpu-sample", str(1 << 20), "--sa-text", "400000", "--sa-patterns", "20000", "--e2e-steps", "2",
         "--c2-sizes", "10,16", "--sa-rep-text", "2000000", "--sa-rep-patterns", "3000", "--c4-log2-keys", "23", "--c4-queries", "5000000", "--c5-text", "3000000", "--c5-patterns", "200000", "--c-reps", "2", "--parity-sample", "2000"],
        capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "exactly one JSON line on stdout"
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
              "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "queries/s" and d["dtype"] == "u32" and d["data"] == "synthetic" and d["vs_baseline"] is None
    assert d["steps"] == 3 and d["gpu_launches"] == 3 and d["results_ok"] is True
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 4 * (1 << 22) and e["d2h_bytes_per_step"] == 4 * (1 << 22) and e["value"] > 0
    assert d["sa"]["binary_ok"] and d["sa"]["mlr_ok"] and d["sa"]["mlr_equals_binary"] and d["sa"]["sa_check_violations"] == 0
    assert d["sa"]["cpu_baseline"]["equals_gpu"] and d["sa"]["ok"] and d["sa"]["e2e"]["equals_device_path"]
    assert "workload" in d["config"]
    assert d["c2"]["ok"] and set(d["c2"]["sizes"]) == {"2^10", "2^16"} and all(v["queries_per_s"] > 0 for v in d["c2"]["sizes"].values())
    rep = d["sa_repetitive"]
    assert rep["ok"] and rep["binary_ok"] and rep["mlr_ok"] and rep["mlr_equals_binary"] and rep["pattern_len"] == [200, 2000]
    # BASELINE configs C4 / C5 ride in the same line, each with an exact host-side parity sample that gates the exit code
    c4, c5 = d["c4"], d["c5"]
    assert c4["ok"] and set(c4["layouts"]) == {"stree16_left_max", "map_b20"} and c4["scaling"] == "strong"
    for lay in c4["layouts"].values():
        assert lay["ok"] and lay["queries_per_s"] > 0 and lay["e2e"]["equals_device_path"] and lay["e2e"]["h2d_bytes_per_step"] == 4 * 5000000
    assert c5["ok"] and c5["binary_ok"] and c5["mlr_ok"] and c5["mlr_equals_binary"] and c5["pattern_len"] == [20, 100]
    assert c5["binary_patterns_per_s"] > 0 and c5["e2e"]["equals_device_path"] and "parity_sample" in c5


@pytest.mark.gpu
def test_bench_line_bucketed(gpu):
    """A configuration large enough for SCHEME_AUTO to take the reordered-batch pipeline: 4 launches per step, stage times."""
    out = subprocess.run(
        [sys.executable, os.path.join(ROOT, "bench.py"), "--n-keys", str(1 << 27), "--queries", str(1 << 24), "--steps", "2", "--warmup", "3",
         "--no-cpu", "--sa-text", "0", "--e2e-steps", "1", "--c4-log2-keys", "0", "--c5-text", "0", "--sa-rep-text", "0", "--c2-sizes", ""],
        capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    d = json.loads([l for l in out.stdout.splitlines() if l.strip()][-1])
    assert d["config"]["scheme"] == "bucketed" and d["gpu_launches"] == 2 * 4 and d["results_ok"] is True
    assert set(d["roofline"]["stage_ms"]) == {"partition", "plan", "search", "unpermute"}
