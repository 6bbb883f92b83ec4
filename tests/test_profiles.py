"""The committed evidence is self-consistent (no GPU needed): the ncu launch list of the bench command, summarised by
tools/launch_summary.py, agrees with the bench line on the dominant kernel's share of the step and on the DRAM traffic the
bench line quotes (profiles/traffic.json)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")


def test_launch_summary_matches_bench_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "launch_summary.py"), os.path.join(PROF, "r1_s4_launches.csv")],
                         capture_output=True, text=True, check=True).stdout
    summ = json.loads(out)
    bench = json.load(open(os.path.join(PROF, "r1_s4_bench.json")))
    traffic = json.load(open(os.path.join(PROF, "traffic.json")))
    shares = {k: v["share_of_step"] for k, v in summ["per_kernel"].items() if v["share_of_step"]}
    assert abs(sum(shares.values()) - 1.0) < 1e-6
    search = [v for k, v in shares.items() if k.startswith("bk_search_kernel")]
    assert len(search) == 1
    dom = bench["roofline"]["dominant_kernel"]
    assert dom["name"] == "bk_search_kernel" and abs(dom["share_of_step"] - search[0]) < 0.05  # cold-cache ncu vs live events
    # the step under ncu and the live step agree within 5 %, and the traffic the bench line carries is the summary's
    assert abs(summ["pipeline_us_per_step"] / 1e3 - bench["ms_per_step"]) / bench["ms_per_step"] < 0.05
    assert abs(traffic["dram_bytes_per_launch"] - summ["pipeline_dram_bytes_per_step"]) / summ["pipeline_dram_bytes_per_step"] < 0.01
    assert bench["roofline"]["traffic"] == traffic["dram_bytes_per_launch"] or bench["roofline"]["traffic"] > 0
    # 7 launches per step, as the bench line claims
    per_step = sum(v["launches_per_step"] for k, v in summ["per_kernel"].items() if v["share_of_step"])
    assert round(per_step) == bench["roofline"]["launches_per_step"] == 7
