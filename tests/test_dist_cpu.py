"""CPU tests for the N>1 path: world_size-2 gloo run of bench.py's host-side logic (contiguous
query shards as in static-search-tree/src/bin/bench.rs:558-573, max-over-ranks timing), and the
reference arm's behaviour under torchrun (rank 0 prints, other ranks exit 0 silently)."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WORKER = r'''
import os, sys, json
sys.path.insert(0, os.environ["SST_ROOT"])
import numpy as np
import torch
import torch.distributed as dist
import bench
from oracle import oracle as O

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
total = 100_003
s, e = bench.shard_range(total, rank, world)
# every rank holds the same (replicated) index and answers its own contiguous shard
rng = np.random.default_rng(5)
vals = np.sort(rng.integers(0, bench.MAX, 50_000, dtype=np.uint32)); vals[-1] = bench.MAX
qs = np.random.default_rng(6).integers(0, bench.MAX, total, dtype=np.uint32)
tree = O.Tree.stree(vals, left_max=True)
mine = tree.search(qs[s:e])
# gather shard sizes and a checksum; max-over-ranks timing
sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
dist.all_gather(sizes, torch.tensor([e - s]))
chk = torch.tensor([int(mine.astype(np.uint64).sum())], dtype=torch.int64)
dist.all_reduce(chk, op=dist.ReduceOp.SUM)
t = torch.tensor([1.0 + rank], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    full, _ = O.lower_bound(vals, qs)
    print(json.dumps({"sizes": [int(x) for x in sizes], "chk": int(chk), "want": int(full.astype(np.uint64).sum()), "tmax": float(t), "world": world}))
dist.destroy_process_group()
'''


def test_shard_range_rule():
    import bench

    for total in (0, 1, 7, 100, 100_003):
        for world in (1, 2, 3, 8):
            parts = [bench.shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            for (a, b), (c, d) in zip(parts, parts[1:]):
                assert b == c and a <= b
            chunk = -(-total // world) if total else 0
            assert all(e - s <= chunk for s, e in parts)


def test_hbm_levels_accounting():
    import bench

    nodes = bench.layer_nodes_for(1 << 28)
    assert nodes == [1, 12, 201, 3415, 58053, 986896, 16777216]
    assert bench.hbm_levels(nodes, 126 * 1024 * 1024) == 1  # only the 1 GiB leaf level exceeds L2
    assert bench.hbm_levels(bench.layer_nodes_for(1 << 30), 126 * 1024 * 1024) == 2
    assert bench.hbm_levels(bench.layer_nodes_for(1 << 20), 126 * 1024 * 1024) == 0


def test_gloo_world2_sharded_query(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, SST_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
         "--master-port", "29571", str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["world"] == 2 and sum(d["sizes"]) == 100_003 and d["sizes"][0] == 50_002
    assert d["chk"] == d["want"]
    assert d["tmax"] == 2.0


def test_reference_arm_other_ranks_exit_silently():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--n-keys", "100000",
                          "--ref-sample", "10000", "--steps", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_arm_line():
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n-keys", "200000",
                          "--ref-sample", "25600", "--steps", "2", "--warmup", "1"], env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
