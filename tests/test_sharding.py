"""Multi-GPU path on CPU: world_size-2 gloo run of the sharding + gather logic used by bench.py / align_sharded.
Pairs shard by contiguous index ranges with no data-path collective; only results are gathered."""
import os
import subprocess
import sys

from _common import ROOT

WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["QB_ROOT"])
import torch, torch.distributed as dist
from quicked_b200.sharding import shard_range, gather_results
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
n = 1001
lo, hi = shard_range(n, rank, world)
scores = list(range(lo, hi))                      # stand-in for per-shard results
cigars = [f"{i}M" for i in range(lo, hi)]
all_scores, all_cigars = gather_results(scores, cigars, rank, world)
if rank == 0:
    assert all_scores == list(range(n)), "scores out of input order"
    assert all_cigars == [f"{i}M" for i in range(n)]
    print("OK", len(all_scores))
t = torch.tensor([float(rank + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)          # the timing reduction bench.py uses
assert t.item() == world
dist.destroy_process_group()
"""


def test_shard_ranges_cover_everything():
    from quicked_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 1000, 1001):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_2_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, QB_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK 1001" in r.stdout


def test_balanced_ranges_mixed_lengths():
    """BASELINE config 5: lengths 100 bp .. 100 kbp, equal bases per class; ranges must tile the batch and balance work"""
    import random
    from quicked_b200.sharding import balanced_ranges, estimated_work
    rnd = random.Random(3)
    lengths = []
    for L, cnt in [(100, 10000), (300, 3333), (1000, 1000), (3000, 333), (10000, 100), (30000, 33), (100000, 10)]:
        lengths += [(L, L)] * cnt
    rnd.shuffle(lengths)
    for world in (1, 2, 4, 8):
        spans = balanced_ranges(lengths, world)
        assert spans[0][0] == 0 and spans[-1][1] == len(lengths)
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        loads = [sum(estimated_work(m, n) for m, n in lengths[lo:hi]) for lo, hi in spans]
        biggest = max(estimated_work(m, n) for m, n in lengths)      # a contiguous split cannot beat one pair's granularity
        assert max(loads) <= sum(loads) / world + biggest, (world, loads)
