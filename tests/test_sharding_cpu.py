"""CPU: the N>1 host logic (contiguous system shards, max-over-ranks timing) with world_size-2 gloo processes."""
import os
import subprocess
import sys
import textwrap

import pytest

from ilupp_b200.sharding import shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_covers_everything_once():
    for total in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_range(total, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == total
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    assert shard_range(4096, 3, 8) == (1536, 2048)
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


WORKER = textwrap.dedent("""
    import os, sys, json
    sys.path.insert(0, %r)
    import torch.distributed as dist
    from ilupp_b200.sharding import init_process_group, shard_range, max_over_ranks, sum_over_ranks, barrier
    rank, local, world = init_process_group("gloo")
    lo, hi = shard_range(4096, rank, world)
    barrier()
    t = max_over_ranks(1.0 + rank)            # slowest rank defines the step time
    units = sum_over_ranks(hi - lo)           # whole-job units
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi))
    if rank == 0:
        print(json.dumps({"t": t, "units": units, "shards": gathered}))
    dist.destroy_process_group()
""")


def test_two_rank_gloo_job(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", str(script)]
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    import json
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    res = json.loads(line)
    assert res["t"] == 2.0 and res["units"] == 4096.0
    assert res["shards"] == [[0, 2048], [2048, 4096]]
