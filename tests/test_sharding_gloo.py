"""CPU: region sharding and the multi-rank plumbing of the benchmark, world_size 2 over gloo on 127.0.0.1."""
import os
import subprocess
import sys

import torch.multiprocessing as mp

from basevar_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_region_matches_the_cpp_twin_properties():
    beg = 5246595
    for g in (1, 2, 3, 4, 8):
        for length in (1, 99999, 100000, 100001, 250000, 6400001, 64000000):
            sh = shard.shard_region(beg, beg + length, g)
            assert sh and len(sh) <= g and sh[0][0] == beg and sh[-1][1] == beg + length
            tasks = []
            for i, (b, e) in enumerate(sh):
                assert e > b and (b - beg) % shard.STEP_REGION_LEN == 0
                if i:
                    assert b == sh[i - 1][1]
                tasks.append((e - b + shard.STEP_REGION_LEN - 1) // shard.STEP_REGION_LEN)
            assert max(tasks) - min(tasks) <= 1
    assert shard.shard_region(10, 10, 4) == []
    assert len(shard.shard_region(0, 50000, 8)) == 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    S = 1000
    lo, hi = shard.rank_site_range(rank, world, S)
    mine = shard.shard_region(0, 430000, world)[rank]
    gathered = [None] * world
    dist.all_gather_object(gathered, (lo, hi, mine))
    t = shard.max_over_ranks(1.0 + rank)          # the slowest rank defines the step time
    dist.barrier()
    q.put((rank, gathered, t))
    dist.destroy_process_group()


def test_two_ranks_over_gloo():
    world, port = 2, 29871
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, gathered, t in res:
        assert t == 2.0                                   # max over ranks
        assert [g[:2] for g in gathered] == [(0, 1000), (1000, 2000)]   # weak scaling: disjoint, contiguous site ranges
        regs = [g[2] for g in gathered]
        assert regs[0][0] == 0 and regs[0][1] == regs[1][0] and regs[1][1] == 430000


def test_reference_arm_runs_on_rank0_only():
    """`bench.py --impl reference` under torchrun: every rank but 0 exits 0 without work or output."""
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0 and p.stdout.strip() == ""
