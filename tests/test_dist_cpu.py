"""CPU tests (gloo, world_size 2): the draw sharding + count all-reduce host logic."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from certifiedgpt_b200.dist import allreduce_counts, rank_world, shard_range

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for world in (1, 2, 3, 4, 8):
        for num in (0, 1, 7, 100, 1000, 1100):
            for base in (0, 100):
                cover = []
                for r in range(world):
                    lo, hi = shard_range(base, num, r, world)
                    assert lo <= hi
                    cover += list(range(lo, hi))
                assert cover == list(range(base, base + num))
    sizes = [shard_range(0, 1000, r, 8)[1] - shard_range(0, 1000, r, 8)[0] for r in range(8)]
    assert max(sizes) - min(sizes) <= 1


def _label_of(sample):  # stand-in for f(x + eps_sample): a pure function of the GLOBAL sample index
    return (sample * 2654435761 >> 7) % 5


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert rank_world(True) == (rank, world)
    total = torch.zeros(5, dtype=torch.int64)
    cursor = 0
    for num in (100, 1000):              # selection + estimation, as in Smooth.certify
        lo, hi = shard_range(cursor, num, rank, world)
        cursor += num
        counts = torch.zeros(5, dtype=torch.int64)
        for s in range(lo, hi):
            counts[_label_of(s)] += 1
        allreduce_counts(counts, True)
        assert counts.sum().item() == num
        total += counts
    if rank == 0:
        torch.save(total, out)
    dist.destroy_process_group()


def test_sharded_counts_equal_single_process(tmp_path):
    out = str(tmp_path / "c.pt")
    mp.spawn(_worker, args=(2, 29533 + os.getpid() % 500, out), nprocs=2, join=True)
    got = torch.load(out).numpy()
    ref = np.zeros(5, dtype=np.int64)
    for s in range(1100):
        ref[_label_of(s)] += 1
    assert np.array_equal(got, ref)


def test_agents_registry_mirrors_reference_protocol():
    from certifiedgpt_b200.agents import AGENTS
    for cls in AGENTS.values():
        assert all(hasattr(cls, m) for m in ("setup_agent", "run", "finalize"))
