"""CPU tests of the multi-GPU host logic (world_size 2, gloo): frames sharded over ranks give the same
per-frame results as one process, whatever the shard count; timing reduce = max over ranks."""
import os
import sys

import numpy as np
import pytest

from helpers import ROOT, oracle, synth


def _load_shard():
    import importlib.util
    spec = importlib.util.spec_from_file_location("orbb200_shard", os.path.join(ROOT, "orb-slam-birdview_b200", "shard.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


shard = _load_shard()


def test_shard_ranges_cover_and_overlap():
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 4, 8):
            r = shard.shard_ranges(n, world, overlap=1)
            assert len(r) == world
            assert r[0][1] == 0 and r[-1][2] == n
            sizes = [e - s for _, s, e in r]
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
            for i in range(1, world):
                assert r[i][1] == r[i - 1][2]                       # owned ranges tile the sequence
                assert r[i][0] == max(0, r[i][1] - 1)               # one predecessor frame is readable
    assert shard.sequences_to_ranks(8, 4) == [[0, 4], [1, 5], [2, 6], [3, 7]]


def _frame_result(i):
    """per-frame work of the hot path, done by the CPU oracle here (no GPU in this container)"""
    img = synth.synth_frame(120, 160, 9000 + i)
    k, d = oracle.Extractor(300, 1.2, 4, 20, 7)(img)
    return shard.digest(k, d)


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        _, s, e = shard.shard_ranges(n_frames, world)[rank]
        mine = {i: _frame_result(i) for i in range(s, e)}
        t = shard.max_over_ranks(1.0 + rank, dist)                  # slowest rank defines the time
        parts = shard.gather_objects(mine, dist)
        dist.barrier()
        if rank == 0:
            merged = {}
            for p in parts:
                merged.update(p)
            q.put((t, merged))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_two_rank_sharding_matches_single_process(world):
    import torch.multiprocessing as mp
    n_frames = 5
    want = {i: _frame_result(i) for i in range(n_frames)}
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    t, merged = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert t == float(world)                                          # max over ranks of (1 + rank)
    assert merged == want                                             # byte-identical per-frame outputs
