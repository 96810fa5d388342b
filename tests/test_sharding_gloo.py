"""CPU, world_size 2 over gloo: the image-sharded inference plumbing (no data-path collective, one gather)."""
import importlib
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_items, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sh = importlib.import_module("jpd-se_b200.sharding")
    idx = sh.shard_indices(num_items, rank, world)
    vals = [float(i) * 10.0 + 1.0 for i in idx]  # stand-in for a per-image metric
    out = sh.gather_results(vals, num_items, rank, world)
    ret[rank] = (idx, out.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_partition_is_exact_cover():
    sh = importlib.import_module("jpd-se_b200.sharding")
    for n in (0, 1, 7, 30):
        for world in (1, 2, 4, 8):
            seen = sorted(i for r in range(world) for i in sh.shard_indices(n, r, world))
            assert seen == list(range(n))


def test_gather_two_ranks_gloo():
    world, n = 2, 7  # ragged: rank 0 gets 4 items, rank 1 gets 3
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n, ret), nprocs=world, join=True)
    want = [float(i) * 10.0 + 1.0 for i in range(n)]
    for r in range(world):
        idx, out = ret[r]
        assert idx == list(range(r, n, world))
        assert out == want


def test_single_rank_needs_no_process_group():
    sh = importlib.import_module("jpd-se_b200.sharding")
    out = sh.gather_results([1.0, 2.0, 3.0], 3, 0, 1)
    assert out.tolist() == [1.0, 2.0, 3.0]
