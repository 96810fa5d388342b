"""CPU, world_size 2 over gloo: the data-parallel training plumbing (jpdse_b200.ddp).

The generator backward is a GPU kernel sequence, so here a stand-in walks the reducer protocol exactly as
GeneratorPlan.backward does (alloc -> fill -> ready, reverse layer order; finish) with per-rank gradients, and the
result must be the mean over ranks, bucket by bucket, with contiguous buckets covering everything exactly once.
"""
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


SHAPES = [("head.bias", (3,)), ("head.weight", (3, 64, 7, 7)), ("up.weight", (128, 64, 3, 3)), ("res.weight", (256, 256, 3, 3)),
          ("down.weight", (128, 64, 3, 3)), ("stem.weight", (64, 39, 7, 7)), ("stem.bias", (64,))]


def _walk(reducer, rank):
    params = [(n, torch.nn.Parameter(torch.zeros(s))) for n, s in SHAPES]
    reducer.begin(params)
    out = {}
    for i, (name, shape) in enumerate(SHAPES):
        t = reducer.alloc(name, shape)
        t.copy_(torch.full(shape, float(i + 1)) * (rank + 1))  # rank r holds (i+1)*(r+1)
        reducer.ready(name, t)
        out[name] = t
    reducer.finish()
    return out


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ddp = importlib.import_module("jpd-se_b200.ddp")
    red = ddp.GradReducer(bucket_bytes=64 * 1024, tail_bytes=48 * 1024, tail_bucket_bytes=4 * 1024)
    grads = _walk(red, rank)
    ok = all(torch.allclose(g, torch.full_like(g, (i + 1) * 1.5)) for i, (n, g) in enumerate(grads.items()))
    # discriminator-style flat all-reduce of .grad
    lin = torch.nn.Linear(4, 3)
    for p in lin.parameters():
        p.grad = torch.full_like(p, float(rank + 1))
    ddp.allreduce_grads(lin.parameters())
    ok_d = all(torch.allclose(p.grad, torch.full_like(p, 1.5)) for p in lin.parameters())
    # weights broadcast from rank 0
    torch.manual_seed(rank)
    m = torch.nn.Conv2d(2, 2, 3)
    ddp.broadcast_parameters(m)
    ret[rank] = (ok, ok_d, red.launched, [p.detach().clone() for p in m.parameters()])
    dist.barrier()
    dist.destroy_process_group()


def test_grad_reducer_two_ranks_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    for r in range(world):
        ok, ok_d, launched, _ = ret[r]
        assert ok and ok_d
        # buckets are contiguous, ordered and cover the flat buffer exactly once; more than one bucket was used
        assert len(launched) >= 3
        assert launched[0][0] == 0 and all(a[1] == b[0] for a, b in zip(launched, launched[1:]))
        total = sum((torch.Size(s).numel() + 3) // 4 * 4 for _, s in SHAPES)
        assert launched[-1][1] == total
        # the gradients produced last leave in small buckets: what is still in flight when the backward ends is tiny
        assert (launched[-1][1] - launched[-1][0]) * 4 <= 4 * 1024 + 64 * 39 * 49 * 4
    assert all(torch.equal(a, b) for a, b in zip(ret[0][3], ret[1][3]))


def test_grad_reducer_single_process_is_a_noop_allocator():
    ddp = importlib.import_module("jpd-se_b200.ddp")
    red = ddp.GradReducer(bucket_bytes=1 << 20)
    grads = _walk(red, 0)
    for i, (n, g) in enumerate(grads.items()):
        assert torch.equal(g, torch.full_like(g, float(i + 1)))
        assert g.data_ptr() % 16 == 0
    assert red.launched[-1][1] == red._offset


def test_grad_reducer_second_backward_without_set_to_none_accumulates():
    """autograd keeps the reducer's views as p.grad; a second backward (gradient accumulation, or
    zero_grad(set_to_none=False)) must neither overwrite those views nor add the flat buffer to itself."""
    ddp = importlib.import_module("jpd-se_b200.ddp")
    red = ddp.GradReducer(bucket_bytes=1 << 20)
    params = [(n, torch.nn.Parameter(torch.zeros(s))) for n, s in SHAPES]

    def backward(scale):
        red.begin(params)
        outs = []
        for i, (name, shape) in enumerate(SHAPES):
            t = red.alloc(name, shape)
            t.copy_(torch.full(shape, float(i + 1) * scale))
            red.ready(name, t)
            outs.append(t)
        red.finish()
        if red.copy_out:  # what _GeneratorFunction.backward does
            outs = [t.clone() for t in outs]
        for (_, p), g in zip(params, outs):  # AccumulateGrad: steal when empty, add in place otherwise
            if p.grad is None:
                p.grad = g
            else:
                p.grad += g

    backward(1.0)
    assert not red.copy_out
    assert all(p.grad.data_ptr() >= red.flat.data_ptr() for _, p in params)  # first backward: zero-copy views
    backward(10.0)
    assert red.copy_out
    lo, hi = red.flat.data_ptr(), red.flat.data_ptr() + red.flat.numel() * 4
    for i, (_, p) in enumerate(params):
        assert not (lo <= p.grad.data_ptr() < hi)
        assert torch.equal(p.grad, torch.full_like(p, 11.0 * (i + 1)))
    for _, p in params:
        p.grad = None
    backward(2.0)
    assert not red.copy_out
