"""Data-parallel generator training on N GPUs of one box: correctness and exposed-communication time.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
      tools/ddp_check.py [--batch 2] [--height 512] [--width 1024] [--steps 5]

1. overlapped all-reduce (jpdse_b200.ddp.GradReducer, launched bucket by bucket from inside the backward) ==
   local backward followed by a plain dist.all_reduce(AVG);
2. mean over ranks of per-rank gradients == the gradient of the N-times-larger batch on one GPU;
3. generator fwd+bwd time with the reducer attached vs detached = exposed communication.
"""
import argparse
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import jpdse_b200  # noqa: F401
    from jpdse_b200 import ddp
    import bench
    nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
    dev = torch.device("cuda", local)
    torch.manual_seed(1234)
    net = nw.define_G(39, 3, 64, "global", 4, 9, 1, 3, "instance", gpu_ids=[local]).train()
    ddp.broadcast_parameters(net)
    B, H, W = args.batch, args.height, args.width
    label, inst, image = [t.to(dev) for t in bench.synth_inputs(B, H, W, seed=1234 + rank)]

    def fwd_bwd(lab, ins, img):
        for p in net.parameters():
            p.grad = None
        y = net.forward_from_maps(lab, ins, img, 35)
        ((y - img).abs().mean() * 10.0).backward()
        return [p.grad for p in net.parameters()]

    def say(*a):
        if rank == 0:
            print(*a, flush=True)

    # ---- 1. overlapped == sequential
    net.grad_reducer = ddp.GradReducer()
    g_overlap = [g.clone() for g in fwd_bwd(label, inst, image)]
    buckets = len(net.grad_reducer.launched)
    net.grad_reducer = None
    g_local = [g.clone() for g in fwd_bwd(label, inst, image)]
    g_seq = [g.clone() for g in g_local]
    for g in g_seq:
        dist.all_reduce(g, op=dist.ReduceOp.AVG)
    worst = max(float((a - b).abs().max() / (b.abs().max() + 1e-20)) for a, b in zip(g_overlap, g_seq))
    say("1. overlapped vs sequential all-reduce: worst relative max-abs difference %.3e over %d tensors, %d buckets (%s)"
        % (worst, len(g_seq), buckets, "OK" if worst < 2e-3 else "MISMATCH"))

    # ---- 2. mean of per-rank gradients == big-batch gradient (rank 0 computes the big batch)
    gathered = []
    for t in (label, inst, image):
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        gathered.append(torch.cat(parts, 0))
    if rank == 0:
        g_big = fwd_bwd(*gathered)

        def cos(a, b):
            a, b = a.double().reshape(-1), b.double().reshape(-1)
            return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
        cs = [cos(a, b) for a, b in zip(g_seq, g_big) if float(b.abs().max()) > 0]
        rel = max(float((a - b).abs().max() / (b.abs().max() + 1e-20)) for a, b in zip(g_seq, g_big))
        say("2. mean over %d ranks vs batch-%d gradient on one GPU: min cosine %.6f, worst relative max-abs diff %.3e (%s)"
            % (world, B * world, min(cs), rel, "OK" if min(cs) > 0.9999 else "MISMATCH"))
    dist.barrier()

    # ---- 3. exposed communication
    def timed(n):
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fwd_bwd(label, inst, image)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    net.grad_reducer = None
    timed(2)
    t_local = timed(args.steps)
    net.grad_reducer = ddp.GradReducer()
    timed(2)
    t_ddp = timed(args.steps)

    def seq():
        gs = fwd_bwd(label, inst, image)
        flat = torch.cat([g.reshape(-1) for g in gs])
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    net.grad_reducer = None
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    seq()
    e0.record()
    for _ in range(args.steps):
        seq()
    e1.record()
    torch.cuda.synchronize()
    t_seq = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    dist.all_reduce(t_seq, op=dist.ReduceOp.MAX)
    say("3. generator fwd+bwd, batch %d/GPU at %dx%d on %d GPUs (max over ranks): no comm %.2f ms | overlapped all-reduce "
        "%.2f ms (exposed %.2f ms) | backward-then-all-reduce %.2f ms; gradient bytes %.0f MB"
        % (B, W, H, world, t_local, t_ddp, t_ddp - t_local, float(t_seq), sum(p.numel() for p in net.parameters()) * 4 / 1e6))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
