"""BASELINE.json configs[2] and configs[4]: batch-size sweep 1..64 at 1024x512, full-res 2048x1024, QF 33/36/39/42
inputs (the QF only changes the synthetic degradation; compute is identical), one GPU per process.

  python tools/sweep.py                      # one GPU
  torchrun --nproc-per-node N tools/sweep.py # image-sharded: every rank runs the same per-GPU batch (weak scaling)
"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import jpdse_b200  # noqa: F401
    import bench
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
    torch.manual_seed(1234)
    net = nw.define_G(39, 3, 64, "global", 4, 9, 1, 3, "instance", gpu_ids=[local]).eval()

    def run(B, H, W, qf=36, iters=10):
        label, inst, image = [t.to(dev) for t in bench.synth_inputs(B, H, W, seed=1234 + rank, qf=qf)]
        with torch.no_grad():
            for _ in range(3):
                net.forward_from_maps(label, inst, image, 35)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                net.forward_from_maps(label, inst, image, 35)
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    if rank == 0:
        print("%-12s %6s %4s %10s %12s   (%d GPU%s, whole-job images/s, max over ranks)" % (
            "image", "batch", "QF", "ms/step", "images/s", world, "s" if world > 1 else ""))
    for B in (1, 2, 4, 8, 16, 32, 64):
        ms = run(B, 512, 1024)
        if rank == 0:
            print("%-12s %6d %4d %10.3f %12.1f" % ("1024x512", B, 36, ms, world * B / ms * 1e3), flush=True)
    for qf in (33, 39, 42):
        ms = run(16, 512, 1024, qf=qf)
        if rank == 0:
            print("%-12s %6d %4d %10.3f %12.1f" % ("1024x512", 16, qf, ms, world * 16 / ms * 1e3), flush=True)
    for B in (1, 4, 16):
        ms = run(B, 1024, 2048, iters=5)
        if rank == 0:
            print("%-12s %6d %4d %10.3f %12.1f" % ("2048x1024", B, 36, ms, world * B / ms * 1e3), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
