"""Per-kernel time breakdown of one generator forward (CUDA events around every C-ABI call).

  python tools/layer_times.py [--batch 16] [--height 512] [--width 1024] [--iters 3]
"""
import argparse
import importlib
import os
import sys

os.environ["JPDSE_SPLIT_STREAMS"] = "1"  # per-kernel events need the single-stream plan

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=3)
    args = ap.parse_args()
    import torch
    import jpdse_b200  # noqa: F401
    from jpdse_b200 import ops
    import bench
    nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
    dev = torch.device("cuda")
    torch.manual_seed(1234)
    net = nw.define_G(39, 3, 64, "global", 4, 9, 1, 3, "instance", gpu_ids=[0]).eval()
    B, H, W = args.batch, args.height, args.width
    label, inst, image = [t.to(dev) for t in bench.synth_inputs(B, H, W)]
    plan = net.plan_for(B, H, W, dev)
    plan.use_graph = False
    names = {id(cv): k for k, cv in plan.convs.items()}
    records = []

    def ev():
        return torch.cuda.Event(enable_timing=True)

    orig_conv_fwd = ops.Conv.forward
    orig_apply = ops.instnorm_apply
    orig_build = ops.build_input

    def conv_fwd(self, x, y, stats=None):
        a, b = ev(), ev()
        a.record()
        r = orig_conv_fwd(self, x, y, stats)
        b.record()
        records.append(("conv " + names.get(id(self), "?"), a, b, self.flops, 0))
        return r

    def apply(raw, stats, out, batch, height, width, channels, pad, relu, residual=None, eps=1e-5):
        a, b = ev(), ev()
        a.record()
        r = orig_apply(raw, stats, out, batch, height, width, channels, pad, relu, residual, eps)
        b.record()
        nbytes = batch * channels * 2 * (height * width + (height + 2 * pad) * (width + 2 * pad) * (2 if residual is not None else 1))
        records.append(("norm c%d %dx%d p%d" % (channels, height, width, pad), a, b, 0, nbytes))
        return r

    def build(*a_, **k_):
        a, b = ev(), ev()
        a.record()
        r = orig_build(*a_, **k_)
        b.record()
        records.append(("build_input", a, b, 0, B * H * W * (4 + 4 + 12) + B * (H + 6) * (W + 6) * 80))
        return r

    ops.Conv.forward = conv_fwd
    ops.instnorm_apply = apply
    ops.build_input = build
    with torch.no_grad():
        for _ in range(2):
            plan.forward_from_maps(label, inst, image, 35)
        torch.cuda.synchronize()
        records.clear()
        t0, t1 = ev(), ev()
        t0.record()
        for _ in range(args.iters):
            plan.forward_from_maps(label, inst, image, 35)
        t1.record()
        torch.cuda.synchronize()
    total = t0.elapsed_time(t1) / args.iters
    agg = {}
    order = []
    for name, a, b, fl, by in records:
        if name not in agg:
            agg[name] = [0.0, 0, fl, by]
            order.append(name)
        agg[name][0] += a.elapsed_time(b)
        agg[name][1] += 1
    print("%-34s %6s %9s %9s %9s" % ("kernel", "calls", "ms/call", "TFLOP/s", "GB/s"))
    sum_ms = 0.0
    for name in order:
        ms, n, fl, by = agg[name]
        per = ms / n
        sum_ms += ms / args.iters
        print("%-34s %6d %9.3f %9.1f %9.1f" % (name, n // args.iters, per, fl / per / 1e9 if fl else 0.0,
                                               by / per / 1e6 if by else 0.0))
    print("sum of kernels %.3f ms/forward; wall (events) %.3f ms/forward; %.1f img/s" % (sum_ms, total, B / total * 1e3))


if __name__ == "__main__":
    main()
