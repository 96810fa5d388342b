"""Per-kernel time breakdown of one generator forward + backward (training plan) and of a whole trainer.step.

  python tools/train_times.py [--batch 2] [--height 512] [--width 1024] [--iters 3] [--no-step]
"""
import argparse
import importlib
import os
os.environ.setdefault("JPDSE_VGG_RANDOM", "1")  # offline box: no pretrained VGG19 checkpoint
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--no-step", action="store_true")
    ap.add_argument("--fp16", action="store_true", help="opt.fp16: netD / VGG under bf16 autocast")
    args = ap.parse_args()
    import torch
    import jpdse_b200  # noqa: F401
    from jpdse_b200 import ops
    import bench
    tr = importlib.import_module("jpd-se_b200.ctu.trainers.pix2pixHD_trainer")
    dev = torch.device("cuda")
    torch.manual_seed(1234)
    opt = bench.make_opt()
    opt.is_train, opt.quiet, opt.fp16 = True, True, args.fp16
    trainer = tr.Pix2PixHDTrainer(opt, mode="train")
    net = trainer.model.netG
    B, H, W = args.batch, args.height, args.width
    label, inst, image = bench.synth_inputs(B, H, W)
    x_dict = {"label": label, "instance": inst, "image": image}
    records = []

    def ev():
        return torch.cuda.Event(enable_timing=True)

    def timed(label_fn, orig, flops_fn=None):
        def f(*a, **k):
            e0, e1 = ev(), ev()
            e0.record()
            r = orig(*a, **k)
            e1.record()
            records.append((label_fn(*a, **k), e0, e1, flops_fn(*a, **k) if flops_fn else 0.0))
            return r
        return f

    kinds = {0: "3x3", 1: "s2", 2: "convT", 3: "7x7", 4: "1x1", 5: "3x3full", 6: "7x7full"}

    def cname(self):
        d = self.desc
        return "%s %d->%d @%dx%d" % (kinds[d.kind], d.cin, d.cout, d.in_h, d.in_w)

    ops.Conv.forward = timed(lambda self, *a, **k: "fwd/dgrad " + cname(self), ops.Conv.forward, lambda self, *a, **k: self.flops)
    ops.Conv.wgrad = timed(lambda self, *a, **k: "wgrad " + cname(self), ops.Conv.wgrad, lambda self, *a, **k: self.flops)
    ops.instnorm_apply = timed(lambda *a, **k: "norm apply c%d" % a[6], ops.instnorm_apply)
    ops.instnorm_backward_reduce = timed(lambda *a, **k: "norm bwd reduce c%d" % a[10], ops.instnorm_backward_reduce)
    ops.instnorm_backward_apply = timed(lambda *a, **k: "norm bwd apply c%d" % a[9], ops.instnorm_backward_apply)
    ops.tanh_backward_nchw = timed(lambda *a, **k: "tanh bwd", ops.tanh_backward_nchw)
    ops.build_input = timed(lambda *a, **k: "build_input", ops.build_input)

    def fwd_bwd():
        y = net.forward_from_maps(label.to(dev), inst.to(dev), image.to(dev), 35)
        loss = (y - image.to(dev)).abs().mean()
        for p in net.parameters():
            p.grad = None
        loss.backward()

    for _ in range(2):
        fwd_bwd()
    torch.cuda.synchronize()
    records.clear()
    t0, t1 = ev(), ev()
    t0.record()
    for _ in range(args.iters):
        fwd_bwd()
    t1.record()
    torch.cuda.synchronize()
    agg = {}
    order = []
    for name, a, b, fl in records:
        if name not in agg:
            agg[name] = [0, 0.0, 0.0]
            order.append(name)
        agg[name][0] += 1
        agg[name][1] += a.elapsed_time(b)
        agg[name][2] += fl
    print("%-44s %6s %9s %9s" % ("kernel", "calls", "ms/iter", "TFLOP/s"))
    tot = 0.0
    for name in order:
        n, ms, fl = agg[name]
        tot += ms
        print("%-44s %6d %9.3f %9.1f" % (name, n // args.iters, ms / args.iters, fl / (ms * 1e-3) / 1e12 if fl else 0.0))
    print("sum of kernels %.3f ms/iter; wall (events) %.3f ms/iter (generator fwd+bwd, batch %d)" % (
        tot / args.iters, t0.elapsed_time(t1) / args.iters, B))
    if not args.no_step:
        for _ in range(2):
            trainer.step(x_dict)
        torch.cuda.synchronize()
        t0.record()
        for _ in range(args.iters):
            trainer.step(x_dict)
        t1.record()
        torch.cuda.synchronize()
        print("trainer.step (G fwd+bwd on jpdse kernels, D/VGG/losses/Adam in PyTorch): %.3f ms/step, batch %d" % (
            t0.elapsed_time(t1) / args.iters, B))


if __name__ == "__main__":
    main()
