"""Per-call GPU time of every jpdse kernel call inside one discriminator / VGG / generator pass, by call site.

Every ops.* wrapper and ops.Conv.forward / wgrad is bracketed with CUDA events; a long device-side sleep is enqueued first so
the host runs ahead and the events bracket kernel time only (no launch gaps).

  python tools/op_times.py [--batch 2] [--what d,vgg,g]
"""
import argparse
import collections
import importlib
import os
import sys

os.environ.setdefault("JPDSE_VGG_RANDOM", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import jpdse_b200  # noqa: E402,F401
from jpdse_b200 import ops  # noqa: E402

KIND = {0: "3x3pad1", 1: "3x3s2", 2: "convT", 3: "7x7", 4: "1x1", 5: "3x3full", 6: "7x7full", 7: "4x4s2", 8: "4x4s1", 9: "4x4s2dgrad",
        10: "4x4s1full", 11: "3x3narrow"}
records = []


def wrap(fn, label):
    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = fn(*a, **k)
        e1.record()
        records.append((label(*a, **k) if callable(label) else label, e0, e1))
        return r
    return inner


def conv_label(tag):
    def f(self, *a, **k):
        d = self.desc
        return "%s %s b%d %dx%d cin%d cout%d epi%d" % (tag, KIND[d.kind], d.batch, d.in_w, d.in_h, d.cin, d.cout, d.epilogue)
    return f


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--what", default="d,vgg")
    args = ap.parse_args()
    tr = importlib.import_module("jpd-se_b200.ctu.trainers.pix2pixHD_trainer")
    opt = bench.make_opt()
    opt.is_train, opt.quiet = True, True
    torch.manual_seed(1234)
    trainer = tr.Pix2PixHDTrainer(opt, mode="train")
    model = trainer.model
    dev = torch.device("cuda", 0)
    B, H, W = args.batch, 512, 1024
    label, inst, image = bench.synth_inputs(B, H, W)
    x = {"label": label.to(dev), "instance": inst.to(dev), "image": image.to(dev)}
    _, nchw = ops.build_input(x["label"], x["instance"], x["image"], 35, nhwc=False, nchw=True)
    input_label = nchw[:, :36].contiguous()
    real = x["image"]
    fake = (real + 0.05 * torch.randn_like(real)).clamp(-1, 1)

    def run_d():
        f = fake.clone().requires_grad_(True)
        l_gan, l_fm, l_real, l_fake = model.netD.fused_losses(None, f, real, ids=(x["label"], x["instance"]), num_labels=35)
        (l_gan + 10.0 * l_fm).backward()
        for p in model.netD.parameters():
            p.grad = None
        ((l_fake + l_real) * 0.5).backward()

    def run_vgg():
        f = fake.clone().requires_grad_(True)
        model.criterionVGG(f, real).backward()

    def run_g():
        for p in model.netG.parameters():
            p.grad = None
        y = model.netG.forward_from_maps(x["label"], x["instance"], x["image"], 35)
        ((y - real).abs().mean() * 10.0).backward()

    runs = {"d": run_d, "vgg": run_vgg, "g": run_g}
    for name in args.what.split(","):
        runs[name]()  # warm-up (plans, packing)
    torch.cuda.synchronize()
    for n in ("d_input", "d_input_backward", "instnorm_apply_act", "act_backward", "l1_pair", "l1_pair_backward", "maxpool2x2",
              "maxpool2x2_backward", "nhwc_pad_to_nchw", "instnorm_backward_reduce_act", "instnorm_backward_apply",
              "instnorm_backward_reduce", "instnorm_apply", "build_input", "tanh_backward_nchw", "nchw_to_nhwc_bf16"):
        def lab(*a, _n=n, **k):
            shp = [tuple(t.shape) for t in a if isinstance(t, torch.Tensor)][:1]
            return "%s %s" % (_n, shp[0] if shp else "")
        setattr(ops, n, wrap(getattr(ops, n), lab))
    ops.Conv.forward = wrap(ops.Conv.forward, conv_label("conv"))
    ops.Conv.wgrad = wrap(ops.Conv.wgrad, conv_label("wgrad"))
    for name in args.what.split(","):
        records.clear()
        torch.cuda.synchronize()
        torch.cuda._sleep(int(2e9 * 0.04))  # ~40 ms: the host enqueues the whole pass behind it
        runs[name]()
        torch.cuda.synchronize()
        agg = collections.OrderedDict()
        for lab_, e0, e1 in records:
            t = e0.elapsed_time(e1)
            if lab_ not in agg:
                agg[lab_] = [0, 0.0]
            agg[lab_][0] += 1
            agg[lab_][1] += t
        total = sum(v[1] for v in agg.values())
        print("==== %s: %d calls, %.3f ms inside jpdse calls (batch %d at %dx%d)" % (name, len(records), total, B, W, H))
        for lab_, (n_, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print("%8.3f ms  %3d x %7.1f us  %s" % (t, n_, 1e3 * t / n_, lab_))


if __name__ == "__main__":
    main()
