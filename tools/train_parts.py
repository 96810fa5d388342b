"""Where a training step's GPU time goes (BASELINE.json configs[3], batch 2 per GPU at 1024x512): CUDA-event times of the
discriminator passes, the VGG loss, the generator and the optimizers, each timed alone, then the whole trainer.step.

  python tools/train_parts.py [--batch 2] [--iters 10] [--kernels]     (--kernels: per-kernel table from the torch profiler)
"""
import argparse
import importlib
import os
import sys

os.environ.setdefault("JPDSE_VGG_RANDOM", "1")  # offline box: no pretrained VGG19 checkpoint
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import jpdse_b200  # noqa: E402,F401


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--kernels", action="store_true")
    args = ap.parse_args()
    tr = importlib.import_module("jpd-se_b200.ctu.trainers.pix2pixHD_trainer")
    opt = bench.make_opt()
    opt.is_train, opt.quiet = True, True
    torch.manual_seed(1234)
    trainer = tr.Pix2PixHDTrainer(opt, mode="train")
    model = trainer.model
    dev = torch.device("cuda", 0)
    B, H, W = args.batch, args.height, args.width
    label, inst, image = bench.synth_inputs(B, H, W)
    x = {"label": label.to(dev), "instance": inst.to(dev), "image": image.to(dev)}
    from jpdse_b200 import ops
    _, nchw = ops.build_input(x["label"], x["instance"], x["image"], 35, nhwc=False, nchw=True)
    input_label = nchw[:, :36].contiguous()
    real = x["image"]
    fake = (real + 0.05 * torch.randn_like(real)).clamp(-1, 1)
    netD, vgg, netG = model.netD, model.criterionVGG, model.netG
    rows = []

    def d_fwd():
        plan = netD.plan_for(B, H, W, dev)
        s = plan.new_slot()
        plan.forward(s, input_label, real)
    rows.append(("netD forward, one pass (2 scales, %d images)" % B, timed(d_fwd, args.iters)))

    def d_losses(which):
        f = fake.clone().requires_grad_(True)
        l_gan, l_fm, l_real, l_fake = netD.fused_losses(input_label, f, real)
        if which == "fwd":
            return
        if which == "G":
            (l_gan + 10.0 * l_fm).backward()
        else:
            for p in netD.parameters():
                p.grad = None
            ((l_fake + l_real) * 0.5).backward()
    t_f = timed(lambda: d_losses("fwd"), args.iters)
    rows.append(("netD fused losses forward (fake + real passes, L1 / MSE terms)", t_f))
    rows.append(("  + backward of loss_G_GAN + 10 loss_G_GAN_Feat (input gradient only)", timed(lambda: d_losses("G"), args.iters) - t_f))
    rows.append(("  + backward of loss_D (parameter gradients, both passes)", timed(lambda: d_losses("D"), args.iters) - t_f))

    def v(which):
        f = fake.clone().requires_grad_(True)
        loss = vgg(f, real)
        if which == "bwd":
            loss.backward()
    t_v = timed(lambda: v("fwd"), args.iters)
    rows.append(("VGG19 loss forward (fake + real = %d images)" % (2 * B), t_v))
    rows.append(("  + backward (data gradients, %d images)" % B, timed(lambda: v("bwd"), args.iters) - t_v))

    def g_fb():
        for p in netG.parameters():
            p.grad = None
        y = netG.forward_from_maps(x["label"], x["instance"], x["image"], 35)
        ((y - real).abs().mean() * 10.0).backward()
    rows.append(("generator forward + backward", timed(g_fb, args.iters)))
    g_fb()
    rows.append(("Adam step, generator (182.6 M parameters)", timed(trainer.optimizer_G.step, args.iters)))
    trainer.step(x)
    rows.append(("Adam step, discriminator (5.6 M parameters)", timed(trainer.optimizer_D.step, args.iters)))
    rows.append(("whole Pix2PixHDTrainer.step", timed(lambda: trainer.step(x), args.iters)))
    print("batch %d at %dx%d" % (B, W, H))
    for name, ms in rows:
        print("%-80s %8.3f ms" % (name, ms))
    if args.kernels:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(2):
                trainer.step(x)
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))


if __name__ == "__main__":
    main()
