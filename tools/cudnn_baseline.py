"""The "existing Blackwell path": the same GlobalGenerator architecture on stock PyTorch modules (ATen/cuDNN) on the
GPU -- fp32 (TF32 allowed, PyTorch's default for convs), and bf16 autocast with channels_last -- timed like bench.py's
device-resident loop. For comparison only; nothing here is on the jpdse_b200 path.

  python tools/cudnn_baseline.py [--batch 16] [--height 512] [--width 1024] [--iters 10]
"""
import argparse

import torch
import torch.nn as nn


def res_block(dim):
    return nn.Sequential(nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, 3), nn.InstanceNorm2d(dim), nn.ReLU(True),
                         nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, 3), nn.InstanceNorm2d(dim))


class Res(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.body = res_block(dim)

    def forward(self, x):
        return x + self.body(x)


def generator(input_nc=39, output_nc=3, ngf=64, n_down=4, n_blocks=9):
    m = [nn.ReflectionPad2d(3), nn.Conv2d(input_nc, ngf, 7), nn.InstanceNorm2d(ngf), nn.ReLU(True)]
    c = ngf
    for _ in range(n_down):
        m += [nn.Conv2d(c, 2 * c, 3, stride=2, padding=1), nn.InstanceNorm2d(2 * c), nn.ReLU(True)]
        c *= 2
    m += [Res(c) for _ in range(n_blocks)]
    for _ in range(n_down):
        m += [nn.ConvTranspose2d(c, c // 2, 3, stride=2, padding=1, output_padding=1), nn.InstanceNorm2d(c // 2), nn.ReLU(True)]
        c //= 2
    m += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, 7), nn.Tanh()]
    return nn.Sequential(*m)


def timed(fn, iters):
    for _ in range(3):
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
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--iters", type=int, default=10)
    a = ap.parse_args()
    dev = torch.device("cuda")
    torch.manual_seed(1234)
    net = generator().to(dev).eval()
    x = torch.randn(a.batch, 39, a.height, a.width, device=dev)
    torch.backends.cudnn.benchmark = True
    with torch.no_grad():
        ms = timed(lambda: net(x), a.iters)
        print("PyTorch/cuDNN fp32 (TF32 convs), NCHW:            %8.2f ms/step  %8.1f images/s" % (ms, a.batch / ms * 1e3))
        net_cl = net.to(memory_format=torch.channels_last)
        x_cl = x.contiguous(memory_format=torch.channels_last)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms = timed(lambda: net_cl(x_cl), a.iters)
        print("PyTorch/cuDNN bf16 autocast, channels_last:       %8.2f ms/step  %8.1f images/s" % (ms, a.batch / ms * 1e3))
    # training-shaped: forward + backward of an L1 loss at a small batch
    net.train()
    xb = torch.randn(2, 39, a.height, a.width, device=dev)
    tgt = torch.rand(2, 3, a.height, a.width, device=dev) - 0.5

    def fb():
        for p in net.parameters():
            p.grad = None
        ((net(xb) - tgt).abs().mean() * 10).backward()
    ms = timed(fb, max(3, a.iters // 2))
    print("PyTorch/cuDNN fp32 (TF32) forward+backward, batch 2: %8.2f ms/step  %8.1f images/s" % (ms, 2 / ms * 1e3))
    with torch.autocast("cuda", dtype=torch.bfloat16):
        ms = timed(fb, max(3, a.iters // 2))
    print("PyTorch/cuDNN bf16 autocast forward+backward, batch 2: %6.2f ms/step  %8.1f images/s" % (ms, 2 / ms * 1e3))


if __name__ == "__main__":
    main()
