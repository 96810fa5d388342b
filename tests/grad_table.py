"""Per-parameter gradient agreement table: sm_100a backward vs CPU oracle autograd (fp32 and bf16-emulated forward).
Usage: python tests/grad_table.py n_down n_blocks B H W"""
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import generator_oracle as orc  # noqa: E402

n_down, n_blocks, B, H, W = (int(a) for a in sys.argv[1:6])
nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
torch.manual_seed(99)
net = nw.define_G(39, 3, 64, "global", n_down, n_blocks, 1, 3, "instance", gpu_ids=[])
sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
gen = torch.Generator().manual_seed(17)
x = torch.randn(B, 39, H, W, generator=gen)
target = torch.rand(B, 3, H, W, generator=gen) - 0.5


def loss_fn(y, t):
    return 10.0 * (y - t).abs().mean() + (y * y).mean()


class RoundSTE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t):
        return t.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


class RoundBoth(torch.autograd.Function):  # rounds the gradient too, like the kernels' bf16 gradient tensors
    @staticmethod
    def forward(ctx, t):
        return t.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def oracle_grads(round_fn):
    for v in sd.values():
        v.grad = None
    loss_fn(orc.generator_forward(sd, x, n_down, n_blocks, round_fn=round_fn), target).backward()
    return {k: v.grad.clone() for k, v in sd.items() if v.grad is not None}


def cos(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


ref32, ref16, ref16b = oracle_grads(None), oracle_grads(RoundSTE.apply), oracle_grads(RoundBoth.apply)
dev = torch.device("cuda", 0)
net = net.to(dev).train()
y = net(x.to(dev))
loss_fn(y, target.to(dev)).backward()
torch.cuda.synchronize()
print("%-34s %9s %9s %9s %9s %9s" % ("parameter", "cos(bf16)", "cos(fp32)", "16v32", "16v16b", "norm-ratio"))
for name, p in net.named_parameters():
    if name not in ref16 or float(ref16[name].abs().max()) < 1e-12:
        continue
    g = p.grad.cpu()
    print("%-34s %9.5f %9.5f %9.5f %9.5f %9.4f" % (name, cos(g, ref16[name]), cos(g, ref32[name]), cos(ref16[name], ref32[name]),
                                                cos(ref16[name], ref16b[name]), float(g.norm() / ref16[name].norm())))
