"""Developer probe: are the backward kernels bit-identical for image 1 of a batch of 2 and the same image alone?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jpdse_b200
from jpdse_b200 import ops
from jpdse_b200._lib import CONV3X3_FULL, CONV3X3_PAD1, CONV3X3_S2, CONVT3X3_S2, EPI_RAW, EPI_RAW_STATS
dev = torch.device("cuda")
torch.manual_seed(0)
C, h, w = 1024, 16, 32


def conv_out(kind, epi, B, x, wt, pad, cin, cout, hh, ww, want_stats=False):
    cv = ops.Conv(kind, epi, B, hh, ww, pad, cin, cin, cout, dev)
    cv.pack(wt)
    oh, ow = cv.out_hw
    y = torch.zeros(B, oh, ow, cout, dtype=torch.bfloat16, device=dev)
    st = torch.zeros(B, cout, 2, dtype=torch.float64, device=dev) if want_stats else None
    cv.forward(x, y, st)
    return y, st


wt = torch.randn(C, C, 3, 3, device=dev) * 0.02
# FULL dgrad
x2 = ops.alloc_nhwc(2, h + 4, w + 4, C, dev); x2.zero_(); x2[:, 2:-2, 2:-2].normal_()
y2, _ = conv_out(CONV3X3_FULL, EPI_RAW, 2, x2, wt, 2, C, C, h, w)
x1 = ops.alloc_nhwc(1, h + 4, w + 4, C, dev); x1.copy_(x2[1:])
y1, _ = conv_out(CONV3X3_FULL, EPI_RAW, 1, x1, wt, 2, C, C, h, w)
print("FULL dgrad   image1 of batch2 == alone:", torch.equal(y2[1:], y1), float((y2[1:].float() - y1.float()).abs().max()))
# forward PAD1 with stats
xp2 = ops.alloc_nhwc(2, h + 2, w + 2, C, dev); xp2.normal_()
r2, s2 = conv_out(CONV3X3_PAD1, EPI_RAW_STATS, 2, xp2, wt, 1, C, C, h, w, True)
xp1 = ops.alloc_nhwc(1, h + 2, w + 2, C, dev); xp1.copy_(xp2[1:])
r1, s1 = conv_out(CONV3X3_PAD1, EPI_RAW_STATS, 1, xp1, wt, 1, C, C, h, w, True)
print("PAD1 forward raw equal:", torch.equal(r2[1:], r1), " stats max rel diff:", float(((s2[1:] - s1).abs() / (s1.abs() + 1e-30)).max()))
# IN backward
g2 = torch.randn(2, h + 2, w + 2, C, device=dev).bfloat16()
dy2 = torch.zeros(2, h, w, C, dtype=torch.bfloat16, device=dev); sm2 = torch.zeros(2, C, 2, dtype=torch.float64, device=dev)
ops.instnorm_backward_reduce(g2, 1, None, r2, s2, dy2, sm2, 2, h, w, C, True)
dx2 = torch.zeros(2, h + 4, w + 4, C, dtype=torch.bfloat16, device=dev)
ops.instnorm_backward_apply(dy2, r2, s2, sm2, dx2, 2, 2, h, w, C)
g1 = g2[1:].contiguous()
dy1 = torch.zeros(1, h, w, C, dtype=torch.bfloat16, device=dev); sm1 = torch.zeros(1, C, 2, dtype=torch.float64, device=dev)
ops.instnorm_backward_reduce(g1, 1, None, r1, s1, dy1, sm1, 1, h, w, C, True)
dx1 = torch.zeros(1, h + 4, w + 4, C, dtype=torch.bfloat16, device=dev)
ops.instnorm_backward_apply(dy1, r1, s1, sm1, dx1, 2, 1, h, w, C)
print("IN bwd dy equal:", torch.equal(dy2[1:], dy1), " sums max rel diff:", float(((sm2[1:] - sm1).abs() / (sm1.abs() + 1e-30)).max()),
      " dx equal:", torch.equal(dx2[1:], dx1), float((dx2[1:].float() - dx1.float()).abs().max()))
# stride-2 kind as dgrad of ConvT (1024 <- 512 at 32x64 -> 16x32)
wt2 = torch.randn(1024, 512, 3, 3, device=dev) * 0.02
xs2 = ops.alloc_nhwc(2, 2 * h, 2 * w, 512, dev); xs2.normal_()
ys2, _ = conv_out(CONV3X3_S2, EPI_RAW, 2, xs2, wt2, 0, 512, 1024, 2 * h, 2 * w)
xs1 = ops.alloc_nhwc(1, 2 * h, 2 * w, 512, dev); xs1.copy_(xs2[1:])
ys1, _ = conv_out(CONV3X3_S2, EPI_RAW, 1, xs1, wt2, 0, 512, 1024, 2 * h, 2 * w)
print("S2 (ConvT dgrad) equal:", torch.equal(ys2[1:], ys1), float((ys2[1:].float() - ys1.float()).abs().max()))
