"""Runs the weight-gradient kernel of one conv shape a few times (for ncu / timing).
  python tools/wgrad_probe.py kind B H W cin cout [iters]     kind in conv3x3|convs2|convt|stem|head"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import jpdse_b200  # noqa: E402,F401
from jpdse_b200 import ops  # noqa: E402
from jpdse_b200._lib import (CONV3X3_PAD1, CONV3X3_S2, CONV7X7_PAD3, CONVT3X3_S2, EPI_BIAS_TANH_NCHW,  # noqa: E402
                             EPI_RAW_STATS)

kind, B, H, W, cin, cout = sys.argv[1], *(int(a) for a in sys.argv[2:7])
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 5
dev = torch.device("cuda")
cfg = {"conv3x3": (CONV3X3_PAD1, EPI_RAW_STATS, 1, 2, (H, W), (cout, cin, 3, 3)),
       "convs2": (CONV3X3_S2, EPI_RAW_STATS, 0, 0, (H // 2, W // 2), (cout, cin, 3, 3)),
       "convt": (CONVT3X3_S2, EPI_RAW_STATS, 0, 0, (2 * H, 2 * W), (cin, cout, 3, 3)),
       "stem": (CONV7X7_PAD3, EPI_RAW_STATS, 3, 0, (H, W), (cout, 39, 7, 7)),
       "head": (CONV7X7_PAD3, EPI_BIAS_TANH_NCHW, 3, 6, (H, W), (cout, cin, 7, 7))}[kind]
k, epi, pad, dy_pad, (oh, ow), wshape = cfg
cv = ops.Conv(k, epi, B, H, W, pad, cin, 39 if kind == "stem" else cin, cout, dev)
x = ops.alloc_nhwc(B, H + 2 * pad, W + 2 * pad, cin, dev)
x.normal_()
dyc = 8 if kind == "head" else cout
dy = ops.alloc_nhwc(B, oh + 2 * dy_pad, ow + 2 * dy_pad, dyc, dev)
dy.normal_()
dw = torch.empty(wshape, device=dev)
for _ in range(2):
    cv.wgrad(x, dy, dy_pad, dw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    cv.wgrad(x, dy, dy_pad, dw)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("%s B%d %dx%d %d->%d: %.3f ms/call (memset + wgrad + finalize), %.1f TFLOP/s" % (kind, B, H, W, cin, cout, ms, cv.flops / ms / 1e9))
