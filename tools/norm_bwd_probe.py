"""The fused InstanceNorm backward on one shape, a few calls (for ncu / timing).
  python tools/norm_bwd_probe.py [B H W C iters]     default: the ResnetBlock shape of a batch-2 step"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import jpdse_b200  # noqa: E402,F401
from jpdse_b200 import ops  # noqa: E402
from jpdse_b200._lib import PAD_SHARED  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
B, H, W, C = (a + [2, 32, 64, 1024])[:4] if len(a) >= 4 else (2, 32, 64, 1024)
iters = a[4] if len(a) > 4 else 20
dev = torch.device("cuda")
g = torch.randn(B, H + 2, W + 2, C, device=dev).bfloat16()
raw = torch.randn(B, H, W, C, device=dev).bfloat16()
skip = torch.randn(B, H, W, C, device=dev).bfloat16()
st = torch.stack((raw.double().sum(dim=(1, 2)), (raw.double() ** 2).sum(dim=(1, 2))), -1).contiguous()
dy = torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16)
dx = torch.empty((B * (H + 2) * (W + 2) + 2 * (W + 2) + 2 + 256) * C, device=dev, dtype=torch.bfloat16)
for name, sk, want_dy, relu in (("relu", None, None, True), ("skip+dy", skip, dy, False)):
    for _ in range(3):
        ops.instnorm_backward_fused(g, 1, sk, raw, st, want_dy, dx, 2 | PAD_SHARED, B, H, W, C, relu)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.instnorm_backward_fused(g, 1, sk, raw, st, want_dy, dx, 2 | PAD_SHARED, B, H, W, C, relu)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    nbytes = (g.numel() + raw.numel() + (skip.numel() if sk is not None else 0) + (dy.numel() if want_dy is not None else 0)
              + B * (H + 2) * (W + 2) * C) * 2
    print("fused norm backward %s (%d,%d,%d,%d): %.1f us/call, %.0f GB/s algorithmic" % (name, B, H, W, C, us, nbytes / us / 1e3))
