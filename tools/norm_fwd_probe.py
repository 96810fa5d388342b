"""The forward InstanceNorm apply kernel on one shape (default: the ResnetBlock shape of a batch-8 half plan), both the
ReLU and the residual variant, a few calls (for ncu / timing).   python tools/norm_fwd_probe.py [B H W C pad iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import jpdse_b200  # noqa: E402,F401
from jpdse_b200 import ops  # noqa: E402

a = [int(v) for v in sys.argv[1:]]
B, H, W, C, pad = a[:5] if len(a) >= 5 else (8, 32, 64, 1024, 1)
iters = a[5] if len(a) > 5 else 20
dev = torch.device("cuda")
raw = torch.randn(B, H, W, C, device=dev).bfloat16()
res = torch.randn(B, H + 2 * pad, W + 2 * pad, C, device=dev).bfloat16()
out = torch.empty(B, H + 2 * pad, W + 2 * pad, C, device=dev, dtype=torch.bfloat16)
st = torch.stack((raw.double().sum(dim=(1, 2)), (raw.double() ** 2).sum(dim=(1, 2))), -1).contiguous()
for name, relu, r in (("relu", True, None), ("residual", False, res)):
    for _ in range(3):
        ops.instnorm_apply(raw, st, out, B, H, W, C, pad, relu, r)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.instnorm_apply(raw, st, out, B, H, W, C, pad, relu, r)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    nbytes = (raw.numel() + out.numel() + (res.numel() if r is not None else 0)) * 2
    print("instnorm_apply %s (%d,%d,%d,%d) pad %d, %d back-to-back calls: %.1f us/call, %.0f GB/s" % (name, B, H, W, C, pad, iters, us, nbytes / us / 1e3))
