"""jpdse_d_input_ids (discriminator operands of [fake; real] from the ids) at batch 2, both scales: time per call and GB/s.
  python tools/d_input_probe.py [iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import jpdse_b200  # noqa: E402,F401
from jpdse_b200 import ops  # noqa: E402

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda")
B, H, W = 2, 512, 1024
l8, i16, _u8, real = bench.synth_inputs_compact(B, H, W, seed=3)
l8, i16, real = l8.to(dev), i16.to(dev), real.to(dev)
fake = (real + 0.1 * torch.randn_like(real)).clamp(-1, 1)
for pool in (False, True):
    Ho, Wo = ((H - 1) // 2 + 1, (W - 1) // 2 + 1) if pool else (H, W)
    oa = ops.alloc_nhwc(B, Ho + 4, Wo + 4, 64, dev)
    ob = ops.alloc_nhwc(B, Ho + 4, Wo + 4, 64, dev)
    for _ in range(3):
        ops.d_input_ids(l8, i16, fake, oa, real, ob, 35, pool)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.d_input_ids(l8, i16, fake, oa, real, ob, 35, pool)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    nbytes = 2 * B * Ho * Wo * 128 + B * H * W * (1 + 2 + 24)
    print("d_input_ids pool=%d: %.1f us/call, %.0f GB/s (outputs %d MB)" % (pool, us, nbytes / us / 1e3, 2 * B * Ho * Wo * 128 >> 20))
