"""Times one forward conv shape (for experiments / ncu).  python tools/conv_probe.py kind B H W cin cout [iters]
kind in conv3x3|convs2|convt|stem (7x7, cin = 40 stored / 39 real)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import jpdse_b200  # noqa: E402,F401
from jpdse_b200 import ops  # noqa: E402
from jpdse_b200._lib import CONV3X3_PAD1, CONV3X3_S2, CONV7X7_PAD3, CONVT3X3_S2, EPI_RAW_STATS  # noqa: E402

kind, B, H, W, cin, cout = sys.argv[1], *(int(a) for a in sys.argv[2:7])
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 10
dev = torch.device("cuda")
k, pad = {"conv3x3": (CONV3X3_PAD1, 1), "convs2": (CONV3X3_S2, 0), "convt": (CONVT3X3_S2, 0), "stem": (CONV7X7_PAD3, 3)}[kind]
cin_real = cin - 1 if kind == "stem" else cin
cv = ops.Conv(k, EPI_RAW_STATS, B, H, W, pad, cin, cin_real, cout, dev)
ks = 7 if kind == "stem" else 3
cv.pack(torch.randn((cin, cout, 3, 3) if kind == "convt" else (cout, cin_real, ks, ks), device=dev) * 0.02)
x = ops.alloc_nhwc(B, H + 2 * pad, W + 2 * pad, cin, dev)
x.normal_()
oh, ow = cv.out_hw
y = torch.empty(B, oh, ow, cout, dtype=torch.bfloat16, device=dev)
st = torch.zeros(B, cout, 2, dtype=torch.float64, device=dev)
for _ in range(3):
    cv.forward(x, y, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    cv.forward(x, y, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("%s B%d %dx%d %d->%d flags=%s: %.3f ms, %.1f TFLOP/s" % (kind, B, H, W, cin, cout, os.environ.get("JPDSE_DEBUG_FLAGS", "0"), ms, cv.flops / ms / 1e9))
if os.environ.get("JPDSE_PROBE_CHECK") and kind == "conv3x3":
    import torch.nn.functional as F
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    wt = torch.randn(cout, cin, 3, 3, device=dev) * 0.02
    cv.pack(wt)
    st.zero_()
    cv.forward(x, y, st)
    nb = min(B, 2)
    ref = F.conv2d(x[:nb].float().permute(0, 3, 1, 2), wt.bfloat16().float())
    got = y[:nb].float().permute(0, 3, 1, 2)
    s1 = got.double().sum(dim=(2, 3))
    print("check: max err %.4g (scale %.3g), stats rel err %.3g, digest %.6f" % (
        float((got - ref).abs().max()), float(ref.abs().max()), float(((st[:nb, :, 0] - s1).abs() / (s1.abs() + 1)).max()),
        float(y.float().abs().sum())))
