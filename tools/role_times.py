"""Per-role wait breakdown of the implicit-GEMM kernel for selected layers (developer tool).

Runs a layer alone at the bench shape and prints, averaged over CTAs, the share of the kernel each role
spent waiting: producer on free stages, MMA on data / on a free accumulator, epilogue on a finished tile.
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import jpdse_b200  # noqa: E402
from jpdse_b200 import ops  # noqa: E402
from jpdse_b200._lib import CONV3X3_PAD1, CONV3X3_S2, CONVT3X3_S2, EPI_RAW_STATS  # noqa: E402

lib = jpdse_b200._lib.load()
lib.jpdse_debug_role_counters.restype = ctypes.c_int
lib.jpdse_debug_role_counters.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
dev = torch.device("cuda")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
CASES = [("convT 128->64 @256x512", CONVT3X3_S2, 256, 512, 0, 128, 64),
         ("convT 256->128 @128x256", CONVT3X3_S2, 128, 256, 0, 256, 128),
         ("convT 512->256 @64x128", CONVT3X3_S2, 64, 128, 0, 512, 256),
         ("conv s2 64->128 @512x1024", CONV3X3_S2, 512, 1024, 0, 64, 128),
         ("conv s2 128->256 @256x512", CONV3X3_S2, 256, 512, 0, 128, 256),
         ("3x3 64->256 @128x256", CONV3X3_PAD1, 128, 256, 1, 64, 256),
         ("res 1024->1024 @32x64", CONV3X3_PAD1, 32, 64, 1, 1024, 1024)]
for name, kind, H, W, pad, cin, cout in CASES:
    cv = ops.Conv(kind, EPI_RAW_STATS, B, H, W, pad, cin, cin, cout, dev)
    wshape = (cin, cout, 3, 3) if kind == CONVT3X3_S2 else (cout, cin, 3, 3)
    cv.pack(torch.randn(wshape, device=dev) * 0.02)
    x = torch.randn(B, H + 2 * pad, W + 2 * pad, cin, device=dev).bfloat16()
    oh, ow = cv.out_hw
    y = torch.empty(B, oh, ow, cout, dtype=torch.bfloat16, device=dev)
    st = torch.zeros(B, cout, 2, dtype=torch.float64, device=dev)
    for _ in range(2):
        cv.forward(x, y, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    cv.forward(x, y, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    lib.jpdse_debug_role_counters(1, None, 0)
    cv.forward(x, y, st)  # for ConvT the counters are those of the LAST phase launch (4 taps)
    buf = (ctypes.c_longlong * (16 * 148))()
    lib.jpdse_debug_role_counters(0, buf, 16 * 148)
    v = torch.tensor(list(buf), dtype=torch.float64).view(148, 16)
    m = v.mean(dim=0)
    print("%-28s %.3f ms | producer: wait-empty %4.0f%% of %8.0f cyc | mma: wait-full %4.0f%% wait-tmem %4.0f%% of %8.0f | "
          "epilogue: wait-tile g0 %4.0f%% g1 %4.0f%% of %8.0f | drain %6.0f cyc/tile over %4.0f tiles/group" % (
              name, ms, 100 * m[0] / max(m[1], 1), m[1], 100 * m[2] / max(m[4], 1), 100 * m[3] / max(m[4], 1), m[4],
              100 * m[5] / max(m[7], 1), 100 * m[6] / max(m[7], 1), m[7], m[8] / max(m[10], 1), m[10]))
