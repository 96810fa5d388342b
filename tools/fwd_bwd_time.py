"""Wall time (CUDA events) of generator forward + backward, no per-kernel instrumentation.
python tools/fwd_bwd_time.py [batch] [iters]   -- JPDSE_WGRAD_STREAM=0 keeps the weight gradients on the main stream"""
import importlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
torch.manual_seed(1234)
net = nw.define_G(39, 3, 64, "global", 4, 9, 1, 3, "instance", gpu_ids=[]).cuda().train()
label, inst, image = (t.cuda() for t in bench.synth_inputs(B, 512, 1024, seed=1))


def step():
    for p in net.parameters():
        p.grad = None
    y = net.forward_from_maps(label, inst, image, 35)
    ((y - image).abs().mean() * 10.0).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    step()
e1.record()
torch.cuda.synchronize()
gsum = sum(float(p.grad.double().abs().sum()) for p in net.parameters())
print("generator fwd+bwd batch %d: %.3f ms/iter (JPDSE_WGRAD_STREAM=%s), grad digest %.6e" % (
    B, e0.elapsed_time(e1) / iters, os.environ.get("JPDSE_WGRAD_STREAM", "1"), gsum))
