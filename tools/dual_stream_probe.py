"""Experiment: one batch-16 plan on one stream vs two batch-8 plans on two streams (norm kernels of one half can
co-reside with the tensor-bound convs of the other)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import jpdse_b200
import bench
from jpdse_b200.engine import GeneratorPlan
nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
dev = torch.device("cuda")
torch.manual_seed(1234)
net = nw.define_G(39, 3, 64, "global", 4, 9, 1, 3, "instance", gpu_ids=[0]).eval()
sd = net.state_dict()
B, H, W = 16, 512, 1024
label, inst, image = [t.to(dev) for t in bench.synth_inputs(B, H, W)]


def timed(fn, iters=10):
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


with torch.no_grad():
    full = GeneratorPlan(39, 3, 64, 4, 9, B, H, W, dev)
    full.load_weights(sd)
    full.use_graph = False
    t_full = timed(lambda: full.forward_from_maps(label, inst, image, 35))
    ref = full.forward_from_maps(label, inst, image, 35).clone()
    for parts in (2, 4):
        plans = [GeneratorPlan(39, 3, 64, 4, 9, B // parts, H, W, dev) for _ in range(parts)]
        for p in plans:
            p.load_weights(sd)
            p.use_graph = False
        streams = [torch.cuda.Stream() for _ in range(parts)]
        chunks = [(label[i::1][i * (B // parts):(i + 1) * (B // parts)], ) for i in range(parts)]
        ins = [(label[i * (B // parts):(i + 1) * (B // parts)].contiguous(), inst[i * (B // parts):(i + 1) * (B // parts)].contiguous(),
                image[i * (B // parts):(i + 1) * (B // parts)].contiguous()) for i in range(parts)]

        def run():
            main = torch.cuda.current_stream()
            ev = torch.cuda.Event()
            ev.record(main)
            for p, s, (a, b, c) in zip(plans, streams, ins):
                s.wait_event(ev)
                with torch.cuda.stream(s):
                    p.forward_from_maps(a, b, c, 35)
            for s in streams:
                main.wait_stream(s)
        t = timed(run)
        out = torch.cat([p.out for p in plans], 0)
        print("%d streams x batch %d: %.3f ms/step (%.1f img/s) vs single stream batch 16: %.3f ms (%.1f img/s); outputs equal: %s" % (
            parts, B // parts, t, B / t * 1e3, t_full, B / t_full * 1e3, torch.equal(out, ref)))
