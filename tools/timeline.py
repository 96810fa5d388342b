"""Kernel timeline of generator forward + backward (or a whole trainer.step with --step) from torch.profiler's CUPTI trace:
per stream busy time, the time no kernel runs on any stream, and the largest kernels. The trace itself is not kept.
  python tools/timeline.py [--step] [batch]"""
import importlib
import json
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
whole = "--step" in sys.argv
B = int(args[0]) if args else 2
if whole:
    os.environ.setdefault("JPDSE_VGG_RANDOM", "1")
    tr = importlib.import_module("jpd-se_b200.ctu.trainers.pix2pixHD_trainer")
    opt = bench.make_opt()
    opt.is_train, opt.quiet = True, True
    torch.manual_seed(1234)
    trainer = tr.Pix2PixHDTrainer(opt, mode="train")
    label, inst, image = bench.synth_inputs(B, 512, 1024)
    batch = {"label": label.cuda(), "instance": inst.cuda(), "image": image.cuda()}

    def step():
        trainer.step(batch)
else:
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
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
t0, t1 = ev[0]["ts"], max(e["ts"] + e["dur"] for e in ev)
print("%d device activities, span %.3f ms" % (len(ev), (t1 - t0) / 1e3))
streams = {}
for e in ev:
    streams.setdefault(e["args"].get("stream"), []).append(e)
for s, es in sorted(streams.items(), key=lambda kv: -sum(e["dur"] for e in kv[1])):
    print("  stream %-4s %4d activities, busy %.3f ms" % (s, len(es), sum(e["dur"] for e in es) / 1e3))
# union of busy intervals
iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in ev)
busy, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
gaps = []
for s, e in iv[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, cur_e - t0))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
print("some kernel running: %.3f ms; nothing running: %.3f ms in %d gaps (largest %s us)" % (
    busy / 1e3, (t1 - t0 - busy) / 1e3, len(gaps), ", ".join("%.0f" % g[0] for g in sorted(gaps, reverse=True)[:8])))
# overlap: time with >= 2 kernels in flight
pts = sorted([(e["ts"], 1) for e in ev] + [(e["ts"] + e["dur"], -1) for e in ev])
depth, last, multi = 0, pts[0][0], 0.0
for t, d in pts:
    if depth >= 2:
        multi += t - last
    depth += d
    last = t
print(">= 2 activities in flight: %.3f ms" % (multi / 1e3))
agg = {}
for e in ev:
    k = e["name"][:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += e["dur"]
for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
    print("  %8.3f ms %4d x %7.1f us  %s" % (d / 1e3, n, d / n, k))
if "--dump" in sys.argv:
    for e in ev:
        print("%9.1f %8.1f s%-3s %s" % (e["ts"] - t0, e["dur"], e["args"].get("stream"), e["name"][:60]))
