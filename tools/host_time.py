"""How long the HOST needs to issue one trainer.step (everything up to the final .item()) against the step's wall time:
if the two are close the step is launch-bound somewhere.   python tools/host_time.py [steps]"""
import importlib
import os
import sys
import time

os.environ.setdefault("JPDSE_VGG_RANDOM", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402

tr = importlib.import_module("jpd-se_b200.ctu.trainers.pix2pixHD_trainer")
opt = bench.make_opt()
opt.is_train, opt.quiet = True, True
torch.manual_seed(1234)
trainer = tr.Pix2PixHDTrainer(opt, mode="train")
label, inst, image = bench.synth_inputs(2, 512, 1024)
x = {"label": label.cuda(), "instance": inst.cuda(), "image": image.cuda()}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
for _ in range(3):
    trainer.step(x)
torch.cuda.synchronize()
marks = []
orig_item = torch.Tensor.item


def item(self):
    marks.append(time.perf_counter())
    return orig_item(self)


torch.Tensor.item = item
host, wall = [], []
for _ in range(n):
    torch.cuda.synchronize()
    marks.clear()
    t0 = time.perf_counter()
    trainer.step(x)
    t1 = time.perf_counter()
    host.append((marks[0] - t0) * 1e3)
    wall.append((t1 - t0) * 1e3)
torch.Tensor.item = orig_item
host.sort()
wall.sort()
print("host issue time per step: median %.2f ms (min %.2f); wall per step: median %.2f ms" % (host[n // 2], host[0], wall[n // 2]))
# phases: forward + losses / backward + optimisers
import cProfile  # noqa: E402
import pstats  # noqa: E402
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    trainer.step(x)
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative")
st.print_stats(28)
