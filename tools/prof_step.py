import os
os.environ.setdefault("JPDSE_VGG_RANDOM", "1")  # offline box: no pretrained VGG19 checkpoint
import importlib, sys, os, torch
sys.path.insert(0, '/root/repo')
import jpdse_b200, bench
tr = importlib.import_module("jpd-se_b200.ctu.trainers.pix2pixHD_trainer")
opt = bench.make_opt(); opt.is_train, opt.quiet = True, True
torch.manual_seed(1234)
trainer = tr.Pix2PixHDTrainer(opt, mode="train")
label, inst, image = bench.synth_inputs(2, 512, 1024)
x = {"label": label, "instance": inst, "image": image}
for _ in range(3): trainer.step(x)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(2): trainer.step(x)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = sum(e.device_time for e in ev if hasattr(e, "device_time")) / 2e3
ours = sum(e.device_time for e in ev if "jpdse" in e.name) / 2e3
import time
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): trainer.step(x)
torch.cuda.synchronize(); wall = (time.perf_counter() - t0) / 5 * 1e3
print("GPU kernel time per step: %.2f ms (jpdse kernels %.2f ms, everything else %.2f ms); wall per step %.2f ms (no profiler)" % (tot, ours, tot - ours, wall))
