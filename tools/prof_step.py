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
