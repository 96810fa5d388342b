import os
os.environ.setdefault("JPDSE_VGG_RANDOM", "1")  # offline box: no pretrained VGG19 checkpoint
"""Developer probe: which tensor shapes the PyTorch-side copies / adds of a trainer.step come from."""
import importlib, sys, os, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
    trainer.step(x)
    torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key in ("aten::copy_", "aten::add_", "aten::mul", "aten::add", "aten::cat", "aten::sub", "aten::abs", "aten::mean", "aten::leaky_relu", "aten::leaky_relu_backward", "aten::threshold_backward", "aten::relu_", "aten::relu","aten::sgn","aten::div","aten::mul_", "aten::avg_pool2d", "aten::avg_pool2d_backward", "aten::native_batch_norm", "aten::instance_norm", "aten::native_batch_norm_backward", "aten::max_pool2d_with_indices", "aten::max_pool2d_with_indices_backward", "aten::fill_", "aten::zero_")]
rows.sort(key=lambda e: -e.self_device_time_total)
for e in rows[:60]:
    print("%-40s %8.1f us x%-3d %s" % (e.key, e.self_device_time_total, e.count, str(e.input_shapes)[:110]))
allk = sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)
print("--- by op (self CUDA)")
for e in allk[:45]:
    print("%-70s %9.1f us x%d" % (e.key[:70], e.self_device_time_total, e.count))
