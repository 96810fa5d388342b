"""Run by tests/test_gpu_reference.py on the GPU box (subprocess: `import ctu` must resolve to baseline/_ref).

1. The UNMODIFIED reference (baseline/_ref) on CUDA in fp32 with TF32 switched off: its own parser ->
   get_trainer(opt)(opt, 'test') -> Pix2PixHDTrainer.get_img (ctu/trainers/pix2pixHD_trainer.py:113-116).
2. `jpdse_b200.install_into_reference()` (INTEGRATION.md route 1), then the SAME reference trainer class built again:
   its Pix2PixHDModel now constructs our GlobalGenerator through networks.define_G and loads the same net_G.pth.
3. Both get_img outputs on the same x_dict: generator gate of tests/test_gpu_parity.py (mean-abs <= 0.02, max-abs <= 0.15,
   PSNR >= 39.2 dB), and the reference's CUDA fp32 output against its own CPU fp32 output (sanity of the checker).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import reference_arm as ra  # noqa: E402
from oracle import generator_oracle as orc  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
H, W = int(sys.argv[1]), int(sys.argv[2])
l8, i16, _u8, img = bench.synth_inputs_compact(1, H, W, seed=77)


def x_dict():
    return {"label": l8.float(), "instance": i16.int(), "image": img.clone(), "path": ["synthetic"]}


ref_gpu, opt = ra.build_test_trainer(gpu_id=0, seed=1234)
sd = {k: v.detach().cpu().clone() for k, v in ref_gpu.model.netG.state_dict().items()}
assert type(ref_gpu.model.netG).__module__ == "ctu.models.pix2pixHD_networks.networks"
y_ref_gpu = ref_gpu.get_img(x_dict()).float().cpu()
del ref_gpu
torch.cuda.empty_cache()

ref_cpu, _ = ra.build_test_trainer(state_dict=sd, gpu_id=-1)
y_ref_cpu = ref_cpu.get_img(x_dict()).float()
d = (y_ref_gpu - y_ref_cpu).abs()
print("reference CUDA fp32 (TF32 off) vs reference CPU fp32: max abs %.3e mean abs %.3e" % (float(d.max()), float(d.mean())))
assert float(d.max()) < 5e-3

import jpdse_b200  # noqa: E402
jpdse_b200.install_into_reference()
ours, _ = ra.build_test_trainer(state_dict=sd, gpu_id=0)
assert type(ours.model.netG).__module__.startswith("jpd-se_b200."), type(ours.model.netG).__module__
y = ours.get_img(x_dict()).float().cpu()
err = (y - y_ref_gpu).abs()
p = orc.psnr(y, y_ref_gpu)
print("drop-in behind the reference trainer vs reference CUDA fp32 at %dx%d: max abs %.4f mean abs %.5f psnr %.2f dB" % (
    W, H, float(err.max()), float(err.mean()), p))
assert float(err.mean()) <= 0.02 and float(err.max()) <= 0.15 and p >= 39.2
print("REFERENCE_GPU_CHECK_OK")
