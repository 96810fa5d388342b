"""Two warm-up steps and N measured Pix2PixHDTrainer.step calls at batch 2, 1024x512 (BASELINE.json configs[3]) -- the
short command the ncu launch list / captures of the training step are taken on.   python tools/one_train_step.py [steps]"""
import importlib
import os
import sys

os.environ.setdefault("JPDSE_VGG_RANDOM", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import bench  # noqa: E402
import jpdse_b200  # noqa: E402,F401

tr = importlib.import_module("jpd-se_b200.ctu.trainers.pix2pixHD_trainer")
opt = bench.make_opt()
opt.is_train, opt.quiet = True, True
torch.manual_seed(1234)
trainer = tr.Pix2PixHDTrainer(opt, mode="train")
dev = torch.device("cuda", 0)
label, inst, image = bench.synth_inputs(2, 512, 1024)
x = {"label": label.to(dev), "instance": inst.to(dev), "image": image.to(dev)}
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1
for _ in range(2):
    trainer.step(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    out = trainer.step(x)
e1.record()
torch.cuda.synchronize()
print("trainer.step: %.3f ms/step (G_Distortion %.5f)" % (e0.elapsed_time(e1) / n, out))
