"""Run by tests/test_gpu_reference.py on the GPU box (subprocess: `import ctu` must resolve to baseline/_ref).

The UNMODIFIED reference Pix2PixHDTrainer in TRAIN mode on CUDA (fp32, TF32 off; its own parser -> get_trainer(opt)(opt,
'train'); the only shim is the offline VGG19: `models.vgg19(pretrained=True)` downloads, so torchvision's constructor is
wrapped to build the same module with random weights, as SURVEY.md Appendix B does) against this repo's mirror trainer
with the SAME initial weights for netG, netD and VGG19 on the same x_dict:
  * the six losses of Pix2PixHDModel.get_train_loss (ctu/models/pix2pixHD_model.py:709-771) within 2 % (bf16 kernels vs fp32)
  * one full Pix2PixHDTrainer.step (ctu/trainers/pix2pixHD_trainer.py:42-85) on both: same returned G_Distortion, and the
    first Adam update of every generator / discriminator weight tensor points the same way (Adam's first step is
    lr * sign(g): agreement of the update signs, weighted by |reference update|, >= 0.9 for netD and >= 0.8 for netG --
    the generator's deepest gradients (stem, first downsampling convs) only reach cosine 0.94 against ANY fp32 path under a
    bf16 forward at random init, see tests/test_gpu_backward.py::test_generator_gradients_vs_oracle; the teacher-forced
    test there is the tight per-layer gate)
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ["JPDSE_VGG_RANDOM"] = "1"
import bench  # noqa: E402
import reference_arm as ra  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
H, W, B = int(sys.argv[1]), int(sys.argv[2]), 2
ra.import_reference()
import torchvision  # noqa: E402
from ctu.models.pix2pixHD_networks import networks as ref_networks  # noqa: E402
_v = torchvision.models.vgg19
ref_networks.models.vgg19 = lambda pretrained=False, **k: _v(weights=None)  # offline: same module, random weights
import ctu.parsers  # noqa: E402
import tempfile  # noqa: E402
tmp = tempfile.mkdtemp(prefix="jpdse_ref_train_")
saved_argv = sys.argv
sys.argv = ["train.py"] + ra.ARGV + ["--gpu_ids", "0", "--save_dir", tmp, "--checkpoints_dir", tmp, "--root_dir", tmp,
                                    "--batch_size", str(B)]
opt = ctu.parsers.CTUTrainParser().parse()
sys.argv = saved_argv
from ctu.trainers import get_trainer  # noqa: E402
torch.manual_seed(11)
ref = get_trainer(opt)(opt, mode="train")
assert type(ref.model.netG).__module__ == "ctu.models.pix2pixHD_networks.networks"

import importlib  # noqa: E402
import jpdse_b200  # noqa: E402,F401
tr = importlib.import_module("jpd-se_b200.ctu.trainers.pix2pixHD_trainer")
import copy  # noqa: E402
opt2 = copy.deepcopy(opt)
opt2.quiet = True
ours = tr.Pix2PixHDTrainer(opt2, mode="train")
ours.model.netG.load_state_dict(ref.model.netG.state_dict())
ours.model.netD.load_state_dict(ref.model.netD.state_dict())
ours.model.criterionVGG.vgg.load_state_dict(ref.model.criterionVGG.vgg.state_dict())

l8, i16, _u8, img = bench.synth_inputs_compact(B, H, W, seed=5)


def x_dict():
    return {"label": l8.float(), "instance": i16.int(), "image": img.clone(), "path": ["a"] * B}


ref.train()
ours.train()
names = ("G_GAN", "G_GAN_Feat", "G_VGG", "G_Distortion", "D_real", "D_fake")
with torch.no_grad():
    pass
lr = [float(v) for v in ref.model(x_dict(), opt, mode="get_train_loss")]
lo = [float(v) for v in ours.model(x_dict(), opt2, mode="get_train_loss")]
for n, a, b in zip(names, lo, lr):
    print("%-13s ours %.5f reference %.5f" % (n, a, b))
    assert abs(a - b) <= 0.02 * abs(b) + 1e-3, (n, a, b)

before_G = {k: v.detach().clone() for k, v in ref.model.netG.state_dict().items()}
before_D = {k: v.detach().clone() for k, v in ref.model.netD.state_dict().items()}
import contextlib  # noqa: E402
import io  # noqa: E402
with contextlib.redirect_stdout(io.StringIO()):
    d_ref = ref.step(x_dict())
d_our = ours.step(x_dict())
print("step: returned G_Distortion ours %.5f reference %.5f" % (d_our, d_ref))
assert abs(d_our - d_ref) <= 0.02 * abs(d_ref)


def agreement(before, ref_net, our_net, what, gate):
    worst = 1.0
    for (k, r), (k2, o) in zip(ref_net.state_dict().items(), our_net.state_dict().items()):
        assert k == k2
        dr, do = (r - before[k]).double().flatten(), (o - before[k]).double().flatten()
        if float(dr.abs().sum()) == 0.0:
            continue
        if k.endswith(".bias") and float(do.abs().sum()) == 0.0:
            continue  # a bias in front of an InstanceNorm: our gradient is exactly zero, the reference's is rounding noise
        w = dr.abs()
        agree = float((w * (torch.sign(dr) == torch.sign(do)).double()).sum() / w.sum())
        worst = min(worst, agree)
        assert agree >= gate, "%s %s: update-sign agreement %.4f" % (what, k, agree)
    return worst


wg = agreement(before_G, ref.model.netG, ours.model.netG, "netG", 0.8)
wd = agreement(before_D, ref.model.netD, ours.model.netD, "netD", 0.9)
print("first Adam step: worst per-tensor update-sign agreement netG %.4f netD %.4f" % (wg, wd))
print("REFERENCE_TRAIN_CHECK_OK")
