"""GPU parity tests of the VGG19 perceptual loss on a B200, through the C ABI.

Reference: Vgg19 / VGGLoss (ctu/models/pix2pixHD_networks/networks.py:474-504, 124-139). Checker:
oracle/discriminator_oracle.py::vgg_forward / vgg_loss (pinned bit-identical to the reference's modules with shared
random weights by oracle/pin_against_reference.py; the pretrained checkpoint is a download and absent offline).

Tolerances: every cut feature map |err| mean <= 2 % of its RMS (bf16 operands, fp32 accumulate, 13 layers deep); the
loss within 2 %; d(loss)/d(fake image): norm within 5 % of the fp32 oracle's autograd and cosine >= 0.98, or -- where bf16
rounding of the forward activations flips ReLU / max-pool choices -- >= what the oracle's own bf16-operand emulation reaches
against its fp32 self minus 0.02 (the calibration used for the generator, tests/test_gpu_backward.py).
"""
import importlib

import pytest
import torch

from oracle import discriminator_oracle as dorc

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.bfloat16().float()


def _cos(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


def _vgg_loss_module(cuda, seed=3):
    nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
    torch.manual_seed(seed)
    vl = nw.VGGLoss([])
    # torchvision's default init (kaiming) keeps activations O(1) through the 13 layers: a fair stand-in for the weights
    sd = {k: v.detach().clone() for k, v in vl.vgg.state_dict().items()}
    return vl.to(cuda), sd


@pytest.mark.parametrize("B,H,W", [(1, 64, 128), (2, 32, 64), (1, 128, 256)])
def test_vgg_features_loss_and_gradient_vs_oracle(cuda, B, H, W):
    vl, sd = _vgg_loss_module(cuda)
    g = torch.Generator().manual_seed(B * 100 + H)
    real = torch.rand(B, 3, H, W, generator=g) - 0.5
    fake = (real + 0.2 * torch.randn(B, 3, H, W, generator=g)).clamp(-1, 1)
    fo = fake.clone().requires_grad_(True)
    ref_loss = dorc.vgg_loss(sd, fo, real)
    ref_loss.backward()

    class _RoundSTE(torch.autograd.Function):  # bf16 rounding with a straight-through gradient
        @staticmethod
        def forward(ctx, t):
            return t.bfloat16().float()

        @staticmethod
        def backward(ctx, g_):
            return g_
    fe = fake.clone().requires_grad_(True)
    dorc.vgg_loss(sd, fe, real, round_fn=_RoundSTE.apply).backward()
    cal = _cos(fe.grad, fo.grad)  # how close ANY bf16-operand path gets to the fp32 gradient (ReLU / max-pool choices flip)
    with torch.no_grad():
        ref_feats = dorc.vgg_forward(sd, torch.cat((fake, real), 0))
    f = fake.clone().to(cuda).requires_grad_(True)
    loss = vl(f, real.to(cuda))
    loss.backward()
    torch.cuda.synchronize()
    plan = vl.plan_for(B, H, W, cuda)
    cuts = [st for st in plan.stages if st.cut is not None]
    assert len(cuts) == 5
    for st, ref in zip(cuts, ref_feats):
        got = st.y.float().cpu()[:, 1:-1, 1:-1].permute(0, 3, 1, 2)
        assert got.shape == ref.shape
        rms = float(ref.pow(2).mean().sqrt())
        assert float((got - ref).abs().mean()) <= 0.02 * rms, (st.key, float((got - ref).abs().mean()), rms)
        ring = st.y.float().cpu().clone()
        ring[:, 1:-1, 1:-1] = 0
        assert float(ring.abs().max()) == 0.0  # zero border = the next conv's padding
    assert abs(float(loss) - float(ref_loss)) <= 0.02 * float(ref_loss), (float(loss), float(ref_loss))
    got_g = f.grad.cpu()
    c = _cos(got_g, fo.grad)
    ratio = float(got_g.norm() / fo.grad.norm())
    print("VGG loss %.5f (oracle %.5f); gradient cosine %.5f, norm ratio %.4f" % (float(loss), float(ref_loss), c, ratio))
    c_emu = _cos(got_g, fe.grad)
    print("bf16-emulation oracle vs fp32 oracle: cosine %.5f; ours vs the emulation: %.5f" % (cal, c_emu))
    assert c >= min(0.98, cal - 0.02) and c_emu >= min(0.98, cal - 0.01) and 0.95 <= ratio <= 1.05


def test_vgg_loss_guards(cuda):
    import jpdse_b200
    vl, _ = _vgg_loss_module(cuda)
    with pytest.raises(jpdse_b200.JpdseError):
        vl(torch.zeros(1, 3, 32, 64), torch.zeros(1, 3, 32, 64))  # CPU tensors: no fallback
    with pytest.raises(jpdse_b200.JpdseError):
        vl.vgg(torch.zeros(1, 3, 32, 64, device=cuda))
    x = torch.zeros(1, 3, 32, 64, device=cuda)
    assert float(vl(x, x)) == 0.0
    with torch.no_grad():
        assert not vl(x, x + 0.1).requires_grad
