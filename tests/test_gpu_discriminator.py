"""GPU parity tests of the PatchGAN discriminator path on a B200, through the C ABI.

Reference: MultiscaleDiscriminator / NLayerDiscriminator (ctu/models/pix2pixHD_networks/networks.py:371-471), GANLoss
(:80-122), discriminate + feature matching (ctu/models/pix2pixHD_model.py:451-460, 715-753). The checker is
oracle/discriminator_oracle.py, pinned bit-identical to the reference (values and autograd) by
oracle/pin_against_reference.py, and torch CPU fp32 ops on the same bf16-rounded operands for the single kernels.

Tolerances
  single conv / data gradient (bf16 operands, fp32 accumulate): |err| <= 2^-7 of the output max (one bf16 rounding)
  weight gradient: |err| <= 2e-3 * max|ref| (fp32 summation order only)
  bandwidth kernels on identical operands: one bf16 rounding of the result; pooling / concat of the input: bit-exact
  whole discriminator (bf16 kernels) vs the fp32 oracle: every feature map |err| mean <= 2 % of its RMS, losses within
  2 %, gradients cosine >= 0.99 and norm within 5 % (w.r.t. the fake image and w.r.t. every netD weight)
"""
import importlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import discriminator_oracle as dorc

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.bfloat16().float()


def _ops():
    import jpdse_b200  # noqa: F401
    from jpdse_b200 import ops
    return ops


def _networks():
    return importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")


def _stored(t, pad, c_pad=None):
    """fp32 NCHW cpu -> bf16 NHWC with a zero border (and zero pad channels)"""
    t = F.pad(t, (pad, pad, pad, pad))
    t = t.permute(0, 2, 3, 1).contiguous()
    if c_pad is not None and c_pad > t.shape[-1]:
        t = F.pad(t, (0, c_pad - t.shape[-1]))
    return t.bfloat16().contiguous()


def _nchw(t, pad=0, c=None):
    t = t.float().cpu()
    if pad:
        t = t[:, pad:-pad, pad:-pad]
    t = t.permute(0, 3, 1, 2)
    return t if c is None else t[:, :c]


def _cos(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


# ------------------------------------------------------------------------------------------------ 4x4 convolutions
@pytest.mark.parametrize("case", [
    ("s2", 2, 16, 32, 39, 64), ("s2", 1, 33, 65, 64, 128), ("s2", 2, 64, 128, 128, 256), ("s2", 1, 257, 513, 64, 128),
    ("s1", 2, 9, 17, 256, 512), ("s1", 1, 33, 65, 128, 128), ("s1", 1, 66, 130, 512, 1),
])
def test_conv4x4_forward(cuda, case):
    ops = _ops()
    from jpdse_b200._lib import CONV4X4_S1, CONV4X4_S2, EPI_BIAS_ACT, EPI_BIAS_NCHW, EPI_RAW_STATS
    name, B, H, W, cin_real, cout = case
    cin = (cin_real + 63) // 64 * 64
    g = torch.Generator().manual_seed(H * 7 + cout)
    x = _bf(torch.randn(B, cin_real, H, W, generator=g))
    w = _bf(torch.randn(cout, cin_real, 4, 4, generator=g) * 0.05)
    bias = torch.randn(cout, generator=g) * 0.1
    stride = 2 if name == "s2" else 1
    ref = F.conv2d(x, w, None, stride=stride, padding=2)
    kind = CONV4X4_S2 if name == "s2" else CONV4X4_S1
    xs = _stored(x, 2, cin).to(cuda)
    oh, ow = ref.shape[2:]
    scale = float(ref.abs().max())
    if cout == 1:
        cv = ops.Conv(kind, EPI_BIAS_NCHW, B, H, W, 2, cin, cin_real, cout, cuda)
        cv.pack(w.to(cuda), bias.to(cuda))
        y = torch.empty(B, 1, oh, ow, device=cuda)
        cv.forward(xs, y)
        assert float((y.cpu() - (ref + bias.view(1, -1, 1, 1))).abs().max()) <= 2e-3 * scale + 1e-5
        return
    # raw + statistics
    cv = ops.Conv(kind, EPI_RAW_STATS, B, H, W, 2, cin, cin_real, cout, cuda)
    assert cv.out_hw == (oh, ow)
    cv.pack(w.to(cuda))
    raw = torch.empty(B, oh, ow, cout, dtype=torch.bfloat16, device=cuda)
    stats = torch.zeros(B, cout, 2, dtype=torch.float64, device=cuda)
    cv.forward(xs, raw, stats)
    got = _nchw(raw)
    assert float((got - ref).abs().max()) <= 2.0 ** -7 * scale
    s1, s2 = got.double().sum(dim=(2, 3)), (got.double() ** 2).sum(dim=(2, 3))
    assert torch.allclose(stats[:, :, 0].cpu(), s1, rtol=1e-5, atol=1e-3) and torch.allclose(stats[:, :, 1].cpu(), s2, rtol=1e-5, atol=1e-3)
    # bias + LeakyReLU into the interior of a zero-bordered tensor
    cv2 = ops.Conv(kind, EPI_BIAS_ACT, B, H, W, 2, cin, cin_real, cout, cuda, out_pad=2, slope=0.2)
    cv2.pack(w.to(cuda), bias.to(cuda))
    out = ops.alloc_nhwc(B, oh + 4, ow + 4, cout, cuda)
    out.fill_(7.0)  # the border must not be touched
    cv2.forward(xs, out)
    want = F.leaky_relu(ref + bias.view(1, -1, 1, 1), 0.2)
    assert float((_nchw(out, 2) - want).abs().max()) <= 2.0 ** -7 * float(want.abs().max())
    border = out.float().cpu().clone()
    border[:, 2:-2, 2:-2] = 7.0
    assert float((border - 7.0).abs().max()) == 0.0


@pytest.mark.parametrize("case", [
    ("s2", 2, 16, 32, 39, 64), ("s2", 1, 33, 65, 64, 128), ("s2", 1, 129, 257, 128, 256), ("s2", 2, 64, 128, 64, 128),
    ("s1", 2, 9, 17, 256, 512), ("s1", 1, 33, 65, 128, 128), ("s1", 1, 66, 130, 512, 1),
])
def test_conv4x4_data_and_weight_gradients(cuda, case):
    ops = _ops()
    from jpdse_b200._lib import CONV4X4_S1, CONV4X4_S1_FULL, CONV4X4_S2, CONV4X4_S2_DGRAD, EPI_RAW, EPI_RAW_STATS
    name, B, H, W, cin_real, cout = case
    cin, cst = (cin_real + 63) // 64 * 64, (cout + 63) // 64 * 64
    g = torch.Generator().manual_seed(H * 11 + cout)
    x = _bf(torch.randn(B, cin_real, H, W, generator=g)).requires_grad_(True)
    w = _bf(torch.randn(cout, cin_real, 4, 4, generator=g) * 0.05).requires_grad_(True)
    stride = 2 if name == "s2" else 1
    y = F.conv2d(x, w, None, stride=stride, padding=2)
    dy = _bf(torch.randn(y.shape, generator=g))
    y.backward(dy)
    oh, ow = y.shape[2:]
    dys = _stored(dy, 2, cst).to(cuda)
    if name == "s2":
        dg = ops.Conv(CONV4X4_S2_DGRAD, EPI_RAW, B, oh, ow, 2, cst, cout, cin, cuda, out_hw=(H, W), cout_real=cin_real)
        fwd = ops.Conv(CONV4X4_S2, EPI_RAW_STATS if cout >= 32 else EPI_RAW, B, H, W, 2, cin, cin_real, cout, cuda)
    else:
        dg = ops.Conv(CONV4X4_S1_FULL, EPI_RAW, B, oh, ow, 2, cst, cout, cin, cuda, cout_real=cin_real)
        from jpdse_b200._lib import EPI_BIAS_NCHW
        fwd = ops.Conv(CONV4X4_S1, EPI_RAW_STATS if cout >= 32 else EPI_BIAS_NCHW, B, H, W, 2, cin, cin_real, cout, cuda)
    dg.pack(w.detach().to(cuda))
    assert dg.out_hw == (H, W)
    dx = torch.full((B, H, W, cin), 3.0, dtype=torch.bfloat16, device=cuda)
    dg.forward(dys, dx)
    got = _nchw(dx)
    assert float((got[:, :cin_real] - x.grad).abs().max()) <= 2.0 ** -7 * float(x.grad.abs().max())
    if cin > cin_real:
        assert float(got[:, cin_real:].abs().max()) == 0.0  # stored pad channels carry no gradient
    xs = _stored(x.detach(), 2, cin).to(cuda)
    dw = torch.empty(cout, cin_real, 4, 4, device=cuda)
    fwd.wgrad(xs, dys, 2, dw)
    assert float((dw.cpu() - w.grad).abs().max()) <= 2e-3 * float(w.grad.abs().max())
    fwd.wgrad(xs, dys, 2, dw, accumulate=True)
    assert float((dw.cpu() - 2 * w.grad).abs().max()) <= 4e-3 * float(w.grad.abs().max())


# ------------------------------------------------------------------------------------------------ bandwidth kernels
@pytest.mark.parametrize("H,W", [(16, 32), (33, 47), (64, 128)])
def test_d_input_and_its_backward(cuda, H, W):
    ops = _ops()
    g = torch.Generator().manual_seed(H)
    B = 2
    a = torch.randn(B, 36, H, W, generator=g)
    b = torch.randn(B, 3, H, W, generator=g)
    x = torch.cat((a, b), 1)
    for pool in (False, True):
        ref = F.avg_pool2d(x, 3, stride=2, padding=1, count_include_pad=False) if pool else x
        Ho, Wo = ref.shape[2:]
        out = ops.alloc_nhwc(B, Ho + 4, Wo + 4, 64, cuda)
        ops.d_input(a.to(cuda), b.to(cuda), out, 64, pool)
        got = out.float().cpu()
        # the pooled mean is s / count in float32 like ATen's; compare after the same bf16 rounding with 1-ulp slack
        want = _bf(ref)
        diff = (_nchw(got, 2, 39) - want).abs()
        assert float(diff.max()) <= (2.0 ** -8 * float(want.abs().max()) if pool else 0.0)
        assert float(got[:, 2:-2, 2:-2, 39:].abs().max()) == 0.0 and float(got[:, :2].abs().max()) == 0.0
    # backward: gradient w.r.t. the image channels (and w.r.t. everything) through identity + pool
    xr = x.clone().requires_grad_(True)
    pooled = F.avg_pool2d(xr, 3, stride=2, padding=1, count_include_pad=False)
    g0 = _bf(torch.randn(B, 39, H, W, generator=g))
    g1 = _bf(torch.randn(pooled.shape, generator=g))
    ((xr * g0).sum() + (pooled * g1).sum()).backward()
    g0s, g1s = _stored(g0, 0, 64).to(cuda), _stored(g1, 0, 64).to(cuda)
    got = ops.d_input_backward(g0s, g1s, 36, 3).cpu()
    assert torch.allclose(got, xr.grad[:, 36:], rtol=1e-5, atol=1e-5)
    got_all = ops.d_input_backward(g0s, g1s, 0, 39).cpu()
    assert torch.allclose(got_all, xr.grad, rtol=1e-5, atol=1e-5)
    assert torch.allclose(ops.d_input_backward(g0s, None, 36, 3).cpu(), g0[:, 36:], rtol=0, atol=0)


@pytest.mark.parametrize("H,W,ldt,idt", [(16, 32, torch.float32, torch.int32), (33, 47, torch.uint8, torch.int16)])
def test_d_input_from_ids_equals_the_float_route(cuda, H, W, ldt, idt):
    """jpdse_d_input_ids (operands straight from class / instance ids) == jpdse_d_input on the reference's float32
    cat(one-hot, edge, image) built by the oracle; both images of a [fake; real] pair in one launch."""
    from oracle import generator_oracle as orc
    ops = _ops()
    g = torch.Generator().manual_seed(W)
    B = 2
    lab = torch.randint(0, 35, (B, 1, H, W), generator=g)
    inst = torch.randint(0, 5, (B, 1, H // 4 + 1, W // 4 + 1), generator=g).repeat_interleave(4, 2).repeat_interleave(4, 3)[:, :, :H, :W]
    fake, real = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 3, H, W, generator=g)
    full = torch.from_numpy(orc.build_input(lab.numpy(), inst.numpy(), fake.numpy(), 35))
    input_label = full[:, :36].contiguous()
    for pool in (False, True):
        Ho, Wo = ((H - 1) // 2 + 1, (W - 1) // 2 + 1) if pool else (H, W)
        want_a, want_b = (ops.alloc_nhwc(B, Ho + 4, Wo + 4, 64, cuda) for _ in range(2))
        ops.d_input(input_label.to(cuda), fake.to(cuda), want_a, 64, pool)
        ops.d_input(input_label.to(cuda), real.to(cuda), want_b, 64, pool)
        got_a, got_b = (ops.alloc_nhwc(B, Ho + 4, Wo + 4, 64, cuda) for _ in range(2))
        ops.d_input_ids(lab.to(ldt).to(cuda), inst.to(idt).to(cuda), fake.to(cuda), got_a, real.to(cuda), got_b, 35, pool)
        assert torch.equal(got_a, want_a) and torch.equal(got_b, want_b)
        single = ops.alloc_nhwc(B, Ho + 4, Wo + 4, 64, cuda)
        ops.d_input_ids(lab.to(ldt).to(cuda), inst.to(idt).to(cuda), real.to(cuda), single, None, None, 35, pool)
        assert torch.equal(single, want_b)


def test_narrow_3x3_conv_for_few_input_channels(cuda):
    """JPDSE_CONV3X3_PAD1_NARROW (the VGG19's RGB conv): 8 stored channels, the 3 pixels under a filter row as one K block."""
    ops = _ops()
    from jpdse_b200._lib import CONV3X3_PAD1_NARROW, EPI_BIAS_ACT
    g = torch.Generator().manual_seed(2)
    for (B, H, W) in ((2, 16, 32), (1, 40, 136)):
        x = _bf(torch.randn(B, 3, H, W, generator=g))
        w = _bf(torch.randn(64, 3, 3, 3, generator=g) * 0.2)
        bias = torch.randn(64, generator=g) * 0.1
        want = F.relu(F.conv2d(x, w, bias, padding=1))
        xs = ops.alloc_nhwc(B, H + 2, W + 2, 8, cuda)
        ops.d_input(x.to(cuda), None, xs, 8, False, out_pad=1)
        assert float((_nchw(xs, 1, 3) - x).abs().max()) == 0.0
        cv = ops.Conv(CONV3X3_PAD1_NARROW, EPI_BIAS_ACT, B, H, W, 1, 8, 3, 64, cuda, out_pad=1, slope=0.0)
        cv.pack(w.to(cuda), bias.to(cuda))
        out = ops.alloc_nhwc(B, H + 2, W + 2, 64, cuda)
        cv.forward(xs, out)
        assert float((_nchw(out, 1) - want).abs().max()) <= 2.0 ** -7 * float(want.abs().max())
        assert float(out.float().cpu()[:, 0].abs().max()) == 0.0


@pytest.mark.parametrize("C,H,W", [(128, 17, 33), (512, 9, 18), (64, 30, 40)])
def test_instnorm_act_forward_and_backward(cuda, C, H, W):
    ops = _ops()
    g = torch.Generator().manual_seed(C + H)
    B = 2
    raw = _bf(torch.randn(B, C, H, W, generator=g) * 2 + 0.3).requires_grad_(True)
    y = F.leaky_relu(F.instance_norm(raw, eps=1e-5), 0.2)
    gy = _bf(torch.randn(B, C, H, W, generator=g))
    sk = _bf(torch.randn(B, C, H, W, generator=g))
    y.backward(gy + sk)
    raws = _stored(raw.detach(), 0).to(cuda)
    st = torch.stack((raw.detach().double().sum(dim=(2, 3)), (raw.detach().double() ** 2).sum(dim=(2, 3))), -1).to(cuda).contiguous()
    out = ops.alloc_nhwc(B, H + 4, W + 4, C, cuda)
    out.fill_(5.0)
    ops.instnorm_apply_act(raws, st, out, B, H, W, C, 2, 0.2)
    got = out.float().cpu()
    assert float((_nchw(got, 2) - y.detach()).abs().max()) <= 2.0 ** -7 * float(y.abs().max())
    ring = got.clone()
    ring[:, 2:-2, 2:-2] = 0
    assert float(ring.abs().max()) == 0.0  # zero border written
    dy = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=cuda)
    sums = torch.zeros(B, C, 2, dtype=torch.float64, device=cuda)
    ops.instnorm_backward_reduce_act(_stored(gy, 0).to(cuda), 0, _stored(sk, 0).to(cuda), raws, st, dy, sums, B, H, W, C, 0.2)
    dx = ops.alloc_nhwc(B, H + 4, W + 4, C, cuda)
    ops.instnorm_backward_apply(dy, raws, st, sums, dx, 2, B, H, W, C)
    assert float((_nchw(dx, 2) - raw.grad).abs().max()) <= 2.0 ** -6 * float(raw.grad.abs().max())


def test_act_backward_l1_pair_and_maxpool(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    B, C, H, W = 2, 64, 18, 26
    pre = _bf(torch.randn(B, C, H, W, generator=g)).requires_grad_(True)
    f = F.leaky_relu(pre, 0.2)
    gy, sk = _bf(torch.randn(B, C, H, W, generator=g)), _bf(torch.randn(B, C, H, W, generator=g))
    f.backward(gy + sk)
    fs = _stored(_bf(f.detach()), 2).to(cuda)
    d_pre = ops.alloc_nhwc(B, H + 4, W + 4, C, cuda)
    d_pre.fill_(9.0)
    db = torch.zeros(C, device=cuda)
    ops.act_backward(_stored(gy, 0).to(cuda), _stored(sk, 0).to(cuda), fs, d_pre, db, B, H, W, C, 2, 2, 0.2)
    got = _nchw(d_pre, 2)
    assert float((got - pre.grad).abs().max()) <= 2.0 ** -7 * float(pre.grad.abs().max())
    assert float(d_pre.float().cpu()[:, :2].abs().max()) == 0.0
    assert torch.allclose(db.cpu(), got.sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-3)
    # ReLU flavour without a second gradient / bias
    ops.act_backward(_stored(gy, 0).to(cuda), None, fs, d_pre, None, B, H, W, C, 2, 2, 0.0)
    want = gy * (f.detach() > 0)
    assert float((_nchw(d_pre, 2) - want).abs().max()) == 0.0
    # L1 between two stored feature maps and its gradient
    a, b = _bf(torch.randn(B, C, H, W, generator=g)), _bf(torch.randn(B, C, H, W, generator=g))
    b[:, :, 0, 0] = a[:, :, 0, 0]  # exact ties -> zero gradient
    sa, sb = _stored(a, 2).to(cuda), _stored(b, 2).to(cuda)
    acc = torch.zeros(1, dtype=torch.float64, device=cuda)
    ops.l1_pair(sa, sb, acc)
    assert abs(float(acc) - float((a - b).abs().double().sum())) <= 1e-5 * float((a - b).abs().double().sum())
    scale = torch.tensor([0.5], device=cuda)
    out = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=cuda)
    ops.l1_pair_backward(sa, sb, out, scale, 0.25, B, H, W, C, 2)
    assert float((_nchw(out) - 0.125 * torch.sign(a - b)).abs().max()) == 0.0
    # MaxPool2d(2, 2) and its backward (first maximum of the window takes the gradient, like ATen)
    x = _bf(torch.randn(B, C, H, W, generator=g))
    x[:, :, 0, 0] = x[:, :, 0, 1]  # a tie inside a window
    xr = x.clone().requires_grad_(True)
    y = F.max_pool2d(xr, 2, 2)
    gp = _bf(torch.randn(y.shape, generator=g))
    y.backward(gp)
    xs = _stored(x, 1).to(cuda)
    ys = ops.alloc_nhwc(B, H // 2 + 2, W // 2 + 2, C, cuda)
    ops.maxpool2x2(xs, ys, B, H, W, C, 1, 1)
    assert float((_nchw(ys, 1) - y.detach()).abs().max()) == 0.0 and float(ys.float().cpu()[:, 0].abs().max()) == 0.0
    dx = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=cuda)
    ops.maxpool2x2_backward(xs, _stored(gp, 0).to(cuda), dx, B, H, W, C, 1)
    assert float((_nchw(dx) - xr.grad).abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------ whole discriminator
def _setup(cuda, B, H, W, seed=7):
    nw = _networks()
    torch.manual_seed(seed)
    netD = nw.define_D(39, 64, 3, "instance", False, 2, True, gpu_ids=[])
    sd = {k: v.detach().clone() for k, v in netD.state_dict().items()}
    g = torch.Generator().manual_seed(seed + 1)
    from oracle import generator_oracle as orc
    lab_ids = torch.randint(0, 35, (B, 1, H // 8, W // 8), generator=g).repeat_interleave(8, 2).repeat_interleave(8, 3)
    inst_ids = torch.randint(0, 9, (B, 1, H // 8, W // 8), generator=g).repeat_interleave(8, 2).repeat_interleave(8, 3)
    real = torch.rand(B, 3, H, W, generator=g) - 0.5
    fake = (real + 0.1 * torch.randn(B, 3, H, W, generator=g)).clamp(-1, 1)
    # input_label = cat(one-hot, edge) exactly as the reference's preprocess makes it (pix2pixHD_model.py:376-396)
    input_label = torch.from_numpy(orc.build_input(lab_ids.numpy(), inst_ids.numpy(), real.numpy(), 35))[:, :36].contiguous()
    _setup.ids = (lab_ids.float(), inst_ids.int())
    return netD.to(cuda), sd, input_label, fake, real


@pytest.mark.parametrize("B,H,W", [(2, 64, 128), (1, 128, 256), (1, 72, 104)])
def test_discriminator_features_vs_oracle(cuda, B, H, W):
    """netD.forward (the reference's API: list over scales of the n_layers + 2 intermediate outputs, float32 NCHW)."""
    netD, sd, input_label, fake, real = _setup(cuda, B, H, W)
    x = torch.cat((input_label, fake), 1)
    with torch.no_grad():
        ref = dorc.discriminator_forward(sd, x, 3, 2)
        emu = dorc.discriminator_forward(sd, x, 3, 2, round_fn=_bf)
        got = netD(x.to(cuda))
    assert len(got) == 2 and all(len(s_) == 5 for s_ in got)
    for i in range(2):
        for j in range(5):
            a, r, e = got[i][j].cpu(), ref[i][j], emu[i][j]
            assert a.shape == r.shape and a.dtype == torch.float32
            rms = float(r.pow(2).mean().sqrt())
            assert float((a - r).abs().mean()) <= 0.02 * rms, (i, j, float((a - r).abs().mean()), rms)
            assert float((a - e).abs().mean()) <= 0.01 * rms, (i, j)  # same-arithmetic emulation: tighter


@pytest.mark.parametrize("B,H,W", [(2, 64, 128), (1, 128, 256)])
def test_discriminator_losses_and_gradients_vs_oracle(cuda, B, H, W, monkeypatch):
    """The discriminator half of get_train_loss (pix2pixHD_model.py:715-753): loss values, d(loss_G)/d(fake image) and
    d(loss_D)/d(every netD parameter), fused route and the reference's call sequence (one autograd node per netD call),
    both against the oracle's autograd (pinned bit-identical to the reference)."""
    netD, sd, input_label, fake, real = _setup(cuda, B, H, W)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    fake_o = fake.clone().requires_grad_(True)
    o_gan, o_fm, o_real, o_fake = dorc.discriminator_losses(sdg, input_label, fake_o, real, 3, 2)
    (o_gan + 10.0 * o_fm).backward()
    ref_gfake = fake_o.grad.clone()
    for v in sdg.values():
        v.grad = None
    ((o_fake + o_real) * 0.5).backward()
    ref_losses = [float(v) for v in (o_gan, o_fm, o_real, o_fake)]

    def run(fused):
        for p in netD.parameters():
            p.grad = None
        f = fake.clone().to(cuda).requires_grad_(True)
        il, re = input_label.to(cuda), real.to(cuda)
        if fused == "ids":
            l_gan, l_fm, l_real, l_fake = netD.fused_losses(None, f, re, ids=ids, num_labels=35)
        elif fused:
            l_gan, l_fm, l_real, l_fake = netD.fused_losses(il, f, re)
        else:
            crit = torch.nn.MSELoss()

            def gan(pred, val):
                return sum(crit(s_[-1], torch.full_like(s_[-1], val)) for s_ in pred)
            pfp = netD(torch.cat((il.detach(), f.detach()), 1))
            l_fake = gan(pfp, 0.0)
            pr = netD(torch.cat((il.detach(), re.detach()), 1))
            l_real = gan(pr, 1.0)
            pf = netD(torch.cat((il, f), 1))
            l_gan = gan(pf, 1.0)
            l_fm = sum(0.5 * F.l1_loss(pf[i][j], pr[i][j].detach()) for i in range(2) for j in range(4))
        (l_gan + 10.0 * l_fm).backward()
        gfake = f.grad.detach().cpu().clone()
        for p in netD.parameters():
            p.grad = None  # optimizer_D.zero_grad() (pix2pixHD_trainer.py:73)
        ((l_fake + l_real) * 0.5).backward()
        torch.cuda.synchronize()
        return [float(v) for v in (l_gan, l_fm, l_real, l_fake)], gfake, {n: p.grad.detach().cpu().clone() for n, p in netD.named_parameters()}

    ids = tuple(t.to(cuda) for t in _setup.ids)  # the class / instance ids input_label was made from
    for fused in (True, "ids", False):
        losses, gfake, pgrads = run(fused)
        for got, want in zip(losses, ref_losses):
            assert abs(got - want) <= 0.02 * abs(want) + 1e-4, (fused, losses, ref_losses)
        c = _cos(gfake, ref_gfake)
        ratio = float(gfake.norm() / ref_gfake.norm())
        assert c >= 0.99 and 0.95 <= ratio <= 1.05, "fused=%s d(loss_G)/d(fake): cosine %.5f norm ratio %.4f" % (fused, c, ratio)
        for name, gr in pgrads.items():
            ref = sdg[name].grad
            if gr.numel() == 1:
                # the output conv's bias: ONE number = mean of (prediction - target) over the map, a cancelling sum
                # that moves with the bf16 rounding of the predictions; absolute gate on the prediction scale
                assert abs(float(gr) - float(ref)) <= 0.02, (fused, name, float(gr), float(ref))
                continue
            if name.endswith(".bias") and any(name.startswith("scale%d_layer%d." % (s_, j)) for s_ in range(2) for j in (1, 2, 3)):
                assert float(gr.abs().max()) == 0.0  # bias in front of InstanceNorm: exactly zero
                continue
            c = _cos(gr, ref)
            ratio = float(gr.norm() / (ref.norm() + 1e-30))
            assert c >= 0.99 and 0.95 <= ratio <= 1.05, "fused=%s %s: cosine %.5f norm ratio %.4f" % (fused, name, c, ratio)


def test_discriminator_golden_and_state_dict(cuda, golden_dir):
    """tests/golden/discriminator_small.npz (outputs of the imported reference): seeded init equality, final maps and losses."""
    g = np.load(os.path.join(golden_dir, "discriminator_small.npz"))
    nw = _networks()
    torch.manual_seed(7)
    netD = nw.define_D(39, 64, 3, "instance", False, 2, True, gpu_ids=[])
    assert abs(float(sum(v.double().sum() for v in netD.state_dict().values())) - float(g["weight_sum"])) < 1e-6
    netD = netD.to(cuda)
    x = torch.from_numpy(g["x"]).to(cuda)
    with torch.no_grad():
        out = netD(x)
    for i, key in enumerate(("final0", "final1")):
        ref = torch.from_numpy(g[key])
        assert float((out[i][-1].cpu() - ref).abs().mean()) <= 0.03 * float(ref.pow(2).mean().sqrt())
    lab, fake, real = x[:, :36], x[:, 36:].clone().requires_grad_(True), torch.from_numpy(g["real"]).to(cuda)
    losses = [float(v) for v in netD.fused_losses(lab, fake, real)]
    for got, want in zip(losses, g["losses"]):
        assert abs(got - float(want)) <= 0.02 * abs(float(want)) + 1e-4


def test_discriminator_guards(cuda):
    import jpdse_b200
    nw = _networks()
    netD = nw.define_D(39, 64, 3, "instance", False, 2, True, gpu_ids=[0])
    with pytest.raises(jpdse_b200.JpdseError):
        netD(torch.zeros(1, 39, 32, 64))  # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        netD(torch.zeros(1, 39, 32, 64, device=cuda), keep_input=True)
    # a pass that was overwritten (more live passes than the plan keeps) must refuse its backward
    x = torch.randn(1, 39, 32, 64, device=cuda)
    first = netD(x)
    for _ in range(3):
        netD(x)
    with pytest.raises(jpdse_b200.JpdseError):
        first[0][-1].sum().backward()


def test_discriminator_and_vgg_losses_at_the_training_size(cuda):
    """BASELINE.json configs[3] image size (1024x512), batch 1: the four discriminator losses (operands built from the ids,
    fake and real as one batch) and the VGG19 loss against the fp32 CPU oracle, and d(loss_G terms)/d(fake) against the
    oracle's autograd. Feature maps here are 257x513, 129x257, 65x129, 66x130, 67x131 and 256x512 ... : the odd sizes the
    reference's PatchGAN produces at its real resolution."""
    from oracle import generator_oracle as orc
    nw = _networks()
    torch.manual_seed(7)
    netD = nw.define_D(39, 64, 3, "instance", False, 2, True, gpu_ids=[])
    sd = {k: v.detach().clone() for k, v in netD.state_dict().items()}
    torch.manual_seed(3)
    vl = nw.VGGLoss([])
    sdv = {k: v.detach().clone() for k, v in vl.vgg.state_dict().items()}
    netD, vl = netD.to(cuda), vl.to(cuda)
    import bench
    l8, i16, _u8, real = bench.synth_inputs_compact(1, 512, 1024, seed=9)
    g = torch.Generator().manual_seed(2)
    fake = (real + 0.1 * torch.randn(real.shape, generator=g)).clamp(-1, 1)
    input_label = torch.from_numpy(orc.build_input(l8.float().numpy(), i16.int().numpy(), real.numpy(), 35))[:, :36].contiguous()
    torch.set_num_threads(os.cpu_count() or 1)
    fo = fake.clone().requires_grad_(True)
    o_gan, o_fm, o_real, o_fake = dorc.discriminator_losses(sd, input_label, fo, real, 3, 2)
    o_vgg = dorc.vgg_loss(sdv, fo, real)
    (o_gan + 10.0 * o_fm + 10.0 * o_vgg).backward()
    f = fake.clone().to(cuda).requires_grad_(True)
    l_gan, l_fm, l_real, l_fake = netD.fused_losses(None, f, real.to(cuda), ids=(l8.to(cuda), i16.to(cuda)), num_labels=35)
    l_vgg = vl(f, real.to(cuda))
    (l_gan + 10.0 * l_fm + 10.0 * l_vgg).backward()
    torch.cuda.synchronize()
    for name, got, want in (("G_GAN", l_gan, o_gan), ("G_GAN_Feat", l_fm, o_fm), ("D_real", l_real, o_real), ("D_fake", l_fake, o_fake),
                            ("G_VGG", l_vgg, o_vgg)):
        print("%-10s ours %.5f oracle %.5f" % (name, float(got), float(want)))
        assert abs(float(got) - float(want)) <= 0.02 * abs(float(want)) + 1e-4, (name, float(got), float(want))
    c = _cos(f.grad.cpu(), fo.grad)
    ratio = float(f.grad.cpu().norm() / fo.grad.norm())
    print("d(G_GAN + 10 G_GAN_Feat + 10 G_VGG)/d(fake) at 1024x512: cosine %.5f norm ratio %.4f" % (c, ratio))
    assert c >= 0.97 and 0.95 <= ratio <= 1.05


@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 66, 130, 256, 512), (1, 33, 65, 128, 128), (2, 34, 66, 512, 1)])
def test_flat_tiles_and_box_shapes_equal_the_rectangular_forms(cuda, monkeypatch, B, H, W, cin, cout):
    """The PatchGAN's stride-1 convs over flat positions (JPDSE_FLAT_S1) and the per-grid weight-gradient pixel boxes
    (JPDSE_WGRAD_BOX) against the rectangular-tile / 64x1-box forms on the same operands: the forward is the same MMAs in
    the same order (bit-identical output), the weight gradient the same sum in another blocking (fp32 summation order)."""
    ops = _ops()
    from jpdse_b200._lib import CONV4X4_S1, EPI_BIAS_NCHW, EPI_RAW_STATS
    g = torch.Generator().manual_seed(W + cout)
    x = _stored(_bf(torch.randn(B, cin, H, W, generator=g)), 2, cin).to(cuda)
    w = _bf(torch.randn(cout, cin, 4, 4, generator=g) * 0.05).to(cuda)
    bias = (torch.randn(cout, generator=g) * 0.1).to(cuda)
    dy = _stored(_bf(torch.randn(B, cout, H + 1, W + 1, generator=g)), 2, (cout + 63) // 64 * 64).to(cuda)

    def run():
        if cout == 1:
            cv = ops.Conv(CONV4X4_S1, EPI_BIAS_NCHW, B, H, W, 2, cin, cin, cout, cuda)
            cv.pack(w, bias)
            y = torch.full((B, 1, H + 1, W + 1), float("nan"), device=cuda)
            cv.forward(x, y)
            st = None
        else:
            cv = ops.Conv(CONV4X4_S1, EPI_RAW_STATS, B, H, W, 2, cin, cin, cout, cuda)
            cv.pack(w)
            y = torch.full((B, H + 1, W + 1, cout), float("nan"), dtype=torch.bfloat16, device=cuda)
            st = torch.zeros(B, cout, 2, dtype=torch.float64, device=cuda)
            cv.forward(x, y, st)
        dw = torch.empty(cout, cin, 4, 4, device=cuda)
        cv.wgrad(x, dy, 2, dw)
        torch.cuda.synchronize()
        return y, st, dw

    y1, st1, dw1 = run()
    monkeypatch.setenv("JPDSE_FLAT_S1", "0")
    monkeypatch.setenv("JPDSE_WGRAD_BOX", "0")
    y0, st0, dw0 = run()
    assert not torch.isnan(y1.float()).any()
    assert torch.equal(y1, y0)
    if st1 is not None:
        assert torch.allclose(st1, st0, rtol=1e-6, atol=1e-4)
    assert float((dw1 - dw0).abs().max()) <= 1e-4 * float(dw0.abs().max())


@pytest.mark.parametrize("B,H,W", [(2, 64, 128), (1, 72, 104)])
def test_output_conv_as_1x1_gemms_equals_the_4x4_conv(cuda, monkeypatch, B, H, W):
    """The PatchGAN's 1-channel output conv as 1x1 GEMMs over the stored pixels (16 tap planes + gather; backward: scatter +
    two 1x1 GEMMs; JPDSE_PATCH_OUT_GEMM, the default) against the 4x4 implicit-GEMM conv and its gradient kinds on the same
    weights and inputs: the four losses, the gradient w.r.t. the fake image and every parameter gradient. Both forms
    multiply bf16 operands and accumulate in float32; only the summation order differs."""
    results = []
    for mode in ("1", "0"):
        monkeypatch.setenv("JPDSE_PATCH_OUT_GEMM", mode)
        netD, sd, input_label, fake, real = _setup(cuda, B, H, W, seed=21)
        f = fake.clone().to(cuda).requires_grad_(True)
        l_gan, l_fm, l_real, l_fake = netD.fused_losses(input_label.to(cuda), f, real.to(cuda))
        (l_gan + 10.0 * l_fm).backward(retain_graph=True)
        gin = f.grad.clone()
        for p in netD.parameters():
            p.grad = None
        ((l_real + l_fake) * 0.5).backward()
        torch.cuda.synchronize()
        results.append(([float(l_gan), float(l_fm), float(l_real), float(l_fake)], gin.cpu(),
                        {k: p.grad.detach().cpu().clone() for k, p in netD.named_parameters()}))
    (la, ga, pa), (lb, gb, pb) = results
    for x, y in zip(la, lb):
        assert abs(x - y) <= 2e-4 * abs(y) + 1e-6, (la, lb)
    assert _cos(ga, gb) >= 0.9995 and abs(float(ga.norm() / gb.norm()) - 1.0) <= 5e-3
    for k in pa:
        if pb[k].numel() > 1 and float(pb[k].abs().max()) > 0:
            assert _cos(pa[k], pb[k]) >= 0.999, k
