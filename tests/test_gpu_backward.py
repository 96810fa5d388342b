"""GPU parity tests of the generator BACKWARD kernels (run with -m gpu on a B200), through the C ABI.

The checker is torch CPU fp32 autograd of the reference's own ops (nn.Conv2d / ConvTranspose2d / InstanceNorm2d /
ReflectionPad2d / ReLU / Tanh: ctu/models/pix2pixHD_networks/networks.py:198-305) on the SAME bf16-rounded operands.

Tolerances
  weight gradient (bf16 operands, fp32 accumulate over up to 2^17 pixels): |err| <= 2e-3 * max|ref| (fp32 summation
      order only -- the operands are identical)
  data gradient: one bf16 rounding of the result (2^-7 relative to the output max)
  InstanceNorm backward: bf16 rounding of dy and dx (2^-6 relative to the max)
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.bfloat16().float()


def _ops():
    import jpdse_b200  # noqa: F401
    from jpdse_b200 import ops
    return ops


def _nhwc(t, pad=0, c_pad=None, mode="constant"):
    """fp32 NCHW cpu -> bf16 NHWC (optionally spatially padded / channel padded)"""
    if pad:
        t = F.pad(t, (pad, pad, pad, pad), mode=mode)
    t = t.permute(0, 2, 3, 1).contiguous()
    if c_pad is not None and c_pad > t.shape[-1]:
        t = F.pad(t, (0, c_pad - t.shape[-1]))
    return t.bfloat16().contiguous()


# ------------------------------------------------------------------------------------------------ weight gradients
@pytest.mark.parametrize("case", [
    ("conv3x3", 2, 8, 16, 64, 64), ("conv3x3", 2, 16, 16, 128, 256), ("conv3x3", 1, 32, 64, 256, 128),
    ("conv3x3", 3, 32, 64, 512, 1024),
    ("conv1x1", 2, 8, 16, 128, 128),
    ("convs2", 2, 16, 32, 64, 128), ("convs2", 1, 64, 128, 128, 256), ("convs2", 2, 32, 32, 512, 1024),
    ("convt", 2, 8, 16, 128, 64), ("convt", 1, 32, 64, 256, 128), ("convt", 2, 8, 16, 1024, 512),
    ("stem", 2, 8, 16, 40, 64), ("stem", 1, 64, 128, 40, 64), ("stem", 2, 70, 96, 40, 64),
    ("head", 2, 8, 16, 64, 3), ("head", 1, 64, 128, 64, 3), ("head", 2, 70, 100, 64, 3),
    ("conv3x3", 2, 6, 10, 64, 64), ("convs2", 2, 12, 40, 64, 128), ("convt", 1, 6, 10, 128, 64), ("conv3x3", 1, 9, 97, 128, 256),
    # "tall" items (M = 256, both accumulators of one item): the ResnetBlock shape of a batch-2 step and split-K forms
    ("conv3x3", 2, 32, 64, 1024, 1024), ("convs2", 2, 128, 256, 256, 512), ("convt", 2, 64, 128, 512, 256),
])
def test_conv_wgrad(cuda, case):
    ops = _ops()
    from jpdse_b200._lib import (CONV1X1, CONV3X3_PAD1, CONV3X3_S2, CONV7X7_PAD3, CONVT3X3_S2, EPI_BIAS_TANH_NCHW,
                                 EPI_RAW_STATS)
    kind_name, B, H, W, cin, cout = case
    g = torch.Generator().manual_seed(B * 1000 + H + cin)
    cin_real = 39 if kind_name == "stem" else cin
    x = _bf(torch.randn(B, cin_real, H, W, generator=g))
    epi, dy_pad = EPI_RAW_STATS, 0
    if kind_name == "conv3x3":
        kind, pad, mode, dy_pad = CONV3X3_PAD1, 1, "reflect", 2
        w = torch.zeros(cout, cin, 3, 3, requires_grad=True)
        y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w)
    elif kind_name == "conv1x1":
        kind, pad, mode = CONV1X1, 0, "constant"
        w = torch.zeros(cout, cin, 1, 1, requires_grad=True)
        y = F.conv2d(x, w)
    elif kind_name == "convs2":
        kind, pad, mode = CONV3X3_S2, 0, "constant"
        w = torch.zeros(cout, cin, 3, 3, requires_grad=True)
        y = F.conv2d(x, w, stride=2, padding=1)
    elif kind_name == "convt":
        kind, pad, mode = CONVT3X3_S2, 0, "constant"
        w = torch.zeros(cin, cout, 3, 3, requires_grad=True)
        y = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
    elif kind_name == "stem":
        kind, pad, mode = CONV7X7_PAD3, 3, "reflect"
        w = torch.zeros(cout, cin_real, 7, 7, requires_grad=True)
        y = F.conv2d(F.pad(x, (3, 3, 3, 3), mode="reflect"), w)
    else:
        kind, pad, mode, epi, dy_pad = CONV7X7_PAD3, 3, "reflect", EPI_BIAS_TANH_NCHW, 6
        w = torch.zeros(cout, cin, 7, 7, requires_grad=True)
        y = F.conv2d(F.pad(x, (3, 3, 3, 3), mode="reflect"), w)
    dy = _bf(torch.randn(y.shape, generator=g))
    y.backward(dy)
    ref = w.grad
    xd = _nhwc(x, pad, c_pad=cin, mode=mode).to(cuda)
    if kind_name == "stem":  # the window view over-reads a few elements behind the tensor: keep zeroed slack
        flat = torch.zeros(xd.numel() + 2048, dtype=torch.bfloat16, device=cuda)
        flat[: xd.numel()] = xd.reshape(-1)
        xd = flat[: xd.numel()].view(xd.shape)
    dyd = _nhwc(dy, dy_pad, c_pad=8 if kind_name == "head" else None).to(cuda)
    cv = ops.Conv(kind, epi, B, H, W, pad, cin, cin_real, cout, cuda)
    dw = torch.full(ref.shape, float("nan"), device=cuda)
    cv.wgrad(xd, dyd, dy_pad, dw)
    torch.cuda.synchronize()
    got = dw.cpu()
    assert not torch.isnan(got).any()
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= 2e-3 * scale, "weight gradient differs from torch autograd"
    cv.wgrad(xd, dyd, dy_pad, dw, accumulate=True)
    assert float((dw.cpu() - 2 * ref).abs().max()) <= 4e-3 * scale


# ------------------------------------------------------------------------------------------------ data gradients
@pytest.mark.parametrize("case", [("conv3x3", 2, 8, 16, 64, 64), ("conv3x3", 1, 32, 64, 256, 128), ("conv3x3", 2, 12, 20, 128, 128),
                                  ("convs2", 2, 16, 32, 64, 128), ("convs2", 1, 32, 64, 256, 512),
                                  ("convs2", 1, 6, 256, 64, 128), ("convs2", 2, 10, 512, 128, 256),  # dgrad through the fused-phase ConvT kernel
                                  ("convt", 2, 8, 16, 128, 64), ("convt", 1, 16, 32, 512, 256),
                                  ("head", 2, 8, 16, 64, 3), ("head", 1, 40, 72, 64, 3),
                                  # pixel grids off the 128-pixel tile grid (overhanging edge tiles)
                                  ("convs2", 2, 12, 40, 64, 128), ("convs2", 1, 6, 192, 64, 128),
                                  ("convt", 1, 6, 10, 128, 64), ("convt", 2, 9, 68, 256, 128)])
def test_conv_dgrad(cuda, case):
    """Gradient w.r.t. the conv input through the same implicit-GEMM kernel (role-swapped kinds)."""
    ops = _ops()
    from jpdse_b200._lib import CONV3X3_FULL, CONV3X3_S2, CONV7X7_FULL, CONVT3X3_S2, EPI_RAW
    kind_name, B, H, W, cin, cout = case
    g = torch.Generator().manual_seed(B * 77 + W + cout)
    if kind_name == "conv3x3":
        xp = torch.zeros(B, cin, H + 2, W + 2, requires_grad=True)  # gradient w.r.t. the PADDED input
        w = _bf(torch.randn(cout, cin, 3, 3, generator=g) * 0.05)
        y = F.conv2d(xp, w)
        dy = _bf(torch.randn(y.shape, generator=g))
        y.backward(dy)
        cv = ops.Conv(CONV3X3_FULL, EPI_RAW, B, H, W, 2, cout, cout, cin, cuda)
        dyd = _nhwc(dy, 2).to(cuda)
    elif kind_name == "head":
        xp = torch.zeros(B, cin, H + 6, W + 6, requires_grad=True)
        w = _bf(torch.randn(cout, cin, 7, 7, generator=g) * 0.05)
        y = F.conv2d(xp, w)
        dy = _bf(torch.randn(y.shape, generator=g))
        y.backward(dy)
        cv = ops.Conv(CONV7X7_FULL, EPI_RAW, B, H, W, 6, 8, cout, cin, cuda)
        dyd = _nhwc(dy, 6, c_pad=8).to(cuda)
    elif kind_name == "convs2":
        xp = torch.zeros(B, cin, H, W, requires_grad=True)
        w = _bf(torch.randn(cout, cin, 3, 3, generator=g) * 0.05)
        y = F.conv2d(xp, w, stride=2, padding=1)
        dy = _bf(torch.randn(y.shape, generator=g))
        y.backward(dy)
        # dgrad of a stride-2 conv == ConvTranspose forward with the same weight memory (Cin_T = cout, Cout_T = cin)
        cv = ops.Conv(CONVT3X3_S2, EPI_RAW, B, H // 2, W // 2, 0, cout, cout, cin, cuda)
        dyd = _nhwc(dy).to(cuda)
    else:
        xp = torch.zeros(B, cin, H, W, requires_grad=True)
        w = _bf(torch.randn(cin, cout, 3, 3, generator=g) * 0.05)
        y = F.conv_transpose2d(xp, w, stride=2, padding=1, output_padding=1)
        dy = _bf(torch.randn(y.shape, generator=g))
        y.backward(dy)
        # dgrad of a ConvTranspose == stride-2 conv with the same weight memory (Cout_c = cin, Cin_c = cout)
        cv = ops.Conv(CONV3X3_S2, EPI_RAW, B, 2 * H, 2 * W, 0, cout, cout, cin, cuda)
        dyd = _nhwc(dy).to(cuda)
    ref = xp.grad
    cv.pack(w.to(cuda))
    flat = torch.zeros(dyd.numel() + 2048, dtype=torch.bfloat16, device=cuda)
    flat[: dyd.numel()] = dyd.reshape(-1)
    dyd = flat[: dyd.numel()].view(dyd.shape)
    out = torch.full((B, ref.shape[2], ref.shape[3], cin), float("nan"), dtype=torch.bfloat16, device=cuda)
    cv.forward(dyd, out)
    torch.cuda.synchronize()
    got = out.float().cpu().permute(0, 3, 1, 2)
    assert not torch.isnan(got).any()
    assert float((got - ref).abs().max()) <= float(ref.abs().max()) * 2.0 ** -7


# ------------------------------------------------------------------------------------------------ InstanceNorm backward
@pytest.mark.parametrize("C,H,W,gpad,relu,skip,zpad", [(64, 16, 24, 0, True, False, 0), (64, 12, 20, 3, True, False, 0),
                                                       (1024, 8, 16, 1, True, False, 2), (1024, 8, 16, 1, False, True, 2),
                                                       (128, 9, 7, 1, True, True, 0), (256, 6, 10, 0, False, False, 2)])
def test_instnorm_backward(cuda, C, H, W, gpad, relu, skip, zpad):
    ops = _ops()
    g = torch.Generator().manual_seed(C + H + gpad)
    B = 2
    raw = _bf(torch.randn(B, C, H, W, generator=g) * 2 + 0.3).requires_grad_(True)
    y = F.instance_norm(raw, eps=1e-5)
    if relu:
        y = F.relu(y)
    yp = F.pad(y, (gpad, gpad, gpad, gpad), mode="reflect") if gpad else y
    gp = _bf(torch.randn(yp.shape, generator=g))
    sk = _bf(torch.randn(y.shape, generator=g)) if skip else None
    loss = (yp * gp).sum() + ((y * sk).sum() if skip else 0.0)
    loss.backward()
    ref_dx = raw.grad
    # expected dy (before the norm backward): fold + skip + mask
    y2 = y.detach().clone().requires_grad_(True)
    yp2 = F.pad(y2, (gpad, gpad, gpad, gpad), mode="reflect") if gpad else y2
    ((yp2 * gp).sum() + ((y2 * sk).sum() if skip else 0.0)).backward()
    ref_dy = y2.grad * ((y.detach() > 0).float() if relu else 1.0)

    rawd = _nhwc(raw.detach()).to(cuda)
    stats = torch.stack([raw.detach().double().sum(dim=(2, 3)), (raw.detach().double() ** 2).sum(dim=(2, 3))], dim=-1).contiguous().to(cuda)
    gd = _nhwc(gp).to(cuda)
    skd = _nhwc(sk).to(cuda) if skip else None
    dy = torch.full((B, H, W, C), float("nan"), dtype=torch.bfloat16, device=cuda)
    sums = torch.zeros(B, C, 2, dtype=torch.float64, device=cuda)
    ops.instnorm_backward_reduce(gd, gpad, skd, rawd, stats, dy, sums, B, H, W, C, relu)
    dx = torch.full((B, H + 2 * zpad, W + 2 * zpad, C), float("nan"), dtype=torch.bfloat16, device=cuda)
    ops.instnorm_backward_apply(dy, rawd, stats, sums, dx, zpad, B, H, W, C)
    torch.cuda.synchronize()
    got_dy = dy.float().cpu().permute(0, 3, 1, 2)
    assert float((got_dy - ref_dy).abs().max()) <= 2.0 ** -7 * max(1.0, float(ref_dy.abs().max()))
    assert torch.allclose(sums[:, :, 0].cpu(), got_dy.double().sum(dim=(2, 3)), rtol=1e-5, atol=1e-3)
    got = dx.float().cpu().permute(0, 3, 1, 2)
    assert not torch.isnan(got).any()
    if zpad:
        inner = got[:, :, zpad:-zpad, zpad:-zpad]
        assert float(got.double().abs().sum() - inner.double().abs().sum()) == 0.0  # zero border
        got = inner
    assert float((got - ref_dx).abs().max()) <= 2.0 ** -6 * max(1.0, float(ref_dx.abs().max()))


def test_tanh_backward(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(2)
    B, C, H, W = 2, 3, 20, 36
    pre = torch.randn(B, C, H, W, generator=g, requires_grad=True)
    out = torch.tanh(pre)
    go = torch.randn(B, C, H, W, generator=g)
    out.backward(go)
    d_pre = torch.full((B, H + 12, W + 12, 8), float("nan"), dtype=torch.bfloat16, device=cuda)
    dbias = torch.zeros(C, device=cuda)
    ops.tanh_backward_nchw(go.to(cuda), out.detach().to(cuda), d_pre, dbias)
    got = d_pre.float().cpu()
    assert not torch.isnan(got).any()
    inner = got[:, 6:-6, 6:-6, :3].permute(0, 3, 1, 2)
    assert float((inner - pre.grad).abs().max()) <= 2.0 ** -8 * float(pre.grad.abs().max())
    assert float(got.double().abs().sum() - inner.double().abs().sum()) == 0.0  # zero border and zero pad channels
    assert torch.allclose(dbias.cpu(), pre.grad.sum(dim=(0, 2, 3)), rtol=1e-4, atol=1e-3)


# ------------------------------------------------------------------------------------------------ whole generator
def _networks():
    import importlib
    return importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")


def _cos(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("n_down,n_blocks,B,H,W", [(2, 1, 2, 32, 64), (4, 2, 1, 128, 256), (2, 1, 2, 36, 60)])
def test_generator_gradients_vs_oracle(cuda, n_down, n_blocks, B, H, W):
    """d(loss)/d(every generator parameter) through the sm_100a backward vs CPU fp32 autograd of the oracle.

    Tolerance, calibrated like the forward's: at random init the gradient is very sensitive to the
    bf16 rounding of the FORWARD activations (ReLU masks / InstanceNorm statistics flip) -- the oracle's own
    bf16-operand emulation only reaches cosine 0.94-0.98 against its fp32 self on the deep layers (measured,
    profiles/r1_grad_cosine_table.txt), so SURVEY's 0.999 is not reachable by ANY bf16 path. Per weight tensor:
      cos(ours, fp32 reference)  >= cos(bf16 oracle, fp32 reference) - 0.02   (as close to fp32 as bf16 arithmetic gets)
      cos(ours, bf16 oracle)     >= cos(bf16 oracle, fp32 reference) - 0.005  (closer to the emulation than that is to fp32)
      the head (no ReLU / IN chaos behind it): cosine >= 0.9995; gradient norms within 5 %.
    The per-kernel tests above pin every backward kernel tightly on identical operands.
    """
    from oracle import generator_oracle as orc
    nw = _networks()
    torch.manual_seed(99)
    net = nw.define_G(39, 3, 64, "global", n_down, n_blocks, 1, 3, "instance", gpu_ids=[])
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(17)
    x = torch.randn(B, 39, H, W, generator=gen)
    target = torch.rand(B, 3, H, W, generator=gen) - 0.5

    def loss_fn(y):  # L1 distortion + a smooth term, like the reference's L1 + feature losses
        return 10.0 * (y - target).abs().mean() + (y * y).mean()

    def oracle_grads(round_fn):
        for v in sd.values():
            v.grad = None
        loss_fn(orc.generator_forward(sd, x, n_down, n_blocks, round_fn=round_fn)).backward()
        return {k: v.grad.clone() for k, v in sd.items() if v.grad is not None}

    class _RoundSTE(torch.autograd.Function):  # bf16 rounding with a straight-through gradient
        @staticmethod
        def forward(ctx, t):
            return t.bfloat16().float()

        @staticmethod
        def backward(ctx, g):
            return g

    ref32 = oracle_grads(None)
    ref16 = oracle_grads(_RoundSTE.apply)

    net = net.to(cuda).train()
    y = net(x.to(cuda))
    assert y.requires_grad
    loss_fn_dev = 10.0 * (y - target.to(cuda)).abs().mean() + (y * y).mean()
    loss_fn_dev.backward()
    torch.cuda.synchronize()
    worst16, worst32 = 1.0, 1.0
    for name, p in net.named_parameters():
        assert p.grad is not None, name
        got = p.grad.cpu()
        assert torch.isfinite(got).all(), name
        if name.endswith(".bias") and not name.startswith("model.%d." % (len(net.model) - 2)):
            assert float(got.abs().max()) == 0.0  # bias in front of InstanceNorm: exactly zero gradient
            continue
        c16, c32 = _cos(got, ref16[name]), _cos(got, ref32[name])
        cal = _cos(ref16[name], ref32[name])
        worst16, worst32 = min(worst16, c16), min(worst32, c32)
        ratio = float(got.norm() / ref16[name].norm())
        assert c32 >= cal - 0.02, "%s: cosine %.5f vs the fp32 reference path (bf16 emulation reaches %.5f)" % (name, c32, cal)
        assert c16 >= cal - 0.005, "%s: cosine %.5f vs the same-arithmetic oracle (calibration %.5f)" % (name, c16, cal)
        if name.startswith("model.%d." % (len(net.model) - 2)):
            assert c16 >= 0.9995 and c32 >= 0.9995, "head gradient cosine %.6f / %.6f" % (c16, c32)
        assert 0.95 <= ratio <= 1.05, "%s: gradient norm ratio %.3f" % (name, ratio)
    print("worst cosine: %.5f (bf16 oracle) %.5f (fp32 reference)" % (worst16, worst32))


@pytest.mark.parametrize("n_down,n_blocks,B,H,W", [(2, 1, 2, 32, 64), (4, 2, 1, 128, 256), (4, 9, 1, 64, 128)])
def test_generator_gradients_teacher_forced(cuda, n_down, n_blocks, B, H, W):
    """Whole-network gradient check at SURVEY.md 8c's stated tolerance (cosine >= 0.999 per tensor).

    The end-to-end comparison above cannot reach 0.999 with ANY bf16 forward (rounding flips ReLU masks / shifts
    InstanceNorm statistics and the difference is amplified layer by layer), so a few-percent scaling bug in one
    layer's data gradient could hide inside its calibrated gate. Here every layer is checked IN CONTEXT instead: the
    checker is CPU fp32 autograd of the reference's ops for that layer (ReflectionPad2d / Conv2d / ConvTranspose2d /
    InstanceNorm2d / ReLU / Tanh, networks.py:198-305) evaluated on OUR saved input activation and raw conv output of
    the real training forward, fed OUR incoming gradient of the real backward; its weight gradient and its gradient
    w.r.t. the layer input must match what the kernels produced. Errors cannot accumulate across layers, so the gate
    is tight: cosine >= 0.999 and norm ratio within 2 % for every weight gradient and every inter-layer gradient."""
    from jpdse_b200._lib import CONV3X3_PAD1, CONV3X3_S2, CONV7X7_PAD3, CONVT3X3_S2
    nw = _networks()
    torch.manual_seed(99)
    net = nw.define_G(39, 3, 64, "global", n_down, n_blocks, 1, 3, "instance", gpu_ids=[0]).train()
    gen = torch.Generator().manual_seed(17)
    x = torch.randn(B, 39, H, W, generator=gen).to(cuda)
    target = (torch.rand(B, 3, H, W, generator=gen) - 0.5).to(cuda)
    y = net(x)
    grad_out = []
    y.register_hook(lambda g_: grad_out.append(g_.detach().clone()))
    plan = net.plan_for(B, H, W, x.device, training=True)
    plan.capture = {}
    (10.0 * (y - target).abs().mean() + (y * y).mean()).backward()
    torch.cuda.synchronize()
    cap, layers = plan.capture, list(plan.layers)
    plan.capture = None
    params = dict(net.named_parameters())

    def nchw(t):
        return t.detach().float().permute(0, 3, 1, 2).cpu().contiguous()

    def check(name, got, ref, what):
        c = _cos(got, ref)
        ratio = float(got.double().norm() / (ref.double().norm() + 1e-30))
        assert c >= 0.999 and 0.98 <= ratio <= 1.02, "%s %s: cosine %.6f, norm ratio %.4f" % (name, what, c, ratio)
        return c

    worst = 1.0
    for si, L in enumerate(layers):
        rec = cap[L.name]
        X = nchw(L.x_in).requires_grad_(True)
        Wt = params[L.name + ".weight"].detach().cpu().clone().requires_grad_(True)
        kind, ip = L.conv.kind, L.conv.desc.in_pad
        if kind == CONV7X7_PAD3:
            conv = F.conv2d(X[:, :39], Wt)
        elif kind == CONV3X3_PAD1:
            conv = F.conv2d(X, Wt)
        else:
            Xi = X[:, :, ip:X.shape[2] - ip, ip:X.shape[3] - ip] if ip else X
            conv = F.conv2d(Xi, Wt, stride=2, padding=1) if kind == CONV3X3_S2 else \
                F.conv_transpose2d(Xi, Wt, stride=2, padding=1, output_padding=1)
        raw = nchw(L.raw)
        assert float((conv.detach() - raw).abs().max()) <= 2.0 ** -7 * float(raw.abs().max()) + 1e-6  # same forward
        yv = conv + (raw - conv).detach()  # masks / statistics from OUR stored raw output, gradient through the conv
        z = F.instance_norm(yv, eps=1e-5)
        if L.relu:
            z = F.relu(z)
        gp = rec["g_pad"]
        G = nchw(rec["g"])
        full = F.pad(z, (gp, gp, gp, gp), mode="reflect") if gp else z
        obj = (full * G).sum()
        if rec["skip"] is not None:
            obj = obj + (z * nchw(rec["skip"])).sum()
        obj.backward()
        worst = min(worst, check(L.name, params[L.name + ".weight"].grad.cpu(), Wt.grad, "weight gradient"))
        if si > 0:
            ref_in = X.grad[:, :, ip:X.shape[2] - ip, ip:X.shape[3] - ip] if (ip and kind != CONV3X3_PAD1) else X.grad
            worst = min(worst, check(L.name, nchw(rec["g_in"]), ref_in, "input gradient"))
    # head: ReflectionPad2d(3) was materialised by the producer; Conv2d(64, 3, 7) + bias + Tanh
    hn = plan.head_name
    X = nchw(plan.head_in).requires_grad_(True)
    Wt = params[hn + ".weight"].detach().cpu().clone().requires_grad_(True)
    bt = params[hn + ".bias"].detach().cpu().clone().requires_grad_(True)
    out = torch.tanh(F.conv2d(X, Wt, bt))
    (out * grad_out[0].cpu()).sum().backward()
    worst = min(worst, check(hn, params[hn + ".weight"].grad.cpu(), Wt.grad, "weight gradient"))
    worst = min(worst, check(hn, params[hn + ".bias"].grad.cpu(), bt.grad, "bias gradient"))
    worst = min(worst, check(hn, nchw(cap[hn]["g_in"]), X.grad, "input gradient"))
    print("teacher-forced worst cosine over %d layers: %.6f" % (len(layers) + 1, worst))


def test_generator_backward_guards(cuda):
    import jpdse_b200
    nw = _networks()
    net = nw.define_G(39, 3, 64, "global", 2, 1, 1, 3, "instance", gpu_ids=[0]).train()
    x = torch.randn(1, 39, 32, 64, device=cuda)
    y1 = net(x)
    net(x)  # a second forward overwrites the activations y1's backward needs
    with pytest.raises(jpdse_b200.JpdseError):
        y1.sum().backward()
    with pytest.raises(NotImplementedError):
        net(x.clone().requires_grad_(True))
    with torch.no_grad():
        assert not net(x).requires_grad


def test_trainer_step_runs_and_learns(cuda, tmp_path, monkeypatch):
    """Pix2PixHDTrainer.step (ctu/trainers/pix2pixHD_trainer.py:42-85) end to end: losses finite, all parameters of G
    and D move, and a few steps on one batch reduce the distortion."""
    import importlib
    import bench
    tr = importlib.import_module("jpd-se_b200.ctu.trainers.pix2pixHD_trainer")
    monkeypatch.setenv("JPDSE_VGG_RANDOM", "1")  # no pretrained VGG19 checkpoint offline: explicit opt-in to random weights
    opt = bench.make_opt()
    opt.is_train, opt.n_downsample_global, opt.n_blocks_global = True, 2, 2
    opt.no_vgg_loss, opt.quiet, opt.save_dir, opt.checkpoints_dir = False, True, str(tmp_path), str(tmp_path)
    torch.manual_seed(5)
    trainer = tr.Pix2PixHDTrainer(opt, mode="train")
    g = torch.Generator().manual_seed(1)
    B, H, W = 2, 64, 128
    x_dict = {"label": torch.randint(0, 35, (B, 1, H // 8, W // 8), generator=g).repeat_interleave(8, 2).repeat_interleave(8, 3).float(),
              "instance": torch.randint(0, 5, (B, 1, H // 8, W // 8), generator=g).repeat_interleave(8, 2).repeat_interleave(8, 3).int(),
              "image": torch.rand(B, 3, H, W, generator=g) - 0.5, "path": ["a", "b"]}
    before = {k: v.detach().clone() for k, v in trainer.model.state_dict().items() if "vgg" not in k}
    first = trainer.step(x_dict)
    for _ in range(7):
        last = trainer.step(x_dict)
    assert first == first and last == last  # finite
    assert last < first, "distortion did not go down (%.4f -> %.4f)" % (first, last)
    moved = [k for k, v in trainer.model.state_dict().items() if k in before and not torch.equal(v, before[k])]
    weights = [k for k in before if k.endswith(".weight")]
    assert all(k in moved for k in weights)
    assert trainer.steps_taken == 8
    ev = trainer.get_eval_loss(x_dict)
    assert ev == ev and ev >= 0
    trainer.save(0, ev)
    opt2 = bench.make_opt()
    opt2.n_downsample_global, opt2.n_blocks_global, opt2.checkpoints_dir = 2, 2, str(tmp_path)
    t2 = tr.Pix2PixHDTrainer(opt2, mode="test")
    a = trainer.get_img(x_dict)
    b = t2.get_img(x_dict)
    assert torch.equal(a, b)


def test_train_loss_reference_route_matches_fused_route(cuda, tmp_path):
    """get_train_loss through preprocess -> _get_img -> netG(input_concat) (the reference's own route, taken for the
    zero_* / compressed switches) gives the same losses as the fused input-build route, and zero_sem changes them."""
    import importlib
    import bench
    p2p = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_model")
    opt = bench.make_opt()
    opt.is_train, opt.n_downsample_global, opt.n_blocks_global, opt.quiet = True, 2, 1, True
    torch.manual_seed(3)
    model = p2p.Pix2PixHDModel(opt).cuda()
    label, inst, image = bench.synth_inputs(1, 64, 128, seed=4)
    x = {"label": label, "instance": inst, "image": image}
    with torch.no_grad():
        fused = [float(v) for v in model(dict(x), opt, mode="get_train_loss")]
        opt.use_compressed = True  # forces the reference route; the "decoded" image is the same tensor
        ref_route = [float(v) for v in model(dict(x, compressed_img=image), opt, mode="get_train_loss")]
        opt.use_compressed = False
        opt.zero_sem = True
        zeroed = [float(v) for v in model(dict(x), opt, mode="get_train_loss")]
        opt.zero_sem = False
    assert all(abs(a - b) <= 2e-3 * max(1.0, abs(a)) for a, b in zip(fused, ref_route)), (fused, ref_route)
    assert abs(zeroed[3] - fused[3]) > 1e-5  # the distortion moves when the semantics are zeroed


def test_full_size_gradient_linearity(cuda):
    """At the BASELINE size (1024x512) the CPU oracle is too slow for a backward, so use a size-independent property:
    every loss term is a batch mean and InstanceNorm is per-sample, hence the gradient of a batch of two
    equals the mean of the two single-image gradients (what data parallelism relies on).

    Every backward kernel is bit-identical for an image inside a batch and alone (tests/batch_consistency_probe.py), and
    the forward is too; what differs between the batch-2 and batch-1 plans is fp32 summation ORDER (grid shapes, split-K),
    and the InstanceNorm backward amplifies such 1e-7 differences layer by layer (it subtracts the common-mode part of
    the gradient). Measured at 1024x512: cosine 0.99994 at the stem, 1.0000000 at the up layers -> gate 0.9999."""
    nw = _networks()
    torch.manual_seed(1234)
    net = nw.define_G(39, 3, 64, "global", 4, 9, 1, 3, "instance", gpu_ids=[0]).train()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 39, 512, 1024, generator=g).to(cuda)
    tgt = (torch.rand(2, 3, 512, 1024, generator=g) - 0.5).to(cuda)

    def grads(xs, ts):
        for p in net.parameters():
            p.grad = None
        (10.0 * (net(xs) - ts).abs().mean()).backward()
        return [p.grad.clone() for p in net.parameters()]

    both = grads(x, tgt)
    g0, g1 = grads(x[:1].contiguous(), tgt[:1].contiguous()), grads(x[1:].contiguous(), tgt[1:].contiguous())
    worst, report = 1.0, []
    for (name, _), a, b0, b1 in zip(net.named_parameters(), both, g0, g1):
        m = 0.5 * (b0 + b1)
        if float(m.abs().max()) == 0.0:
            assert float(a.abs().max()) == 0.0
            continue
        c = _cos(a.cpu(), m.cpu())
        report.append("%s %.7f %.3e" % (name, c, float((a - m).abs().max() / m.abs().max())))
        worst = min(worst, c)
    print("\n".join(report))
    assert worst >= 0.9999, worst


def test_binarizer_train_mode_forward_and_backward(cuda):
    """Binarizer in train() (ctu/quantizers/binarize.py:44-65): stochastic sign given the SAME uniform draw, and the
    straight-through backward (identity through the sign, tanh', conv weight / data gradients).

    Forward: bit-exact against the oracle's soft_sign applied to tanh of OUR (bf16-operand) pre-activations with the
    noise the module drew; against the fp32 reference path only symbols whose threshold lies within the bf16 rounding
    of the pre-activation may differ (< 1 %). Backward vs torch autograd on the bf16-rounded operands: weight gradient
    2e-3 of the max, data gradient one bf16 rounding."""
    import importlib
    from oracle import quantizer_oracle as qorc
    bz = importlib.import_module("jpd-se_b200.ctu.quantizers.binarize")
    torch.manual_seed(21)
    m = bz.Binarizer(128, 64).to(cuda).train()
    g = torch.Generator().manual_seed(4)
    x = _bf(torch.randn(2, 128, 16, 32, generator=g))
    with torch.no_grad():
        m.conv.weight.copy_(_bf(m.conv.weight.cpu() * 3).to(cuda))
    xd = x.to(cuda).requires_grad_(True)
    torch.manual_seed(77)
    y = m(xd)
    torch.manual_seed(77)
    u = xd.new(2, 64, 16, 32).float().uniform_().cpu()  # the draw the module made
    w = m.conv.weight.detach().cpu()
    xr = x.clone().requires_grad_(True)
    wr = w.clone().requires_grad_(True)
    t_ref = torch.tanh(F.conv2d(xr, wr))
    y_ref = torch.from_numpy(qorc.soft_sign(t_ref.detach().numpy(), u.numpy()))
    got = y.detach().cpu()
    assert set(got.unique().tolist()) <= {-1.0, 1.0}
    assert float((got != y_ref).float().mean()) < 0.01
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy.to(cuda))
    t_ref.backward(gy)  # straight-through: d(sign)/dt = 1
    assert float((m.conv.weight.grad.cpu() - wr.grad).abs().max()) <= 1e-2 * float(wr.grad.abs().max())
    assert float((xd.grad.cpu() - xr.grad).abs().max()) <= 2.0 ** -6 * float(xr.grad.abs().max())


def test_wide_image_forward_backward_vs_oracle(cuda):
    """A 128x1024 image (strip-shaped: many column strips, few rows) through the row-stationary stem / head, the fused
    ConvTranspose kernel and their backward forms; forward and head / last-up-layer gradients against the oracle."""
    from oracle import generator_oracle as orc
    import bench
    nw = _networks()
    torch.manual_seed(31)
    net = nw.define_G(39, 3, 64, "global", 3, 1, 1, 3, "instance", gpu_ids=[])
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    label, inst, image = bench.synth_inputs(1, 128, 1024, seed=6)
    x = torch.from_numpy(orc.build_input(label.numpy(), inst.numpy(), image.numpy(), 35))
    ref = orc.generator_forward(sd, x, 3, 1)
    (10.0 * (ref - image).abs().mean()).backward()
    net = net.to(cuda).train()
    y = net.forward_from_maps(label.to(cuda), inst.to(cuda), image.to(cuda), 35)
    (10.0 * (y - image.to(cuda)).abs().mean()).backward()
    err = (y.detach().cpu() - ref.detach()).abs()
    assert float(err.mean()) <= 0.02 and orc.psnr(y.detach().cpu(), ref.detach()) >= 39.2
    names = [n for n, _ in net.named_parameters() if n.endswith(".weight")]
    for n in names[-2:]:  # head and last ConvTranspose: no chaotic depth behind them
        c = _cos(dict(net.named_parameters())[n].grad.cpu(), sd[n].grad)
        assert c >= 0.999, (n, c)
    for n, p in net.named_parameters():
        assert torch.isfinite(p.grad).all(), n


def test_weight_gradients_on_their_own_stream_equal_the_sequential_backward(cuda, monkeypatch):
    """engine.GeneratorPlan.backward runs the weight-gradient kernels on a side stream (they only depend on dx_k and a
    saved activation); with JPDSE_WGRAD_STREAM=0 they stay in line. Same kernels, same operands: the gradients must
    agree to the split-K atomics' reordering noise, three backward passes in a row (buffer reuse across calls)."""
    import copy
    nw = _networks()
    torch.manual_seed(5)
    net_a = nw.define_G(39, 3, 64, "global", 4, 2, 1, 3, "instance", gpu_ids=[])
    net_b = copy.deepcopy(net_a)
    gen = torch.Generator().manual_seed(3)
    xs = [torch.randn(2, 39, 128, 256, generator=gen).to(cuda) for _ in range(3)]
    grads = []
    for net, flag in ((net_a, "1"), (net_b, "0")):
        monkeypatch.setenv("JPDSE_WGRAD_STREAM", flag)
        net = net.to(cuda).train()
        per_pass = []
        for x in xs:
            for p in net.parameters():
                p.grad = None
            y = net(x)
            (y * y).mean().backward()
            per_pass.append({n: p.grad.clone() for n, p in net.named_parameters()})
        torch.cuda.synchronize()
        grads.append(per_pass)
    for pa, pb in zip(*grads):
        for n in pa:
            scale = float(pb[n].abs().max()) + 1e-12
            assert float((pa[n] - pb[n]).abs().max()) <= 1e-5 * scale, n


def test_training_step_side_streams_equal_in_line(cuda, tmp_path, monkeypatch):
    """The training step's side streams (VGG19 on its own stream, the discriminator's own loss backward beside the
    generator backward; Pix2PixHDModel.side_stream) only change WHEN kernels run: losses and every gradient of netG /
    netD must equal the in-line run (JPDSE_TRAIN_STREAMS=0) up to the run-to-run noise of the atomic reductions."""
    import importlib
    import bench
    tr = importlib.import_module("jpd-se_b200.ctu.trainers.pix2pixHD_trainer")
    g = torch.Generator().manual_seed(3)
    B, H, W = 2, 64, 128
    x_dict = {"label": torch.randint(0, 35, (B, 1, H // 8, W // 8), generator=g).repeat_interleave(8, 2).repeat_interleave(8, 3).float(),
              "instance": torch.randint(0, 5, (B, 1, H // 8, W // 8), generator=g).repeat_interleave(8, 2).repeat_interleave(8, 3).int(),
              "image": torch.rand(B, 3, H, W, generator=g) - 0.5, "path": ["a", "b"]}
    results = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("JPDSE_TRAIN_STREAMS", mode)
        opt = bench.make_opt()
        opt.is_train, opt.n_downsample_global, opt.n_blocks_global = True, 2, 2
        opt.quiet, opt.save_dir, opt.checkpoints_dir = True, str(tmp_path), str(tmp_path)
        torch.manual_seed(5)
        trainer = tr.Pix2PixHDTrainer(opt, mode="train")
        grads = {}
        # capture the gradients the optimizers see (hooks fire when a gradient is accumulated)
        for n, p in list(trainer.model.netG.named_parameters()) + [("D." + k, v) for k, v in trainer.model.netD.named_parameters()]:
            p.register_hook(lambda g_, n=n: grads.__setitem__(n, g_.detach().clone()))
        out = trainer.step(x_dict)
        torch.cuda.synchronize()
        results[mode] = (out, grads)
    (o0, g0), (o1, g1) = results["0"], results["1"]
    assert abs(o0 - o1) <= 1e-5 * abs(o0)
    assert set(g0) == set(g1) and len(g0) > 20
    for n in g0:
        a, b = g0[n].double().flatten(), g1[n].double().flatten()
        if float(a.norm()) == 0.0:
            assert float(b.norm()) == 0.0, n
            continue
        c = float((a * b).sum() / (a.norm() * b.norm()))
        assert c >= 0.99999 and abs(float(b.norm() / a.norm()) - 1.0) <= 1e-3, (n, c)


def test_binarizing_generator_training(cuda):
    """binarize_generator=True is the reference parser's default (networks.py:219-238): conv1x1 -> tanh -> STOCHASTIC sign in
    train() mode (ctu/quantizers/binarize.py:13-65), straight-through backward. The plan draws the uniform noise from torch's
    generator like the reference; fed the same draw, the oracle (pinned bit-identical to the reference in train() mode) must
    give the same codes wherever the comparison is not within bf16 rounding of the threshold, an output inside the bf16
    gate, and -- teacher-forced on OUR codes' decisions -- matching gradients for the Binarizer weight and the layers
    on both sides of it."""
    from oracle import generator_oracle as orc
    nw = _networks()
    torch.manual_seed(21)
    n_down, n_blocks, B, H, W = 2, 2, 2, 32, 64
    net = nw.define_G(39, 3, 64, "global", n_down, n_blocks, 1, 3, "instance", gpu_ids=[0], binarize_generator=True,
                      bin_generator_before_res=False, generator_binarizer_out_channels=128).train()
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(4)
    x = torch.randn(B, 39, H, W, generator=gen)
    target = torch.rand(B, 3, H, W, generator=gen) - 0.5
    y = net(x.to(cuda))
    (10.0 * (y - target.to(cuda)).abs().mean()).backward()
    torch.cuda.synchronize()
    plan = net.plan_for(B, H, W, x.device if x.is_cuda else torch.device("cuda", 0), training=True)
    noise = plan.bin_noise.cpu().clone()
    codes = plan.codes.cpu().clone()
    assert set(codes.unique().tolist()) <= {-1.0, 1.0}
    # oracle with the same draw, bf16-operand emulation (the codes are threshold decisions: compare like with like)
    collect = {}
    with torch.no_grad():
        orc.generator_forward(sd, x, n_down, n_blocks, round_fn=_bf, binarize=True, noise=noise, collect=collect)
    key = [k for k in collect if k.endswith(str(4 + 3 * n_down + n_blocks))][0]
    agree = float((collect[key] == codes).float().mean())
    print("stochastic codes equal to the same-arithmetic oracle's on %.3f %% of the symbols" % (100 * agree))
    assert agree >= 0.99
    # gradients: oracle autograd with OUR codes forced (y_codes = t + (ours - t).detach() keeps the straight-through path)
    sdg = {k: v.clone().requires_grad_(True) for k, v in sd.items()}

    class _Force(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t, u):
            return codes.clone()

        @staticmethod
        def backward(ctx, g_):
            return g_, None
    saved = orc._SoftSignSTE
    orc._SoftSignSTE = _Force
    try:
        yo = orc.generator_forward(sdg, x, n_down, n_blocks, binarize=True, noise=noise)
    finally:
        orc._SoftSignSTE = saved
    (10.0 * (yo - target).abs().mean()).backward()
    err = (y.detach().cpu() - yo.detach()).abs()
    assert float(err.mean()) <= 0.02 and float(err.max()) <= 0.15
    bname = "model.%d.conv.weight" % (4 + 3 * n_down + n_blocks)
    checked = 0
    for name, p in net.named_parameters():
        if name.endswith(".bias") and not name.startswith("model.%d." % (len(net.model) - 2)):
            continue
        got, ref = p.grad.cpu(), sdg[name].grad
        c = _cos(got, ref)
        ratio = float(got.norm() / (ref.norm() + 1e-30))
        # behind the binarizer (decoder side) the forward is identical given the codes: tight; the Binarizer weight and the
        # encoder side see the tanh' / ReLU-mask / statistics differences of a bf16 forward: calibrated like the generator's
        tight = name == bname or int(name.split(".")[1]) > 4 + 3 * n_down + n_blocks
        assert c >= (0.99 if tight else 0.90) and 0.9 <= ratio <= 1.1, "%s: cosine %.5f norm ratio %.4f" % (name, c, ratio)
        checked += 1
    assert checked >= 10 and dict(net.named_parameters())[bname].grad.abs().sum() > 0
    # eval() + autograd is refused, eval() inference still gives the deterministic sign
    net.eval()
    with pytest.raises(NotImplementedError):
        net(x.to(cuda))
    with torch.no_grad():
        assert net(x.to(cuda)).shape == (B, 3, H, W)


@pytest.mark.parametrize("C,H,W,gpad,relu,use_skip,want_dy", [(1024, 32, 64, 1, True, False, False), (1024, 32, 64, 1, False, True, True),
                                                          (256, 20, 36, 0, True, False, True), (64, 9, 13, 3, True, True, False)])
def test_fused_instnorm_backward_equals_reduce_plus_apply(cuda, C, H, W, gpad, relu, use_skip, want_dy):
    """jpdse_instnorm_backward_fused (one CTA per (image, 8 channels), small maps) against the reduce + apply pair and
    against torch autograd of InstanceNorm2d (+ReLU) (+ReflectionPad2d fold-back) on the same operands."""
    ops = _ops()
    g = torch.Generator().manual_seed(C + H)
    B, z = 2, 2
    raw = _bf(torch.randn(B, C, H, W, generator=g) * 1.5 + 0.2).requires_grad_(True)
    y = F.instance_norm(raw, eps=1e-5)
    if relu:
        y = F.relu(y)
    yp = F.pad(y, (gpad, gpad, gpad, gpad), mode="reflect") if gpad else y
    gy = _bf(torch.randn(yp.shape, generator=g))
    sk = _bf(torch.randn(B, C, H, W, generator=g)) if use_skip else None
    ((yp * gy).sum() + ((y * sk).sum() if use_skip else 0.0)).backward()
    raws = _nhwc(raw.detach()).to(cuda)
    st = torch.stack((raw.detach().double().sum(dim=(2, 3)), (raw.detach().double() ** 2).sum(dim=(2, 3))), -1).to(cuda).contiguous()
    gs = _nhwc(gy).to(cuda)
    sks = _nhwc(sk).to(cuda) if use_skip else None
    dx_f = torch.full((B, H + 2 * z, W + 2 * z, C), 5.0, dtype=torch.bfloat16, device=cuda)
    dy_f = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=cuda) if want_dy else None
    ops.instnorm_backward_fused(gs, gpad, sks, raws, st, dy_f, dx_f, z, B, H, W, C, relu)
    dy_p = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=cuda)
    sums = torch.zeros(B, C, 2, dtype=torch.float64, device=cuda)
    ops.instnorm_backward_reduce(gs, gpad, sks, raws, st, dy_p, sums, B, H, W, C, relu)
    dx_p = torch.empty_like(dx_f)
    ops.instnorm_backward_apply(dy_p, raws, st, sums, dx_p, z, B, H, W, C)
    if want_dy:
        assert torch.equal(dy_f, dy_p)  # the masked gradient is the same arithmetic
    scale = float(raw.grad.abs().max())
    inner = dx_f.float().cpu()[:, z:-z, z:-z].permute(0, 3, 1, 2)
    assert float((inner - raw.grad).abs().max()) <= 2.0 ** -6 * scale
    assert float((dx_f.float() - dx_p.float()).abs().max()) <= 2.0 ** -7 * scale
    ring = dx_f.float().cpu().clone()
    ring[:, z:-z, z:-z] = 0
    assert float(ring.abs().max()) == 0.0  # zero border written


def _to_shared(t_nhwc, p):
    """(B,H,W,C) -> the JPDSE_PAD_SHARED layout as a flat (positions, C) tensor: pitch W+p, image stride (H+p) rows."""
    B, H, W, C = t_nhwc.shape
    P, S = W + p, (H + p) * (W + p)
    flat = torch.zeros(B * S + p * P + p + 256, C, dtype=t_nhwc.dtype, device=t_nhwc.device)
    for b in range(B):
        rows = flat[b * S + p * P + p: b * S + p * P + p + H * P].view(H, P, C)
        rows[:, :W] = t_nhwc[b]
    return flat


@pytest.mark.parametrize("B,H,W,C", [(2, 32, 64, 1024), (1, 8, 16, 64), (3, 12, 20, 128), (4, 32, 64, 256)])
def test_shared_border_gradient_layout(cuda, B, H, W, C):
    """JPDSE_PAD_SHARED / JPDSE_CONV3X3_FULL_SHARED (the ResnetBlock gradients of the backward walk): the InstanceNorm
    backward kernels write the layout (every stored position, zeros included), and the data-gradient conv and the weight
    gradient read it -- each BIT-IDENTICAL to the per-image bordered form on the same operands."""
    ops = _ops()
    from jpdse_b200._lib import CONV3X3_FULL, CONV3X3_FULL_SHARED, CONV3X3_PAD1, EPI_RAW, EPI_RAW_STATS, PAD_SHARED
    g = torch.Generator().manual_seed(B * 1000 + C + W)
    z = 2
    P, S = W + z, (H + z) * (W + z)
    stored = B * S + z * P + z
    # ---- producers
    raw = _bf(torch.randn(B, C, H, W, generator=g) * 1.5 + 0.2)
    raws = _nhwc(raw).to(cuda)
    st = torch.stack((raw.double().sum(dim=(2, 3)), (raw.double() ** 2).sum(dim=(2, 3))), -1).to(cuda).contiguous()
    gs = _nhwc(_bf(torch.randn(B, C, H + 2, W + 2, generator=g))).to(cuda)
    dy = torch.empty(B, H, W, C, dtype=torch.bfloat16, device=cuda)
    sums = torch.zeros(B, C, 2, dtype=torch.float64, device=cuda)
    ops.instnorm_backward_reduce(gs, 1, None, raws, st, dy, sums, B, H, W, C, True)
    dx_ref = torch.full((B, H + 2 * z, W + 2 * z, C), 5.0, dtype=torch.bfloat16, device=cuda)
    ops.instnorm_backward_apply(dy, raws, st, sums, dx_ref, z, B, H, W, C)
    want = _to_shared(dx_ref[:, z:-z, z:-z], z)
    dx_sh = torch.full_like(want, 5.0)
    ops.instnorm_backward_apply(dy, raws, st, sums, dx_sh, z | PAD_SHARED, B, H, W, C)
    assert torch.equal(dx_sh[:stored], want[:stored])
    assert float((dx_sh[stored:].float() - 5.0).abs().max()) == 0.0  # nothing written behind the layout
    if H * W <= ops.FUSED_NORM_BACKWARD_MAX_PIXELS:
        dx_ref_f = torch.full_like(dx_ref, 5.0)
        ops.instnorm_backward_fused(gs, 1, None, raws, st, None, dx_ref_f, z, B, H, W, C, True)
        dx_sh_f = torch.full_like(want, 5.0)
        ops.instnorm_backward_fused(gs, 1, None, raws, st, None, dx_sh_f, z | PAD_SHARED, B, H, W, C, True)
        assert torch.equal(dx_sh_f[:stored], _to_shared(dx_ref_f[:, z:-z, z:-z], z)[:stored])
        assert float((dx_sh_f[stored:].float() - 5.0).abs().max()) == 0.0
    # ---- consumers
    w = _bf(torch.randn(C, C, 3, 3, generator=g) * 0.05).to(cuda)
    flat = torch.zeros(dx_ref.numel() + 2048, dtype=torch.bfloat16, device=cuda)
    flat[: dx_ref.numel()] = dx_ref.reshape(-1)
    dx_b = flat[: dx_ref.numel()].view(dx_ref.shape)
    full = ops.Conv(CONV3X3_FULL, EPI_RAW, B, H, W, 2, C, C, C, cuda)
    full.pack(w)
    out_ref = torch.full((B, H + 2, W + 2, C), float("nan"), dtype=torch.bfloat16, device=cuda)
    full.forward(dx_b, out_ref)
    shared = ops.Conv(CONV3X3_FULL_SHARED, EPI_RAW, B, H, W, 2, C, C, C, cuda)
    shared.pack(w)
    out_sh = torch.full((B * (H + 2) * (W + 2) + 64, C), float("nan"), dtype=torch.bfloat16, device=cuda)
    shared.forward(dx_sh, out_sh)
    torch.cuda.synchronize()
    assert torch.equal(out_sh[: B * (H + 2) * (W + 2)].view(out_ref.shape), out_ref)
    assert torch.isnan(out_sh[B * (H + 2) * (W + 2):].float()).all()  # no store behind the last output pixel
    fwd = ops.Conv(CONV3X3_PAD1, EPI_RAW_STATS, B, H, W, 1, C, C, C, cuda)
    x_in = _nhwc(_bf(torch.randn(B, C, H + 2, W + 2, generator=g))).to(cuda)
    xf = torch.zeros(x_in.numel() + 2048, dtype=torch.bfloat16, device=cuda)
    xf[: x_in.numel()] = x_in.reshape(-1)
    x_in = xf[: x_in.numel()].view(x_in.shape)
    dw_ref = torch.empty(C, C, 3, 3, device=cuda)
    dw_sh = torch.empty(C, C, 3, 3, device=cuda)
    fwd.wgrad(x_in, dx_b, z, dw_ref)
    fwd.wgrad(x_in, dx_sh, z | PAD_SHARED, dw_sh)
    torch.cuda.synchronize()
    # (split-K shapes add their partial sums with fp32 atomics in arrival order: equal up to fp32 summation order)
    assert float((dw_sh - dw_ref).abs().max()) <= 1e-5 * float(dw_ref.abs().max())
