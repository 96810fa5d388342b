"""GPU parity tests (run with -m gpu on a B200). Every call goes through the C ABI of libjpdse_b200.so and is
compared with the CPU oracle / the golden vectors generated from the reference.

Tolerances
  integer / index outputs (one-hot, edges, concat, round, sign, soft-sign, S2HVQ indices): bit-exact
  single conv (bf16 operands, fp32 accumulate) vs torch fp32 conv of the same bf16-rounded operands:
      |err| <= 1 bf16 ulp of the output scale (2^-7 relative to the output max)
  generator (bf16 kernels) vs reference fp32: mean-abs <= 0.02, max-abs <= 0.15, PSNR >= 39.2 dB
      (= reference-autocast-bf16 PSNR 40.2 dB - 1 dB)
"""
import importlib
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import generator_oracle as orc
from oracle import quantizer_oracle as qorc

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.bfloat16().float()


def _ops():
    import jpdse_b200  # noqa: F401
    from jpdse_b200 import ops
    return ops


def _networks():
    return importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")


# ------------------------------------------------------------------------------------------------ input build
def test_build_input_golden(cuda, golden_dir):
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    L = int(g["num_labels"])
    lab, inst, img = (torch.from_numpy(g[k]).to(cuda) for k in ("label", "instance", "image"))
    nhwc, nchw = ops.build_input(lab, inst, img, L, pad=3, c_pad=40, nhwc=True, nchw=True)
    assert np.array_equal(nchw.cpu().numpy(), g["input_concat"])  # bit-exact vs the reference's input_concat
    want = _bf(torch.from_numpy(orc.reflect_pad_nhwc(g["input_concat"], 3, 40))).numpy()
    assert np.array_equal(nhwc.float().cpu().numpy(), want)


@pytest.mark.parametrize("ldt,idt", [(torch.uint8, torch.int16), (torch.int64, torch.int64), (torch.float32, torch.float32)])
def test_build_input_dtypes_and_ragged_shapes(cuda, ldt, idt):
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    B, H, W, L = 3, 17, 29, 35  # odd sizes, not a multiple of the 256-pixel block
    lab = torch.randint(0, L, (B, 1, H, W), generator=g)
    inst = torch.randint(0, 4, (B, 1, H, W), generator=g)
    img = torch.rand(B, 3, H, W, generator=g) - 0.5
    ref = orc.build_input(lab.numpy(), inst.numpy(), img.numpy(), L)
    nhwc, nchw = ops.build_input(lab.to(ldt).to(cuda), inst.to(idt).to(cuda), img.to(cuda), L, pad=3, nhwc=True, nchw=True)
    assert np.array_equal(nchw.cpu().numpy(), ref)
    assert np.array_equal(nhwc.float().cpu().numpy(), _bf(torch.from_numpy(orc.reflect_pad_nhwc(ref, 3, 40))).numpy())


def test_build_input_counts_out_of_range_labels(cuda):
    ops = _ops()
    lab = torch.zeros(1, 1, 8, 8)
    lab[0, 0, 0, 0], lab[0, 0, 3, 3], lab[0, 0, 7, 7] = 35.0, -1.0, float("nan")  # scatter_ would raise on these
    bad = torch.zeros(1, dtype=torch.int32, device=cuda)
    _, nchw = ops.build_input(lab.to(cuda), torch.zeros(1, 1, 8, 8, dtype=torch.int32, device=cuda),
                              torch.zeros(1, 3, 8, 8, device=cuda), 35, nhwc=False, nchw=True, bad_count=bad)
    assert int(bad.item()) == 3
    assert float(nchw[0, :35, 0, 0].sum()) == 0.0 and float(nchw[0, :35, 1, 1].sum()) == 1.0


def test_get_edges_and_preprocess_mirror(cuda, golden_dir):
    p2p = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_model")
    import bench
    g = np.load(os.path.join(golden_dir, "preprocess.npz"))
    opt = bench.make_opt()
    opt.n_downsample_global, opt.n_blocks_global = 1, 0
    model = p2p.Pix2PixHDModel(opt)
    e = model.get_edges(torch.from_numpy(g["instance"]))
    assert np.array_equal(e.cpu().numpy(), g["input_concat"][:, 35:36])
    pre = model.preprocess({k: torch.from_numpy(g[k]) for k in ("label", "instance", "image")})
    assert np.array_equal(pre["input_label"].cpu().numpy(), g["input_concat"][:, :36])
    with pytest.raises(RuntimeError):
        bad = torch.from_numpy(g["label"]).clone()
        bad[0, 0, 0, 0] = 99
        model.preprocess({"label": bad, "instance": torch.from_numpy(g["instance"]), "image": torch.from_numpy(g["image"])})


# ------------------------------------------------------------------------------------------------ convolutions
def _conv_case(ops, cuda, kind_name, B, H, W, cin, cout, cin_real=None, seed=0):
    from jpdse_b200._lib import CONV1X1, CONV3X3_PAD1, CONV3X3_S2, CONV7X7_PAD3, CONVT3X3_S2, EPI_RAW_STATS
    g = torch.Generator().manual_seed(seed)
    cin_real = cin if cin_real is None else cin_real
    x = _bf(torch.randn(B, cin_real, H, W, generator=g))
    if kind_name == "conv1x1":
        kind, pad = CONV1X1, 0
        w = _bf(torch.randn(cout, cin_real, 1, 1, generator=g) * 0.05)
        ref = F.conv2d(x, w)
    elif kind_name == "conv3x3":
        kind, pad = CONV3X3_PAD1, 1
        w = _bf(torch.randn(cout, cin_real, 3, 3, generator=g) * 0.05)
        ref = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w)
    elif kind_name == "convs2":
        kind, pad = CONV3X3_S2, 0
        w = _bf(torch.randn(cout, cin_real, 3, 3, generator=g) * 0.05)
        ref = F.conv2d(x, w, stride=2, padding=1)
    elif kind_name == "convt":
        kind, pad = CONVT3X3_S2, 0
        w = _bf(torch.randn(cin_real, cout, 3, 3, generator=g) * 0.05)
        ref = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
    else:
        kind, pad = CONV7X7_PAD3, 3
        w = _bf(torch.randn(cout, cin_real, 7, 7, generator=g) * 0.05)
        ref = F.conv2d(F.pad(x, (3, 3, 3, 3), mode="reflect"), w)
    xd = ops.nchw_to_nhwc_bf16(x.to(cuda), pad_reflect=pad, c_pad=cin)
    cv = ops.Conv(kind, EPI_RAW_STATS, B, H, W, pad, cin, cin_real, cout, cuda)
    cv.pack(w.to(cuda))
    oh, ow = cv.out_hw
    y = torch.full((B, oh, ow, cout), float("nan"), dtype=torch.bfloat16, device=cuda)
    stats = torch.zeros(B, cout, 2, dtype=torch.float64, device=cuda)
    cv.forward(xd, y, stats)
    torch.cuda.synchronize()
    got = y.float().cpu().permute(0, 3, 1, 2)
    assert not torch.isnan(got).any()
    scale = float(ref.abs().max())
    assert float((got - ref).abs().max()) <= scale * 2.0 ** -7, "conv output beyond one bf16 ulp of the output scale"
    # statistics are those of the bf16-rounded tensor the kernel stored
    s1 = got.double().sum(dim=(2, 3))
    s2 = (got.double() ** 2).sum(dim=(2, 3))
    assert torch.allclose(stats[:, :, 0].cpu(), s1, rtol=1e-5, atol=1e-3)
    assert torch.allclose(stats[:, :, 1].cpu(), s2, rtol=1e-5, atol=1e-3)


@pytest.mark.parametrize("case", [
    ("conv1x1", 2, 8, 16, 64, 64), ("conv1x1", 1, 16, 32, 128, 256),
    ("conv3x3", 2, 8, 16, 64, 64), ("conv3x3", 2, 16, 16, 128, 256), ("conv3x3", 1, 32, 64, 1024, 1024),
    ("convs2", 2, 16, 32, 64, 128), ("convs2", 1, 32, 64, 128, 256),
    ("convt", 2, 8, 16, 128, 64), ("convt", 1, 16, 32, 256, 128),
    ("convt", 2, 5, 128, 128, 64), ("convt", 1, 9, 256, 256, 128), ("convt", 2, 3, 128, 64, 64),  # fused-phase kernel (W % 128 == 0, Cout <= 128)
    ("conv7x7", 2, 8, 16, 40, 64, 39), ("conv7x7", 1, 16, 128, 40, 64, 39),
    ("conv7x7", 2, 140, 256, 40, 64, 39),  # row-stationary stem path: 2 strips, a full and a ragged row chunk
    # pixel grids that are not a whole number of 128-pixel tiles (overhanging edge tiles: TMA zero fill / clipped
    # stores, masked statistics) -- any H, W the reference accepts
    ("conv1x1", 1, 5, 7, 64, 64), ("conv3x3", 2, 6, 10, 64, 64), ("conv3x3", 1, 24, 48, 128, 256),
    ("conv3x3", 1, 5, 192, 64, 128), ("conv3x3", 2, 3, 5, 1024, 1024),
    ("convs2", 2, 12, 40, 64, 128), ("convs2", 1, 10, 400, 128, 256),
    ("convt", 2, 6, 10, 128, 64), ("convt", 1, 7, 192, 256, 128), ("convt", 1, 3, 5, 1024, 512),
    ("conv7x7", 2, 9, 20, 40, 64, 39), ("conv7x7", 1, 6, 200, 40, 64, 39),
])
def test_conv_kinds(cuda, case):
    _conv_case(_ops(), cuda, *case, seed=len(case) + case[2])


@pytest.mark.parametrize("shape", [(2, 16, 64), (2, 16, 128), (2, 140, 256),  # generic path / row path / ragged chunks
                                   (2, 9, 20), (1, 6, 100), (1, 5, 200)])       # overhanging tiles / ragged last strip
def test_head_conv_bias_tanh(cuda, shape):
    ops = _ops()
    from jpdse_b200._lib import CONV7X7_PAD3, EPI_BIAS_TANH_NCHW
    g = torch.Generator().manual_seed(11)
    B, H, W = shape
    x = _bf(torch.randn(B, 64, H, W, generator=g))
    w = _bf(torch.randn(3, 64, 7, 7, generator=g) * 0.02)
    bias = torch.randn(3, generator=g) * 0.1
    ref = torch.tanh(F.conv2d(F.pad(x, (3, 3, 3, 3), mode="reflect"), w, bias))
    cv = ops.Conv(CONV7X7_PAD3, EPI_BIAS_TANH_NCHW, B, H, W, 3, 64, 64, 3, cuda)
    cv.pack(w.to(cuda), bias.to(cuda))
    y = torch.full((B, 3, H, W), float("nan"), device=cuda)
    cv.forward(ops.nchw_to_nhwc_bf16(x.to(cuda), pad_reflect=3), y)
    assert float((y.cpu() - ref).abs().max()) < 1e-4


@pytest.mark.parametrize("kind_name,cin,cout", [("pad1", 64, 64), ("pad1", 1024, 1024), ("pad1", 128, 256), ("s2", 64, 128),
                                                ("s2", 512, 1024), ("full", 64, 64), ("full", 1024, 1024), ("full", 256, 128)])
def test_coalesced_weight_packers_equal_generic_gather(cuda, monkeypatch, kind_name, cin, cout):
    """The shared-memory-staged packers of the 3x3 layouts write the same bytes as the one-element-per-thread kernel."""
    ops = _ops()
    from jpdse_b200._lib import CONV3X3_FULL, CONV3X3_PAD1, CONV3X3_S2, EPI_RAW, EPI_RAW_STATS
    kind, pad, epi = {"pad1": (CONV3X3_PAD1, 1, EPI_RAW_STATS), "s2": (CONV3X3_S2, 0, EPI_RAW_STATS),
                      "full": (CONV3X3_FULL, 2, EPI_RAW)}[kind_name]
    shape = (cin, cout, 3, 3) if kind_name == "full" else (cout, cin, 3, 3)
    w = torch.randn(shape, generator=torch.Generator().manual_seed(cin + cout)).to(cuda)
    cv = ops.Conv(kind, epi, 1, 16, 32, pad, cin, cin, cout, cuda)
    cv.w_packed.fill_(float("nan"))
    cv.pack(w)
    fast = cv.w_packed.clone()
    monkeypatch.setenv("JPDSE_GENERIC_PACK", "1")
    cv.w_packed.fill_(float("nan"))
    cv.pack(w)
    torch.cuda.synchronize()
    assert not torch.isnan(fast.float()).any()
    assert torch.equal(fast.view(torch.int16), cv.w_packed.view(torch.int16))


def test_conv_rejects_bad_arguments(cuda):
    import jpdse_b200
    ops = _ops()
    from jpdse_b200._lib import CONV3X3_PAD1, EPI_RAW_STATS
    with pytest.raises(jpdse_b200.JpdseError):
        ops.Conv(CONV3X3_PAD1, EPI_RAW_STATS, 1, 8, 16, 1, 100, 100, 64, cuda)  # cin not a multiple of 64
    with pytest.raises(jpdse_b200.JpdseError):
        ops.Conv(CONV3X3_PAD1, EPI_RAW_STATS, 1, 8, 16, 1, 64, 64, 96, cuda)   # raw+stats output not a multiple of the N tile
    from jpdse_b200._lib import CONV3X3_S2
    with pytest.raises(jpdse_b200.JpdseError):
        ops.Conv(CONV3X3_S2, EPI_RAW_STATS, 1, 7, 16, 0, 64, 64, 64, cuda)     # odd height under a stride-2 conv


# ------------------------------------------------------------------------------------------------ InstanceNorm apply
@pytest.mark.parametrize("C,H,W,pad,relu,res", [(64, 16, 24, 0, True, False), (64, 12, 20, 3, True, False),
                                                (1024, 8, 16, 1, True, False), (1024, 8, 16, 1, False, True),
                                                (128, 9, 7, 1, False, False)])
def test_instnorm_apply(cuda, C, H, W, pad, relu, res):
    ops = _ops()
    g = torch.Generator().manual_seed(C + H)
    B = 2
    raw = _bf(torch.randn(B, C, H, W, generator=g) * 3 + 0.7)
    resid = _bf(torch.randn(B, C, H, W, generator=g)) if res else None
    y = F.instance_norm(raw, eps=1e-5)
    if relu:
        y = F.relu(y)
    if res:
        y = y + resid
    if pad:
        y = F.pad(y, (pad, pad, pad, pad), mode="reflect")
    raw_d = raw.permute(0, 2, 3, 1).contiguous().bfloat16().to(cuda)
    stats = torch.stack([raw.double().sum(dim=(2, 3)), (raw.double() ** 2).sum(dim=(2, 3))], dim=-1).contiguous().to(cuda)
    res_d = None
    if res:
        res_d = F.pad(resid, (pad, pad, pad, pad), mode="reflect").permute(0, 2, 3, 1).contiguous().bfloat16().to(cuda)
    out = torch.full((B, H + 2 * pad, W + 2 * pad, C), float("nan"), dtype=torch.bfloat16, device=cuda)
    ops.instnorm_apply(raw_d, stats, out, B, H, W, C, pad, relu, residual=res_d)
    got = out.float().cpu().permute(0, 3, 1, 2)
    assert not torch.isnan(got).any()
    assert float((got - y).abs().max()) <= 2.0 ** -7 * max(1.0, float(y.abs().max()))  # one bf16 rounding of the result


# ------------------------------------------------------------------------------------------------ generator
def _golden_net(golden_dir):
    g = np.load(os.path.join(golden_dir, "generator_small.npz"))
    nw = _networks()
    torch.manual_seed(int(g["seed"]))
    net = nw.define_G(39, 3, 64, "global", int(g["n_down"]), int(g["n_blocks"]), 1, 3, "instance", gpu_ids=[])
    return g, net


def test_generator_vs_oracle_and_golden_weights(cuda, golden_dir):
    g, net = _golden_net(golden_dir)  # same seeded weights as the golden file (checked on CPU by test_oracle_golden)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(5)
    x = torch.randn(2, 39, 128, 256, generator=gen)  # smallest image the 128-pixel tiling admits at 4 downsamplings
    with torch.no_grad():
        ref = orc.generator_forward(sd, x, int(g["n_down"]), int(g["n_blocks"]))
        emu = orc.generator_forward(sd, x, int(g["n_down"]), int(g["n_blocks"]), round_fn=_bf)
        y = net.to(cuda).eval()(x.to(cuda)).cpu()
    err = (y - ref).abs()
    assert float(err.mean()) <= 0.02 and float(err.max()) <= 0.15 and orc.psnr(y, ref) >= 39.2
    # against the same-arithmetic model (bf16 operands, fp32 accumulate) the agreement is much tighter
    assert float((y - emu).abs().mean()) <= 0.008


def test_generator_full_architecture(cuda):
    nw = _networks()
    torch.manual_seed(1234)
    net = nw.define_G(39, 3, 64, "global", 4, 9, 1, 3, "instance", gpu_ids=[]).eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    gen = torch.Generator().manual_seed(6)
    B, H, W = 2, 128, 256
    label = torch.randint(0, 35, (B, 1, H // 16, W // 16), generator=gen).repeat_interleave(16, 2).repeat_interleave(16, 3).float()
    inst = torch.randint(0, 9, (B, 1, H // 16, W // 16), generator=gen).repeat_interleave(16, 2).repeat_interleave(16, 3).int()
    image = torch.rand(B, 3, H, W, generator=gen) - 0.5
    x = torch.from_numpy(orc.build_input(label.numpy(), inst.numpy(), image.numpy(), 35))
    with torch.no_grad():
        ref = orc.generator_forward(sd, x, 4, 9)
        net = net.to(cuda)
        y_maps = net.forward_from_maps(label.to(cuda), inst.to(cuda), image.to(cuda), 35).cpu()
        y_nchw = net(x.to(cuda)).cpu()
    assert torch.equal(y_maps, y_nchw) or float((y_maps - y_nchw).abs().max()) < 2e-3  # fused input build == concat path
    err = (y_maps - ref).abs()
    assert float(err.mean()) <= 0.02 and float(err.max()) <= 0.15 and orc.psnr(y_maps, ref) >= 39.2
    assert float(y_maps.abs().max()) < 1.0  # tanh range


@pytest.mark.parametrize("H,W", [(512, 1024), (1024, 2048)])
def test_generator_vs_oracle_at_baseline_sizes(cuda, H, W):
    """BASELINE.json configs[1] / configs[2] image sizes (1024x512 and 2048x1024), full 4-down / 9-block generator with
    the bench's weights (seed 1234) and synthetic inputs, batch 1: the sm_100a path against the fp32 CPU oracle
    (oracle/generator_oracle.py, pinned bit-identical to the reference) under the stated bf16 gate. The image is also
    run as image 1 of a batch of 2 (the batch the timed plans run is a multiple of this; InstanceNorm is per sample) and
    must come out bit-identical, which carries the oracle check over to every image slot of a batch."""
    import bench
    nw = _networks()
    torch.manual_seed(1234)
    net = nw.define_G(39, 3, 64, "global", 4, 9, 1, 3, "instance", gpu_ids=[]).eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    label, inst, image = bench.synth_inputs(2, H, W, seed=1234)
    x = torch.from_numpy(orc.build_input(label[1:].numpy(), inst[1:].numpy(), image[1:].numpy(), 35))
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = orc.generator_forward(sd, x, 4, 9)
        net = net.to(cuda)
        one = net.forward_from_maps(label[1:].to(cuda), inst[1:].to(cuda), image[1:].to(cuda), 35).cpu()
        two = net.forward_from_maps(label.to(cuda), inst.to(cuda), image.to(cuda), 35).cpu()
    assert torch.equal(two[1:], one)
    err = (one - ref).abs()
    p = orc.psnr(one, ref)
    print("%dx%d vs fp32 oracle: max abs %.4f mean abs %.5f psnr %.2f dB" % (W, H, float(err.max()), float(err.mean()), p))
    assert float(err.mean()) <= 0.02 and float(err.max()) <= 0.15 and p >= 39.2
    assert float(one.abs().max()) < 1.0


@pytest.mark.parametrize("B,H,W", [(2, 48, 80), (1, 144, 272), (1, 96, 1552)])
def test_generator_sizes_off_the_tile_grid(cuda, B, H, W):
    """Image sizes the reference accepts (multiples of 16) whose pixel grids are not whole 128-pixel tiles at some or
    all of the five resolutions: 3x5, 9x17 and 6x97 bottlenecks."""
    nw = _networks()
    torch.manual_seed(77)
    net = nw.define_G(39, 3, 64, "global", 4, 2, 1, 3, "instance", gpu_ids=[]).eval()
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.randn(B, 39, H, W, generator=torch.Generator().manual_seed(H + W))
    with torch.no_grad():
        ref = orc.generator_forward(sd, x, 4, 2)
        emu = orc.generator_forward(sd, x, 4, 2, round_fn=_bf)
        y = net.to(cuda)(x.to(cuda)).cpu()
    err = (y - ref).abs()
    assert float(err.mean()) <= 0.02 and float(err.max()) <= 0.15 and orc.psnr(y, ref) >= 39.2
    assert float((y - emu).abs().mean()) <= 0.008


def test_generator_state_dict_round_trip_and_repack(cuda, tmp_path):
    nw = _networks()
    torch.manual_seed(3)
    a = nw.define_G(39, 3, 64, "global", 2, 1, 1, 3, "instance", gpu_ids=[0]).eval()
    torch.save(a.state_dict(), tmp_path / "net_G.pth")  # BaseModel.save_network format (base_model.py:54-59)
    torch.manual_seed(4)
    b = nw.define_G(39, 3, 64, "global", 2, 1, 1, 3, "instance", gpu_ids=[0]).eval()
    x = torch.randn(1, 39, 64, 64, device=cuda)
    with torch.no_grad():
        ya, yb = a(x), b(x)
        assert float((ya - yb).abs().max()) > 1e-3
        b.load_state_dict(torch.load(tmp_path / "net_G.pth"))
        yb2 = b(x)  # weights changed -> packed copies must be rebuilt
    assert torch.equal(ya, yb2)


def test_unsupported_training_inputs_fail_loudly(cuda):
    import jpdse_b200
    nw = _networks()
    net = nw.define_G(39, 3, 64, "global", 1, 0, 1, 3, "instance", gpu_ids=[0])
    with pytest.raises(NotImplementedError):  # the generator input takes no gradient on this path: refuse, no fallback
        net(torch.zeros(1, 39, 128, 128, device=cuda, requires_grad=True))
    wide = nw.define_G(39, 3, 128, "global", 1, 0, 1, 3, "instance", gpu_ids=[0])
    with pytest.raises(jpdse_b200.JpdseError):  # backward kernels are built for ngf == 64
        wide(torch.zeros(1, 39, 128, 128, device=cuda))
    with torch.no_grad():
        assert wide(torch.zeros(1, 39, 128, 128, device=cuda)).shape == (1, 3, 128, 128)


def test_trainer_get_img_matches_oracle(cuda):
    import bench
    trainers = importlib.import_module("jpd-se_b200.ctu.trainers")
    opt = bench.make_opt()
    opt.n_blocks_global = 1
    torch.manual_seed(8)
    trainer = trainers.get_trainer(opt)(opt, "test")
    label, inst, image = bench.synth_inputs(1, 128, 256, seed=2)
    x_dict = {"label": label, "instance": inst, "image": image, "path": ["synthetic"]}
    y = trainer.get_img(x_dict).cpu()
    sd = {k: v.cpu() for k, v in trainer.model.netG.state_dict().items()}
    with torch.no_grad():
        ref = orc.generator_forward(sd, torch.from_numpy(orc.build_input(label.numpy(), inst.numpy(), image.numpy(), 35)), 4, 1)
    assert y.shape == (1, 3, 128, 256)
    assert float((y - ref).abs().mean()) <= 0.02 and orc.psnr(y, ref) >= 39.2


# ------------------------------------------------------------------------------------------------ full-size properties
def test_full_size_batch_independence(cuda):
    """At the BASELINE size (1024x512) the oracle is too slow for a batch, so check a size-independent
    property: InstanceNorm is per-sample, hence image i of a batch must equal the same image run alone."""
    import bench
    nw = _networks()
    torch.manual_seed(1234)
    net = nw.define_G(39, 3, 64, "global", 4, 9, 1, 3, "instance", gpu_ids=[0]).eval()
    label, inst, image = [t.to(cuda) for t in bench.synth_inputs(3, 512, 1024, seed=9)]
    with torch.no_grad():
        yb = net.forward_from_maps(label, inst, image, 35)
        y1 = net.forward_from_maps(label[1:2].contiguous(), inst[1:2].contiguous(), image[1:2].contiguous(), 35)
    assert yb.shape == (3, 3, 512, 1024) and torch.isfinite(yb).all()
    # identical arithmetic; only the fp64 atomic order of the statistics may differ
    assert float((yb[1:2] - y1).abs().max()) < 1e-3
    assert float((yb[0] - yb[1]).abs().max()) > 1e-2  # different images really give different outputs


def test_full_size_input_build_properties(cuda):
    ops = _ops()
    import bench
    label, inst, image = bench.synth_inputs(2, 512, 1024, seed=4)
    nhwc, nchw = ops.build_input(label.to(cuda), inst.to(cuda), image.to(cuda), 35, nhwc=True, nchw=True)
    assert torch.equal(nchw[:, :35].sum(dim=1), torch.ones(2, 512, 1024, device=cuda))  # exactly one class per pixel
    assert torch.equal(nchw[:, 36:], image.to(cuda))
    assert np.array_equal(nchw[:, 35:36].cpu().numpy(), orc.get_edges(inst.numpy()))
    inner = nhwc[:, 3:-3, 3:-3, :39].float().permute(0, 3, 1, 2)
    assert torch.equal(inner, _bf(nchw))
    assert torch.equal(nhwc[:, 0], nhwc[:, 6]) and torch.equal(nhwc[:, :, 1], nhwc[:, :, 5])  # reflection
    assert float(nhwc[..., 39].abs().max()) == 0.0


# ------------------------------------------------------------------------------------------------ quantisers
def test_quantizers_golden(cuda, golden_dir):
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "quantizers.npz"))
    q = torch.from_numpy(g["q"]).to(cuda)
    assert np.array_equal(ops.round_f32(q).cpu().numpy(), g["round"], equal_nan=True)
    assert np.array_equal(ops.sign_f32(q).cpu().numpy(), g["sign"])
    ss = ops.softsign_f32(torch.from_numpy(g["ss_x"]).to(cuda), torch.from_numpy(g["ss_u"]).to(cuda))
    assert np.array_equal(ss.cpu().numpy(), g["ss_y"])
    assert np.array_equal(ops.sign_to_bits(torch.from_numpy(g["ss_y"]).to(cuda)).cpu().numpy(), qorc.code_bits(g["ss_y"]))
    # empty and unaligned / ragged sizes
    assert ops.round_f32(torch.empty(0, device=cuda)).numel() == 0
    odd = torch.randn(1027, device=cuda)[1:]
    assert torch.equal(ops.round_f32(odd.contiguous()).cpu(), torch.round(odd.cpu()))


def test_quantizer_mirror_modules(cuda, golden_dir):
    rnd = importlib.import_module("jpd-se_b200.ctu.quantizers.round")
    bz = importlib.import_module("jpd-se_b200.ctu.quantizers.binarize")
    g = np.load(os.path.join(golden_dir, "quantizers.npz"))
    x = torch.tensor([1.5, 1.4, 1.6], device=cuda, requires_grad=True)  # the reference's own smoke (round.py:17-32)
    y = rnd.RoundedIdentity.apply(x)
    y.sum().backward()
    assert y.tolist() == [2.0, 1.0, 2.0] and x.grad.tolist() == [1.0, 1.0, 1.0]
    ds = bz.DifferentiableSign().eval()
    assert np.array_equal(ds(torch.from_numpy(g["q"]).to(cuda)).cpu().numpy(), g["sign"])
    ds.train()
    xs = torch.from_numpy(g["ss_x"]).to(cuda).requires_grad_(True)
    ys = ds(xs)
    ys.sum().backward()
    assert set(ys.detach().cpu().unique().tolist()) <= {-1.0, 1.0} and float(xs.grad.min()) == 1.0  # straight-through
    p_plus = float((ys == 1).float().mean())  # P(+1) = (1+x)/2 on average
    assert abs(p_plus - float(((1 + xs.detach()) / 2).mean())) < 0.05


def test_binarizer_eval(cuda, golden_dir):
    bz = importlib.import_module("jpd-se_b200.ctu.quantizers.binarize")
    g = np.load(os.path.join(golden_dir, "quantizers.npz"))
    m = bz.Binarizer(64, 16).eval()
    w = _bf(torch.from_numpy(g["bin_w"]))
    x = _bf(torch.from_numpy(g["bin_x"]))
    m.conv.weight.data.copy_(w)
    with torch.no_grad():
        y = m.to(cuda)(x.to(cuda)).cpu()
    pre = F.conv2d(x, w)
    ref = torch.sign(torch.tanh(pre))
    assert set(y.unique().tolist()) <= {-1.0, 0.0, 1.0}
    decided = pre.abs() > 1e-4  # away from fp32 summation-order noise around 0 the codes are bit-exact
    assert torch.equal(y[decided], ref[decided]) and float(decided.float().mean()) > 0.99


def test_s2hvq_golden_and_ties(cuda, golden_dir):
    s2h = importlib.import_module("jpd-se_b200.ctu.quantizers.s2h_vq")
    g = np.load(os.path.join(golden_dir, "quantizers.npz"))
    code_len = int(g["vq_code_len"])
    vq = s2h.S2HVQ(torch.from_numpy(g["vq_cb"]).to(cuda), sigma=float(g["vq_sigma"]))
    x = torch.from_numpy(g["vq_x"]).to(cuda)
    with torch.no_grad():
        assert np.array_equal(vq._get_score_mtrx(vq._vec2mtrx(x, code_len)).cpu().numpy(), g["vq_scores"])
        assert np.array_equal(vq.encode(x, code_len, train=False, raw=True).cpu().numpy(), g["vq_hard"])
        idx = vq.encode(x, code_len, train=False, raw=False)
        assert idx.dtype == torch.int64 and np.array_equal(idx.cpu().numpy(), g["vq_index"])  # incl. duplicated center
        assert np.allclose(vq.encode(x, code_len, train=True, raw=True).cpu().numpy(), g["vq_soft"], atol=1e-6)
        assert np.array_equal(vq.decode(torch.from_numpy(g["vq_hard"]).to(cuda)).cpu().numpy(), g["vq_decoded"])


def test_s2hvq_refuses_to_drop_gradients(cuda):
    s2h = importlib.import_module("jpd-se_b200.ctu.quantizers.s2h_vq")
    vq = s2h.S2HVQ(torch.randn(8, 4).to(cuda), sigma=2.0)
    x = torch.randn(3, 8, device=cuda)
    with pytest.raises(NotImplementedError):  # code book is a Parameter: the reference would backprop through softmax
        vq.encode(x, 2, train=True, raw=True)
    with torch.no_grad():
        assert vq.encode(x, 2, train=True, raw=True).shape == (3, 2, 8)


def test_s2hvq_random_floats_match_up_to_near_ties(cuda):
    ops = _ops()
    g = torch.Generator().manual_seed(12)
    x = torch.randn(5000, 12, generator=g)
    cb = torch.randn(64, 12, generator=g)
    out = ops.s2hvq_encode(x.to(cuda), cb.to(cuda), 3.0, want_scores=True, want_index=True, want_soft=True)
    sc = qorc.s2hvq_scores(x.unsqueeze(0), cb)[0]
    assert torch.allclose(out["scores"].cpu(), sc, rtol=1e-5, atol=1e-5)
    idx = out["index"].cpu()
    ref_idx = sc.argmin(dim=-1)
    same = idx == ref_idx
    # a different index is only acceptable on a near-tie of the two scores (fp32 summation order)
    gap = (sc.gather(1, idx[:, None]) - sc.gather(1, ref_idx[:, None])).abs().squeeze(1)
    assert bool(((gap <= 1e-5 * sc.min(dim=-1).values.abs().clamp_min(1.0)) | same).all())
    assert torch.allclose(out["soft"].cpu(), torch.softmax(-3.0 * sc, dim=-1), atol=1e-5)


# ------------------------------------------------------------------------------------------------ eval-metric path
def test_tensor2im_and_distortion_bit_exact(cuda):
    """tensor2im (float64 de-normalise, clip, TRUNCATE) and the L1 / MSE taken on its bytes."""
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    B, H, W = 3, 37, 53
    a = (torch.rand(B, 3, H, W, generator=g) * 1.3 - 0.65)   # beyond [-0.5, 0.5]: exercises the clip on both sides
    b = torch.tanh(torch.randn(B, 3, H, W, generator=g))
    a[0, 0, 0, :6] = torch.tensor([-0.5, 0.5, 0.0, 0.4999999, 127.5 / 255 - 0.5, 1.0 / 255 - 0.5])
    for mean, std in (((0.5, 0.5, 0.5), (1.0, 1.0, 1.0)), ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))):
        ua = np.stack([orc.tensor2im_uint8(t, mean, std) for t in a])
        ub = np.stack([orc.tensor2im_uint8(t, mean, std) for t in b])
        got = ops.tensor2im_u8(a.to(cuda), mean, std).cpu().numpy()
        assert np.array_equal(got, ua)
        for mode in ("l1", "mse"):
            d = ua.astype(np.float64) - ub.astype(np.float64)
            want = np.abs(d).mean() if mode == "l1" else (d * d).mean()
            val = float(ops.distortion_u8(a.to(cuda), b.to(cuda), mode, mean, std))
            assert val == want, (mode, val, want)


def test_eval_metric_matches_reference_golden(cuda, golden_dir):
    """The reference's own numbers (tests/golden/eval_metric.npz, written by oracle/pin_against_reference.py from
    ctu.utils.misc.tensor2im + nn.L1Loss / nn.MSELoss): device bytes bit-exact, losses equal as float32."""
    ops = _ops()
    g = np.load(os.path.join(golden_dir, "eval_metric.npz"))
    a, b = torch.from_numpy(g["a"]).to(cuda), torch.from_numpy(g["b"]).to(cuda)
    assert np.array_equal(ops.tensor2im_u8(a).cpu().numpy(), g["a_u8"])
    assert np.array_equal(ops.tensor2im_u8(b).cpu().numpy(), g["b_u8"])
    assert np.float32(float(ops.distortion_u8(a, b, "l1"))) == g["l1"]
    assert np.float32(float(ops.distortion_u8(a, b, "mse"))) == g["mse"]


def test_cuda_graph_replay_equals_eager_launches(cuda):
    """Inference plans replay a captured CUDA graph; the result must be bit-identical to the eager launch sequence, for
    changing inputs, both entry points, and across a weight update (graphs are re-captured)."""
    import bench
    nw = _networks()
    torch.manual_seed(11)
    net = nw.define_G(39, 3, 64, "global", 4, 2, 1, 3, "instance", gpu_ids=[0]).eval()
    outs = {}
    for use_graph in (True, False):
        with torch.no_grad():
            res = []
            for seed in (1, 2, 1):
                label, inst, image = [t.to(cuda) for t in bench.synth_inputs(2, 128, 256, seed=seed)]
                plan = net.plan_for(2, 128, 256, cuda)
                plan.use_graph = use_graph
                res.append(net.forward_from_maps(label, inst, image, 35).clone())
                res.append(net(torch.randn(2, 39, 128, 256, generator=torch.Generator().manual_seed(seed)).to(cuda)).clone())
            outs[use_graph] = res
    assert all(torch.equal(a, b) for a, b in zip(outs[True], outs[False]))
    assert torch.equal(outs[True][0], outs[True][4]) and not torch.equal(outs[True][0], outs[True][2])
    with torch.no_grad():
        label, inst, image = [t.to(cuda) for t in bench.synth_inputs(2, 128, 256, seed=1)]
        net.plan_for(2, 128, 256, cuda).use_graph = True
        before = net.forward_from_maps(label, inst, image, 35).clone()
        for p in net.parameters():
            p.mul_(1.01)
        after = net.forward_from_maps(label, inst, image, 35)
    assert not torch.equal(before, after)


def test_full_res_2048x1024_batch_independence(cuda):
    """BASELINE.json configs[2] size: image i of a batch equals the same image run alone."""
    import bench
    nw = _networks()
    torch.manual_seed(1234)
    net = nw.define_G(39, 3, 64, "global", 4, 9, 1, 3, "instance", gpu_ids=[0]).eval()
    label, inst, image = [t.to(cuda) for t in bench.synth_inputs(2, 1024, 2048, seed=5)]
    with torch.no_grad():
        both = net.forward_from_maps(label, inst, image, 35).clone()
        one = net.forward_from_maps(label[1:].contiguous(), inst[1:].contiguous(), image[1:].contiguous(), 35)
    assert both.shape == (2, 3, 1024, 2048) and torch.isfinite(both).all()
    assert torch.equal(both[1:], one)


def test_binarizing_generator_inference(cuda):
    """binarize_generator=True, bin_before_res=False (the parser default when --no_generator_binarization is absent):
    Binarizer behind the res blocks, codes in {-1,+1}; mode='get_binary_code' returns them (networks.py:252-261).
    Codes are signs of pre-activations, so against the fp32 reference only symbols whose pre-activation is within the
    bf16 error of zero may flip: agreement >= 97 %; the decoded image is checked through the same-codes property
    (image from OUR codes == oracle up-path on OUR codes within the bf16 tolerance)."""
    nw = _networks()
    torch.manual_seed(11)
    net = nw.define_G(39, 3, 64, "global", 4, 2, 1, 3, "instance", gpu_ids=[], binarize_generator=True,
                      bin_generator_before_res=False)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    x = torch.randn(2, 39, 128, 256, generator=torch.Generator().manual_seed(6))
    net = net.to(cuda).eval()
    with torch.no_grad():
        codes = net(x.to(cuda), mode="get_binary_code").cpu()
        img = net(x.to(cuda)).cpu()
        ref_codes = orc.generator_forward(sd, x, 4, 2, binarize=True, codes_only=True)
    assert codes.shape == (2, 128, 8, 16) and set(codes.unique().tolist()) <= {-1.0, 0.0, 1.0}
    assert float((codes == ref_codes).float().mean()) >= 0.97
    # up-path of the oracle on our codes
    up_sd = dict(sd)
    with torch.no_grad():
        t = codes
        idx = 4 + 3 * 4 + 2 + 1
        for i in range(4):
            t = F.relu(F.instance_norm(F.conv_transpose2d(t, sd["model.%d.weight" % idx], sd["model.%d.bias" % idx], stride=2,
                                                          padding=1, output_padding=1)))
            idx += 3
        want = torch.tanh(F.conv2d(F.pad(t, (3, 3, 3, 3), mode="reflect"), sd["model.%d.weight" % (idx + 1)],
                                   sd["model.%d.bias" % (idx + 1)]))
    assert float((img - want).abs().mean()) <= 0.02 and orc.psnr(img, want) >= 39.2
    with pytest.raises(AttributeError):
        nw.define_G(39, 3, 64, "global", 1, 0, 1, 3, "instance", gpu_ids=[0])(x[:, :, :128, :128].to(cuda), mode="get_binary_code")


def test_trainer_get_code_and_eval_rate(cuda):
    """Binarizing generator behind the ctu trainer API: get_code returns the {0,1} code bits (B, 128*h*w) and
    get_eval_rate the (Shannon, raw) bits per pixel (pix2pixHD_trainer.py:100-110, pix2pixHD_model.py:466-490)."""
    import bench
    trainers = importlib.import_module("jpd-se_b200.ctu.trainers")
    opt = bench.make_opt()
    opt.no_generator_binarization, opt.n_blocks_global = False, 1
    torch.manual_seed(8)
    trainer = trainers.get_trainer(opt)(opt, "test")
    label, inst, image = bench.synth_inputs(2, 128, 256, seed=2)
    x_dict = {"label": label, "instance": inst, "image": image, "path": ["a", "b"]}
    code = trainer.get_code(x_dict)
    assert code.shape == (2, 128 * 8 * 16) and set(code.unique().tolist()) <= {0.0, 0.5, 1.0}
    shannon, actual = trainer.get_eval_rate(x_dict)
    assert abs(actual - 128 * 8 * 16 / (128 * 256)) < 1e-9
    assert 0.0 < float(shannon) <= actual + 1e-6
    img = trainer.get_img(x_dict)
    assert img.shape == (2, 3, 128, 256) and torch.isfinite(img).all()
    bits = _ops().sign_to_bits(code * 2 - 1)
    assert bits.dtype == torch.uint8 and int(bits.max()) <= 1


def test_split_stream_plan_equals_single_stream(cuda, monkeypatch):
    """Batches >= 8 run as two half-batch plans on two CUDA streams inside one captured graph; the result must be
    bit-identical to the single-stream plan (InstanceNorm is per-sample), eagerly and replayed."""
    import bench
    nw = _networks()
    torch.manual_seed(12)
    net = nw.define_G(39, 3, 64, "global", 4, 2, 1, 3, "instance", gpu_ids=[0]).eval()
    label, inst, image = [t.to(cuda) for t in bench.synth_inputs(8, 128, 256, seed=9)]
    outs = []
    for split, use_graph in (("2", True), ("2", False), ("1", True)):
        monkeypatch.setenv("JPDSE_SPLIT_STREAMS", split)
        net._plans = {}
        with torch.no_grad():
            plan = net.plan_for(8, 128, 256, cuda)
            assert hasattr(plan, "parts") == (split == "2")
            plan.use_graph = use_graph
            a = net.forward_from_maps(label, inst, image, 35).clone()
            b = net.forward_from_maps(label, inst, image, 35).clone()   # second call: graph replay
            x = torch.randn(8, 39, 128, 256, generator=torch.Generator().manual_seed(1)).to(cuda)
            c = net(x).clone()
        assert torch.equal(a, b)
        outs.append((a, c))
    assert all(torch.equal(outs[0][0], o[0]) and torch.equal(outs[0][1], o[1]) for o in outs[1:])


def test_cta_pair_conv_equals_single_cta_kernel(cuda, monkeypatch):
    """The res-block conv runs on CTA pairs (tcgen05 cta_group::2, conv_pair.cu) when the problem is large enough; same
    packed weights, same K order, fp32 accumulate -> bit-identical raw output and matching statistics vs conv_igemm.cu,
    and within one bf16 ulp of torch."""
    ops = _ops()
    from jpdse_b200._lib import CONV3X3_PAD1, EPI_RAW_STATS
    g = torch.Generator().manual_seed(5)
    B, H, W, C = 10, 32, 64, 1024   # 160 m-tiles -> 80 pairs x 4 n-tiles = 320 pair tiles (>= 148: pair path), ragged waves
    x = _bf(torch.randn(B, C, H, W, generator=g))
    w = _bf(torch.randn(C, C, 3, 3, generator=g) * 0.02)
    xd = ops.nchw_to_nhwc_bf16(x.to(cuda), pad_reflect=1)
    outs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("JPDSE_PAIR_CONV", flag)
        cv = ops.Conv(CONV3X3_PAD1, EPI_RAW_STATS, B, H, W, 1, C, C, C, cuda)
        cv.pack(w.to(cuda))
        y = torch.full((B, H, W, C), float("nan"), dtype=torch.bfloat16, device=cuda)
        st = torch.zeros(B, C, 2, dtype=torch.float64, device=cuda)
        cv.forward(xd, y, st)
        torch.cuda.synchronize()
        outs.append((y, st))
    assert torch.equal(outs[0][0], outs[1][0])
    assert torch.allclose(outs[0][1], outs[1][1], rtol=1e-6, atol=1e-3)
    ref = F.conv2d(F.pad(x[:2], (1, 1, 1, 1), mode="reflect"), w)
    got = outs[0][0][:2].float().cpu().permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= float(ref.abs().max()) * 2.0 ** -7


def test_compact_uint8_inputs_bit_exact_with_loader_normalisation(cuda):
    """uint8 label / int16 instance / uint8 RGB go straight to the input-build kernel; the loader's
    ToTensor + Normalize (x/255, then (x - mean)/std, float32) is fused in and must be bit-exact with torch's."""
    import bench
    ops = _ops()
    g = torch.Generator().manual_seed(13)
    B, H, W, L = 2, 33, 47, 35
    lab = torch.randint(0, L, (B, 1, H, W), generator=g)
    inst = torch.randint(0, 9, (B, 1, H, W), generator=g)
    u8 = torch.randint(0, 256, (B, 3, H, W), generator=g, dtype=torch.uint8)
    for mean, std in (((0.5, 0.5, 0.5), (1.0, 1.0, 1.0)), ((0.485, 0.456, 0.406), (0.229, 0.224, 0.225))):
        m = torch.tensor(mean).view(1, 3, 1, 1)
        s = torch.tensor(std).view(1, 3, 1, 1)
        img = (u8.float().div(255) - m) / s  # torchvision ToTensor + Normalize
        ref = orc.build_input(lab.numpy(), inst.numpy(), img.numpy(), L)
        nhwc, nchw = ops.build_input(lab.to(torch.uint8).to(cuda), inst.to(torch.int16).to(cuda), u8.to(cuda), L, pad=3, nhwc=True,
                                     nchw=True, mean=mean, std=std)
        assert np.array_equal(nchw.cpu().numpy(), ref)
        assert np.array_equal(nhwc.float().cpu().numpy(), _bf(torch.from_numpy(orc.reflect_pad_nhwc(ref, 3, 40))).numpy())
    # and through the trainer API
    trainers = importlib.import_module("jpd-se_b200.ctu.trainers")
    opt = bench.make_opt()
    opt.n_blocks_global = 1
    torch.manual_seed(8)
    trainer = trainers.get_trainer(opt)(opt, "test")
    label, inst2, image = bench.synth_inputs(1, 128, 256, seed=2)
    u = ((image + 0.5) * 255).round().clamp(0, 255).to(torch.uint8)
    a = trainer.get_img({"label": label, "instance": inst2, "image": u.float().div(255) - 0.5})
    b = trainer.get_img({"label": label.to(torch.uint8), "instance": inst2.to(torch.int16), "image": u})
    assert torch.equal(a, b)


def test_second_device_in_the_same_process(cuda):
    """One process per GPU is the design, but a process that moves to another device (`--gpu_ids 1` after a first call
    on device 0) must work: the kernels' shared-memory attributes are set per device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    nw = _networks()
    torch.manual_seed(3)
    net = nw.define_G(39, 3, 64, "global", 2, 1, 1, 3, "instance", gpu_ids=[]).eval()
    x = torch.randn(1, 39, 32, 64, generator=torch.Generator().manual_seed(1))
    outs = []
    for dev in (0, 1):
        with torch.cuda.device(dev), torch.no_grad():
            import copy
            outs.append(copy.deepcopy(net).to("cuda:%d" % dev)(x.to("cuda:%d" % dev)).cpu())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("kind_name,B,H,W", [("pad1", 2, 6, 256), ("pad1", 1, 3, 200), ("pad1_act", 2, 5, 128), ("pad1_act", 1, 4, 330),
                                             ("full", 2, 8, 16), ("full", 1, 20, 300), ("full_shared", 2, 12, 40),
                                             ("full_shared", 3, 9, 130)])
def test_row_mode_equals_tap_by_tap(cuda, monkeypatch, kind_name, B, H, W):
    """Row mode of the implicit-GEMM kernel (64 -> 64 channels: weights resident in shared memory, one 130-pixel A box per
    filter row, the kw taps as shifted descriptors -- the VGG19's first block) against the generic tap-by-tap K loop
    (JPDSE_ROW_MODE=0): the same MMAs in the same order, so the outputs are BIT-IDENTICAL; and against torch."""
    ops = _ops()
    from jpdse_b200._lib import (CONV3X3_FULL, CONV3X3_FULL_SHARED, CONV3X3_PAD1, EPI_BIAS_ACT, EPI_RAW, EPI_RAW_STATS)
    g = torch.Generator().manual_seed(B * 100 + W)
    w = _bf(torch.randn(64, 64, 3, 3, generator=g) * 0.05)
    bias = torch.randn(64, generator=g) * 0.1

    def run():
        if kind_name.startswith("pad1"):
            act = kind_name == "pad1_act"
            x = _bf(torch.randn(B, 64, H + 2, W + 2, generator=torch.Generator().manual_seed(5)))
            xd = torch.zeros(B * (H + 2) * (W + 2) * 64 + 4096, dtype=torch.bfloat16, device=cuda)
            xd[: x.numel()] = x.permute(0, 2, 3, 1).reshape(-1).to(cuda)
            xd = xd[: x.numel()].view(B, H + 2, W + 2, 64)
            if act:
                cv = ops.Conv(CONV3X3_PAD1, EPI_BIAS_ACT, B, H, W, 1, 64, 64, 64, cuda, out_pad=1, slope=0.0)
                cv.pack(w.to(cuda), bias.to(cuda))
                y = torch.zeros(B, H + 2, W + 2, 64, dtype=torch.bfloat16, device=cuda)
                cv.forward(xd, y)
                ref = F.relu(F.conv2d(x, w, bias))
                got = y[:, 1:-1, 1:-1]
            else:
                cv = ops.Conv(CONV3X3_PAD1, EPI_RAW_STATS, B, H, W, 1, 64, 64, 64, cuda)
                cv.pack(w.to(cuda))
                y = torch.full((B, H, W, 64), float("nan"), dtype=torch.bfloat16, device=cuda)
                st = torch.zeros(B, 64, 2, dtype=torch.float64, device=cuda)
                cv.forward(xd, y, st)
                ref = F.conv2d(x, w)
                got = y
            torch.cuda.synchronize()
            return y.clone(), got.float().cpu().permute(0, 3, 1, 2), ref
        dy = _bf(torch.randn(B, 64, H, W, generator=torch.Generator().manual_seed(6)))
        xp = torch.zeros(B, 64, H + 2, W + 2, requires_grad=True)
        F.conv2d(xp, w).backward(dy)
        ref = xp.grad
        dyn = dy.permute(0, 2, 3, 1).contiguous().to(cuda)
        if kind_name == "full":
            cv = ops.Conv(CONV3X3_FULL, EPI_RAW, B, H, W, 2, 64, 64, 64, cuda)
            buf = torch.zeros(B * (H + 4) * (W + 4) * 64 + 4096, dtype=torch.bfloat16, device=cuda)
            xv = buf[: B * (H + 4) * (W + 4) * 64].view(B, H + 4, W + 4, 64)
            xv[:, 2:-2, 2:-2] = dyn
        else:
            cv = ops.Conv(CONV3X3_FULL_SHARED, EPI_RAW, B, H, W, 2, 64, 64, 64, cuda)
            P, S = W + 2, (H + 2) * (W + 2)
            xv = torch.zeros(B * S + 2 * P + 2 + 256, 64, dtype=torch.bfloat16, device=cuda)
            for b in range(B):
                xv[b * S + 2 * P + 2: b * S + 2 * P + 2 + H * P].view(H, P, 64)[:, :W] = dyn[b]
        cv.pack(w.to(cuda))
        y = torch.full((B, H + 2, W + 2, 64), float("nan"), dtype=torch.bfloat16, device=cuda)
        cv.forward(xv, y)
        torch.cuda.synchronize()
        return y.clone(), y.float().cpu().permute(0, 3, 1, 2), ref

    y_row, got, ref = run()
    assert not torch.isnan(got).any()
    assert float((got - ref).abs().max()) <= float(ref.abs().max()) * 2.0 ** -7
    monkeypatch.setenv("JPDSE_ROW_MODE", "0")
    y_tap, _, _ = run()
    assert torch.equal(y_row, y_tap)
