"""CPU: the oracle restatement reproduces the golden vectors generated from the imported reference
(oracle/pin_against_reference.py). Integer / index work bit-exact, floating point within 1e-5 (oneDNN may
pick another conv kernel on a different host CPU)."""
import os

import numpy as np
import torch

from oracle import generator_oracle as orc
from oracle import quantizer_oracle as qorc


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name))


def test_preprocess_golden(golden_dir):
    g = _load(golden_dir, "preprocess.npz")
    got = orc.build_input(g["label"], g["instance"], g["image"], int(g["num_labels"]))
    assert got.dtype == np.float32
    assert np.array_equal(got, g["input_concat"])


def test_preprocess_edge_semantics():
    inst = np.zeros((1, 1, 5, 6), dtype=np.int32)
    inst[0, 0, 2, 3] = 7
    e = orc.get_edges(inst)[0, 0]
    want = np.zeros((5, 6), dtype=np.float32)
    for h, w in [(2, 3), (1, 3), (3, 3), (2, 2), (2, 4)]:
        want[h, w] = 1
    assert np.array_equal(e, want)  # 4-neighbour cross, no diagonals, borders never an edge by themselves
    assert orc.get_edges(np.full((1, 1, 4, 4), 3, dtype=np.int16)).sum() == 0


def test_one_hot_truncates_and_rejects_out_of_range():
    lab = np.array([[[[0.99, 1.5], [33.999, 34.0]]]], dtype=np.float32)
    oh = orc.one_hot(lab, 35)
    assert oh[0, :, 0, 0].argmax() == 0 and oh[0, :, 0, 1].argmax() == 1
    assert oh[0, :, 1, 0].argmax() == 33 and oh[0, :, 1, 1].argmax() == 34
    assert np.array_equal(oh.sum(axis=1), np.ones((1, 2, 2), dtype=np.float32))
    try:
        orc.one_hot(np.full((1, 1, 2, 2), 35.0, dtype=np.float32), 35)
        assert False, "id 35 must be rejected for 35 channels (scatter_ raises in the reference)"
    except IndexError:
        pass


def test_generator_golden(golden_dir):
    import importlib
    g = _load(golden_dir, "generator_small.npz")
    nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
    torch.manual_seed(int(g["seed"]))
    net = nw.define_G(int(g["input_nc"]), int(g["output_nc"]), int(g["ngf"]), "global", int(g["n_down"]),
                      int(g["n_blocks"]), 1, 3, "instance", gpu_ids=[])
    sd = net.state_dict()
    wsum = float(sum(v.double().sum() for v in sd.values()))
    assert abs(wsum - float(g["weight_sum"])) < 1e-6 * max(1.0, abs(wsum)), "seeded init differs from the reference's"
    with torch.no_grad():
        y = orc.generator_forward(sd, torch.from_numpy(g["x"]), int(g["n_down"]), int(g["n_blocks"]))
    assert np.allclose(y.numpy(), g["y"], atol=1e-5, rtol=0)


def test_generator_gradient_golden(golden_dir):
    """Oracle autograd reproduces the reference's gradient checksums (sum, abs-sum per parameter) for L1 x 10."""
    import importlib
    g = _load(golden_dir, "generator_small.npz")
    gg = _load(golden_dir, "generator_small_grads.npz")
    nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
    torch.manual_seed(int(g["seed"]))
    net = nw.define_G(39, 3, 64, "global", int(g["n_down"]), int(g["n_blocks"]), 1, 3, "instance", gpu_ids=[])
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    y = orc.generator_forward(sd, torch.from_numpy(g["x"]), int(g["n_down"]), int(g["n_blocks"]))
    (10.0 * (y - torch.from_numpy(gg["target"])).abs().mean()).backward()
    for name, (s1, s2) in zip(gg["names"], gg["sums"]):
        grad = sd[str(name)].grad.double()
        assert abs(float(grad.abs().sum()) - s2) <= 1e-4 * max(s2, 1e-6) + 1e-9, name
        assert abs(float(grad.sum()) - s1) <= 1e-4 * max(s2, 1e-6) + 1e-9, name


def test_generator_bf16_emulation_close_to_fp32(golden_dir):
    import importlib
    g = _load(golden_dir, "generator_small.npz")
    nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
    torch.manual_seed(int(g["seed"]))
    sd = nw.define_G(39, 3, 64, "global", 4, 2, 1, 3, "instance", gpu_ids=[]).state_dict()
    x = torch.from_numpy(g["x"])
    with torch.no_grad():
        y = orc.generator_forward(sd, x, 4, 2, round_fn=lambda t: t.bfloat16().float())
    assert orc.psnr(y, torch.from_numpy(g["y"])) > 35.0


def test_quantizers_golden(golden_dir):
    g = _load(golden_dir, "quantizers.npz")
    assert np.array_equal(qorc.rounded_identity(g["q"]), g["round"], equal_nan=True)
    assert np.array_equal(qorc.sign(g["q"]), g["sign"])
    assert np.array_equal(qorc.soft_sign(g["ss_x"], g["ss_u"]), g["ss_y"])
    b = qorc.binarizer_eval(torch.from_numpy(g["bin_x"]), torch.from_numpy(g["bin_w"]))
    assert np.array_equal(b.numpy(), g["bin_y"])
    # the reference's own smoke values, ctu/quantizers/round.py:17-32
    assert qorc.rounded_identity(np.array([1.5, 1.4, 1.6], dtype=np.float32)).tolist() == [2.0, 1.0, 2.0]


def test_s2hvq_golden(golden_dir):
    g = _load(golden_dir, "quantizers.npz")
    code_len = int(g["vq_code_len"])
    x = torch.from_numpy(g["vq_x"])
    xm = x.view(-1, code_len, x.size(1) // code_len)
    cb = torch.from_numpy(g["vq_cb"])
    assert np.array_equal(qorc.s2hvq_scores(xm, cb).numpy(), g["vq_scores"])
    idx, hard = qorc.s2hvq_hard(xm, cb)
    assert np.array_equal(idx.numpy(), g["vq_index"]) and np.array_equal(hard.numpy(), g["vq_hard"])
    assert np.allclose(qorc.s2hvq_soft(xm, cb, float(g["vq_sigma"])).numpy(), g["vq_soft"], atol=1e-7)
    dec, _ = qorc.s2hvq_decode(torch.from_numpy(g["vq_hard"]), cb)
    assert np.array_equal(dec.numpy(), g["vq_decoded"])


def test_code_bits():
    assert qorc.code_bits(np.array([-1.0, 1.0, 0.0], dtype=np.float32)).tolist() == [0, 1, 0]


def test_eval_metric_golden(golden_dir):
    """tensor2im bytes and the eval L1 / MSE of the reference (ctu/utils/misc.py:64-95, pix2pixHD_model.py:636-641),
    stored by oracle/pin_against_reference.py from the imported reference."""
    g = np.load(os.path.join(golden_dir, "eval_metric.npz"))
    a, b = torch.from_numpy(g["a"]), torch.from_numpy(g["b"])
    assert np.array_equal(np.stack([orc.tensor2im_uint8(t) for t in a]), g["a_u8"])
    assert np.array_equal(np.stack([orc.tensor2im_uint8(t) for t in b]), g["b_u8"])
    assert orc.eval_distortion(a, b, mode="l1") == g["l1"]   # sums below 2^24: the reference's float32 mean is exact
    assert orc.eval_distortion(a, b, mode="mse") == g["mse"]


def test_discriminator_golden(golden_dir):
    """netD / GANLoss / feature matching of the reference (networks.py:371-471, 80-122; pix2pixHD_model.py:715-753):
    outputs stored by oracle/pin_against_reference.py from the imported reference, reproduced by the oracle."""
    import importlib
    from oracle import discriminator_oracle as dorc
    nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
    g = np.load(os.path.join(golden_dir, "discriminator_small.npz"))
    torch.manual_seed(7)
    sd = nw.define_D(39, 64, 3, "instance", False, 2, True, gpu_ids=[]).state_dict()
    assert abs(float(sum(v.double().sum() for v in sd.values())) - float(g["weight_sum"])) < 1e-6
    x, real = torch.from_numpy(g["x"]), torch.from_numpy(g["real"])
    with torch.no_grad():
        feats = dorc.discriminator_forward(sd, x, 3, 2)
    assert np.allclose(feats[0][-1].numpy(), g["final0"], atol=1e-5) and np.allclose(feats[1][-1].numpy(), g["final1"], atol=1e-5)
    sums = np.array([float(t.double().sum()) for s_ in feats for t in s_])
    assert np.allclose(sums, g["feat_sums"], rtol=1e-5, atol=1e-2)
    fake = x[:, 36:].clone().requires_grad_(True)
    l_gan, l_fm, l_real, l_fake = dorc.discriminator_losses(sd, x[:, :36], fake, real, 3, 2)
    assert np.allclose([float(l_gan), float(l_fm), float(l_real), float(l_fake)], g["losses"], rtol=1e-5)
    (l_gan + 10.0 * l_fm).backward()
    assert np.allclose([float(fake.grad.double().sum()), float(fake.grad.double().abs().sum())], g["g_fake_sum"], rtol=1e-4, atol=1e-6)
