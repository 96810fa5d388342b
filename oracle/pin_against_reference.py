"""Pins the oracle to the reference and (re)generates tests/golden/*.npz.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/pin_against_reference.py

It imports the UNMODIFIED reference modules, feeds both the reference and the oracle restatement the same
seeded inputs, asserts bit-identical results on CPU, and stores small input/output vectors that the CPU
test-suite (tests/test_oracle_golden.py) re-checks anywhere. Nothing here is product code.
"""
import functools
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("JPDSE_REFERENCE", "/root/reference")
GOLDEN = os.environ.get("JPDSE_GOLDEN_DIR", os.path.join(ROOT, "tests", "golden"))


def import_reference():
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    for m in ("skimage", "skimage.io"):  # imported but unused by ctu/data/ctu_dataset.py:14
        sys.modules.setdefault(m, types.ModuleType(m))
    from ctu.models.pix2pixHD_networks import networks
    from ctu.models import pix2pixHD_model
    from ctu.quantizers import binarize, round as round_mod, s2h_vq
    return networks, pix2pixHD_model, binarize, round_mod, s2h_vq


def main():
    networks, p2p, binarize, round_mod, s2h_vq = import_reference()
    from oracle import generator_oracle as orc
    from oracle import quantizer_oracle as qorc
    import importlib
    ours = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
    os.makedirs(GOLDEN, exist_ok=True)
    g = torch.Generator().manual_seed(20201234)

    # ---------------------------------------------------------------- preprocess: one-hot / edges / concat
    B, H, W, L = 2, 24, 40, 35
    label = torch.randint(0, L, (B, 1, H // 4, W // 4), generator=g).repeat_interleave(4, 2).repeat_interleave(4, 3).float()
    label[0, 0, 0, :5] = torch.tensor([0.0, 34.0, 33.999, 1.5, 0.99])  # .long() truncation cases
    inst = torch.randint(0, 6, (B, 1, H // 8, W // 8), generator=g).repeat_interleave(8, 2).repeat_interleave(8, 3).int()
    inst[1, 0, 5, 7] = 1000  # isolated pixel: 4-neighbour cross
    image = torch.rand(B, 3, H, W, generator=g) - 0.5
    opt = types.SimpleNamespace(use_compressed=False, no_label=False, num_labels=L, contain_dontcare_label=False,
                                data_type=32, no_instance=False, sem_masking=False)
    stub = types.SimpleNamespace(opt=opt, use_gpu=lambda: False, FloatTensor=torch.FloatTensor,
                                 ByteTensor=torch.ByteTensor)
    stub.get_edges = functools.partial(p2p.Pix2PixHDModel.get_edges, stub)
    ref_pre = p2p.Pix2PixHDModel.preprocess(stub, {"label": label.clone(), "instance": inst.clone(), "image": image.clone()})
    ref_concat = torch.cat((ref_pre["input_label"], ref_pre["real_image"]), dim=1)  # pix2pixHD_model.py:595
    got = orc.build_input(label.numpy(), inst.numpy(), image.numpy(), L)
    assert np.array_equal(got, ref_concat.numpy()), "oracle build_input != reference preprocess+concat"
    np.savez_compressed(os.path.join(GOLDEN, "preprocess.npz"), label=label.numpy(), instance=inst.numpy(),
                        image=image.numpy(), input_concat=ref_concat.numpy().astype(np.float32), num_labels=L)
    print("preprocess: oracle == reference (bit-exact); golden written")

    # ---------------------------------------------------------------- generator forward
    cfg = dict(input_nc=39, output_nc=3, ngf=64, n_down=4, n_blocks=2, seed=1234)
    torch.manual_seed(cfg["seed"])
    G = networks.define_G(cfg["input_nc"], cfg["output_nc"], cfg["ngf"], "global", cfg["n_down"], cfg["n_blocks"], 1, 3,
                          "instance", gpu_ids=[]).eval()
    torch.manual_seed(cfg["seed"])
    G2 = ours.define_G(cfg["input_nc"], cfg["output_nc"], cfg["ngf"], "global", cfg["n_down"], cfg["n_blocks"], 1, 3,
                       "instance", gpu_ids=[])
    sd, sd2 = G.state_dict(), G2.state_dict()
    assert list(sd.keys()) == list(sd2.keys()), "state-dict keys differ from the reference"
    assert all(torch.equal(sd[k], sd2[k]) for k in sd), "same seed must give the reference's weights"
    x = torch.randn(1, 39, 32, 64, generator=g)
    with torch.no_grad():
        y_ref = G(x)
        y_orc = orc.generator_forward(sd, x, cfg["n_down"], cfg["n_blocks"])
    assert torch.equal(y_ref, y_orc), "oracle generator != reference generator"
    wsum = float(sum(v.double().sum() for v in sd.values()))
    np.savez_compressed(os.path.join(GOLDEN, "generator_small.npz"), x=x.numpy(), y=y_ref.numpy(),
                        weight_sum=wsum, **{k: np.int64(v) for k, v in cfg.items()})
    print("generator: oracle == reference (bit-exact), our define_G reproduces the reference init; golden written")

    # full-size architecture, tiny image: checks the 9-block / 1024-channel wiring
    torch.manual_seed(1234)
    Gf = networks.define_G(39, 3, 64, "global", 4, 9, 1, 3, "instance", gpu_ids=[]).eval()
    xf = torch.randn(1, 39, 32, 64, generator=g)
    with torch.no_grad():
        assert torch.equal(Gf(xf), orc.generator_forward(Gf.state_dict(), xf, 4, 9))
    print("generator (full 182.6M-param architecture): oracle == reference (bit-exact)")

    # binarizing generator (parser default when --no_generator_binarization is absent): Binarizer behind the res blocks
    torch.manual_seed(11)
    Gb = networks.define_G(39, 3, 64, "global", 4, 2, 1, 3, "instance", gpu_ids=[], binarize_generator=True,
                           bin_generator_before_res=False).eval()
    torch.manual_seed(11)
    Gb2 = ours.define_G(39, 3, 64, "global", 4, 2, 1, 3, "instance", gpu_ids=[], binarize_generator=True,
                        bin_generator_before_res=False)
    assert list(Gb.state_dict().keys()) == list(Gb2.state_dict().keys())
    assert all(torch.equal(v, Gb2.state_dict()[k]) for k, v in Gb.state_dict().items())
    with torch.no_grad():
        assert torch.equal(Gb(x), orc.generator_forward(Gb.state_dict(), x, 4, 2, binarize=True))
        assert torch.equal(Gb(x, mode="get_binary_code"),
                           orc.generator_forward(Gb.state_dict(), x, 4, 2, binarize=True, codes_only=True))
    # train() mode: the stochastic sign draws ONE uniform tensor per forward (binarize.py:20); same generator state -> same draw
    Gb.train()
    for p_ in Gb.parameters():
        p_.grad = None
    torch.manual_seed(123)
    yb = Gb(x)
    (10.0 * (yb - torch.zeros_like(yb)).abs().mean()).backward()
    torch.manual_seed(123)
    nb = torch.empty(1, 128, x.shape[2] // 16, x.shape[3] // 16).uniform_()
    sdb = {k: v.detach().clone().requires_grad_(True) for k, v in Gb.state_dict().items()}
    yo = orc.generator_forward(sdb, x, 4, 2, binarize=True, noise=nb)
    assert torch.equal(yo, yb), "oracle stochastic binarizing generator != reference (train mode)"
    (10.0 * yo.abs().mean()).backward()
    for name, p_ in Gb.named_parameters():
        assert torch.equal(p_.grad, sdb[name].grad), "oracle gradient != reference for %s (binarizing generator)" % name
    Gb.eval()
    print("binarizing generator: oracle == reference (image and get_binary_code, bit-exact), same keys and init; train() mode "
          "with the same uniform draw: outputs and autograd gradients bit-exact")

    # ---------------------------------------------------------------- generator backward (autograd through the reference)
    # loss_G.backward() (pix2pixHD_trainer.py:69) is autograd over the same modules: the oracle's functional forward
    # must give the reference's gradients bit for bit on CPU.
    G.train()
    tgt = torch.rand(1, 3, 32, 64, generator=g) - 0.5
    for p_ in G.parameters():
        p_.grad = None
    (10.0 * (G(x) - tgt).abs().mean()).backward()
    sdg = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    (10.0 * (orc.generator_forward(sdg, x, cfg["n_down"], cfg["n_blocks"]) - tgt).abs().mean()).backward()
    gsum = {}
    for name, p_ in G.named_parameters():
        assert torch.equal(p_.grad, sdg[name].grad), "oracle gradient != reference gradient for %s" % name
        gsum[name] = np.array([float(p_.grad.double().sum()), float(p_.grad.double().abs().sum())])
    np.savez_compressed(os.path.join(GOLDEN, "generator_small_grads.npz"), target=tgt.numpy(),
                        names=np.array(list(gsum.keys())), sums=np.stack(list(gsum.values())))
    G.eval()
    print("generator backward: oracle autograd == reference autograd (bit-exact); golden gradient checksums written")

    # ---------------------------------------------------------------- eval-metric path
    # tensor2im (ctu/utils/misc.py:64-95) and the distortion get_eval_loss takes on its bytes (pix2pixHD_model.py:636-641)
    from ctu.utils import misc
    eopt = types.SimpleNamespace(normalize_mean=[0.5, 0.5, 0.5], normalize_std=[1.0, 1.0, 1.0])
    for (eb, eh, ew) in ((2, 24, 40), (1, 256, 512)):
        ra = (torch.rand(eb, 3, eh, ew, generator=g) - 0.5) * 1.2   # some values outside [-0.5, 0.5]: the clip matters
        rb = (ra + torch.randn(eb, 3, eh, ew, generator=g) * 0.05)
        ra[0, :, 0, :4] = torch.tensor([-0.5, 0.5, 0.49999997, 1.0 / 255 - 0.5])  # exact byte boundaries
        ia, ib = misc.tensor2im(ra, eopt), misc.tensor2im(rb, eopt)
        oa = np.stack([orc.tensor2im_uint8(x_, eopt.normalize_mean, eopt.normalize_std) for x_ in ra])
        ob = np.stack([orc.tensor2im_uint8(x_, eopt.normalize_mean, eopt.normalize_std) for x_ in rb])
        assert ia.dtype == np.uint8 and np.array_equal(ia, oa) and np.array_equal(ib, ob), "oracle tensor2im != reference"
        ta = torch.tensor(ia.transpose(0, 3, 1, 2)).to(torch.float)
        tb = torch.tensor(ib.transpose(0, 3, 1, 2)).to(torch.float)
        for mode, crit in (("l1", torch.nn.L1Loss()), ("mse", torch.nn.MSELoss())):
            ref_loss = np.float32(crit(ta, tb).item())
            got_loss = orc.eval_distortion(ra, rb, eopt.normalize_mean, eopt.normalize_std, mode)
            ulp = abs(int(ref_loss.view(np.int32)) - int(got_loss.view(np.int32)))
            assert ulp <= 2, "eval %s loss: oracle %r vs reference %r (%d ulp)" % (mode, got_loss, ref_loss, ulp)
            print("eval-loss %s at %dx%dx%d: reference %.9g, exact-integer oracle %.9g (%d float32 ulp)" % (
                mode, eb, eh, ew, ref_loss, got_loss, ulp))
        if eb == 2:
            np.savez_compressed(os.path.join(GOLDEN, "eval_metric.npz"), a=ra.numpy(), b=rb.numpy(), a_u8=ia, b_u8=ib,
                                l1=np.float32(torch.nn.L1Loss()(ta, tb).item()),
                                mse=np.float32(torch.nn.MSELoss()(ta, tb).item()))
    print("eval metric: oracle tensor2im == reference (bit-exact bytes); L1 / MSE within 2 float32 ulp; golden written")

    # ---------------------------------------------------------------- training-step networks: discriminator, losses, VGG
    from oracle import discriminator_oracle as dorc
    torch.manual_seed(7)
    Dr = networks.define_D(39, 64, 3, "instance", False, 2, True, gpu_ids=[])
    torch.manual_seed(7)
    Do = ours.define_D(39, 64, 3, "instance", False, 2, True, gpu_ids=[])
    sr, so = Dr.state_dict(), Do.state_dict()
    assert list(sr.keys()) == list(so.keys()) and all(torch.equal(sr[k], so[k]) for k in sr), "netD init/keys differ"
    xd = torch.randn(2, 39, 64, 128, generator=g)
    fr = Dr(xd)
    fo = dorc.discriminator_forward(sr, xd, 3, 2)
    assert len(fr) == len(fo) == 2 and all(len(a) == len(b) == 5 for a, b in zip(fr, fo))
    assert all(torch.equal(a, b) for s_, t_ in zip(fr, fo) for a, b in zip(s_, t_)), "oracle netD != reference netD"
    gr = networks.GANLoss(use_lsgan=True)
    for target_is_real in (True, False):
        assert torch.equal(gr(Dr(xd), target_is_real), dorc.gan_loss(dorc.discriminator_forward(sr, xd, 3, 2), target_is_real))
    # the discriminator half of get_train_loss (pix2pixHD_model.py:715-753) incl. its autograd: losses and the gradients
    # w.r.t. every netD parameter (loss_D) and w.r.t. the fake image (loss_G_GAN + 10 * loss_G_GAN_Feat)
    lab, fake, real = xd[:, :36].clone(), xd[:, 36:].clone().requires_grad_(True), torch.randn(2, 3, 64, 128, generator=g)
    crit_feat = torch.nn.L1Loss()
    pfp = Dr(torch.cat((lab.detach(), fake.detach()), 1))
    l_d_fake = gr(pfp, False)
    pr = Dr(torch.cat((lab.detach(), real.detach()), 1))
    l_d_real = gr(pr, True)
    pf = Dr(torch.cat((lab, fake), 1))
    l_g_gan = gr(pf, True)
    l_fm = 0
    for i in range(2):
        for j in range(len(pf[i]) - 1):
            l_fm += 0.5 * crit_feat(pf[i][j], pr[i][j].detach())
    (l_g_gan + 10.0 * l_fm).backward()
    g_fake_ref = fake.grad.clone()
    Dr.zero_grad()
    ((l_d_fake + l_d_real) * 0.5).backward()
    sdg = {k: v.detach().clone().requires_grad_(True) for k, v in sr.items()}
    fake2 = fake.detach().clone().requires_grad_(True)
    o_gan, o_fm, o_real, o_fake = dorc.discriminator_losses(sdg, lab, fake2, real, 3, 2)
    assert torch.equal(o_gan, l_g_gan) and torch.equal(o_fm, l_fm) and torch.equal(o_real, l_d_real) and torch.equal(o_fake, l_d_fake)
    (o_gan + 10.0 * o_fm).backward()
    assert torch.equal(fake2.grad, g_fake_ref), "oracle d(loss_G)/d(fake) != reference"
    for v in sdg.values():
        v.grad = None
    ((o_fake + o_real) * 0.5).backward()
    for name, p_ in Dr.named_parameters():
        assert torch.equal(p_.grad, sdg[name].grad), "oracle netD gradient != reference for %s" % name
    np.savez_compressed(os.path.join(GOLDEN, "discriminator_small.npz"), x=xd.numpy(), real=real.numpy(),
                        losses=np.array([float(l_g_gan), float(l_fm), float(l_d_real), float(l_d_fake)]),
                        final0=fr[0][-1].detach().numpy(), final1=fr[1][-1].detach().numpy(),
                        feat_sums=np.array([float(t_.double().sum()) for s_ in fr for t_ in s_]),
                        g_fake_sum=np.array([float(g_fake_ref.double().sum()), float(g_fake_ref.double().abs().sum())]),
                        weight_sum=float(sum(v.double().sum() for v in sr.values())))
    print("discriminator: oracle netD / GANLoss / feature matching == reference (values and autograd, bit-exact); same keys "
          "and init as our define_D; golden written")
    # VGG19 perceptual loss (networks.py:124-139, 474-504) with random weights (the pretrained checkpoint is a download)
    import torchvision
    _v = torchvision.models.vgg19
    networks.models.vgg19 = lambda pretrained=False, **k: _v(weights=None)
    torch.manual_seed(3)
    vl = networks.VGGLoss([])
    vx, vy = torch.randn(1, 3, 32, 64, generator=g), torch.randn(1, 3, 32, 64, generator=g)
    vsd = vl.vgg.state_dict()
    with torch.no_grad():
        assert all(torch.equal(a, b) for a, b in zip(vl.vgg(vx), dorc.vgg_forward(vsd, vx))), "oracle VGG19 != reference"
        assert torch.equal(vl(vx, vy), dorc.vgg_loss(vsd, vx, vy)), "oracle VGGLoss != reference"
    os.environ["JPDSE_VGG_RANDOM"] = "1"
    torch.manual_seed(3)
    vo = ours.Vgg19()
    assert list(vo.state_dict().keys()) == list(vsd.keys()) and all(torch.equal(vo.state_dict()[k], vsd[k]) for k in vsd)
    print("VGG19 / VGGLoss: oracle == reference (bit-exact); our Vgg19 has the same keys and (seeded) init")

    # ---------------------------------------------------------------- quantisers
    q = (torch.randn(4099, generator=g) * 3).float()
    q[:10] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, 1.4, 1.6, 0.0, -0.0, float("nan")])
    r_ref = round_mod.RoundedIdentity.apply(q)
    assert np.array_equal(qorc.rounded_identity(q.numpy()), r_ref.numpy(), equal_nan=True)
    ds = binarize.DifferentiableSign().eval()
    s_ref = ds(q)
    assert np.array_equal(qorc.sign(q.numpy()), s_ref.numpy())
    # stochastic sign: same RNG stream -> same uniform draw
    xs = torch.tanh(torch.randn(4099, generator=g))
    torch.manual_seed(99)
    ss_ref = binarize.SoftSignFunction.apply(xs)
    torch.manual_seed(99)
    u = xs.new(xs.size()).uniform_()
    assert np.array_equal(qorc.soft_sign(xs.numpy(), u.numpy()), ss_ref.numpy())
    # binarizer
    torch.manual_seed(5)
    bz = binarize.Binarizer(64, 16).eval()
    xb = torch.randn(1, 64, 8, 16, generator=g)
    with torch.no_grad():
        b_ref = bz(xb)
    assert torch.equal(qorc.binarizer_eval(xb, bz.conv.weight.detach()), b_ref)
    # S2HVQ: dyadic inputs -> every fp32 operation is exact, so indices/scores are summation-order independent
    n, code_len, csz, ncen = 6, 5, 8, 16
    cb = torch.randint(-8, 9, (ncen, csz), generator=g).float() / 4
    cb[3] = cb[1]  # duplicated center: argmin/argmax ties must resolve to the first index
    xv = torch.randint(-8, 9, (n, code_len * csz), generator=g).float() / 4
    vq = s2h_vq.S2HVQ(cb.clone(), sigma=1.5)
    with torch.no_grad():
        xm = vq._vec2mtrx(xv, code_len)
        sc_ref = vq._get_score_mtrx(xm)
        hard_ref = vq.encode(xv, code_len, train=False, raw=True)
        idx_ref = vq.encode(xv, code_len, train=False, raw=False)
        soft_ref = vq.encode(xv, code_len, train=True, raw=True)
        dec_ref = vq.decode(hard_ref)
    assert torch.equal(qorc.s2hvq_scores(xm, cb), sc_ref)
    o_idx, o_hard = qorc.s2hvq_hard(xm, cb)
    assert torch.equal(o_idx, idx_ref) and torch.equal(o_hard, hard_ref)
    assert torch.equal(qorc.s2hvq_soft(xm, cb, 1.5), soft_ref)
    assert torch.equal(qorc.s2hvq_decode(hard_ref, cb)[0], dec_ref)
    np.savez_compressed(os.path.join(GOLDEN, "quantizers.npz"), q=q.numpy(), round=r_ref.numpy(), sign=s_ref.numpy(),
                        ss_x=xs.numpy(), ss_u=u.numpy(), ss_y=ss_ref.numpy(), bin_x=xb.numpy(),
                        bin_w=bz.conv.weight.detach().numpy(), bin_y=b_ref.numpy(), vq_x=xv.numpy(), vq_cb=cb.numpy(),
                        vq_code_len=code_len, vq_sigma=1.5, vq_scores=sc_ref.numpy(), vq_hard=hard_ref.numpy(),
                        vq_index=idx_ref.numpy(), vq_soft=soft_ref.numpy(), vq_decoded=dec_ref.numpy())
    # the reference's own smoke values (round.py:17-32): 1.5 -> 2, 1.4 -> 1, 1.6 -> 2
    assert qorc.rounded_identity(np.array([1.5, 1.4, 1.6], dtype=np.float32)).tolist() == [2.0, 1.0, 2.0]
    print("quantisers: oracle == reference (bit-exact); golden written")

    # ---------------------------------------------------------------- command-line surface of the model plugin
    # (ctu/models/pix2pixHD_model.py:22-101, reached through ctu.models.get_option_setter from base_parser.py:142-144)
    import argparse
    import json
    table = {}
    for is_train in (True, False):
        ap = p2p.Pix2PixHDModel.modify_commandline_options(argparse.ArgumentParser(), is_train)
        table["train" if is_train else "test"] = sorted(
            [a.dest, type(a).__name__, None if a.type is None else a.type.__name__, a.default,
             None if a.choices is None else list(a.choices)] for a in ap._actions if a.dest != "help")
    with open(os.path.join(GOLDEN, "model_options.json"), "w") as f:
        json.dump(table, f, indent=0, sort_keys=True)
    print("model options: %d flags; golden written" % len(table["train"]))


if __name__ == "__main__":
    main()
