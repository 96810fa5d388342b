"""ORACLE -- TEST INFRASTRUCTURE ONLY. Never imported by the product package (jpd-se_b200/).

CPU restatement of the training step's discriminator side (file:line relative to /root/reference):
  MultiscaleDiscriminator / NLayerDiscriminator   ctu/models/pix2pixHD_networks/networks.py:371-471
  GANLoss (LSGAN: MSE against a constant target)  networks.py:80-122
  feature matching                                ctu/models/pix2pixHD_model.py:746-753
  VGGLoss / Vgg19                                 networks.py:124-139, 474-504
written functionally over a state dict with torch's CPU fp32 ops (the arithmetic library the reference itself calls).
`oracle/pin_against_reference.py` asserts it is bit-identical to the imported reference modules on CPU (forward values and
autograd gradients).
"""
import torch
import torch.nn.functional as F


def patchgan_forward(sd, prefix, x, n_layers=3, round_fn=None):
    """NLayerDiscriminator(getIntermFeat=True).forward (networks.py:462-468): the n_layers + 2 intermediate outputs.
    `prefix` = 'scale%d' (MultiscaleDiscriminator registers the stages as scale{i}_layer{j}, networks.py:381-383).
    round_fn: optional bf16-operand emulation hook (see generator_oracle.generator_forward)."""
    r = (lambda t: t) if round_fn is None else round_fn
    feats = []
    for j in range(n_layers + 2):
        w, b = sd["%s_layer%d.0.weight" % (prefix, j)], sd["%s_layer%d.0.bias" % (prefix, j)]
        stride = 2 if j < n_layers else 1
        norm = 0 < j <= n_layers
        if round_fn is not None and norm:
            b = None  # cancels under the affine-free InstanceNorm; the kernels skip it before rounding
        x = F.conv2d(r(x), r(w), b, stride=stride, padding=2)
        if norm:
            x = F.instance_norm(r(x), eps=1e-5)
        if j <= n_layers:
            x = F.leaky_relu(x, 0.2)
            x = r(x)
        feats.append(x)
    return feats


def discriminator_forward(sd, x, n_layers=3, num_D=2, round_fn=None):
    """MultiscaleDiscriminator.forward (networks.py:404-419): result[i] = scale{num_D-1-i} applied to the input
    downsampled i times with AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False) (networks.py:387)."""
    result = []
    for i in range(num_D):
        result.append(patchgan_forward(sd, "scale%d" % (num_D - 1 - i), x, n_layers, round_fn))
        if i != num_D - 1:
            x = F.avg_pool2d(x, 3, stride=2, padding=1, count_include_pad=False)
    return result


def gan_loss(pred, target_is_real):
    """GANLoss.__call__ with use_lsgan=True (networks.py:112-122): sum over scales of MSE(last output, 1 or 0)."""
    loss = 0
    for scale in pred:
        t = torch.full_like(scale[-1], 1.0 if target_is_real else 0.0)
        loss = loss + F.mse_loss(scale[-1], t)
    return loss


def feature_matching(pred_fake, pred_real, num_D=2):
    """pix2pixHD_model.py:746-753 (without the keep_input / raw-feature variant): L1 over every intermediate output."""
    loss = 0
    for i in range(num_D):
        for j in range(len(pred_fake[i]) - 1):
            loss = loss + (1.0 / num_D) * F.l1_loss(pred_fake[i][j], pred_real[i][j].detach())
    return loss


def discriminator_losses(sd, input_label, fake, real, n_layers=3, num_D=2, round_fn=None):
    """The discriminator half of Pix2PixHDModel.get_train_loss (pix2pixHD_model.py:715-753):
    returns (loss_G_GAN, loss_G_GAN_Feat, loss_D_real, loss_D_fake)."""
    pred_fake_pool = discriminator_forward(sd, torch.cat((input_label.detach(), fake.detach()), 1), n_layers, num_D, round_fn)
    loss_D_fake = gan_loss(pred_fake_pool, False)
    pred_real = discriminator_forward(sd, torch.cat((input_label.detach(), real.detach()), 1), n_layers, num_D, round_fn)
    loss_D_real = gan_loss(pred_real, True)
    pred_fake = discriminator_forward(sd, torch.cat((input_label, fake), 1), n_layers, num_D, round_fn)
    loss_G_GAN = gan_loss(pred_fake, True)
    return loss_G_GAN, feature_matching(pred_fake, pred_real, num_D), loss_D_real, loss_D_fake


# --------------------------------------------------------------------------------------------- VGG19 perceptual loss
VGG_CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512)  # up to conv5_1
VGG_CUTS = (2, 7, 12, 21, 30)   # torchvision feature indices where Vgg19 cuts its five slices (networks.py:479-491)
VGG_WEIGHTS = (1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0)  # VGGLoss.weights (networks.py:132)


def vgg_forward(sd, x, round_fn=None):
    """Vgg19.forward (networks.py:496-504): relu1_1, relu2_1, relu3_1, relu4_1, relu5_1 of torchvision's VGG19 `features`.
    sd keys follow the module: slice{k}.{torchvision index}.weight / .bias."""
    r = (lambda t: t) if round_fn is None else round_fn
    outs, idx, k = [], 0, 0
    x = r(x)
    for v in VGG_CFG:
        if v == "M":
            x = F.max_pool2d(x, 2, 2)
            idx += 1
        else:
            key = "slice%d.%d" % (k + 1, idx)
            x = r(F.relu(F.conv2d(x, r(sd[key + ".weight"]), sd[key + ".bias"], padding=1)))
            idx += 2
        if k < 5 and idx == VGG_CUTS[k]:
            outs.append(x)
            k += 1
    return outs


def vgg_loss(sd, x, y, round_fn=None):
    """VGGLoss.forward (networks.py:134-139)."""
    fx, fy = vgg_forward(sd, x, round_fn), vgg_forward(sd, y, round_fn)
    loss = 0
    for wgt, a, b in zip(VGG_WEIGHTS, fx, fy):
        loss = loss + wgt * F.l1_loss(a, b.detach())
    return loss
