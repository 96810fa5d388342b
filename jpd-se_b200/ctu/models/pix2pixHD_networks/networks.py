"""Drop-in for the generator half of ``ctu.models.pix2pixHD_networks.networks``.

Same names, signatures, state-dict keys and error behaviour as the reference
(ctu/models/pix2pixHD_networks/networks.py:19-56 ``weights_init`` / ``get_norm_layer`` / ``define_G``,
:198-263 ``GlobalGenerator``, :266-305 ``ResnetBlock``), but ``GlobalGenerator.forward`` runs the
hand-written sm_100a kernels of libjpdse_b200.so instead of ATen/cuDNN. The ``nn`` modules below only
hold the parameters (so ``net_G.pth`` loads unchanged and ``print(netG)`` looks the same); none of their
``forward`` methods is on the path.
"""
import functools

import torch
import torch.nn as nn

from ....engine import GeneratorPlan
from ...._lib import JpdseError


def weights_init(m):
    # networks.py:19-25 -- N(0, 0.02) on every *Conv* weight, biases keep PyTorch's default init
    classname = m.__class__.__name__
    if classname.find('Conv') != -1:
        m.weight.data.normal_(0.0, 0.02)
    elif classname.find('BatchNorm2d') != -1:
        m.weight.data.normal_(1.0, 0.02)
        m.bias.data.fill_(0)


def get_norm_layer(norm_type='instance'):
    # networks.py:27-36; only the affine-free InstanceNorm is implemented by the kernels
    if norm_type == 'instance':
        return functools.partial(nn.InstanceNorm2d, affine=False)
    if norm_type in ('batch', 'identity'):
        raise NotImplementedError('jpdse_b200: normalization layer [%s] is outside the accelerated path' % norm_type)
    raise NotImplementedError('normalization layer [%s] is not found' % norm_type)


def define_G(input_nc, output_nc, ngf, netG, n_downsample_global=3, n_blocks_global=9, n_local_enhancers=1,
             n_blocks_local=3, norm='instance', gpu_ids=[], binarize_encoder=False, encoder_binarizer_out_channels=128,
             encoder_groups=1, binarize_generator=False, bin_generator_before_res=True,
             generator_binarizer_out_channels=128):
    norm_layer = get_norm_layer(norm_type=norm)
    if netG == 'global':
        netG = GlobalGenerator(input_nc, output_nc, ngf, n_downsample_global, n_blocks_global, norm_layer,
                               binarize=binarize_generator, bin_before_res=bin_generator_before_res,
                               binarizer_out_channels=generator_binarizer_out_channels)
    elif netG in ('local', 'encoder'):
        # LocalEnhancer / Encoder are not instantiated by the shipped scripts
        raise NotImplementedError('jpdse_b200: netG=%r is outside the accelerated path' % netG)
    else:
        raise TypeError('generator not implemented!')  # the reference's bare raise('...') is a TypeError
    if len(gpu_ids) > 0:
        assert (torch.cuda.is_available())
        netG.cuda(gpu_ids[0])
    netG.apply(weights_init)
    return netG


class ResnetBlock(nn.Module):
    """Parameter holder with the reference's key layout: conv_block.1 and conv_block.5."""

    def __init__(self, dim, padding_type, norm_layer, activation=nn.ReLU(True), use_dropout=False):
        super(ResnetBlock, self).__init__()
        if padding_type != 'reflect':
            raise NotImplementedError('padding [%s] is not implemented' % padding_type)
        if use_dropout:
            raise NotImplementedError('jpdse_b200: dropout in ResnetBlock is never enabled by the reference')
        self.conv_block = nn.Sequential(
            nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0), norm_layer(dim), activation,
            nn.ReflectionPad2d(1), nn.Conv2d(dim, dim, kernel_size=3, padding=0), norm_layer(dim))

    def forward(self, x):
        raise JpdseError('ResnetBlock is executed inside GlobalGenerator.forward by the jpdse_b200 kernels')


class GlobalGenerator(nn.Module):
    def __init__(self, input_nc, output_nc, ngf=64, n_downsampling=3, n_blocks=9, norm_layer=nn.Identity,
                 padding_type='reflect', binarize=False, binarizer_out_channels=128, bin_before_res=True):
        assert (n_blocks >= 0)
        super(GlobalGenerator, self).__init__()
        if binarize:
            raise NotImplementedError('jpdse_b200: a Binarizer inside the generator is not wired yet; '
                                      'use ctu.quantizers.binarize.Binarizer stand-alone')
        activation = nn.ReLU(True)
        self.binarize = binarize
        self.bin_before_res = bin_before_res
        self.n_downsampling = n_downsampling
        self.n_blocks = n_blocks
        self.input_nc, self.output_nc, self.ngf = input_nc, output_nc, ngf

        model = [nn.ReflectionPad2d(3), nn.Conv2d(input_nc, ngf, kernel_size=7, padding=0, bias=True), norm_layer(ngf),
                 activation]
        for i in range(n_downsampling):
            mult = 2 ** i
            model += [nn.Conv2d(ngf * mult, ngf * mult * 2, kernel_size=3, stride=2, padding=1),
                      norm_layer(ngf * mult * 2), activation]
        mult = 2 ** n_downsampling
        for i in range(n_blocks):
            model += [ResnetBlock(ngf * mult, padding_type=padding_type, activation=activation, norm_layer=norm_layer)]
        for i in range(n_downsampling):
            mult = 2 ** (n_downsampling - i)
            model += [nn.ConvTranspose2d(ngf * mult, int(ngf * mult / 2), kernel_size=3, stride=2, padding=1,
                                         output_padding=1),
                      norm_layer(int(ngf * mult / 2)), activation]
        model += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0, bias=True), nn.Tanh()]
        self.model = nn.Sequential(*model)
        self._plans = {}
        self._packed_version = {}

    # ------------------------------------------------------------------ engine plumbing
    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def plan_for(self, batch, height, width, device):
        """GeneratorPlan (buffers + packed weights) for this problem size, re-packed when weights change."""
        key = (batch, height, width, str(device))
        plan = self._plans.get(key)
        if plan is None:
            plan = GeneratorPlan(self.input_nc, self.output_nc, self.ngf, self.n_downsampling, self.n_blocks, batch,
                                 height, width, device)
            self._plans = {key: plan}  # one live plan: activations at batch 16 are several GB
            self._packed_version = {}
        ver = self._weights_version()
        if self._packed_version.get(key) != ver:
            plan.load_weights({k: v for k, v in self.state_dict().items()})
            self._packed_version[key] = ver
        return plan

    def _check_runnable(self, input):
        if not input.is_cuda:
            raise JpdseError('jpdse_b200 GlobalGenerator runs on a B200 only; got a %s tensor (no CPU fallback)'
                             % input.device)
        if torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError('jpdse_b200: generator backward is not implemented yet; call under '
                                      'torch.no_grad() (trainer.get_img does)')

    def forward(self, input, mode='get_continuous_img'):
        if mode == 'get_continuous_img':
            self._check_runnable(input)
            B, _, H, W = input.shape
            plan = self.plan_for(B, H, W, input.device)
            return plan.forward_nchw(input.contiguous().float()).clone()
        elif mode == 'get_binary_code':
            if not self.binarize:
                raise AttributeError('Generator: no binarizer found')
            raise NotImplementedError('jpdse_b200: generator binarizer not wired')
        else:
            raise ValueError('Invalid generator mode: {}'.format(mode))

    def forward_from_maps(self, label, instance, image, num_labels):
        """Fused preprocess + generator: skips the (B,39,H,W) float tensor of pix2pixHD_model.py:595."""
        self._check_runnable(image)
        B, _, H, W = image.shape
        plan = self.plan_for(B, H, W, image.device)
        return plan.forward_from_maps(label, instance, image, num_labels).clone()
