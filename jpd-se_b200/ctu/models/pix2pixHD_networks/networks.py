"""Drop-in for ``ctu.models.pix2pixHD_networks.networks``.

Same names, signatures, state-dict keys and error behaviour as the reference
(ctu/models/pix2pixHD_networks/networks.py:19-56 ``weights_init`` / ``get_norm_layer`` / ``define_G``,
:198-263 ``GlobalGenerator``, :266-305 ``ResnetBlock``), but ``GlobalGenerator.forward`` -- and, when gradients
are enabled, its backward -- run the hand-written sm_100a kernels of libjpdse_b200.so instead of ATen/cuDNN.
The generator's ``nn`` modules only hold the parameters (so ``net_G.pth`` loads unchanged and ``print(netG)``
looks the same); none of their ``forward`` methods is on the path.

The training step's other networks (SURVEY.md section 8f "next" rows, not the accelerated path) are plain
PyTorch modules with the reference's parameter names so ``net_D.pth`` loads unchanged:
``define_D`` / ``MultiscaleDiscriminator`` / ``NLayerDiscriminator`` (:58-66, :371-471), ``GANLoss`` (:80-122),
``VGGLoss`` / ``Vgg19`` (:124-139, :474-504).
"""
import functools
import os

import torch
import torch.nn as nn

from ....engine import GeneratorPlan, make_inference_plan
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


class _GeneratorFunction(torch.autograd.Function):
    """Autograd node of the whole generator: forward and backward are GeneratorPlan kernel sequences."""

    @staticmethod
    def forward(ctx, module, plan, run, *params):
        out = run(plan)
        ctx.module, ctx.plan, ctx.generation = module, plan, plan.generation
        return out.clone()

    @staticmethod
    def backward(ctx, grad_out):
        module, plan = ctx.module, ctx.plan
        if plan.generation != ctx.generation:
            raise JpdseError('jpdse_b200: the generator ran another forward before this backward; its saved '
                             'activations are gone (run one forward per backward)')
        named = list(module.named_parameters())
        shapes = {n[:-len('.weight')]: tuple(p.shape) for n, p in named if n.endswith('.weight')}
        reducer = module.grad_reducer
        if reducer is not None:
            reducer.begin([(n, p) for n, p in named])
        with torch.cuda.device(plan.device):
            grads = plan.backward(grad_out.contiguous().float(), shapes,
                                  on_grad=None if reducer is None else reducer.ready,
                                  alloc=None if reducer is None else reducer.alloc)
        out = []
        for n, p in named:
            g = grads.get(n)
            if g is None and p.requires_grad:
                # a conv bias in front of an affine-free InstanceNorm: its gradient is exactly zero
                g = reducer.alloc(n, tuple(p.shape)).zero_() if reducer is not None else torch.zeros_like(p)
                if reducer is not None:
                    reducer.ready(n, g)
            out.append(g if p.requires_grad else None)
        if reducer is not None:
            reducer.finish()
            if reducer.copy_out:  # accumulating into an existing .grad: never hand out views of the flat buffer
                out = [None if g is None else g.clone() for g in out]
        return (None, None, None) + tuple(out)


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
        if binarize and bin_before_res:
            # the reference builds ResnetBlock(binarizer_out_channels) followed by ResnetBlock(ngf * mult)
            # (networks.py:226-229): it only runs when the two widths coincide, and no shipped script enables it
            raise NotImplementedError('jpdse_b200: bin_before_res=True is outside the accelerated path (the reference '
                                      'wiring is only consistent for binarizer_out_channels == ngf * 2**n_downsampling)')
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
        self.binarizer_out_channels = None
        if binarize:  # behind the res blocks (networks.py:231-238)
            from ....ctu.quantizers.binarize import Binarizer
            model += [Binarizer(in_channels=ngf * mult, out_channels=binarizer_out_channels)]
            self.binarizer_out_channels = binarizer_out_channels
        for i in range(n_downsampling):
            mult = 2 ** (n_downsampling - i)
            in_channels = binarizer_out_channels if (i == 0 and binarize) else ngf * mult
            model += [nn.ConvTranspose2d(in_channels, int(ngf * mult / 2), kernel_size=3, stride=2, padding=1,
                                         output_padding=1),
                      norm_layer(int(ngf * mult / 2)), activation]
        model += [nn.ReflectionPad2d(3), nn.Conv2d(ngf, output_nc, kernel_size=7, padding=0, bias=True), nn.Tanh()]
        self.model = nn.Sequential(*model)
        self._plans = {}
        self._packed_version = {}
        # data-parallel hook: an object with begin / alloc / ready / finish (jpdse_b200.ddp.GradReducer) that
        # all-reduces each gradient while the rest of the backward is still running
        self.grad_reducer = None

    # ------------------------------------------------------------------ engine plumbing
    def _weights_version(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def plan_for(self, batch, height, width, device, training=False):
        """GeneratorPlan (buffers + packed weights) for this problem size, re-packed when weights change."""
        key = (batch, height, width, str(device), bool(training))
        plan = self._plans.get(key)
        if plan is None:
            if training:
                plan = GeneratorPlan(self.input_nc, self.output_nc, self.ngf, self.n_downsampling, self.n_blocks, batch,
                                     height, width, device, training=True,
                                     binarizer_out_channels=self.binarizer_out_channels)
            else:
                plan = make_inference_plan(self.input_nc, self.output_nc, self.ngf, self.n_downsampling, self.n_blocks,
                                           batch, height, width, device,
                                           binarizer_out_channels=self.binarizer_out_channels)
            # one live plan per mode: activations at batch 16 are several GB
            self._plans = {k: v for k, v in self._plans.items() if k[4] != bool(training)}
            self._plans[key] = plan
            self._packed_version.pop(key, None)
        ver = self._weights_version()
        if self._packed_version.get(key) != ver:
            plan.load_weights({k: v for k, v in self.state_dict().items()})
            self._packed_version[key] = ver
        return plan

    def _check_runnable(self, input):
        if not input.is_cuda:
            raise JpdseError('jpdse_b200 GlobalGenerator runs on a B200 only; got a %s tensor (no CPU fallback)'
                             % input.device)
        if torch.is_grad_enabled() and input.requires_grad:
            raise NotImplementedError('jpdse_b200: the generator input (labels + decoded image) takes no gradient; '
                                      'detach it (the reference path never differentiates it)')

    def _wants_grad(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def train(self, mode=True):
        # nn.Module.train flips the Binarizer to its stochastic mode; inside the generator plan it always runs the
        # deterministic sign (inference-only there), so training a binarizing generator raises in plan_for instead
        return super(GlobalGenerator, self).train(mode)

    def _run(self, batch, height, width, device, run):
        """Run `run(plan)`; with gradients enabled the call becomes one autograd node over all parameters."""
        # define_G(..., gpu_ids=[k]) puts the net on cuda:k whatever the current device is (networks.py:52-53): launch on
        # the tensors' device, like the reference's .cuda(gpu_ids[0]) path does
        with torch.cuda.device(device):
            if not self._wants_grad():
                return run(self.plan_for(batch, height, width, device)).clone()
            plan = self.plan_for(batch, height, width, device, training=True)
            return _GeneratorFunction.apply(self, plan, run, *self.parameters())

    def forward(self, input, mode='get_continuous_img'):
        if mode == 'get_continuous_img':
            self._check_runnable(input)
            B, _, H, W = input.shape
            x = input.detach().contiguous().float()
            return self._run(B, H, W, input.device, lambda plan: plan.forward_nchw(x))
        elif mode == 'get_binary_code':
            if not self.binarize:
                raise AttributeError('Generator: no binarizer found')
            self._check_runnable(input)
            B, _, H, W = input.shape
            with torch.cuda.device(input.device):
                plan = self.plan_for(B, H, W, input.device)
                return plan.binary_code_nchw(input.detach().contiguous().float()).clone()
        else:
            raise ValueError('Invalid generator mode: {}'.format(mode))

    def forward_from_maps(self, label, instance, image, num_labels, mean=(0.5, 0.5, 0.5), std=(1.0, 1.0, 1.0),
                          bad_count=None):
        """Fused preprocess + generator: skips the (B,39,H,W) float tensor of pix2pixHD_model.py:595. A uint8 `image`
        (raw decoder output) is normalised with (x/255 - mean)/std inside the input-build kernel. `bad_count`: optional
        int32 device counter of class ids outside [0, num_labels) (scatter_ raises on those in the reference)."""
        self._check_runnable(image)
        B, _, H, W = image.shape
        image = image.detach()
        return self._run(B, H, W, image.device,
                         lambda plan: plan.forward_from_maps(label, instance, image, num_labels, mean, std, bad_count))


# ================================================================================================== training-only parts
# Plain PyTorch (cuDNN) -- outside the accelerated path. Parameter names follow the
# reference so checkpoints interchange.
def define_D(input_nc, ndf, n_layers_D, norm='instance', use_sigmoid=False, num_D=1, getIntermFeat=False, gpu_ids=[]):
    # networks.py:58-66
    netD = MultiscaleDiscriminator(input_nc, ndf, n_layers_D, get_norm_layer(norm_type=norm), use_sigmoid, num_D,
                                   getIntermFeat)
    if len(gpu_ids) > 0:
        assert (torch.cuda.is_available())
        netD.cuda(gpu_ids[0])
    netD.apply(weights_init)
    return netD


def _patchgan_stages(input_nc, ndf, n_layers, norm_layer, use_sigmoid):
    """The PatchGAN as a list of stages (networks.py:428-454): 4x4 convs, pad 2, stride 2 for the first
    n_layers stages then stride 1, widths doubling up to 512, LeakyReLU(0.2), norm on all but the first / last."""
    widths = [ndf]
    for _ in range(1, n_layers + 1):
        widths.append(min(widths[-1] * 2, 512))
    stages = [[nn.Conv2d(input_nc, ndf, kernel_size=4, stride=2, padding=2), nn.LeakyReLU(0.2, True)]]
    for n in range(1, n_layers + 1):
        stride = 2 if n < n_layers else 1
        stages.append([nn.Conv2d(widths[n - 1], widths[n], kernel_size=4, stride=stride, padding=2),
                       norm_layer(widths[n]), nn.LeakyReLU(0.2, True)])
    stages.append([nn.Conv2d(widths[-1], 1, kernel_size=4, stride=1, padding=2)])
    if use_sigmoid:
        stages.append([nn.Sigmoid()])
    return stages


class NLayerDiscriminator(nn.Module):
    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_sigmoid=False, getIntermFeat=False):
        super(NLayerDiscriminator, self).__init__()
        self.getIntermFeat = getIntermFeat
        self.n_layers = n_layers
        stages = _patchgan_stages(input_nc, ndf, n_layers, norm_layer, use_sigmoid)
        if getIntermFeat:
            for n, st in enumerate(stages):
                setattr(self, 'model' + str(n), nn.Sequential(*st))
        else:
            self.model = nn.Sequential(*[m for st in stages for m in st])

    def forward(self, input):
        if not self.getIntermFeat:
            return self.model(input)
        feats, x = [], input
        for n in range(self.n_layers + 2):
            x = getattr(self, 'model' + str(n))(x)
            feats.append(x)
        return feats


class MultiscaleDiscriminator(nn.Module):
    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=nn.BatchNorm2d, use_sigmoid=False, num_D=3,
                 getIntermFeat=False):
        super(MultiscaleDiscriminator, self).__init__()
        self.num_D, self.n_layers, self.getIntermFeat = num_D, n_layers, getIntermFeat
        for i in range(num_D):
            netD = NLayerDiscriminator(input_nc, ndf, n_layers, norm_layer, use_sigmoid, getIntermFeat)
            if getIntermFeat:
                for j in range(n_layers + 2):
                    setattr(self, 'scale%d_layer%d' % (i, j), getattr(netD, 'model' + str(j)))
            else:
                setattr(self, 'layer' + str(i), netD.model)
        self.downsample = nn.AvgPool2d(3, stride=2, padding=[1, 1], count_include_pad=False)

    def singleD_forward(self, model, input, keep_input=False):
        if self.getIntermFeat:
            result = [input]
            for stage in model:
                result.append(stage(result[-1]))
            return result if keep_input else result[1:]
        return [input, model(input)] if keep_input else [model(input)]

    def forward(self, input, keep_input=False):
        result, x = [], input
        for i in range(self.num_D):
            s = self.num_D - 1 - i  # the last-registered scale sees the full-resolution input
            if self.getIntermFeat:
                model = [getattr(self, 'scale%d_layer%d' % (s, j)) for j in range(self.n_layers + 2)]
            else:
                model = getattr(self, 'layer' + str(s))
            result.append(self.singleD_forward(model, x, keep_input=keep_input))
            if i != self.num_D - 1:
                x = self.downsample(x)
        return result


class GANLoss(nn.Module):
    """networks.py:80-122: MSE (LSGAN) or BCE against a constant target, summed over the discriminator scales."""

    def __init__(self, use_lsgan=True, target_real_label=1.0, target_fake_label=0.0, tensor=torch.FloatTensor):
        super(GANLoss, self).__init__()
        self.real_label, self.fake_label = target_real_label, target_fake_label
        self.loss = nn.MSELoss() if use_lsgan else nn.BCELoss()

    def get_target_tensor(self, input, target_is_real):
        return torch.full_like(input, self.real_label if target_is_real else self.fake_label, requires_grad=False)

    def __call__(self, input, target_is_real):
        if isinstance(input[0], list):
            loss = 0
            for scale in input:
                loss = loss + self.loss(scale[-1], self.get_target_tensor(scale[-1], target_is_real))
            return loss
        return self.loss(input[-1], self.get_target_tensor(input[-1], target_is_real))


class Vgg19(nn.Module):
    """networks.py:474-504: torchvision VGG19 features cut after relu1_1, 2_1, 3_1, 4_1, 5_1, frozen."""
    CUTS = (2, 7, 12, 21, 30)

    def __init__(self, requires_grad=False):
        super(Vgg19, self).__init__()
        from torchvision import models
        # The reference calls models.vgg19(pretrained=True) (networks.py:477), which downloads on first use. Offline the
        # checkpoint must already be in the torch hub cache: silently training against a random-feature G_VGG loss
        # (weighted by lambda_feat = 10) would diverge from the reference, so a missing checkpoint RAISES. Random
        # weights are an explicit opt-in for benchmarks and tests (JPDSE_VGG_RANDOM=1; bench.py records it).
        wts = models.VGG19_Weights.IMAGENET1K_V1
        cached = os.path.join(torch.hub.get_dir(), 'checkpoints', os.path.basename(wts.url))
        self.pretrained = True
        if os.path.isfile(cached) or os.environ.get('JPDSE_ALLOW_DOWNLOAD'):
            feats = models.vgg19(weights=wts).features
        elif os.environ.get('JPDSE_VGG_RANDOM', '0') == '1':
            self.pretrained = False
            feats = models.vgg19(weights=None).features
        else:
            raise JpdseError('jpdse_b200: the pretrained VGG19 checkpoint %s is not in the torch hub cache and this box '
                             'is offline; the reference trains against vgg19(pretrained=True) (networks.py:477). Put the '
                             'file there, or set JPDSE_VGG_RANDOM=1 to run with random VGG weights (benchmarks / tests '
                             'only)' % cached)
        lo = 0
        for k, hi in enumerate(self.CUTS):
            seq = nn.Sequential()
            for x in range(lo, hi):
                seq.add_module(str(x), feats[x])
            setattr(self, 'slice%d' % (k + 1), seq)
            lo = hi
        if not requires_grad:
            for param in self.parameters():
                param.requires_grad = False

    def forward(self, X):
        out = []
        for k in range(5):
            X = getattr(self, 'slice%d' % (k + 1))(X)
            out.append(X)
        return out


class VGGLoss(nn.Module):
    def __init__(self, gpu_ids):
        super(VGGLoss, self).__init__()
        self.vgg = Vgg19().cuda() if len(gpu_ids) else Vgg19()
        self.criterion = nn.L1Loss()
        self.weights = [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0]

    def forward(self, x, y):
        x_vgg, y_vgg = self.vgg(x), self.vgg(y)
        loss = 0
        for wgt, fx, fy in zip(self.weights, x_vgg, y_vgg):
            loss = loss + wgt * self.criterion(fx, fy.detach())
        return loss
