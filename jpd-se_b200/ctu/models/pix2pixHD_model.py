"""Hot-path mirror of ``ctu.models.pix2pixHD_model.Pix2PixHDModel``.

Follows the reference for the configuration its scripts ship (scripts/pix2pixHD_bpg_test.sh:
``--no_label_encoding --no_feat_encoding --no_generator_binarization``, no ``--sem_masking``):
  __init__      derives netG's channel counts like pix2pixHD_model.py:118-150 and calls define_G
  forward       the mode dispatcher of :231-245 (modes off the accelerated path raise)
  preprocess    :362-412 -- one-hot scatter (:376-382), get_edges (:392), cat (:394), on-device
  get_edges     :774-783
  _get_img      :508-618 -- the zero_vis / zero_sem / zero_ins switches (:583-595), concat, netG call
  get_img       :463-465
  get_train_loss :709-771 -- fake/real discrimination, GAN + feature-matching + VGG + distortion losses; the
                generator forward/backward inside it are the sm_100a kernels (one autograd node), netD / VGG /
                the losses are stock PyTorch (SURVEY.md section 8f "next" rows)
  get_eval_loss :636-641 -- distortion after de-normalisation and uint8 truncation (ctu/utils/misc.py:64-95)
  create_optimizers :247-281, discriminate :451-460, save / load_network (base_model.py:54-97)

The default ``get_img`` path never materialises the reference's (B,39,H,W) float tensors: label ids,
instance ids and the image go straight into the fused input-build kernel (jpdse_build_input), which
writes the reflect-padded NHWC bf16 operand of the stem conv. ``preprocess`` / ``get_edges`` remain
available with the reference's outputs for callers (and tests) that want them.
"""
import os

import torch
import torch.nn as nn

from ... import ops
from ..._lib import JpdseError
from .pix2pixHD_networks import networks


def _opt(opt, name, default):
    return getattr(opt, name, default)


# The model's command-line surface (pix2pixHD_model.py:22-101): (flag, type or None for store_true, default, choices).
# Every flag of the reference is accepted so its scripts parse unchanged; the ones outside the accelerated path are
# rejected in __init__ with NotImplementedError, not silently ignored.
_FLAGS = (
    # architecture
    ('num_D', int, 2, None), ('n_layers_D', int, 3, None), ('ndf', int, 64, None), ('no_lsgan', None, False, None),
    ('pool_size', int, 0, None), ('no_instance', None, False, None), ('no_label', None, False, None),
    ('sem_masking', None, False, None), ('binary_mask', None, False, None), ('netE_groups', int, 1, None),
    ('inst_wise_pool', None, False, None), ('norm', str, 'instance', None), ('use_dropout', None, False, None),
    # objective
    ('lambda_feat', float, 10.0, None), ('lambda_distortion', float, 10.0, None), ('anneal_lambda', None, False, None),
    ('anneal_interval', int, 5000, None), ('anneal_factor', float, 5., None), ('match_raw_feat', None, False, None),
    ('no_gan_feat_loss', None, False, None), ('no_vgg_loss', None, False, None), ('no_distortion_loss', None, False, None),
    ('no_g_gan_loss', None, False, None), ('no_d_gan_loss', None, False, None),
    # data I/O
    ('data_type', int, 32, [8, 16, 32]), ('fp16', None, False, None), ('local_rank', int, 0, None), ('input_nc', int, 3, None),
    ('use_compressed', None, False, None), ('ext', str, 'jpg', ['jpg', 'j2k', 'bpg', 'webp']), ('quality', str, '100', None),
    ('zero_sem', None, False, None), ('zero_ins', None, False, None), ('zero_vis', None, False, None),
    # model I/O
    ('checkpoints_dir', str, None, None),
    # generator
    ('netG', str, 'global', None), ('ngf', int, 64, None), ('n_downsample_global', int, 4, None),
    ('n_blocks_global', int, 9, None), ('n_blocks_local', int, 3, None), ('n_local_enhancers', int, 1, None),
    ('niter_fix_global', int, 0, None),
    # feature / label encoders (outside the accelerated path)
    ('no_feat_encoding', None, False, None), ('no_feat', None, False, None), ('feat_num', int, 3, None),
    ('n_downsample_E', int, 4, None), ('nef', int, 64, None), ('use_netE_output', None, False, None),
    ('no_label_encoding', None, False, None), ('label_encoder_out_channels', int, 36, None),
    ('n_downsample_E4label', int, 4, None), ('ne4lf', int, 64, None),
    ('no_encoder_binarization', None, False, None), ('encoder_binarizer_out_channels', int, 128, None),
    ('no_label_encoder_binarization', None, False, None), ('label_encoder_binarizer_out_channels', int, 128, None),
    # generator binarisation
    ('no_generator_binarization', None, False, None), ('bin_generator_before_res', None, False, None),
    ('generator_binarizer_out_channels', int, 128, None),
)


class Pix2PixHDModel(nn.Module):
    loss_names = ('G_GAN', 'G_GAN_Feat', 'G_VGG', 'G_Distortion', 'D_real', 'D_fake')  # pix2pixHD_model.py:213

    @staticmethod
    def modify_commandline_options(parser, train):
        """pix2pixHD_model.py:22-101: same flags, types, defaults and choices (pinned by tests/golden/model_options.json)."""
        for name, typ, default, choices in _FLAGS:
            if typ is None:
                parser.add_argument('--' + name, action='store_true', default=default)
            elif choices is not None:
                parser.add_argument('--' + name, type=typ, default=default, choices=choices)
            else:
                parser.add_argument('--' + name, type=typ, default=default)
        return parser

    def __init__(self, opt):
        super(Pix2PixHDModel, self).__init__()
        self.opt = opt
        self.gpu_ids = list(_opt(opt, 'gpu_ids', [0]))
        self.is_train = _opt(opt, 'is_train', False)
        self.use_features = not _opt(opt, 'no_feat', False)
        for flag, want in (('no_label_encoding', True), ('no_feat_encoding', True), ('sem_masking', False),
                           ('no_label', False)):
            if _opt(opt, flag, want) != want:
                raise NotImplementedError('jpdse_b200 Pix2PixHDModel: option %s=%r is outside the accelerated path '
                                          '(shipped scripts use %r)' % (flag, getattr(opt, flag), want))
        if self.is_train and _opt(opt, 'niter_fix_global', 0) > 0:
            raise NotImplementedError('jpdse_b200 Pix2PixHDModel: --niter_fix_global is outside the accelerated path')
        # --fp16 in the reference = apex AMP O1 around the whole step (pix2pixHD_trainer.py:65-67,75-76). Here every
        # network of the step (generator, netD, VGG19) already computes with bf16 operands / fp32 accumulation on the
        # sm_100a kernels, with fp32 master weights and no loss scaling needed: the flag is accepted and changes nothing.
        self.amp = bool(self.is_train and _opt(opt, 'fp16', False))
        self.num_labels = opt.num_labels + 1 if _opt(opt, 'contain_dontcare_label', False) else opt.num_labels
        netG_input_nc = self.num_labels
        if not _opt(opt, 'no_instance', False):
            netG_input_nc += 1
        if self.use_features:
            netG_input_nc += _opt(opt, 'input_nc', 3)
        self.netG = networks.define_G(
            netG_input_nc, _opt(opt, 'num_out_channels', 3), _opt(opt, 'ngf', 64), _opt(opt, 'netG', 'global'),
            _opt(opt, 'n_downsample_global', 4), _opt(opt, 'n_blocks_global', 9), _opt(opt, 'n_local_enhancers', 1),
            _opt(opt, 'n_blocks_local', 3), _opt(opt, 'norm', 'instance'), gpu_ids=self.gpu_ids,
            binarize_generator=not _opt(opt, 'no_generator_binarization', True),
            bin_generator_before_res=_opt(opt, 'bin_generator_before_res', False),
            generator_binarizer_out_channels=_opt(opt, 'generator_binarizer_out_channels', 128))
        if self.is_train:
            # pix2pixHD_model.py:151-162: D sees semantics (+edge) + image
            netD_input_nc = self.num_labels + _opt(opt, 'num_out_channels', 3)
            if not _opt(opt, 'no_instance', False):
                netD_input_nc += 1
            self.netD = networks.define_D(netD_input_nc, _opt(opt, 'ndf', 64), _opt(opt, 'n_layers_D', 3),
                                          _opt(opt, 'norm', 'instance'), _opt(opt, 'no_lsgan', False),
                                          _opt(opt, 'num_D', 2), True, gpu_ids=self.gpu_ids)
        if not self.is_train or _opt(opt, 'load_model', False):
            if _opt(opt, 'checkpoints_dir', None) is not None:
                self.load_network(self.netG, 'G', opt)
                if self.is_train:
                    self.load_network(self.netD, 'D', opt)
        if self.is_train:
            if _opt(opt, 'pool_size', 0) > 0:
                raise NotImplementedError('jpdse_b200: ImagePool (pool_size > 0) is outside the accelerated path')
            self.criterionGAN = networks.GANLoss(use_lsgan=not _opt(opt, 'no_lsgan', False))
            self.criterionFeat = torch.nn.L1Loss()
            self.criterionVGG = networks.VGGLoss(self.gpu_ids)
            # netD and VGG19 run on the sm_100a kernels (jpdse_b200.discriminator / jpdse_b200.vgg)
        else:
            self.loss_names = ('G_Distortion')  # sic: a plain string in the reference (:215)
        fn = _opt(opt, 'distortion_loss_fn', 'l1')
        self.criterionDistortion = torch.nn.L1Loss() if fn == 'l1' else torch.nn.MSELoss()

    # ------------------------------------------------------------------ checkpoints (base_model.py:54-97)
    def load_network(self, network, network_label, opt):
        load_path = os.path.join(opt.checkpoints_dir, 'net_%s.pth' % network_label)
        if not os.path.isfile(load_path):
            print('%s does not exist' % load_path)
            if network_label == 'G':
                raise TypeError('generator must exist')  # the reference's bare raise('...') is a TypeError
            return
        sd = torch.load(load_path, map_location='cpu')
        try:
            network.load_state_dict(sd)
        except RuntimeError:
            own = network.state_dict()
            network.load_state_dict({k: v for k, v in sd.items() if k in own and v.size() == own[k].size()}, strict=False)
            print('pretrained network %s does not match exactly; loaded the layers that do' % network_label)

    def save_network(self, network, network_label, save_dir):
        torch.save({k: v.cpu() for k, v in network.state_dict().items()}, os.path.join(save_dir, 'net_%s.pth' % network_label))

    def save(self):
        self.save_network(self.netG, 'G', self.opt.save_dir)
        if self.is_train:
            self.save_network(self.netD, 'D', self.opt.save_dir)

    def create_optimizers(self, opt):
        # pix2pixHD_model.py:247-281: two Adams, same lr / betas
        kw = dict(lr=_opt(opt, 'lr', 0.0002), betas=(_opt(opt, 'beta1', 0.5), _opt(opt, 'beta2', 0.999)))
        # same update rule; the multi-tensor "fused" implementation makes one pass over the 730 MB of generator
        # parameters and moments instead of one pass per elementary operation (2.7 -> ~1 ms a step)
        if self.use_gpu() and os.environ.get('JPDSE_NO_FUSED_ADAM', '0') != '1':
            kw['fused'] = True
        return torch.optim.Adam(list(self.netG.parameters()), **kw), torch.optim.Adam(list(self.netD.parameters()), **kw)

    def use_gpu(self):
        return len(self.gpu_ids) > 0

    # ------------------------------------------------------------------ dispatcher (pix2pixHD_model.py:231-245)
    def forward(self, x_dict, opt, mode):
        if mode == 'get_img':
            return self.get_img(x_dict)
        if mode == 'get_train_loss':
            return self.get_train_loss(x_dict)
        if mode == 'get_eval_loss':
            return self.get_eval_loss(x_dict)
        if mode == 'get_code':
            return self.get_code(self.preprocess(x_dict))
        if mode == 'get_eval_rate':
            return self.get_eval_rate(self.preprocess(x_dict))
        raise ValueError('Invalid forward mode: {}'.format(mode))

    # ------------------------------------------------------------------ preprocessing with reference outputs
    def get_edges(self, t):
        """(B,1,H,W) instance ids -> float32 (B,1,H,W) edge map, bit-exact with pix2pixHD_model.py:774-783."""
        t = t.cuda() if not t.is_cuda else t
        B, _, H, W = t.shape
        dummy_label = torch.zeros((B, 1, H, W), dtype=torch.uint8, device=t.device)
        dummy_image = torch.zeros((B, 3, H, W), dtype=torch.float32, device=t.device)
        _, nchw = ops.build_input(dummy_label, t.contiguous(), dummy_image, 1, nhwc=False, nchw=True)
        return nchw[:, 1:2].contiguous()

    def preprocess(self, x_dict):
        """Reference-format outputs: input_label = cat(one-hot, edge) float32 (pix2pixHD_model.py:375-396)."""
        label = x_dict['label'].cuda().contiguous()
        inst = x_dict['instance'].cuda().contiguous()
        image = x_dict['image'].cuda().float().contiguous()
        if _opt(self.opt, 'no_instance', False):
            raise NotImplementedError('jpdse_b200: --no_instance is outside the accelerated path')
        bad = torch.zeros(1, dtype=torch.int32, device=image.device)
        _, nchw = ops.build_input(label, inst, image, self.num_labels, nhwc=False, nchw=True, bad_count=bad)
        if int(bad.item()) != 0:
            # the reference's scatter_ raises on an out-of-range class id
            raise RuntimeError('index out of range in label map (%d pixels outside [0,%d))' % (int(bad.item()),
                                                                                               self.num_labels))
        return {'input_label': nchw[:, :self.num_labels + 1], 'real_image': image, 'instance_ids': inst,
                '_label_ids': label}

    # ------------------------------------------------------------------ generator call (pix2pixHD_model.py:508-618)
    def _get_img(self, x_dict, mode='get_continuous_img'):
        if mode not in ('get_continuous_img', 'get_binary_code'):
            raise ValueError('Invalid forward mode: {}'.format(mode))
        if mode == 'get_binary_code' and _opt(self.opt, 'no_generator_binarization', True):
            return []  # no encoders on this path: the generator's Binarizer is the only code source (:581-582)
        input_label, real_image = x_dict['input_label'], x_dict['real_image']
        if _opt(self.opt, 'use_compressed', False):
            real_image = x_dict['compressed_img']
        feat_map = real_image
        if _opt(self.opt, 'zero_vis', False):
            feat_map = feat_map.new_zeros(feat_map.size())
        if _opt(self.opt, 'zero_sem', False):
            input_concat = torch.cat((input_label.new_zeros(input_label.size()), feat_map), dim=1)
        elif _opt(self.opt, 'zero_ins', False):
            input_label[:, -1:, ...].mul_(0.)
            input_concat = torch.cat((input_label, feat_map), dim=1)
        else:
            input_concat = torch.cat((input_label, feat_map), dim=1)
        if mode == 'get_binary_code':
            # pix2pixHD_model.py:611-617: codes flattened per image and mapped {-1,+1} -> {0,1}
            code = self.netG.forward(input_concat, mode='get_binary_code')
            return [(code.view(code.size(0), -1) + 1) / 2.]
        fake_image = self.netG.forward(input_concat)
        return fake_image, input_label

    def get_code(self, x_dict):
        # pix2pixHD_model.py:493-505
        with torch.no_grad():
            return self._get_img(x_dict, mode='get_binary_code')

    def get_eval_rate(self, x_dict):
        """pix2pixHD_model.py:466-490: per-image Shannon bits-per-pixel of the binary code and the raw bpp."""
        with torch.no_grad():
            real_image = x_dict['real_image']
            shannon_bpp_total, actual_bpp_total = 0., 0.
            for codes_ in self.get_code(x_dict):
                for j in range(real_image.size(0)):
                    code = codes_[j]
                    original_img_size = real_image[j].size(-2) * real_image[j].size(-1)
                    code_p = torch.mean(code)
                    code_entropy = - code_p * torch.log(code_p) - (1 - code_p) * torch.log(1 - code_p)
                    shannon_bpp_total = shannon_bpp_total + code_entropy * code.size(-1) / original_img_size
                    actual_bpp_total += code.size(-1) / original_img_size
            return shannon_bpp_total / real_image.size(0), actual_bpp_total / real_image.size(0)

    def get_img(self, x_dict):
        """pix2pixHD_model.py:463-465. Fast path: ids + image -> fused input build -> generator."""
        opt = self.opt
        plain = not (_opt(opt, 'zero_vis', False) or _opt(opt, 'zero_sem', False) or _opt(opt, 'zero_ins', False)
                     or _opt(opt, 'no_instance', False))
        if not plain:
            fake, _ = self._get_img(self.preprocess(x_dict))
            return fake
        label = x_dict['label'].cuda(non_blocking=True).contiguous()
        inst = x_dict['instance'].cuda(non_blocking=True).contiguous()
        key = 'compressed_img' if _opt(opt, 'use_compressed', False) else 'image'
        if key not in x_dict:
            raise JpdseError("x_dict['compressed_img'] is required with --use_compressed: libbpg is outside this "
                             "path, supply the decoded image tensor")
        image = x_dict[key].cuda(non_blocking=True)
        if image.dtype != torch.uint8:  # uint8 = the compact loader format: normalised on the device (extension)
            image = image.float()
        self._poll_bad_labels()
        out = self.netG.forward_from_maps(label, inst, image.contiguous(), self.num_labels,
                                          _opt(opt, 'normalize_mean', (0.5, 0.5, 0.5)),
                                          _opt(opt, 'normalize_std', (1.0, 1.0, 1.0)), bad_count=self._bad_counter(image.device))
        self._post_bad_labels()
        return out

    # ------------------------------------------------------------------ out-of-range class ids (scatter_ raises, :381-382)
    # The reference's one-hot scatter_ raises on an id outside [0, num_labels) (e.g. Cityscapes id 255 -> 35,
    # ctu/data/ctu_dataset.py:105). The fused input-build kernel counts those pixels in a device counter instead; the
    # fast path copies it to pinned host memory behind every call and raises LAZILY -- at the next call whose copy has
    # landed, or in check_labels() -- so no step pays a host sync for it.
    def _bad_counter(self, device):
        st = getattr(self, '_bad_state', None)
        if st is None or st['dev'].device != device:
            st = {'dev': torch.zeros(1, dtype=torch.int32, device=device),
                  'host': torch.zeros(1, dtype=torch.int32).pin_memory(), 'event': None}
            self._bad_state = st
        return st['dev']

    def _post_bad_labels(self):
        st = self._bad_state
        st['host'].copy_(st['dev'], non_blocking=True)
        st['event'] = torch.cuda.Event()
        st['event'].record(torch.cuda.current_stream(st['dev'].device))

    def _raise_bad_labels(self, n):
        st = self._bad_state
        st['dev'].zero_()
        st['host'].zero_()
        st['event'] = None
        raise RuntimeError('index out of range in label map (%d pixels outside [0,%d) in an earlier batch; the '
                           'reference\'s scatter_ raises on these)' % (n, self.num_labels))

    def _poll_bad_labels(self):
        st = getattr(self, '_bad_state', None)
        if st is not None and st['event'] is not None and st['event'].query():
            n = int(st['host'][0])
            if n:
                self._raise_bad_labels(n)

    def check_labels(self):
        """Synchronising form of the lazy check: raises if any batch since the last check had an out-of-range id."""
        st = getattr(self, '_bad_state', None)
        if st is not None:
            n = int(st['dev'].item())
            if n:
                self._raise_bad_labels(n)

    # ------------------------------------------------------------------ training (pix2pixHD_model.py:451-460, 709-771)
    def _fast_inputs(self, x_dict):
        label = x_dict['label'].cuda(non_blocking=True).contiguous()
        inst = x_dict['instance'].cuda(non_blocking=True).contiguous()
        image = x_dict['image'].cuda(non_blocking=True).float().contiguous()
        return label, inst, image

    def discriminate(self, input_label, test_image, use_pool=False, keep_input=False):
        # cuts the graph of both inputs: this is the discriminator's own loss
        input_concat = torch.cat((input_label.detach(), test_image.detach()), dim=1)
        return self.netD.forward(input_concat, keep_input)

    def get_train_loss(self, x_dict):
        opt = self.opt
        if _opt(opt, 'use_compressed', False) and 'compressed_img' not in x_dict:
            raise JpdseError("x_dict['compressed_img'] is required with --use_compressed: libbpg is outside this path, "
                             "supply the decoded image tensor")
        plain = not (_opt(opt, 'zero_vis', False) or _opt(opt, 'zero_sem', False) or _opt(opt, 'zero_ins', False)
                     or _opt(opt, 'no_instance', False) or _opt(opt, 'use_compressed', False))
        keep_input = bool(_opt(opt, 'match_raw_feat', False))
        ids = None
        if plain:
            label, inst, real_image = self._fast_inputs(x_dict)
            # out-of-range ids are counted on the device; Pix2PixHDTrainer.step checks the counter behind its own
            # .item() sync (check_labels), so the training step pays no extra synchronisation
            bad = self._bad_counter(real_image.device)
            if self._fused_d(keep_input):
                # the discriminator operands are built straight from the ids too: the reference's float32 (B,36,H,W)
                # input_label (pix2pixHD_model.py:376-396) is never materialised
                input_label, ids = None, (label, inst)
                fake_image = self.netG.forward_from_maps(label, inst, real_image, self.num_labels, bad_count=bad)
            else:
                _, nchw = ops.build_input(label, inst, real_image, self.num_labels, nhwc=False, nchw=True, bad_count=bad)
                input_label = nchw[:, :self.num_labels + 1]
                fake_image = self.netG.forward_from_maps(label, inst, real_image, self.num_labels)
        else:
            # the reference's own route (pix2pixHD_model.py:711-712): preprocess -> _get_img (zeroing switches,
            # compressed input) -> netG on the (B,39,H,W) tensor; the generator still runs on the sm_100a kernels
            pre = self.preprocess(x_dict)
            if _opt(opt, 'use_compressed', False):
                pre['compressed_img'] = x_dict['compressed_img'].cuda(non_blocking=True).float()
            real_image = pre['real_image']
            fake_image, input_label = self._get_img(pre)
        return self._losses(input_label, fake_image, real_image, keep_input, ids)

    def _fused_d(self, keep_input):
        return (not keep_input and not _opt(self.opt, 'no_lsgan', False) and _opt(self.opt, 'num_D', 2) <= 2
                and os.environ.get('JPDSE_NO_FUSED_D', '0') != '1')

    def _losses(self, input_label, fake_image, real_image, keep_input, ids=None):
        opt = self.opt
        join_vgg = self._start_vgg_loss(fake_image, real_image)
        if self._fused_d(keep_input):
            # pix2pixHD_model.py:715-753 in one kernel-side pass over [fake; real] (see MultiscaleDiscriminator.fused_losses)
            loss_G_GAN, loss_G_GAN_Feat, loss_D_real, loss_D_fake = self.netD.fused_losses(
                input_label, fake_image, real_image, ids=ids, num_labels=self.num_labels,
                d_stream=self.side_stream('dloss', fake_image.device))
        else:
            # the reference's call sequence, one autograd node per netD call
            pred_fake_pool = self.discriminate(input_label, fake_image, use_pool=True)
            loss_D_fake = self.criterionGAN(pred_fake_pool, False)
            pred_real = self.discriminate(input_label, real_image, keep_input=keep_input)
            loss_D_real = self.criterionGAN(pred_real, True)
            pred_fake = self.netD.forward(torch.cat((input_label, fake_image), dim=1), keep_input=keep_input)
            loss_G_GAN = self.criterionGAN(pred_fake, True)
            loss_G_GAN_Feat = 0.
            D_weights = 1.0 / _opt(opt, 'num_D', 2)
            for i in range(_opt(opt, 'num_D', 2)):
                for j in range(len(pred_fake[i]) - 1):
                    loss_G_GAN_Feat = loss_G_GAN_Feat + D_weights * self.criterionFeat(pred_fake[i][j], pred_real[i][j].detach())
        loss_G_distortion = self.criterionDistortion(fake_image, real_image)
        loss_G_VGG = join_vgg()
        return loss_G_GAN, loss_G_GAN_Feat, loss_G_VGG, loss_G_distortion, loss_D_real, loss_D_fake

    # ------------------------------------------------------------------ side streams of the training step
    # Autograd runs a node's backward on the CUDA stream its forward ran on. The three loss chains of a step only meet at
    # the generator output, so each gets its own stream and their kernels fill each other's wave tails:
    #   main    generator forward -> discriminator pass -> loss_G backward through D into the generator backward
    #   vgg     VGG19 forward (beside the discriminator forward) and backward (beside D's input-gradient walk)
    #   dloss   the discriminator's own loss: its parameter-gradient walk runs beside the generator backward
    # Same kernels, same values; JPDSE_TRAIN_STREAMS=0 keeps everything on the current stream.
    def side_stream(self, name, device):
        if os.environ.get('JPDSE_TRAIN_STREAMS', '1') == '0':
            return None
        streams = getattr(self, '_side_streams', None)
        if streams is None:
            streams = self._side_streams = {}
        st = streams.get((name, str(device)))
        if st is None:
            st = streams[(name, str(device))] = torch.cuda.Stream(device=device)
        return st

    def _start_vgg_loss(self, fake_image, real_image):
        """criterionVGG(fake, real) (pix2pixHD_model.py:756) enqueued on the `vgg` stream; returns a function that joins
        the stream and hands back the loss."""
        side = self.side_stream('vgg', fake_image.device) if fake_image.is_cuda else None
        if side is None:
            loss = self.criterionVGG(fake_image, real_image)
            return lambda: loss
        main = torch.cuda.current_stream(fake_image.device)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            loss = self.criterionVGG(fake_image, real_image)
        fake_image.record_stream(side)
        real_image.record_stream(side)

        def join():
            main.wait_stream(side)
            return loss
        return join

    def get_eval_loss(self, x_dict):
        """pix2pixHD_model.py:621-641: distortion AFTER de-normalisation and uint8 truncation (tensor2im,
        ctu/utils/misc.py:64-95), computed on-device by jpdse_distortion_u8 -- no GPU->CPU->numpy->GPU round trip."""
        recon = self.get_img(x_dict)
        real = x_dict['image'].cuda(non_blocking=True).float().contiguous()
        mode = 'l1' if _opt(self.opt, 'distortion_loss_fn', 'l1') == 'l1' else 'mse'
        return ops.distortion_u8(recon.contiguous(), real, mode, _opt(self.opt, 'normalize_mean', (0.5, 0.5, 0.5)),
                                 _opt(self.opt, 'normalize_std', (1.0, 1.0, 1.0))).float()
