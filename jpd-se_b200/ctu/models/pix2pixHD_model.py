"""Hot-path mirror of ``ctu.models.pix2pixHD_model.Pix2PixHDModel`` (inference side).

Follows the reference for the configuration its scripts ship (scripts/pix2pixHD_bpg_test.sh:
``--no_label_encoding --no_feat_encoding --no_generator_binarization``, no ``--sem_masking``):
  __init__      derives netG's channel counts like pix2pixHD_model.py:118-150 and calls define_G
  forward       the mode dispatcher of :231-245 (modes off the accelerated path raise)
  preprocess    :362-412 -- one-hot scatter (:376-382), get_edges (:392), cat (:394), on-device
  get_edges     :774-783
  _get_img      :508-618 -- the zero_vis / zero_sem / zero_ins switches (:583-595), concat, netG call
  get_img       :463-465

The default ``get_img`` path never materialises the reference's (B,39,H,W) float tensors: label ids,
instance ids and the image go straight into the fused input-build kernel (jpdse_build_input), which
writes the reflect-padded NHWC bf16 operand of the stem conv. ``preprocess`` / ``get_edges`` remain
available with the reference's outputs for callers (and tests) that want them.
"""
import torch
import torch.nn as nn

from ... import ops
from ..._lib import JpdseError
from .pix2pixHD_networks import networks


def _opt(opt, name, default):
    return getattr(opt, name, default)


class Pix2PixHDModel(nn.Module):
    loss_names = ('G_GAN', 'G_GAN_Feat', 'G_VGG', 'G_Distortion', 'D_real', 'D_fake')  # pix2pixHD_model.py:213

    def __init__(self, opt):
        super(Pix2PixHDModel, self).__init__()
        self.opt = opt
        self.gpu_ids = list(_opt(opt, 'gpu_ids', [0]))
        self.is_train = _opt(opt, 'is_train', False)
        self.use_features = not _opt(opt, 'no_feat', False)
        for flag, want in (('no_label_encoding', True), ('no_feat_encoding', True), ('no_generator_binarization', True),
                           ('sem_masking', False), ('no_label', False)):
            if _opt(opt, flag, want) != want:
                raise NotImplementedError('jpdse_b200 Pix2PixHDModel: option %s=%r is outside the accelerated path '
                                          '(shipped scripts use %r)' % (flag, getattr(opt, flag), want))
        if self.is_train:
            raise NotImplementedError('jpdse_b200 Pix2PixHDModel: training mode needs the generator backward '
                                      '(not implemented yet)')
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
            binarize_generator=False)

    def use_gpu(self):
        return len(self.gpu_ids) > 0

    # ------------------------------------------------------------------ dispatcher (pix2pixHD_model.py:231-245)
    def forward(self, x_dict, opt, mode):
        if mode == 'get_img':
            return self.get_img(x_dict)
        if mode in ('get_code', 'get_train_loss', 'get_eval_loss', 'get_eval_rate'):
            raise NotImplementedError('jpdse_b200 Pix2PixHDModel: mode %r is not on the accelerated path yet' % mode)
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
        if mode != 'get_continuous_img':
            raise NotImplementedError('jpdse_b200: mode %r not on the accelerated path' % mode)
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
        fake_image = self.netG.forward(input_concat)
        return fake_image, input_label

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
        image = x_dict[key].cuda(non_blocking=True).float().contiguous()
        return self.netG.forward_from_maps(label, inst, image, self.num_labels)
