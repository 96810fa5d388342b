"""Mirror of ``ctu.trainers.pix2pixHD_trainer.Pix2PixHDTrainer`` for the accelerated (inference) path.

Reference: ctu/trainers/pix2pixHD_trainer.py:11-30 (ctor), :113-116 (get_img) and the BaseTrainer
counters of ctu/trainers/base_trainer.py:10-12. ``step`` and the eval/rate modes need pieces that are
not on the accelerated path yet and raise.
"""
import torch
import torch.nn as nn

from ..models.pix2pixHD_model import Pix2PixHDModel


class BaseTrainer(nn.Module):
    def __init__(self, opt):
        super(BaseTrainer, self).__init__()
        self.opt = opt
        self.start_epoch = 1
        self.best_val_loss = float('inf')
        self.steps_taken = 0


class Pix2PixHDTrainer(BaseTrainer):
    def __init__(self, opt, mode):
        super(Pix2PixHDTrainer, self).__init__(opt)
        if mode not in ('train', 'test'):
            raise ValueError('Invalid trainer mode: {}'.format(mode))
        if mode == 'train':
            raise NotImplementedError('jpdse_b200 Pix2PixHDTrainer: train mode needs the generator backward '
                                      '(not implemented yet)')
        self.model = Pix2PixHDModel(opt)

    def get_img(self, x_dict):
        # pix2pixHD_trainer.py:113-116
        self.eval()
        with torch.no_grad():
            return self.model(x_dict, self.opt, mode='get_img')

    def step(self, x_dict):
        raise NotImplementedError('jpdse_b200 Pix2PixHDTrainer.step: generator backward not implemented yet')

    def get_eval_loss(self, x_dict):
        raise NotImplementedError('jpdse_b200 Pix2PixHDTrainer.get_eval_loss: not on the accelerated path yet')
