"""Mirror of ``ctu.trainers.pix2pixHD_trainer.Pix2PixHDTrainer``.

Reference: ctu/trainers/pix2pixHD_trainer.py:11-30 (ctor), :42-85 (step), :88-116 (eval / get_img), :119-141 (save)
and the BaseTrainer counters of ctu/trainers/base_trainer.py:6-24. ``step`` keeps the reference's order -- generator
loss backward + Adam step, then discriminator loss backward + Adam step -- with the generator forward/backward on the
sm_100a kernels. Under ``torch.distributed`` (one process per GPU) the generator gradients are all-reduced while the
backward is still running (jpdse_b200.ddp.GradReducer) and the discriminator gradients in one flat call.
"""
import os

import torch
import torch.nn as nn

from ... import ddp
from ..models.pix2pixHD_model import Pix2PixHDModel


def _opt(opt, name, default):
    return getattr(opt, name, default)


class BaseTrainer(nn.Module):
    def __init__(self, opt, mode='train'):
        super(BaseTrainer, self).__init__()
        self.opt = opt
        if mode == 'train':
            self.steps_taken = 0
            self.start_epoch = 0
            self.best_val_loss = 1e12
            if _opt(opt, 'tf_log', False):
                raise NotImplementedError('jpdse_b200: --tf_log (tensorflow summaries) is outside the accelerated path')
        elif mode != 'test':
            raise ValueError('Invalid trainer mode: {}'.format(mode))
        self.mode = mode

    def load(self):
        pass

    def log_loss_values(self, loss_dict):
        # base_trainer.py:80-88 writes tensorflow summaries; train.py:121-123 only calls it under --tf_log, which the
        # constructor above already refuses
        raise NotImplementedError('jpdse_b200: log_loss_values needs --tf_log (tensorflow summaries), outside the accelerated path')


class Pix2PixHDTrainer(BaseTrainer):
    def __init__(self, opt, mode='train'):
        super(Pix2PixHDTrainer, self).__init__(opt, mode)
        self.model = Pix2PixHDModel(opt)
        if len(_opt(opt, 'gpu_ids', [0])) > 0:
            self.model = self.model.cuda()
        if mode == 'train':
            self.optimizer_G, self.optimizer_D = self.model.create_optimizers(opt)
            if _opt(opt, 'schedule_lr', False):
                from torch.optim.lr_scheduler import ReduceLROnPlateau
                kw = dict(factor=_opt(opt, 'lr_decay_factor', .1), patience=_opt(opt, 'lr_decay_patience', 5))
                self.scheduler_G = ReduceLROnPlateau(self.optimizer_G, 'min', **kw)
                self.scheduler_D = ReduceLROnPlateau(self.optimizer_D, 'min', **kw)
            self.lambda_distortion_weight = 1.
            if ddp._world() > 1:
                ddp.broadcast_parameters(self.model.netG)
                ddp.broadcast_parameters(self.model.netD)
                self.model.netG.grad_reducer = ddp.GradReducer()

    def _get_train_loss(self, x_dict):
        return self.model(x_dict, self.opt, mode='get_train_loss')

    def scheduler_step(self, val_loss_value):
        self.scheduler_G.step(val_loss_value)
        self.scheduler_D.step(val_loss_value)

    def step(self, x_dict):
        # pix2pixHD_trainer.py:42-85
        opt = self.opt
        self.train()
        loss_dict = dict(zip(self.model.loss_names, self._get_train_loss(x_dict)))

        def term(name, flag, scale=1.0):
            if _opt(opt, flag, False):
                return loss_dict[name].new_zeros(1, requires_grad=True)
            return loss_dict[name] * scale if scale != 1.0 else loss_dict[name]

        loss_D = (loss_dict['D_fake'] + loss_dict['D_real']) * 0.5 if not _opt(opt, 'no_d_gan_loss', False) \
            else loss_dict['D_fake'].new_zeros(1, requires_grad=True)
        lam_feat = _opt(opt, 'lambda_feat', 10.0)
        loss_G = (term('G_GAN', 'no_g_gan_loss') + term('G_VGG', 'no_vgg_loss', lam_feat)
                  + term('G_GAN_Feat', 'no_gan_feat_loss', lam_feat)
                  + term('G_Distortion', 'no_distortion_loss',
                         _opt(opt, 'lambda_distortion', 10.0) * self.lambda_distortion_weight))
        if not _opt(opt, 'quiet', False):
            print('g_gan: {:.4f}, g_gan_feat_match: {:.4f}, g_vgg: {:.4f}, g_distortion ({}): {:.4f}, d_real: {:.4f}, '
                  'd_fake: {:.4f}'.format(loss_dict['G_GAN'].item(), loss_dict['G_GAN_Feat'].item(),
                                          loss_dict['G_VGG'].item(), _opt(opt, 'distortion_loss_fn', 'l1'),
                                          loss_dict['G_Distortion'].item(), loss_dict['D_real'].item(),
                                          loss_dict['D_fake'].item()))
        d_stream = self.model.side_stream('dloss', loss_G.device) if loss_G.is_cuda and self.model._fused_d(False) else None
        if d_stream is None:
            # the reference's order (pix2pixHD_trainer.py:64-78)
            self.optimizer_G.zero_grad()
            loss_G.backward()  # generator gradients come back already averaged over the ranks
            self.optimizer_G.step()
            self.optimizer_D.zero_grad()  # drops the netD gradients loss_G.backward() produced
            loss_D.backward()
            ddp.allreduce_grads(self.model.netD.parameters())
            self.optimizer_D.step()
        else:
            # Same two updates, other streams: the discriminator's gradients depend on nothing the generator update
            # touches (both losses come from ONE forward, and the fused route's loss_G.backward() deposits no netD
            # gradients to drop), so loss_D.backward() and the all-reduce of its 22 MB of gradients are enqueued FIRST, on
            # their own stream, and run beside the generator's backward + Adam; the discriminator's Adam step comes last.
            # (JPDSE_D_BACKWARD_FIRST=0 issues the generator's backward first -- same streams, same results; measured
            # 18.30 vs 18.17 ms per step, so the discriminator stays first.)
            main = torch.cuda.current_stream(loss_G.device)
            d_stream.wait_stream(main)
            d_first = os.environ.get('JPDSE_D_BACKWARD_FIRST', '1') != '0'
            self.optimizer_D.zero_grad()

            def d_backward():
                with torch.cuda.stream(d_stream):
                    loss_D.backward()
                    ddp.allreduce_grads(self.model.netD.parameters())

            if d_first:
                d_backward()
            self.optimizer_G.zero_grad()
            loss_G.backward()  # generator gradients come back already averaged over the ranks
            if not d_first:
                d_backward()
            self.optimizer_G.step()
            main.wait_stream(d_stream)
            self.optimizer_D.step()
        self.steps_taken += 1
        if _opt(opt, 'anneal_lambda', False) and not (self.steps_taken % _opt(opt, 'anneal_interval', 5000)):
            self.lambda_distortion_weight *= _opt(opt, 'anneal_factor', 5.)
        value = loss_dict['G_Distortion'].item()
        self.model.check_labels()  # the device is already synchronised by .item(): raise like the reference's scatter_
        return value

    def get_eval_loss(self, x_dict):
        self.eval()
        with torch.no_grad():
            value = self.model(x_dict, self.opt, mode='get_eval_loss').item()
        self.model.check_labels()
        return value

    def get_img(self, x_dict):
        # pix2pixHD_trainer.py:113-116
        self.eval()
        with torch.no_grad():
            return self.model(x_dict, self.opt, mode='get_img')

    def _get_code(self, x_dict):
        self.eval()
        with torch.no_grad():
            return self.model(x_dict, self.opt, mode='get_code')

    def get_code(self, x_dict):
        # pix2pixHD_trainer.py:100-103 (raises on an empty list like the reference's torch.cat)
        return torch.cat([c for c in self._get_code(x_dict) if c is not None], dim=-1)

    def get_eval_rate(self, x_dict):
        # pix2pixHD_trainer.py:106-110
        self.eval()
        with torch.no_grad():
            return self.model(x_dict, self.opt, mode='get_eval_rate')

    def save(self, epoch, val_loss_value):
        # pix2pixHD_trainer.py:119-141
        self.best_val_loss = val_loss_value
        states = {'epoch': epoch, 'steps_taken': self.steps_taken,
                  'optimizer_G_state_dict': self.optimizer_G.state_dict(),
                  'optimizer_D_state_dict': self.optimizer_D.state_dict(), 'best_val_loss': self.best_val_loss}
        if _opt(self.opt, 'schedule_lr', False):
            states['scheduler_G_state_dict'] = self.scheduler_G.state_dict()
            states['scheduler_D_state_dict'] = self.scheduler_D.state_dict()
        if _opt(self.opt, 'anneal_lambda', False):
            states['lambda_distortion_weight'] = self.lambda_distortion_weight
        torch.save(states, os.path.join(self.opt.save_dir, 'stats_and_optim.pt'))
        self.model.save()

    def load(self):
        path = os.path.join(self.opt.checkpoints_dir, "stats_and_optim.pt")
        states = torch.load(path, map_location='cpu')
        if self.mode == 'train':
            self.start_epoch = states['epoch'] + 1
            self.steps_taken = states['steps_taken']
            self.best_val_loss = states['best_val_loss']
            self.optimizer_G.load_state_dict(states['optimizer_G_state_dict'])
            self.optimizer_D.load_state_dict(states['optimizer_D_state_dict'])
            # pix2pixHD_trainer.py:158-175: schedulers under --schedule_lr, lambda weight under --anneal_lambda, both
            # tolerant of checkpoints written without them
            if _opt(self.opt, 'schedule_lr', False):
                try:
                    self.scheduler_G.load_state_dict(states['scheduler_G_state_dict'])
                    self.scheduler_D.load_state_dict(states['scheduler_D_state_dict'])
                except KeyError:
                    print('Found no saved learning rate schedulers. New ones will be constructed...')
            if _opt(self.opt, 'anneal_lambda', False):
                try:
                    self.lambda_distortion_weight = states['lambda_distortion_weight']
                except KeyError:
                    print('Found no saved lambda_distortion_weight. Resetting this weight to 1...')
