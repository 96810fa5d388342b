"""Mirror of ctu/quantizers/binarize.py (binary quantiser of Toderici et al.).

  SoftSignFunction   stochastic +-1 with straight-through gradient (binarize.py:13-28); the uniform noise is
                     drawn with torch's generator exactly like the reference (``input.new(size).uniform_()``)
                     and the thresholding runs in jpdse_softsign_f32
  DifferentiableSign train = SoftSignFunction, eval = sign (binarize.py:31-41)
  Binarizer          1x1 conv (no bias) -> tanh -> sign (binarize.py:44-65); in eval mode the three are ONE
                     tcgen05 implicit-GEMM launch with a sign epilogue
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from ... import ops
from ..._lib import CONV1X1, EPI_SIGN_NCHW, JpdseError


class SoftSignFunction(Function):
    @staticmethod
    def forward(ctx, input):
        prob = input.new(input.size()).uniform_()
        return ops.softsign_f32(input.contiguous().float(), prob.float())

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output


class DifferentiableSign(nn.Module):
    def __init__(self):
        super(DifferentiableSign, self).__init__()

    def forward(self, x):
        # Apply quantization noise while only training
        if self.training:
            return SoftSignFunction.apply(x)
        return ops.sign_f32(x.contiguous().float())


class Binarizer(nn.Module):
    def __init__(self, in_channels, out_channels, groups=1):
        super(Binarizer, self).__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False, groups=groups)
        self.differentiable_sign = DifferentiableSign()
        self._conv_cache = None

    def forward(self, x):
        if self.training or self.conv.groups != 1:
            raise NotImplementedError('jpdse_b200 Binarizer: only the eval-mode forward with groups=1 is on the '
                                      'accelerated path (training needs the conv backward)')
        if not x.is_cuda:
            raise JpdseError('jpdse_b200 Binarizer runs on a B200 only (no CPU fallback)')
        B, C, H, W = x.shape
        cout = self.conv.out_channels
        key = (B, C, H, W, self.conv.weight.data_ptr(), self.conv.weight._version)
        if self._conv_cache is None or self._conv_cache[0] != key:
            cv = ops.Conv(CONV1X1, EPI_SIGN_NCHW, B, H, W, 0, C, C, cout, x.device)
            cv.pack(self.conv.weight.detach().float().contiguous())
            self._conv_cache = (key, cv)
        cv = self._conv_cache[1]
        xh = ops.nchw_to_nhwc_bf16(x.contiguous().float())
        y = torch.empty((B, cout, H, W), dtype=torch.float32, device=x.device)
        cv.forward(xh, y)
        return y
