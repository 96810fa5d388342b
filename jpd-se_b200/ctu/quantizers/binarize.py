"""Mirror of ctu/quantizers/binarize.py (binary quantiser of Toderici et al.).

  SoftSignFunction   stochastic +-1 with straight-through gradient (binarize.py:13-28); the uniform noise is
                     drawn with torch's generator exactly like the reference (``input.new(size).uniform_()``)
                     and the thresholding runs in jpdse_softsign_f32
  DifferentiableSign train = SoftSignFunction, eval = sign (binarize.py:31-41)
  Binarizer          1x1 conv (no bias) -> tanh -> sign (binarize.py:44-65); in eval mode the three are ONE
                     tcgen05 implicit-GEMM launch with a sign epilogue. In train mode the conv is the same kernel with
                     a raw bf16 epilogue, tanh + stochastic sign one bandwidth kernel, and the backward (identity
                     through the sign, tanh', conv weight / data gradients) runs on jpdse_conv_wgrad and a 1x1 conv
                     with the transposed weight -- one autograd node.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from ... import ops
from ..._lib import CONV1X1, EPI_RAW, EPI_SIGN_NCHW, JpdseError, check
from ... import _lib


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


class _BinarizerTrainFunction(Function):
    """conv1x1 -> tanh -> SoftSign (train mode) and its backward on the jpdse_b200 kernels."""

    @staticmethod
    def forward(ctx, x, weight):
        lib = _lib.load()
        B, C, H, W = x.shape
        cout = weight.shape[0]
        dev = x.device
        cv = ops.Conv(CONV1X1, EPI_RAW, B, H, W, 0, C, C, cout, dev)
        w32 = weight.detach().float().contiguous()
        cv.pack(w32)
        xh = ops.nchw_to_nhwc_bf16(x.detach().contiguous().float())
        pre = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device=dev)
        cv.forward(xh, pre)
        noise = x.new(B, cout, H, W).float().uniform_()  # same generator call as the reference's input.new(size).uniform_()
        y = torch.empty((B, cout, H, W), dtype=torch.float32, device=dev)
        t = torch.empty_like(y)
        check(lib.jpdse_binarizer_train_forward(ops._ptr(pre), ops._ptr(noise), ops._ptr(y), ops._ptr(t), B, cout, H, W,
                                                ops._stream()))
        ops._count()
        ctx.save_for_backward(xh, t, w32)
        ctx.cv = cv
        ctx.noise = noise
        return y

    @staticmethod
    def backward(ctx, grad_y):
        lib = _lib.load()
        xh, t, w32 = ctx.saved_tensors
        B, cout, H, W = t.shape
        C = xh.shape[-1]
        dev = t.device
        dpre = torch.empty((B, H, W, cout), dtype=torch.bfloat16, device=dev)
        check(lib.jpdse_binarizer_train_backward(ops._ptr(grad_y.contiguous().float()), ops._ptr(t), ops._ptr(dpre), B, cout,
                                                 H, W, ops._stream()))
        ops._count()
        dw = torch.empty((cout, C, 1, 1), dtype=torch.float32, device=dev)
        ctx.cv.wgrad(xh, dpre, 0, dw)
        # data gradient of a 1x1 conv = 1x1 conv with the transposed weight
        cvt = ops.Conv(CONV1X1, EPI_RAW, B, H, W, 0, cout, cout, C, dev)
        cvt.pack(w32.view(cout, C).t().contiguous().view(C, cout, 1, 1))
        dxh = torch.empty((B, H, W, C), dtype=torch.bfloat16, device=dev)
        cvt.forward(dpre, dxh)
        return ops.nhwc_bf16_to_nchw(dxh), dw


class Binarizer(nn.Module):
    def __init__(self, in_channels, out_channels, groups=1):
        super(Binarizer, self).__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=1, bias=False, groups=groups)
        self.differentiable_sign = DifferentiableSign()
        self._conv_cache = None

    def forward(self, x):
        if self.conv.groups != 1:
            raise NotImplementedError('jpdse_b200 Binarizer: grouped 1x1 convs are outside the accelerated path')
        if not x.is_cuda:
            raise JpdseError('jpdse_b200 Binarizer runs on a B200 only (no CPU fallback)')
        if self.training:
            if x.shape[1] % 64 or self.conv.out_channels % 64:
                raise JpdseError('jpdse_b200 Binarizer (train mode): channel counts must be multiples of 64')
            return _BinarizerTrainFunction.apply(x, self.conv.weight)
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
