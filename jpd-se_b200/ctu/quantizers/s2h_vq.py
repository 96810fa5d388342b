"""Mirror of ctu/quantizers/s2h_vq.py ``S2HVQ`` (forward passes) on the jpdse_b200 kernels.

Same constructor, properties, method names, shapes and ValueErrors as the reference
(s2h_vq.py:13-295). Scores, argmin, one-hot, softmax and the decode gather are single fused kernels: the
(n, code_len, n_center, center_size) temporary of ``_get_score_mtrx`` (:85) is never materialised.
"""
import torch
import torch.nn as nn

from ... import ops


class S2HVQ(nn.Module):
    def __init__(self, code_book, sigma=10., **kwargs):
        super(S2HVQ, self).__init__()
        assert sigma > 0, "sigma must be greater than 0, got {}".format(sigma)
        self._center_size = code_book.size(1)
        self._code_book = nn.Parameter(code_book)
        self._sigma = sigma

    # read-only views of the constructor arguments, and sigma with the reference's positivity check (s2h_vq.py:39-63)
    center_size = property(lambda self: self._center_size)
    code_book = property(lambda self: self._code_book)

    def _set_sigma(self, value):
        assert value > 0, "sigma must be greater than 0, got {}".format(value)
        self._sigma = value

    sigma = property(lambda self: self._sigma, _set_sigma)

    def _rows(self, x_mtrx):
        if torch.is_grad_enabled() and (x_mtrx.requires_grad or self._code_book.requires_grad):
            # the reference's soft quantisation is differentiable through autograd (s2h_vq.py:91-109); these kernels
            # are forward-only, so refuse instead of silently cutting the graph (no caller in the reference trains it)
            raise NotImplementedError('jpdse_b200 S2HVQ: forward-only kernels; call under torch.no_grad() '
                                      '(or detach the input and freeze the code book)')
        return x_mtrx.detach().contiguous().float().view(-1, x_mtrx.size(-1))

    def _cb(self):
        return self._code_book.detach().contiguous().float()

    def _get_score_mtrx(self, x_mtrx):
        out = ops.s2hvq_encode(self._rows(x_mtrx), self._cb(), self.sigma, want_scores=True)
        return out["scores"].view(x_mtrx.size(0), x_mtrx.size(1), -1)

    def _soft_quantize(self, x_mtrx):
        out = ops.s2hvq_encode(self._rows(x_mtrx), self._cb(), self.sigma, want_soft=True)
        return out["soft"].view(x_mtrx.size(0), x_mtrx.size(1), -1)

    def _hard_quantize(self, x_mtrx):
        out = ops.s2hvq_encode(self._rows(x_mtrx), self._cb(), self.sigma, want_one_hot=True)
        return out["one_hot"].view(x_mtrx.size(0), x_mtrx.size(1), -1)

    def _vec2mtrx(self, x, code_len):
        return x.view(-1, code_len, x.size(1) // code_len)

    def _mtrx2vec(self, x_mtrx):
        return x_mtrx.view(-1, x_mtrx.size(1) * x_mtrx.size(2))

    def _decode_mtrx(self, code_raw):
        rows = code_raw.detach().contiguous().float().view(-1, code_raw.size(-1))
        out = ops.s2hvq_decode(rows, self._cb())
        return out.view(code_raw.size(0), code_raw.size(1), -1)

    def decode(self, code_raw):
        return self._mtrx2vec(x_mtrx=self._decode_mtrx(code_raw=code_raw))

    def _encode_vctr(self, x, code_len, train=True):
        x_mtrx = self._vec2mtrx(x=x, code_len=code_len)
        return self._soft_quantize(x_mtrx=x_mtrx) if train else self._hard_quantize(x_mtrx=x_mtrx)

    def _encode_sclr(self, x, code_len, train=True):
        x_mtrx = self._vec2mtrx(x=x, code_len=code_len)
        if train:
            # torch.max of the soft scores == torch.min of the distances, ties to the first index
            code_raw = self._soft_quantize(x_mtrx)
            _, code = torch.max(code_raw, dim=-1)
            return code
        out = ops.s2hvq_encode(self._rows(x_mtrx), self._cb(), self.sigma, want_index=True)
        return out["index"].view(x_mtrx.size(0), x_mtrx.size(1))

    def encode(self, x, code_len, train=True, raw=True):
        if x.size(1) < code_len:
            raise ValueError("x.size(1) must be greater than or equal to code_len, got " +
                             "{} and {}, respectively".format(x.size(1), code_len))
        if x.size(1) % code_len != 0:
            raise ValueError("code_len must divide x.size(1), got {} and {}, respectively".format(code_len, x.size(1)))
        if x.size(1) // code_len != self.code_book.size(1):
            raise ValueError("Illegal code_len. x.size(1) // code_len must equal " +
                             "code_book.size(0), got {}, {}, and {}, respectively.".format(
                                 x.size(1), code_len, self.code_book.size(1)))
        if raw:
            return self._encode_vctr(x=x, code_len=code_len, train=train)
        return self._encode_sclr(x=x, code_len=code_len, train=train)
