"""Mirror of ctu/quantizers/round.py: round-to-nearest-even forward, identity backward (round.py:8-15)."""
import torch

from ... import ops


class RoundedIdentity(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input):
        return ops.round_f32(input.contiguous())

    @staticmethod
    def backward(ctx, grad_output):
        return grad_output.clone()
