"""Forward plan of the pix2pixHD GlobalGenerator on the sm_100a kernels.

Mirrors the 40-module Sequential of the reference (ctu/models/pix2pixHD_networks/networks.py:198-263,
ResnetBlock :266-305) as a fixed sequence of C-ABI calls on pre-allocated NHWC bf16 buffers:

    build/convert -> [7x7 stem conv -> IN+ReLU] -> n x [3x3 s2 conv -> IN+ReLU]
                  -> blocks x [3x3 conv -> IN+ReLU+reflect pad -> 3x3 conv -> IN + skip + reflect pad]
                  -> n x [ConvT 3x3 s2 -> IN+ReLU] -> [7x7 head conv + bias + tanh] (fp32 NCHW)

Every InstanceNorm's statistics come out of the producing conv's epilogue; the "IN" boxes above are
the one-read/one-write apply kernel, which also writes the reflect border the next conv needs.
Conv biases in front of an affine-free InstanceNorm cancel exactly and are not applied.
"""
import torch

from . import ops
from ._lib import (CONV3X3_PAD1, CONV3X3_S2, CONV7X7_PAD3, CONVT3X3_S2, EPI_BIAS_TANH_NCHW, EPI_RAW_STATS,
                   JpdseError)


def _round_up(x, m):
    return (x + m - 1) // m * m


class GeneratorPlan:
    """Buffers + conv descriptors for one (batch, H, W) problem size."""

    def __init__(self, input_nc, output_nc, ngf, n_downsampling, n_blocks, batch, height, width, device):
        if ngf % 64:
            raise JpdseError("jpdse_b200 generator needs ngf %% 64 == 0 (got %d)" % ngf)
        if output_nc > 128:
            raise JpdseError("output_nc > 128 is not supported")
        if n_downsampling < 1:
            raise JpdseError("n_downsampling must be >= 1")
        div = 1 << n_downsampling
        if height % div or width % div:
            raise JpdseError("image size %dx%d is not divisible by 2^%d" % (height, width, n_downsampling))
        self.input_nc, self.output_nc, self.ngf = input_nc, output_nc, ngf
        self.n_down, self.n_blocks = n_downsampling, n_blocks
        self.B, self.H, self.W = batch, height, width
        self.device = device
        self.c_in_pad = _round_up(input_nc, 8)
        B, H, W = batch, height, width

        # ---- convolutions (weights are packed later by load_weights)
        self.convs = {}  # state-dict prefix -> ops.Conv
        self.steps = []  # executable plan
        self.stem = ops.Conv(CONV7X7_PAD3, EPI_RAW_STATS, B, H, W, 3, self.c_in_pad, input_nc, ngf, device)
        self.convs["model.1"] = self.stem
        idx = 4
        self.down = []
        c, h, w = ngf, H, W
        for _ in range(n_downsampling):
            cv = ops.Conv(CONV3X3_S2, EPI_RAW_STATS, B, h, w, 0, c, c, 2 * c, device)
            self.convs["model.%d" % idx] = cv
            self.down.append(cv)
            idx += 3
            c, h, w = 2 * c, h // 2, w // 2
        self.cb, self.hb, self.wb = c, h, w  # bottleneck
        self.res = []
        for _ in range(n_blocks):
            c1 = ops.Conv(CONV3X3_PAD1, EPI_RAW_STATS, B, h, w, 1, c, c, c, device)
            c2 = ops.Conv(CONV3X3_PAD1, EPI_RAW_STATS, B, h, w, 1, c, c, c, device)
            self.convs["model.%d.conv_block.1" % idx] = c1
            self.convs["model.%d.conv_block.5" % idx] = c2
            self.res.append((c1, c2))
            idx += 1
        self.up = []
        pad_in = 1 if n_blocks > 0 else 0  # the last res block leaves a reflect border we skip over
        for i in range(n_downsampling):
            cv = ops.Conv(CONVT3X3_S2, EPI_RAW_STATS, B, h, w, pad_in if i == 0 else 0, c, c, c // 2, device)
            self.convs["model.%d" % idx] = cv
            self.up.append(cv)
            idx += 3
            c, h, w = c // 2, 2 * h, 2 * w
        self.head = ops.Conv(CONV7X7_PAD3, EPI_BIAS_TANH_NCHW, B, H, W, 3, ngf, ngf, output_nc, device)
        self.convs["model.%d" % (idx + 1)] = self.head
        self.flops = sum(cv.flops for cv in self.convs.values())

        # ---- buffers
        # largest raw conv output / largest (padded) activation, in bf16 elements
        raw_elems, act_elems = B * H * W * ngf, B * (H + 6) * (W + 6) * ngf
        self.raw = torch.empty(raw_elems, dtype=torch.bfloat16, device=device)
        self.act = [torch.zeros(act_elems + 2048, dtype=torch.bfloat16, device=device) for _ in range(3)]
        self.x0 = ops.alloc_nhwc(B, H + 6, W + 6, self.c_in_pad, device)
        n_norm = 1 + n_downsampling + 2 * n_blocks + n_downsampling
        cmax = max(ngf << n_downsampling, ngf)
        self.stats = torch.zeros((n_norm, B, cmax, 2), dtype=torch.float64, device=device)
        self.out = torch.empty((B, output_nc, H, W), dtype=torch.float32, device=device)
        self.weights_version = None

    # ---- weights
    def load_weights(self, state_dict):
        """Pack float32 reference-layout weights (keys as in net_G.pth) into the kernels' layout."""
        for prefix, cv in self.convs.items():
            w = state_dict[prefix + ".weight"]
            b = state_dict.get(prefix + ".bias") if cv is self.head else None
            if w.device != self.device:
                w = w.to(self.device)
            if b is not None and b.device != self.device:
                b = b.to(self.device)
            cv.pack(w.contiguous().float(), None if b is None else b.contiguous().float())

    def _view(self, buf, B, H, W, C):
        return buf[: B * H * W * C].view(B, H, W, C)

    def _stats(self, i, C):
        # contiguous (B, C, 2) slice of layer i's statistics
        return self.stats[i].view(-1)[: self.B * C * 2].view(self.B, C, 2)

    # ---- forward
    def forward_from_x0(self):
        """Runs the generator on self.x0 (bf16 NHWC, reflect-padded by 3); returns fp32 NCHW (B,out,H,W)."""
        B = self.B
        self.stats.zero_()
        ops._count()
        si = 0
        # stem
        c, h, w = self.ngf, self.H, self.W
        raw = self._view(self.raw, B, h, w, c)
        self.stem.forward(self.x0, raw, self._stats(si, c))
        cur = 0
        x = self._view(self.act[cur], B, h, w, c)
        ops.instnorm_apply(raw, self._stats(si, c), x, B, h, w, c, 0, True)
        si += 1
        # downsampling
        for i, cv in enumerate(self.down):
            c, h, w = 2 * c, h // 2, w // 2
            raw = self._view(self.raw, B, h, w, c)
            cv.forward(x, raw, self._stats(si, c))
            last = i == self.n_down - 1
            pad = (1 if self.n_blocks > 0 else 0) if last else 0
            cur ^= 1
            x = self._view(self.act[cur], B, h + 2 * pad, w + 2 * pad, c)
            ops.instnorm_apply(raw, self._stats(si, c), x, B, h, w, c, pad, True)
            si += 1
        # residual blocks: x (padded by 1) lives in act[cur]; t and the new x use the other two buffers
        for c1, c2 in self.res:
            raw = self._view(self.raw, B, h, w, c)
            c1.forward(x, raw, self._stats(si, c))
            t_idx = (cur + 1) % 3
            t = self._view(self.act[t_idx], B, h + 2, w + 2, c)
            ops.instnorm_apply(raw, self._stats(si, c), t, B, h, w, c, 1, True)
            si += 1
            c2.forward(t, raw, self._stats(si, c))
            n_idx = (cur + 2) % 3
            xn = self._view(self.act[n_idx], B, h + 2, w + 2, c)
            ops.instnorm_apply(raw, self._stats(si, c), xn, B, h, w, c, 1, False, residual=x)
            si += 1
            x, cur = xn, n_idx
        # upsampling
        for i, cv in enumerate(self.up):
            c, h, w = c // 2, 2 * h, 2 * w
            raw = self._view(self.raw, B, h, w, c)
            cv.forward(x, raw, self._stats(si, c))
            last = i == self.n_down - 1
            pad = 3 if last else 0
            cur = (cur + 1) % 3
            x = self._view(self.act[cur], B, h + 2 * pad, w + 2 * pad, c)
            ops.instnorm_apply(raw, self._stats(si, c), x, B, h, w, c, pad, True)
            si += 1
        # head
        self.head.forward(x, self.out)
        return self.out

    def forward_nchw(self, inp):
        """inp: float32 (B, input_nc, H, W) -- the tensor the reference feeds netG (pix2pixHD_model.py:609)."""
        if tuple(inp.shape) != (self.B, self.input_nc, self.H, self.W):
            raise JpdseError("plan built for %s, got %s" % ((self.B, self.input_nc, self.H, self.W), tuple(inp.shape)))
        ops.nchw_to_nhwc_bf16(inp, pad_reflect=3, c_pad=self.c_in_pad, out=self.x0)
        return self.forward_from_x0()

    def forward_from_maps(self, label, instance, image, num_labels):
        """Fused preprocessing path: label ids + instance ids + image -> generator output."""
        if num_labels + 4 != self.input_nc:
            raise JpdseError("num_labels + 4 must equal input_nc")
        ops.build_input(label, instance, image, num_labels, pad=3, c_pad=self.c_in_pad, out_nhwc=self.x0)
        return self.forward_from_x0()
