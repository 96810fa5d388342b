"""VGG19 perceptual loss on the sm_100a kernels.

Reference: ``Vgg19`` (ctu/models/pix2pixHD_networks/networks.py:474-504: torchvision VGG19 ``features`` cut after relu1_1,
relu2_1, relu3_1, relu4_1, relu5_1, frozen) and ``VGGLoss`` (:124-139: sum_k w_k * L1(vgg(x)_k, vgg(y)_k.detach()),
w = [1/32, 1/16, 1/8, 1/4, 1]), called as ``criterionVGG(fake_image, real_image)`` (ctu/models/pix2pixHD_model.py:756).

The fake and the real batch run as ONE batch of 2B images through the 13 convolutions:
    3x3 zero-pad conv + bias + ReLU = ONE implicit-GEMM launch (JPDSE_CONV3X3_PAD1 on a zero-bordered NHWC bf16 tensor,
    JPDSE_EPI_BIAS_ACT epilogue with slope 0, output written into the interior of the next conv's zero-bordered operand);
    MaxPool2d(2, 2) = jpdse_maxpool2x2; the five weighted L1 terms = jpdse_l1_pair between the two halves of a feature map.
Backward (weights are frozen: data gradients only, w.r.t. the fake half): per layer jpdse_act_backward (ReLU mask + the L1
gradient at a cut) -> JPDSE_CONV3X3_FULL data gradient; pools by jpdse_maxpool2x2_backward.
"""
import torch

from . import ops
from ._lib import CONV3X3_FULL, CONV3X3_PAD1, CONV3X3_PAD1_NARROW, EPI_BIAS_ACT, EPI_RAW, JpdseError

# torchvision vgg19 `features` up to relu5_1: (feature index, 'conv' cin cout | 'pool'), cut after indices 1, 6, 11, 20, 29
CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512)
CUTS = (2, 7, 12, 21, 30)
WEIGHTS = (1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0)  # VGGLoss.weights, networks.py:132


class _Stage:
    __slots__ = ("kind", "key", "conv", "dgrad", "cin", "cout", "h", "w", "x", "y", "cut")


class VGGPlan:
    def __init__(self, batch, height, width, device, in_channels=3):
        if height % 16 or width % 16:
            raise JpdseError("jpdse_b200 VGG19 needs an image size divisible by 16 (got %dx%d)" % (height, width))
        self.B, self.H, self.W, self.device, self.in_channels = batch, height, width, device, in_channels
        B2 = 2 * batch
        # the RGB input is stored with 8 channels (16 B per pixel); its conv takes the 3 pixels under a filter row as one
        # 64-element K block (JPDSE_CONV3X3_PAD1_NARROW: K = 3 x 64 instead of 9 x 64 with a 64-channel padding)
        if in_channels > 8:
            raise JpdseError("jpdse_b200 VGG19 takes up to 8 input channels")
        self.c_in = 8
        self.x_in = ops.alloc_nhwc(B2, height + 2, width + 2, self.c_in, device)
        self.stages = []
        h, w, cin, cin_real = height, width, self.c_in, in_channels
        x = self.x_in
        idx, k = 0, 0
        for v in CFG:
            st = _Stage()
            st.h, st.w, st.cin, st.x, st.cut = h, w, cin, x, None
            if v == "M":
                st.kind, st.cout = "pool", cin
                st.y = ops.alloc_nhwc(B2, h // 2 + 2, w // 2 + 2, cin, device)
                h, w = h // 2, w // 2
                idx += 1
            else:
                st.kind, st.cout = "conv", v
                st.key = "slice%d.%d" % (k + 1, idx)
                first = not self.stages
                st.conv = ops.Conv(CONV3X3_PAD1_NARROW if first else CONV3X3_PAD1, EPI_BIAS_ACT, B2, h, w, 1, cin, cin_real, v,
                                   device, out_pad=1, slope=0.0)
                # data gradient w.r.t. the (zero-)padded input, fake half only (the input conv's: 64 stored channels, 3 real)
                st.dgrad = ops.Conv(CONV3X3_FULL, EPI_RAW, batch, h, w, 2, v, v, 64 if first else cin, device,
                                    cout_real=cin_real)
                st.y = ops.alloc_nhwc(B2, h + 2, w + 2, v, device)
                cin, cin_real = v, v
                idx += 2
            if k < 5 and idx == CUTS[k]:
                st.cut = k
                k += 1
            x = st.y
            self.stages.append(st)
        self.flops = sum(st.conv.flops for st in self.stages if st.kind == "conv")
        self.generation = 0
        self._scratch = {}
        self._loss_w = None

    def _buf(self, name, shape):
        key = (name, tuple(shape))
        t = self._scratch.get(key)
        if t is None:
            n = 1
            for v in shape:
                n *= v
            flat = torch.zeros(n + 2048, dtype=torch.bfloat16, device=self.device)
            t = flat[:n].view(*shape)
            self._scratch[key] = t
        return t

    def load_weights(self, state_dict):
        """state_dict keys as in the reference's Vgg19 module: slice{k}.{torchvision index}.weight / .bias."""
        for st in self.stages:
            if st.kind != "conv":
                continue
            w = state_dict[st.key + ".weight"].detach().to(self.device).contiguous().float()
            b = state_dict[st.key + ".bias"].detach().to(self.device).contiguous().float()
            st.conv.pack(w, b)
            st.dgrad.pack(w)

    # ------------------------------------------------------------------ forward: features of [fake; real] and the loss
    def forward(self, fake, real):
        """fake, real: float32 (B,3,H,W). Returns the VGGLoss value (0-dim float32 tensor); keeps the activations."""
        B = self.B
        if tuple(fake.shape) != (B, self.in_channels, self.H, self.W) or fake.shape != real.shape:
            raise JpdseError("VGG plan built for %s, got %s / %s" % ((B, self.in_channels, self.H, self.W), tuple(fake.shape),
                                                                     tuple(real.shape)))
        self.generation += 1
        acc = torch.zeros(5, dtype=torch.float64, device=self.device)
        numel = [0.0] * 5
        with ops.stream_cached():
            ops.d_input(fake, None, self.x_in[:B], self.c_in, pool=False, out_pad=1)
            ops.d_input(real, None, self.x_in[B:], self.c_in, pool=False, out_pad=1)
            for st in self.stages:
                if st.kind == "conv":
                    st.conv.forward(st.x, st.y)
                else:
                    ops.maxpool2x2(st.x, st.y, 2 * B, st.h, st.w, st.cin, 1, 1)
                if st.cut is not None:
                    ops.l1_pair(st.y[:B], st.y[B:], acc[st.cut:st.cut + 1])
                    numel[st.cut] = float(B * st.cout * st.h * st.w)
        self._numel = numel
        if self._loss_w is None:
            # a per-plan constant, built once: torch.tensor(..., device=cuda) is a BLOCKING pageable copy -- inside the step
            # it held the host until this stream had drained (2.6 ms per training step with nothing being launched)
            self._loss_w = torch.tensor([WEIGHTS[k] / numel[k] for k in range(5)], dtype=torch.float64, device=self.device)
        return (acc * self._loss_w).sum().float()

    # ------------------------------------------------------------------ backward: d(loss)/d(fake)
    def backward(self, generation, g_loss):
        """g_loss: upstream gradient of the loss (0-dim / 1-element float32 device tensor). Returns float32 (B,3,H,W)."""
        if generation != self.generation:
            raise JpdseError("jpdse_b200: the VGG plan ran another forward before this backward")
        B = self.B
        scale = g_loss.detach().reshape(1).float().contiguous()
        g, g_pad = None, 0  # gradient w.r.t. the output of the stage being processed (fake half), maybe with a border
        with ops.stream_cached():
            for si in range(len(self.stages) - 1, -1, -1):
                st = self.stages[si]
                skip = None
                if st.cut is not None:
                    skip = self._buf("l1_%d" % si, (B, st.h, st.w, st.cout))
                    ops.l1_pair_backward(st.y[:B], st.y[B:], skip, scale, WEIGHTS[st.cut] / self._numel[st.cut], B, st.h, st.w,
                                         st.cout, 1)
                if st.kind == "pool":
                    dx = self._buf("pool_%d" % si, (B, st.h, st.w, st.cin))
                    ops.maxpool2x2_backward(st.x[:B], g, dx, B, st.h, st.w, st.cin, 1, g_pad=g_pad)
                    g, g_pad = dx, 0
                    continue
                if g is None:
                    g, skip = skip, None
                d_pre = self._buf("dpre_%d" % si, (B, st.h + 4, st.w + 4, st.cout))
                ops.act_backward(g, skip, st.y[:B], d_pre, None, B, st.h, st.w, st.cout, 1, 2, 0.0, g_pad=g_pad)
                g = self._buf("g_%d" % si, (B, st.h + 2, st.w + 2, st.dgrad.cout))
                st.dgrad.forward(d_pre, g)
                g_pad = 1
            return ops.nhwc_pad_to_nchw(g, self.in_channels, 1)
