"""Forward and backward plan of the multi-scale PatchGAN discriminator on the sm_100a kernels.

Mirrors ``MultiscaleDiscriminator`` / ``NLayerDiscriminator`` of the reference
(ctu/models/pix2pixHD_networks/networks.py:371-471) as a fixed sequence of C-ABI calls on pre-allocated NHWC bf16 buffers:

    per scale i (scale 0 = the input, scale 1 = AvgPool2d(3, 2, 1, count_include_pad=False) of it, networks.py:387):
      cat(input_label, image) [+ pool]  ->  x_in   (zero border 2, 39 -> 64 stored channels)            jpdse_d_input
      layer 0      4x4 s2 conv + bias + LeakyReLU(0.2)                     ONE launch (JPDSE_EPI_BIAS_ACT epilogue)
      layers 1..n  4x4 conv (s2, last one s1) -> raw + InstanceNorm statistics in the epilogue, then
                   InstanceNorm apply + LeakyReLU -> zero-bordered feature map                jpdse_instnorm_apply_act
      output       4x4 s1 conv 512 -> 1 + bias, float32 NCHW                                    (JPDSE_EPI_BIAS_NCHW)

Every stored feature map IS the operand of the next conv (its zero border is that conv's padding) and the tensor the
feature-matching loss reads (ctu/models/pix2pixHD_model.py:746-753). Conv biases in front of an InstanceNorm cancel
exactly (forward) and have an exactly-zero gradient, as in the generator.

Backward walks the layers in reverse; per layer the LeakyReLU / InstanceNorm backward, the weight gradient (MN-major
tcgen05 GEMM, conv_wgrad.cu) and the data gradient (4x4 stride-2: four output-phase GEMMs; stride-1: full correlation).
What is computed follows what the caller needs: the discriminator's own loss needs parameter gradients and no input
gradient; the generator's GAN / feature-matching losses need the gradient w.r.t. the image channels of the input and --
because ``optimizer_D.zero_grad()`` discards them (ctu/trainers/pix2pixHD_trainer.py:73) -- no parameter gradients.
"""
import os

import torch

from . import ops
from ._lib import (CONV1X1, CONV4X4_S1, CONV4X4_S1_FULL, CONV4X4_S2, CONV4X4_S2_DGRAD, EPI_BIAS_ACT, EPI_BIAS_NCHW, EPI_RAW,
                   EPI_RAW_STATS, JpdseError)

SLOPE = 0.2  # nn.LeakyReLU(0.2, True), networks.py:430-445
PAD = 2      # int(np.ceil((4 - 1) / 2)), networks.py:429


def _round_up(x, m):
    return (x + m - 1) // m * m


class _DLayer:
    __slots__ = ("key", "conv", "dgrad", "dgrad_half", "stride", "cin", "cin_real", "cout", "in_h", "in_w", "out_h", "out_w",
                 "norm", "final", "z_conv", "z_wconv", "dz_dgrad", "dz_dgrad_half", "zero_bias", "bias_dev")


class _Slot:
    """Saved activations of one forward pass (one per live autograd graph: fake / real / ...)."""

    def __init__(self):
        self.generation = -1
        self.x_in, self.raw, self.feat, self.stats, self.final, self.z = [], [], [], [], [], []


class DiscriminatorPlan:
    def __init__(self, input_nc, ndf, n_layers, num_D, batch, height, width, device, n_slots=3, pair=False):
        """pair=True: `batch` = 2 * B images, the first B fake and the last B real, run as ONE batch (the fused-loss route):
        the generator-side backward then only walks the first half (extra data-gradient descriptors at batch B)."""
        if num_D < 1 or num_D > 2:
            raise JpdseError("jpdse_b200 discriminator supports num_D in {1, 2} (the reference's default is 2); got %d" % num_D)
        if ndf % 64:
            raise JpdseError("jpdse_b200 discriminator needs ndf %% 64 == 0 (got %d)" % ndf)
        if input_nc > 64:
            raise JpdseError("jpdse_b200 discriminator supports up to 64 input channels (got %d)" % input_nc)
        self.input_nc, self.ndf, self.n_layers, self.num_D = input_nc, ndf, n_layers, num_D
        self.B, self.H, self.W, self.device = batch, height, width, device
        if pair and batch % 2:
            raise JpdseError("a pair plan holds an even number of images")
        self.pair, self.half = pair, batch // 2
        self.c_in = 64
        self.patch_gemm = os.environ.get("JPDSE_PATCH_OUT_GEMM", "1") != "0"
        B = batch
        widths = [ndf]
        for _ in range(1, n_layers + 1):
            widths.append(min(widths[-1] * 2, 512))
        self.scales = []  # scale i (as executed): list of _DLayer; parameters come from scale{num_D-1-i}_layer{j}
        h, w = height, width
        for i in range(num_D):
            if i > 0:
                h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            layers = []
            lh, lw, cin, cin_real = h, w, self.c_in, input_nc
            for j in range(n_layers + 2):
                L = _DLayer()
                L.key = "scale%d_layer%d.0" % (num_D - 1 - i, j)
                L.final = j == n_layers + 1
                L.norm = 0 < j <= n_layers
                L.stride = 2 if j < n_layers else 1
                L.cin, L.cin_real, L.in_h, L.in_w = cin, cin_real, lh, lw
                L.cout = 1 if L.final else widths[j]
                kind = CONV4X4_S2 if L.stride == 2 else CONV4X4_S1
                if j == 0:
                    L.conv = ops.Conv(kind, EPI_BIAS_ACT, B, lh, lw, PAD, cin, cin_real, L.cout, device, out_pad=PAD, slope=SLOPE)
                elif L.final:
                    L.conv = ops.Conv(kind, EPI_BIAS_NCHW, B, lh, lw, PAD, cin, cin_real, L.cout, device)
                else:
                    L.conv = ops.Conv(kind, EPI_RAW_STATS, B, lh, lw, PAD, cin, cin_real, L.cout, device)
                L.out_h, L.out_w = L.conv.out_hw
                cst = _round_up(L.cout, 64)  # stored channels of the gradient w.r.t. this conv's output
                if L.stride == 2:
                    L.dgrad = ops.Conv(CONV4X4_S2_DGRAD, EPI_RAW, B, L.out_h, L.out_w, PAD, cst, L.cout, cin, device,
                                       out_hw=(lh, lw), cout_real=cin_real)
                else:
                    L.dgrad = ops.Conv(CONV4X4_S1_FULL, EPI_RAW, B, L.out_h, L.out_w, PAD, cst, L.cout, cin, device,
                                       cout_real=cin_real)
                L.dgrad_half = None
                if L.final and self.patch_gemm:
                    # the 1-channel output conv as 1x1 GEMMs over the stored pixels (include/jpdse_b200.h, jpdse_patch_out_*):
                    # the 512-channel activation is read once instead of once per tap
                    # (the stored pixels of an image are contiguous: to the two GEMMs that run over ALL of them the image is one
                    # row of (lh+4)*(lw+4) pixels -- whole 128-pixel tiles instead of one and a bit per 134-pixel row)
                    npx = (lh + 2 * PAD) * (lw + 2 * PAD)
                    L.z_conv = ops.Conv(CONV1X1, EPI_BIAS_NCHW, B, 1, npx, 0, cin, cin_real, 16, device)
                    L.z_wconv = ops.Conv(CONV1X1, EPI_RAW, B, 1, npx, 0, cin, cin_real, 64, device)
                    L.dz_dgrad = ops.Conv(CONV1X1, EPI_RAW, B, lh, lw, PAD, 64, 16, cin, device)
                    L.dz_dgrad_half = ops.Conv(CONV1X1, EPI_RAW, self.half, lh, lw, PAD, 64, 16, cin, device) if pair else None
                    L.zero_bias = torch.zeros(16, dtype=torch.float32, device=device)
                if pair:
                    if L.stride == 2:
                        L.dgrad_half = ops.Conv(CONV4X4_S2_DGRAD, EPI_RAW, self.half, L.out_h, L.out_w, PAD, cst, L.cout, cin,
                                                device, out_hw=(lh, lw), cout_real=cin_real)
                    else:
                        L.dgrad_half = ops.Conv(CONV4X4_S1_FULL, EPI_RAW, self.half, L.out_h, L.out_w, PAD, cst, L.cout, cin,
                                                device, cout_real=cin_real)
                layers.append(L)
                lh, lw, cin, cin_real = L.out_h, L.out_w, L.cout, L.cout
            self.scales.append(layers)
        self.flops = sum(L.conv.flops for sc in self.scales for L in sc)
        self.slots = [self._make_slot() for _ in range(n_slots)]
        self._next_slot = 0
        self._generation = 0
        self._scratch = {}

    # ------------------------------------------------------------------ buffers
    def _make_slot(self):
        s, B, dev = _Slot(), self.B, self.device
        for layers in self.scales:
            L0 = layers[0]
            s.x_in.append(ops.alloc_nhwc(B, L0.in_h + 2 * PAD, L0.in_w + 2 * PAD, self.c_in, dev))
            raws, feats, stats = [], [], []
            for L in layers:
                if L.final:
                    continue
                raws.append(torch.empty((B, L.out_h, L.out_w, L.cout), dtype=torch.bfloat16, device=dev) if L.norm else None)
                feats.append(ops.alloc_nhwc(B, L.out_h + 2 * PAD, L.out_w + 2 * PAD, L.cout, dev))
                stats.append(torch.zeros((B, L.cout, 2), dtype=torch.float64, device=dev) if L.norm else None)
            s.raw.append(raws)
            s.feat.append(feats)
            s.stats.append(stats)
            Lf = layers[-1]
            s.final.append(torch.empty((B, 1, Lf.out_h, Lf.out_w), dtype=torch.float32, device=dev))
            s.z.append(torch.empty((B, 16, Lf.in_h + 2 * PAD, Lf.in_w + 2 * PAD), dtype=torch.float32, device=dev)
                       if self.patch_gemm else None)
        return s

    def _buf(self, name, shape, dtype=torch.bfloat16, zero=False):
        key = (name, tuple(shape), dtype)
        t = self._scratch.get(key)
        if t is None:
            n = 1
            for v in shape:
                n *= v
            flat = torch.zeros(n + 2048, dtype=dtype, device=self.device)  # slack: TMA boxes may over-read the last rows
            t = flat[:n].view(*shape)
            self._scratch[key] = t
        elif zero:
            t.zero_()
        return t

    # ------------------------------------------------------------------ weights
    def load_weights(self, state_dict):
        """Pack float32 reference-layout weights (keys scale{s}_layer{j}.0.weight / .bias as in net_D.pth)."""
        for layers in self.scales:
            for L in layers:
                w = state_dict[L.key + ".weight"].detach().to(self.device).contiguous().float()
                b = state_dict[L.key + ".bias"].detach().to(self.device).contiguous().float()
                if L.final and self.patch_gemm:
                    L.z_conv.pack(w[0].permute(1, 2, 0).reshape(16, L.cin_real, 1, 1).contiguous(), L.zero_bias)
                    wt = w[0].reshape(L.cin_real, 16, 1, 1).contiguous()
                    L.dz_dgrad.pack(wt)
                    if L.dz_dgrad_half is not None:
                        L.dz_dgrad_half.pack(wt)
                    L.bias_dev = b
                    continue
                L.conv.pack(w, None if L.norm else b)  # a bias in front of an InstanceNorm cancels exactly
                L.dgrad.pack(w)
                if L.dgrad_half is not None:
                    L.dgrad_half.pack(w)

    # ------------------------------------------------------------------ forward
    def new_slot(self):
        i = self._next_slot
        self._next_slot = (i + 1) % len(self.slots)
        return i

    def forward(self, slot, a, b=None):
        """a: float32 (B,ca,H,W) [input_label, or the whole 39-channel input]; b: float32 (B,cb,H,W) [the image] or None.
        Fills slot `slot`; returns its generation (backward refuses a slot that was overwritten since)."""
        if tuple(a.shape[0:1] + a.shape[2:]) != (self.B, self.H, self.W) or a.shape[1] + (0 if b is None else b.shape[1]) != self.input_nc:
            raise JpdseError("discriminator plan built for (%d,%d,%d,%d), got %s%s" % (
                self.B, self.input_nc, self.H, self.W, tuple(a.shape), "" if b is None else " + %s" % (tuple(b.shape),)))
        s = self.slots[slot]
        with ops.stream_cached():
            for i in range(self.num_D):
                ops.d_input(a, b, s.x_in[i], self.c_in, pool=i > 0, out_pad=PAD)
        return self._run_layers(slot)

    def forward_pair_from_ids(self, slot, label, instance, fake, real, num_labels):
        """Pair plan: operands of [fake; real] straight from the class / instance ids (jpdse_d_input_ids), then the layers."""
        if not self.pair:
            raise JpdseError("forward_pair_from_ids needs a pair plan")
        h = self.half
        if tuple(fake.shape) != (h, self.input_nc - num_labels - 1, self.H, self.W) or fake.shape != real.shape:
            raise JpdseError("pair plan built for 2 x (%d,%d,%d,%d) images, got %s / %s" % (
                h, self.input_nc - num_labels - 1, self.H, self.W, tuple(fake.shape), tuple(real.shape)))
        s = self.slots[slot]
        with ops.stream_cached():
            for i in range(self.num_D):
                ops.d_input_ids(label, instance, fake, s.x_in[i][:h], real, s.x_in[i][h:], num_labels, pool=i > 0, out_pad=PAD)
        return self._run_layers(slot)

    def forward_pair(self, slot, input_label, fake, real):
        """Pair plan from the reference's float32 input_label (B,36,H,W) (routes that do not have the ids)."""
        if not self.pair:
            raise JpdseError("forward_pair needs a pair plan")
        h = self.half
        s = self.slots[slot]
        with ops.stream_cached():
            for i in range(self.num_D):
                ops.d_input(input_label, fake, s.x_in[i][:h], self.c_in, pool=i > 0, out_pad=PAD)
                ops.d_input(input_label, real, s.x_in[i][h:], self.c_in, pool=i > 0, out_pad=PAD)
        return self._run_layers(slot)

    def _run_layers(self, slot):
        s = self.slots[slot]
        self._generation += 1
        s.generation = self._generation
        with ops.stream_cached():
            for i, layers in enumerate(self.scales):
                x = s.x_in[i]
                for j, L in enumerate(layers):
                    if L.final and self.patch_gemm:
                        L.z_conv.forward(x, s.z[i])
                        ops.patch_out_gather(s.z[i], L.bias_dev, s.final[i], self.B, L.in_h, L.in_w)
                    elif L.final:
                        L.conv.forward(x, s.final[i])
                    elif not L.norm:
                        L.conv.forward(x, s.feat[i][j])
                        x = s.feat[i][j]
                    else:
                        st = s.stats[i][j]
                        st.zero_()
                        ops._count()
                        L.conv.forward(x, s.raw[i][j], st)
                        ops.instnorm_apply_act(s.raw[i][j], st, s.feat[i][j], self.B, L.out_h, L.out_w, L.cout, PAD, SLOPE)
                        x = s.feat[i][j]
        return s.generation

    def feature_nchw(self, slot, i, j):
        """Feature j of scale i as the reference returns it: float32 NCHW."""
        s = self.slots[slot]
        L = self.scales[i][j]
        if L.final:
            return s.final[i].clone()
        return ops.nhwc_pad_to_nchw(s.feat[i][j], L.cout, PAD)

    # ------------------------------------------------------------------ backward
    def backward(self, slot, generation, feat_grads, final_grads, need_input, need_params, param_grads=None, accumulate=False,
                 input_channels=None, first_half=False):
        """Backward of the pass stored in `slot`.

        feat_grads[i][j]: dense bf16 (B,h,w,C) gradient w.r.t. feature j < n_layers+1 of scale i, or None.
        final_grads[i]: float32 (B,1,h,w) gradient w.r.t. the output map of scale i, or None.
        need_params: fill `param_grads` {state-dict key: float32 tensor in the torch layout} (`accumulate`: add to them).
        need_input: returns float32 (B,c,H,W) = gradient w.r.t. input channels `input_channels` = (c0, c) (default: all).
        first_half (pair plans): only the first half of the batch (the fake images) is walked; every gradient tensor handed
        in has that batch size.
        """
        s = self.slots[slot]
        if s.generation != generation:
            raise JpdseError("jpdse_b200: this discriminator pass was overwritten by a later forward before its backward "
                             "(the plan keeps %d passes alive)" % len(self.slots))
        B = self.half if first_half else self.B
        if first_half and (not self.pair or need_params):
            raise JpdseError("first_half backward is the input-gradient walk of a pair plan")
        tag = "h" if first_half else ""
        g_in = [None] * self.num_D
        with ops.stream_cached():
            for i, layers in enumerate(self.scales):
                n = len(layers)
                g = None  # dense gradient w.r.t. the feature below the layer being processed
                for j in range(n - 1, -1, -1):
                    L = layers[j]
                    x = s.x_in[i] if j == 0 else s.feat[i][j - 1]
                    dgrad = L.dgrad_half if first_half else L.dgrad
                    if L.final and self.patch_gemm:
                        if final_grads[i] is None:
                            continue
                        # gradient of the 16 tap planes: the P operand of the weight gradient AND the data gradient's input
                        dz = self._buf("dz%s%d" % (tag, i), (B, L.in_h + 2 * PAD, L.in_w + 2 * PAD, 64))
                        ops.patch_out_scatter(final_grads[i].contiguous(), dz, B, L.in_h, L.in_w)
                        if need_params:
                            self._bias_grad(param_grads, L.key + ".bias", final_grads[i].sum().reshape(1), accumulate)
                            dw64 = self._buf("dw64_%d" % i, (64, L.cin_real, 1, 1), torch.float32)
                            L.z_wconv.wgrad(x, dz, 0, dw64)
                            dw_new = dw64[:16, :, 0, 0].t().reshape(1, L.cin_real, 4, 4)
                            key = L.key + ".weight"
                            if accumulate and key in param_grads:
                                param_grads[key] += dw_new
                            else:
                                param_grads[key] = dw_new.contiguous()
                        if j > 0 or need_input:
                            g = self._buf("g%s%d_%d" % (tag, i, j), (B, L.in_h, L.in_w, L.cin))
                            (L.dz_dgrad_half if first_half else L.dz_dgrad).forward(dz, g)
                        else:
                            g = None
                        continue
                    if L.final:
                        if final_grads[i] is None:
                            continue
                        # float32 (B,1,h,w) -> zero-bordered bf16 with 64 stored channels (the GEMMs' K granularity)
                        d_out = self._buf("dfin%s%d" % (tag, i), (B, L.out_h + 2 * PAD, L.out_w + 2 * PAD, 64))
                        ops.d_input(final_grads[i].contiguous(), None, d_out, 64, pool=False, out_pad=PAD)
                        if need_params:
                            self._bias_grad(param_grads, L.key + ".bias", final_grads[i].sum().reshape(1), accumulate)
                    else:
                        skip = feat_grads[i][j]
                        if g is None and skip is None:
                            continue  # nothing flows into this layer (and hence into none below it)
                        if g is None:
                            g, skip = skip, None
                        d_out = self._buf("dout%s%d_%d" % (tag, i, j), (B, L.out_h + 2 * PAD, L.out_w + 2 * PAD, L.cout))
                        if L.norm:
                            dy = self._buf("dy%s%d_%d" % (tag, i, j), (B, L.out_h, L.out_w, L.cout))
                            sums = self._buf("sums%s%d_%d" % (tag, i, j), (B, L.cout, 2), torch.float64, zero=True)
                            ops.instnorm_backward_reduce_act(g, 0, skip, s.raw[i][j][:B], s.stats[i][j][:B], dy, sums, B, L.out_h,
                                                             L.out_w, L.cout, SLOPE)
                            ops.instnorm_backward_apply(dy, s.raw[i][j][:B], s.stats[i][j][:B], sums, d_out, PAD, B, L.out_h,
                                                        L.out_w, L.cout)
                            if need_params:  # analytically zero (the reference's is rounding noise)
                                self._bias_grad(param_grads, L.key + ".bias", None, accumulate, L.cout)
                        else:
                            db = None
                            if need_params:
                                db = self._buf("db%d_%d" % (i, j), (L.cout,), torch.float32, zero=True)
                            ops.act_backward(g, skip, s.feat[i][j][:B], d_out, db, B, L.out_h, L.out_w, L.cout, PAD, PAD, SLOPE)
                            if need_params:
                                self._bias_grad(param_grads, L.key + ".bias", db, accumulate)
                    if need_params:
                        key = L.key + ".weight"
                        dw = param_grads.get(key)
                        fresh = dw is None
                        if fresh:
                            dw = torch.empty((L.cout, L.cin_real, 4, 4), dtype=torch.float32, device=self.device)
                            param_grads[key] = dw
                        L.conv.wgrad(x, d_out, PAD, dw, accumulate=accumulate and not fresh)
                    if j > 0 or need_input:
                        oh, ow = dgrad.out_hw
                        g = self._buf("g%s%d_%d" % (tag, i, j), (B, oh, ow, L.cin))
                        dgrad.forward(d_out, g)
                    else:
                        g = None
                g_in[i] = g
            if not need_input:
                return None
            c0, c = (0, self.input_nc) if input_channels is None else input_channels
            if g_in[0] is None:
                return torch.zeros((B, c, self.H, self.W), dtype=torch.float32, device=self.device)
            g1 = g_in[1] if self.num_D > 1 else None
            return ops.d_input_backward(g_in[0], g1, c0, c)

    def _bias_grad(self, param_grads, key, value, accumulate, n=None):
        if value is None:
            value = torch.zeros(n, dtype=torch.float32, device=self.device)
        cur = param_grads.get(key)
        if cur is None or not accumulate:
            param_grads[key] = value.clone()
        else:
            cur += value
