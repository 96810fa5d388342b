"""Forward and backward plans of the pix2pixHD GlobalGenerator on the sm_100a kernels.

Mirrors the 40-module Sequential of the reference (ctu/models/pix2pixHD_networks/networks.py:198-263,
ResnetBlock :266-305) as a fixed sequence of C-ABI calls on pre-allocated NHWC bf16 buffers:

    build/convert -> [7x7 stem conv -> IN+ReLU] -> n x [3x3 s2 conv -> IN+ReLU]
                  -> blocks x [3x3 conv -> IN+ReLU+reflect pad -> 3x3 conv -> IN + skip + reflect pad]
                  -> n x [ConvT 3x3 s2 -> IN+ReLU] -> [7x7 head conv + bias + tanh] (fp32 NCHW)

Every InstanceNorm's statistics come out of the producing conv's epilogue; the "IN" boxes above are
the one-read/one-write apply kernel, which also writes the reflect border the next conv needs.
Conv biases in front of an affine-free InstanceNorm cancel exactly and are not applied.

Backward (``training=True`` plans; what autograd does for ``loss_G.backward()`` in the reference,
ctu/trainers/pix2pixHD_trainer.py:69) walks the same layers in reverse. Per layer:

    g (grad w.r.t. the layer's reflect-padded output) --reduce--> dy = relu-mask * (fold(g) + skip), (S1, S2)
      --apply--> dx = IN backward, zero-bordered --wgrad--> dW      (MN-major tcgen05 GEMM, conv_wgrad.cu)
                                                 --dgrad--> g of the previous layer (the forward igemm kernel with
                                                            role-swapped kinds: s2 <-> ConvT, 3x3/7x7 -> FULL)
"""
import os

import torch

from . import ops
from ._lib import (CONV1X1, CONV3X3_FULL, CONV3X3_FULL_SHARED, CONV3X3_PAD1, CONV3X3_S2, CONV7X7_FULL, CONV7X7_PAD3,
                   CONVT3X3_S2, PAD_SHARED,
                   EPI_BIAS_TANH_NCHW, EPI_RAW, EPI_RAW_STATS, EPI_SIGN_NCHW, JpdseError)


def _round_up(x, m):
    return (x + m - 1) // m * m


class _Layer:
    """One conv + InstanceNorm stage of the plan (forward record used by backward)."""
    __slots__ = ("name", "conv", "dgrad", "x_in", "raw", "stats", "out", "out_pad", "h", "w", "c", "relu", "residual",
                 "consumer_reads_border")


class GeneratorPlan:
    """Buffers + conv descriptors for one (batch, H, W) problem size."""

    def __init__(self, input_nc, output_nc, ngf, n_downsampling, n_blocks, batch, height, width, device,
                 training=False, binarizer_out_channels=None, out=None):
        if ngf % 64:
            raise JpdseError("jpdse_b200 generator needs ngf %% 64 == 0 (got %d)" % ngf)
        if output_nc > 128:
            raise JpdseError("output_nc > 128 is not supported")
        if n_downsampling < 1:
            raise JpdseError("n_downsampling must be >= 1")
        div = 1 << n_downsampling
        if height % div or width % div:
            raise JpdseError("image size %dx%d is not divisible by 2^%d" % (height, width, n_downsampling))
        if training and (ngf != 64 or output_nc > 8 or _round_up(input_nc, 8) != 40):
            raise JpdseError("jpdse_b200 generator backward supports ngf == 64, output_nc <= 8 and 33..40 input channels")
        self.input_nc, self.output_nc, self.ngf = input_nc, output_nc, ngf
        self.n_down, self.n_blocks = n_downsampling, n_blocks
        self.B, self.H, self.W = batch, height, width
        self.device = device
        self.training = training
        self.shared_border = os.environ.get("JPDSE_SHARED_BORDER", "1") != "0"  # ResnetBlock gradient layout (backward)
        self.dgrad_first = os.environ.get("JPDSE_DGRAD_FIRST", "1") != "0"      # launch order of the two gradients of a conv
        self.c_in_pad = _round_up(input_nc, 8)
        B, H, W = batch, height, width

        # ---- convolutions (weights are packed later by load_weights)
        self.convs = {}   # state-dict prefix -> forward ops.Conv
        self.dgrads = {}  # state-dict prefix -> data-gradient ops.Conv (training plans; same weight tensor, other packing)
        self.stem = ops.Conv(CONV7X7_PAD3, EPI_RAW_STATS, B, H, W, 3, self.c_in_pad, input_nc, ngf, device)
        self.convs["model.1"] = self.stem
        idx = 4
        self.down = []
        c, h, w = ngf, H, W
        for _ in range(n_downsampling):
            name = "model.%d" % idx
            cv = ops.Conv(CONV3X3_S2, EPI_RAW_STATS, B, h, w, 0, c, c, 2 * c, device)
            self.convs[name] = cv
            if training:  # dgrad of a stride-2 conv == ConvTranspose forward on the same weight memory
                self.dgrads[name] = ops.Conv(CONVT3X3_S2, EPI_RAW, B, h // 2, w // 2, 0, 2 * c, 2 * c, c, device)
            self.down.append((name, cv))
            idx += 3
            c, h, w = 2 * c, h // 2, w // 2
        self.cb, self.hb, self.wb = c, h, w  # bottleneck
        self.res = []
        for _ in range(n_blocks):
            n1, n2 = "model.%d.conv_block.1" % idx, "model.%d.conv_block.5" % idx
            c1 = ops.Conv(CONV3X3_PAD1, EPI_RAW_STATS, B, h, w, 1, c, c, c, device)
            c2 = ops.Conv(CONV3X3_PAD1, EPI_RAW_STATS, B, h, w, 1, c, c, c, device)
            self.convs[n1], self.convs[n2] = c1, c2
            if training:
                # data gradients of the ResnetBlock convs read dx in the shared-border layout (JPDSE_PAD_SHARED): M is exactly
                # the B*(h+2)*(w+2) output pixels -- 144 tiles instead of 152 at batch 2, one wave on 148 SMs instead of two
                full = CONV3X3_FULL_SHARED if self.shared_border else CONV3X3_FULL
                self.dgrads[n1] = ops.Conv(full, EPI_RAW, B, h, w, 2, c, c, c, device)
                self.dgrads[n2] = ops.Conv(full, EPI_RAW, B, h, w, 2, c, c, c, device)
            self.res.append((n1, c1, n2, c2))
            idx += 1
        self.up = []
        pad_in = 1 if n_blocks > 0 else 0  # the last res block leaves a reflect border we skip over
        # Binarizer behind the res blocks (networks.py:231-238, bin_before_res=False): 1x1 conv + tanh + sign as ONE
        # implicit-GEMM launch; the +-1 codes are then the input of the first ConvTranspose
        self.binarizer = None
        if binarizer_out_channels is not None:
            if binarizer_out_channels % 64:
                raise JpdseError("generator binarizer needs out_channels %% 64 == 0")
            self.binarizer_name = "model.%d.conv" % idx
            # inference: sign(tanh(conv)) in the epilogue; training plans: raw conv output, then the stochastic sign
            # kernel (ctu/quantizers/binarize.py:13-41) and its straight-through backward
            self.binarizer = ops.Conv(CONV1X1, EPI_RAW if training else EPI_SIGN_NCHW, B, h, w, pad_in, c, c,
                                      binarizer_out_channels, device)
            self.convs[self.binarizer_name] = self.binarizer
            self.codes = torch.empty((B, binarizer_out_channels, h, w), dtype=torch.float32, device=device)
            self.codes_nhwc = ops.alloc_nhwc(B, h, w, binarizer_out_channels, device)
            if training:
                cb = binarizer_out_channels
                # data gradient of the 1x1 conv = 1x1 conv with the transposed weight
                self.binarizer_dgrad = ops.Conv(CONV1X1, EPI_RAW, B, h, w, 0, cb, cb, c, device)
                self.bin_pre = torch.empty((B, h, w, cb), dtype=torch.bfloat16, device=device)
                self.bin_tanh = torch.empty((B, cb, h, w), dtype=torch.float32, device=device)
                self.bin_noise = torch.empty((B, cb, h, w), dtype=torch.float32, device=device)
                self.bin_dpre = torch.empty((B, h, w, cb), dtype=torch.bfloat16, device=device)
                self.bin_gx = torch.empty((B, h, w, c), dtype=torch.bfloat16, device=device)
                self.bin_in = None
                self.stochastic = True  # nn.Module.training of the generator (DifferentiableSign, binarize.py:37-41)
            idx += 1
            pad_in = 0
        c_up = c if self.binarizer is None else binarizer_out_channels
        for i in range(n_downsampling):
            name = "model.%d" % idx
            cv = ops.Conv(CONVT3X3_S2, EPI_RAW_STATS, B, h, w, pad_in if i == 0 else 0, c_up if i == 0 else c,
                          c_up if i == 0 else c, c // 2, device)
            self.convs[name] = cv
            if training:  # dgrad of a ConvTranspose == stride-2 conv on the same weight memory
                self.dgrads[name] = ops.Conv(CONV3X3_S2, EPI_RAW, B, 2 * h, 2 * w, 0, c // 2, c // 2,
                                             c_up if i == 0 else c, device)
            self.up.append((name, cv))
            idx += 3
            c, h, w = c // 2, 2 * h, 2 * w
        self.head_name = "model.%d" % (idx + 1)
        self.head = ops.Conv(CONV7X7_PAD3, EPI_BIAS_TANH_NCHW, B, H, W, 3, ngf, ngf, output_nc, device)
        self.convs[self.head_name] = self.head
        if training:
            self.dgrads[self.head_name] = ops.Conv(CONV7X7_FULL, EPI_RAW, B, H, W, 6, 8, output_nc, ngf, device)
        self.flops = sum(cv.flops for cv in self.convs.values())

        # ---- buffers
        self.x0 = ops.alloc_nhwc(B, H + 6, W + 6, self.c_in_pad, device)
        n_norm = 1 + n_downsampling + 2 * n_blocks + n_downsampling
        cmax = max(ngf << n_downsampling, ngf)
        self.stats = torch.zeros((n_norm, B, cmax, 2), dtype=torch.float64, device=device)
        self.out = out if out is not None else torch.empty((B, output_nc, H, W), dtype=torch.float32, device=device)
        raw_elems, act_elems = B * H * W * ngf, B * (H + 6) * (W + 6) * ngf
        if not training:
            # largest raw conv output / largest (padded) activation, in bf16 elements
            self.raw = torch.empty(raw_elems, dtype=torch.bfloat16, device=device)
            self.act = [torch.zeros(act_elems + 2048, dtype=torch.bfloat16, device=device) for _ in range(3)]
        else:
            self._saved = {}  # (kind, layer index) -> per-layer tensors kept for backward
            self.bwd_sums = torch.zeros((n_norm, B, cmax, 2), dtype=torch.float64, device=device)
            self.g_buf = [torch.zeros(act_elems + 2048, dtype=torch.bfloat16, device=device) for _ in range(2)]
            self.dy_buf = [torch.zeros(raw_elems, dtype=torch.bfloat16, device=device) for _ in range(3)]
            dx_elems = max(raw_elems, B * (self.hb + 4) * (self.wb + 4) * self.cb)
            # two dx buffers: the weight gradient of layer k (own stream, below) still reads dx_k while the main
            # stream writes dx_{k-1}
            self.dx_buf = [torch.zeros(dx_elems + 2048, dtype=torch.bfloat16, device=device) for _ in range(2)]
            self.overlap_wgrad = os.environ.get("JPDSE_WGRAD_STREAM", "1") != "0"
            self.fused_norm_backward = os.environ.get("JPDSE_FUSED_NORM_BACKWARD", "1") != "0"
            self._wgrad_stream = torch.cuda.Stream(device=device)
            self._wgrad_ws = None  # float32 scratch of the weight-gradient kernels on that stream
            self.d_pre = torch.zeros(B * (H + 12) * (W + 12) * 8 + 2048, dtype=torch.bfloat16, device=device)
        self.layers = []
        self.generation = 0
        # tests: a dict here makes backward() keep copies of every layer's incoming / outgoing gradient (the rotating
        # gradient buffers are overwritten as the walk proceeds), keyed by layer name -- used by the teacher-forced
        # whole-network gradient check
        self.capture = None
        self._codes_only = False  # mode='get_binary_code': stop behind the Binarizer
        # CUDA-graph replay of the inference forward (63 launches are host-bound below batch ~4: 2.5 -> ~1.2 ms at
        # batch 1); JPDSE_NO_GRAPH=1 or plan.use_graph = False runs every launch eagerly
        self.use_graph = not training and os.environ.get("JPDSE_NO_GRAPH", "0") != "1"
        self._graphs = {}

    # ---- weights
    def load_weights(self, state_dict):
        """Pack float32 reference-layout weights (keys as in net_G.pth) into the kernels' layout."""
        self._graphs = {}  # captured graphs hold the old bias pointer
        for prefix, cv in self.convs.items():
            w = state_dict[prefix + ".weight"]
            b = state_dict.get(prefix + ".bias") if cv is self.head else None
            if w.device != self.device:
                w = w.to(self.device)
            if b is not None and b.device != self.device:
                b = b.to(self.device)
            w = w.detach().contiguous().float()
            cv.pack(w, None if b is None else b.detach().contiguous().float())
            dg = self.dgrads.get(prefix)
            if dg is not None:
                dg.pack(w)
            if cv is self.binarizer and self.training:
                cout, cin = w.shape[0], w.shape[1]
                self.binarizer_dgrad.pack(w.view(cout, cin).t().contiguous().view(cin, cout, 1, 1))

    def _view(self, buf, B, H, W, C):
        return buf[: B * H * W * C].view(B, H, W, C)

    def _stats(self, i, C, which=None):
        # contiguous (B, C, 2) slice of layer i's statistics
        src = self.stats if which is None else which
        return src[i].view(-1)[: self.B * C * 2].view(self.B, C, 2)

    def _raw_buf(self, si, B, h, w, c):
        if not self.training:
            return self._view(self.raw, B, h, w, c)
        key = ("raw", si)
        if key not in self._saved:
            self._saved[key] = torch.empty((B, h, w, c), dtype=torch.bfloat16, device=self.device)
        return self._saved[key]

    def _act_buf(self, si, slot, B, h, w, c):
        """Activation (B,h,w,c) incl. border: rotating buffers for inference, one per layer for training."""
        if not self.training:
            return self._view(self.act[slot], B, h, w, c)
        key = ("act", si)
        if key not in self._saved:
            self._saved[key] = ops.alloc_nhwc(B, h, w, c, self.device)
        return self._saved[key]

    def _norm(self, si, name, conv, x_in, h, w, c, pad, relu, slot, residual=None, consumer_reads_border=True):
        """conv -> raw (+stats) -> InstanceNorm apply (+ReLU)(+residual)(+reflect pad). Returns the padded output."""
        B = self.B
        raw = self._raw_buf(si, B, h, w, c)
        st = self._stats(si, c)
        conv.forward(x_in, raw, st)
        out = self._act_buf(si, slot, B, h + 2 * pad, w + 2 * pad, c)
        ops.instnorm_apply(raw, st, out, B, h, w, c, pad, relu, residual=residual)
        if self.training:
            L = _Layer()
            L.name, L.conv, L.dgrad, L.x_in, L.raw, L.stats, L.out = name, conv, self.dgrads.get(name), x_in, raw, st, out
            L.out_pad, L.h, L.w, L.c, L.relu, L.residual = pad, h, w, c, relu, residual is not None
            L.consumer_reads_border = consumer_reads_border
            self.layers.append(L)
        return out

    # ---- forward
    def forward_from_x0(self):
        """Runs the generator on self.x0 (bf16 NHWC, reflect-padded by 3); returns fp32 NCHW (B,out,H,W)."""
        with ops.stream_cached():
            return self._forward_from_x0()

    def _forward_from_x0(self):
        self.stats.zero_()
        ops._count()
        self.layers = []
        self.generation += 1
        si = 0
        c, h, w = self.ngf, self.H, self.W
        cur = 0
        x = self._norm(si, "model.1", self.stem, self.x0, h, w, c, 0, True, cur)
        si += 1
        for i, (name, cv) in enumerate(self.down):
            c, h, w = 2 * c, h // 2, w // 2
            last = i == self.n_down - 1
            pad = (1 if self.n_blocks > 0 else 0) if last else 0
            cur ^= 1
            x = self._norm(si, name, cv, x, h, w, c, pad, True, cur)
            si += 1
        # residual blocks: x (padded by 1) lives in slot cur; t and the new x use the other two
        for k, (n1, c1, n2, c2) in enumerate(self.res):
            t = self._norm(si, n1, c1, x, h, w, c, 1, True, (cur + 1) % 3)
            si += 1
            n_idx = (cur + 2) % 3
            # the last block's border is skipped over by the first ConvTranspose
            x = self._norm(si, n2, c2, t, h, w, c, 1, False, n_idx, residual=x,
                           consumer_reads_border=k != self.n_blocks - 1)
            si += 1
            cur = n_idx
        if self.binarizer is not None and self.training:
            # Binarizer in train() mode (binarize.py:44-65): conv1x1 -> tanh -> stochastic sign, the uniform noise drawn
            # from torch's generator exactly like the reference's `input.new(input.size()).uniform_()`
            if not self.stochastic:
                raise NotImplementedError("jpdse_b200: gradients through a binarizing generator need train() mode (the "
                                          "stochastic sign); eval() + autograd is outside the accelerated path")
            self.bin_in = x
            self.binarizer.forward(x, self.bin_pre)
            self.bin_noise.uniform_()
            ops.binarizer_train_forward(self.bin_pre, self.bin_noise, self.codes, self.bin_tanh)
            x = ops.nchw_to_nhwc_bf16(self.codes, out=self.codes_nhwc)
        elif self.binarizer is not None:
            self.binarizer.forward(x, self.codes)  # sign(tanh(conv1x1(x))), float32 NCHW like the reference's codes
            if self._codes_only:
                return self.codes
            x = ops.nchw_to_nhwc_bf16(self.codes, out=self.codes_nhwc)
        for i, (name, cv) in enumerate(self.up):
            c, h, w = c // 2, 2 * h, 2 * w
            last = i == self.n_down - 1
            cur = (cur + 1) % 3
            x = self._norm(si, name, cv, x, h, w, c, 3 if last else 0, True, cur)
            si += 1
        self.head_in = x
        self.head.forward(x, self.out)
        return self.out

    def _graphed(self, key, inputs, eager):
        """Run `eager(*static_inputs)` through a captured CUDA graph (captured on first use per input signature)."""
        entry = self._graphs.get(key)
        if entry is None:
            static = [torch.empty_like(t) for t in inputs]
            for s_, t in zip(static, inputs):
                s_.copy_(t)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):  # warm-up outside capture (one-time cudaFuncSetAttribute calls etc.)
                eager(*static)
            torch.cuda.current_stream(self.device).wait_stream(side)
            before = ops.launch_count
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                eager(*static)
            entry = (graph, static, ops.launch_count - before)
            self._graphs[key] = entry
        graph, static, launches = entry
        for s_, t in zip(static, inputs):
            s_.copy_(t)
        graph.replay()
        ops._count(launches)
        self.generation += 1
        return self.out

    def _eager_nchw(self, inp):
        ops.nchw_to_nhwc_bf16(inp, pad_reflect=3, c_pad=self.c_in_pad, out=self.x0)
        return self.forward_from_x0()

    def binary_code_nchw(self, inp):
        """GlobalGenerator.forward(mode='get_binary_code') (networks.py:252-261): the prefix up to the Binarizer."""
        if self.binarizer is None:
            raise JpdseError("this plan has no binarizer")
        self._codes_only = True
        try:
            return self._eager_nchw(inp)
        finally:
            self._codes_only = False

    def forward_nchw(self, inp):
        """inp: float32 (B, input_nc, H, W) -- the tensor the reference feeds netG (pix2pixHD_model.py:609)."""
        if tuple(inp.shape) != (self.B, self.input_nc, self.H, self.W):
            raise JpdseError("plan built for %s, got %s" % ((self.B, self.input_nc, self.H, self.W), tuple(inp.shape)))
        if self.use_graph:
            return self._graphed(("nchw",), [inp], self._eager_nchw)
        return self._eager_nchw(inp)

    def forward_from_maps(self, label, instance, image, num_labels, mean=(0.5, 0.5, 0.5), std=(1.0, 1.0, 1.0),
                          bad_count=None):
        """Fused preprocessing path: label ids + instance ids + image -> generator output. `image` is float32 (already
        normalised) or uint8 (raw decoder output: (x/255 - mean)/std is applied inside the input-build kernel).
        `bad_count`: optional persistent int32 device counter bumped for every class id outside [0, num_labels) (the
        reference's scatter_ raises on those, pix2pixHD_model.py:381-382); the caller checks it lazily."""
        if num_labels + 4 != self.input_nc:
            raise JpdseError("num_labels + 4 must equal input_nc")

        def eager(lab, ins, img):
            ops.build_input(lab, ins, img, num_labels, pad=3, c_pad=self.c_in_pad, out_nhwc=self.x0, mean=mean, std=std,
                            bad_count=bad_count)
            return self.forward_from_x0()
        if self.use_graph:
            key = ("maps", label.dtype, instance.dtype, image.dtype, num_labels, tuple(mean), tuple(std),
                   None if bad_count is None else bad_count.data_ptr())
            return self._graphed(key, [label, instance, image], eager)
        return eager(label, instance, image)

    # ---- backward
    def backward(self, grad_out, weight_shapes, on_grad=None, alloc=None):
        """grad_out: float32 (B,output_nc,H,W) = dL/d(generator output) of the LAST forward of this plan.

        Returns {state-dict key: float32 gradient} for every conv weight and the head bias (the biases in front of an
        InstanceNorm have an exactly-zero gradient and are not returned). `weight_shapes` maps prefix -> torch weight
        shape. `on_grad(key, tensor)` is called as soon as a gradient's kernels are enqueued, in reverse layer order
        (the hook the data-parallel all-reduce overlaps on). `alloc(key, shape)` may supply the gradient tensors (the
        reducer hands out slices of one flat buffer so buckets are contiguous).
        """
        if not self.training:
            raise JpdseError("this GeneratorPlan was built for inference (training=False)")
        if not self.layers:
            raise JpdseError("backward() called before forward()")
        with ops.stream_cached():
            return self._backward(grad_out, weight_shapes, on_grad, alloc)

    def _backward(self, grad_out, weight_shapes, on_grad, alloc):
        B, H, W = self.B, self.H, self.W
        dev = self.device
        grads = {}

        def new(key, shape):
            if alloc is not None:
                return alloc(key, tuple(shape))
            return torch.empty(shape, dtype=torch.float32, device=dev)

        def emit(key, t):
            grads[key] = t
            if on_grad is not None:
                on_grad(key, t)

        # Weight gradients run on their own stream: they depend only on dx_k and the saved forward activation, nothing
        # downstream depends on them, and the main chain (data gradient -> InstanceNorm backward -> ...) leaves SMs idle
        # at every wave tail (152 data-gradient tiles on 148 SMs at batch 2). The main stream only waits for them
        # before it overwrites a dx buffer they read, and at the end.
        main = torch.cuda.current_stream(dev)
        side = self._wgrad_stream if self.overlap_wgrad else None
        if self._wgrad_ws is None:
            need = max([self.head.wgrad_workspace_bytes(6)] +
                       [L.conv.wgrad_workspace_bytes(2 if L.conv.kind == CONV3X3_PAD1 else 0) for L in self.layers])
            self._wgrad_ws = torch.empty((need + 3) // 4, dtype=torch.float32, device=dev)
        ws = self._wgrad_ws

        def mark():
            """Event on the main stream: 'everything issued so far' (None when not overlapping)."""
            if side is None:
                return None
            ev = torch.cuda.Event()
            ev.record(main)
            return ev

        def weight_grad(fn, ready=None):
            """Enqueue fn (wgrad + emit) behind `ready` (default: everything the main stream has issued so far); returns
            the event the main stream must wait on before it overwrites fn's inputs (None when not overlapping)."""
            if side is None:
                fn()
                return None
            ev = ready if ready is not None else mark()
            side.wait_event(ev)
            with torch.cuda.stream(side), ops.stream_cached():
                fn()
                done = torch.cuda.Event()
                done.record(side)
            return done

        self.bwd_sums.zero_()
        ops._count()
        # ---- head: tanh', bias, wgrad, dgrad
        d_pre = self._view(self.d_pre, B, H + 12, W + 12, 8)
        dbias = new(self.head_name + ".bias", (self.output_nc,)).zero_()
        ops.tanh_backward_nchw(grad_out, self.out, d_pre, dbias)
        dw_head = new(self.head_name + ".weight", weight_shapes[self.head_name])

        def head_grads():
            self.head.wgrad(self.head_in, d_pre, 6, dw_head, workspace=ws)
            emit(self.head_name + ".bias", dbias)
            emit(self.head_name + ".weight", dw_head)

        weight_grad(head_grads)
        gi = 0
        g = self._view(self.g_buf[gi], B, H + 6, W + 6, self.ngf)
        self.dgrads[self.head_name].forward(d_pre, g)
        if self.capture is not None:
            self.capture[self.head_name] = {"g_in": g.clone()}
        g_pad = 3
        skip, skip_idx = None, None  # second gradient of the current layer's output (ResnetBlock skip connection)
        pending_idx = None           # dy buffer holding dL/dx_{k+1} until the walk reaches the producer of x_k
        dx_free = [None, None]       # per dx buffer: event of the weight gradient that last read it
        for step, si in enumerate(range(len(self.layers) - 1, -1, -1)):
            L = self.layers[si]
            h, w, c = L.h, L.w, L.c
            if not L.consumer_reads_border:
                g_pad = 0  # the consumer skipped the border: its dgrad produced an unpadded gradient
            dy_idx = [i for i in range(3) if i != skip_idx and i != pending_idx][0]
            dy = self._view(self.dy_buf[dy_idx], B, h, w, c)
            sums = self._stats(si, c, self.bwd_sums)
            if self.capture is not None:
                self.capture[L.name] = {"g": g.clone(), "g_pad": g_pad, "skip": None if skip is None else skip.clone()}
            z = 2 if L.conv.kind == CONV3X3_PAD1 else 0
            slot = step & 1
            if dx_free[slot] is not None:
                main.wait_event(dx_free[slot])
            if z and self.shared_border:
                # B*(h+2)*(w+2) + 2*(w+2) + 2 positions of c channels; the kernels only see the pointer
                dx = self.dx_buf[slot][:(B * (h + z) * (w + z) + z * (w + z) + z) * c].view(-1, c)
                z |= PAD_SHARED
            else:
                dx = self._view(self.dx_buf[slot], B, h + 2 * z, w + 2 * z, c)
            if h * w <= ops.FUSED_NORM_BACKWARD_MAX_PIXELS and self.fused_norm_backward:
                # small maps (the 1024-channel bottleneck): reduce + apply in ONE launch, dy stays in registers and is only
                # written where the ResnetBlock skip connection needs it
                ops.instnorm_backward_fused(g, g_pad, skip, L.raw, L.stats, dy if L.residual else None, dx, z, B, h, w, c, L.relu)
            else:
                ops.instnorm_backward_reduce(g, g_pad, skip, L.raw, L.stats, dy, sums, B, h, w, c, L.relu)
                ops.instnorm_backward_apply(dy, L.raw, L.stats, sums, dx, z, B, h, w, c)
            dw = new(L.name + ".weight", weight_shapes[L.name])

            def layer_grad(L=L, dx=dx, z=z, dw=dw):
                L.conv.wgrad(L.x_in, dx, z, dw, workspace=ws)
                emit(L.name + ".weight", dw)

            if si == 0:
                dx_free[slot] = weight_grad(layer_grad)
                break  # the stem's input (labels + decoded image) needs no gradient
            # Both gradients of this conv only wait for dx. The data gradient is LAUNCHED first: it is the one the chain
            # waits for, and the weight gradient (75 registers, one CTA per SM) then shares the SMs with the next layer's
            # InstanceNorm backward, which the data gradient (136 registers) cannot. Launched the other way round the
            # weight gradient took the SMs first and sat on the critical path (profiles/r2_timeline_g_b2.txt).
            ready = mark() if self.dgrad_first else None
            if not self.dgrad_first:
                dx_free[slot] = weight_grad(layer_grad)
            gi ^= 1
            oh, ow = L.dgrad.out_hw
            g = self._view(self.g_buf[gi], B, oh, ow, L.dgrad.cout)
            L.dgrad.forward(dx, g)
            if self.dgrad_first:
                dx_free[slot] = weight_grad(layer_grad, ready)
            if self.capture is not None:
                self.capture[L.name]["g_in"] = g.clone()  # gradient w.r.t. this layer's input as the conv saw it
            if self.binarizer is not None and L.x_in is self.codes_nhwc:
                # Binarizer backward (binarize.py:26-28): identity through the stochastic sign, tanh', then the 1x1
                # conv's weight gradient and its data gradient (1x1 conv with the transposed weight)
                ops.binarizer_train_backward(ops.nhwc_bf16_to_nchw(g), self.bin_tanh, self.bin_dpre)
                dwb = new(self.binarizer_name + ".weight", weight_shapes[self.binarizer_name])

                def bin_grad(dwb=dwb):
                    self.binarizer.wgrad(self.bin_in, self.bin_dpre, 0, dwb, workspace=ws)
                    emit(self.binarizer_name + ".weight", dwb)

                bin_ev = weight_grad(bin_grad)
                if bin_ev is not None:
                    main.wait_event(bin_ev)  # bin_dpre / bin_in are single buffers: keep it simple, wait
                g = self.bin_gx
                self.binarizer_dgrad.forward(self.bin_dpre, g)
            g_pad = self.layers[si - 1].out_pad
            skip, skip_idx = None, None
            if L.residual:
                # x_{k+1} = x_k + IN(...): this dy IS dL/dx_{k+1}; it reaches x_k through the skip connection,
                # two layers further down (below conv_block.1)
                pending_idx = dy_idx
            elif L.conv.kind == CONV3X3_PAD1 and pending_idx is not None:
                # conv_block.1: the layer below produced the block input x_k, gradient = fold(g) + dL/dx_{k+1}
                skip_idx, pending_idx = pending_idx, None
                skip = self._view(self.dy_buf[skip_idx], B, h, w, c)
        if side is not None:
            main.wait_stream(side)
        return grads


class SplitGeneratorPlan:
    """Inference plan that runs the batch as `parts` half-batches on separate CUDA streams.

    The forward is a strict chain of full-GPU kernels, but they bind on different resources: the convs on the tensor
    pipe / the L2 fabric with one 200 KB-shared-memory CTA per SM, the InstanceNorm passes on HBM with no shared memory
    at all. Two independent half-batches let a norm kernel of one half co-reside with a conv of the other, and let the
    next conv's CTAs start while the previous one's tail drains. Measured on B200 at batch 16: 17.3 -> 16.5 ms with two
    streams (four streams: 16.7 ms); outputs are bit-identical (InstanceNorm is per-sample). The fork / join is captured
    into ONE CUDA graph like the single-stream plan's launches.
    """

    def __init__(self, input_nc, output_nc, ngf, n_downsampling, n_blocks, batch, height, width, device, parts=2,
                 binarizer_out_channels=None):
        if batch % parts:
            raise JpdseError("SplitGeneratorPlan: batch %d is not a multiple of %d" % (batch, parts))
        self.B, self.H, self.W, self.device = batch, height, width, device
        self.input_nc = input_nc
        self.training = False
        self.out = torch.empty((batch, output_nc, height, width), dtype=torch.float32, device=device)
        pb = batch // parts
        self.parts = [GeneratorPlan(input_nc, output_nc, ngf, n_downsampling, n_blocks, pb, height, width, device,
                                    binarizer_out_channels=binarizer_out_channels, out=self.out[i * pb:(i + 1) * pb])
                      for i in range(parts)]
        for p in self.parts:
            p.use_graph = False  # the parent captures the whole fork / join
        self.streams = [torch.cuda.Stream(device=device) for _ in range(parts)]
        self.parallel = True   # False: run the parts one after the other on the current stream (instrumented runs)
        self.use_graph = os.environ.get("JPDSE_NO_GRAPH", "0") != "1"
        self._graphs = {}
        self.generation = 0
        self.flops = sum(p.flops for p in self.parts)
        self.binarizer = self.parts[0].binarizer

    def load_weights(self, state_dict):
        self._graphs = {}
        for p in self.parts:
            p.load_weights(state_dict)

    def _fan_out(self, fn, inputs):
        pb = self.B // len(self.parts)
        if not self.parallel:
            for i, p in enumerate(self.parts):
                fn(p, *[t[i * pb:(i + 1) * pb] for t in inputs])
            return self.out
        main = torch.cuda.current_stream(self.device)
        fork = torch.cuda.Event()
        fork.record(main)
        for i, (p, s) in enumerate(zip(self.parts, self.streams)):
            s.wait_event(fork)
            with torch.cuda.stream(s):
                fn(p, *[t[i * pb:(i + 1) * pb] for t in inputs])
        for s in self.streams:
            main.wait_stream(s)
        return self.out

    def _run(self, key, inputs, fn):
        self.generation += 1
        if not self.use_graph:
            return self._fan_out(fn, inputs)
        entry = self._graphs.get((key, self.parallel))
        if entry is None:
            static = [torch.empty_like(t) for t in inputs]
            for s_, t in zip(static, inputs):
                s_.copy_(t)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                self._fan_out(fn, static)  # warm-up outside capture
            torch.cuda.current_stream(self.device).wait_stream(side)
            before = ops.launch_count
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._fan_out(fn, static)
            entry = (graph, static, ops.launch_count - before)
            self._graphs[(key, self.parallel)] = entry
        graph, static, launches = entry
        for s_, t in zip(static, inputs):
            s_.copy_(t)
        graph.replay()
        ops._count(launches)
        return self.out

    def forward_nchw(self, inp):
        if tuple(inp.shape) != (self.B, self.input_nc, self.H, self.W):
            raise JpdseError("plan built for %s, got %s" % ((self.B, self.input_nc, self.H, self.W), tuple(inp.shape)))
        return self._run(("nchw",), [inp], lambda p, x: p._eager_nchw(x))

    def forward_from_maps(self, label, instance, image, num_labels, mean=(0.5, 0.5, 0.5), std=(1.0, 1.0, 1.0),
                          bad_count=None):
        if num_labels + 4 != self.input_nc:
            raise JpdseError("num_labels + 4 must equal input_nc")

        def fn(p, lab, ins, img):
            ops.build_input(lab, ins, img, num_labels, pad=3, c_pad=p.c_in_pad, out_nhwc=p.x0, mean=mean, std=std,
                            bad_count=bad_count)
            return p.forward_from_x0()
        key = ("maps", label.dtype, instance.dtype, image.dtype, num_labels, tuple(mean), tuple(std),
               None if bad_count is None else bad_count.data_ptr())
        return self._run(key, [label, instance, image], fn)

    def binary_code_nchw(self, inp):
        pb = self.B // len(self.parts)
        return torch.cat([p.binary_code_nchw(inp[i * pb:(i + 1) * pb]).clone() for i, p in enumerate(self.parts)], 0)


def make_inference_plan(input_nc, output_nc, ngf, n_downsampling, n_blocks, batch, height, width, device,
                        binarizer_out_channels=None):
    """Two half-batch plans on two streams when the batch is large enough for the halves to still fill the GPU."""
    parts = int(os.environ.get("JPDSE_SPLIT_STREAMS", "2"))
    if parts > 1 and batch % parts == 0 and batch // parts >= 4:
        return SplitGeneratorPlan(input_nc, output_nc, ngf, n_downsampling, n_blocks, batch, height, width, device, parts,
                                  binarizer_out_channels)
    return GeneratorPlan(input_nc, output_nc, ngf, n_downsampling, n_blocks, batch, height, width, device,
                         binarizer_out_channels=binarizer_out_channels)
