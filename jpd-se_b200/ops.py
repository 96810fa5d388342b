"""Torch-tensor front end of the C ABI (include/jpdse_b200.h).

PyTorch only supplies device memory and the current stream; every computation is a kernel of
libjpdse_b200.so. Each wrapper validates device / dtype / contiguity and raises on any error --
there is no fallback path.
"""
import ctypes

import torch

from . import _lib
from ._lib import (CONV1X1, CONV3X3_FULL, CONV3X3_FULL_SHARED, CONV3X3_PAD1, CONV3X3_S2, CONV4X4_S1, CONV4X4_S1_FULL, CONV4X4_S2,
                   CONV4X4_S2_DGRAD, CONV7X7_FULL, CONV7X7_PAD3, CONVT3X3_S2, EPI_BIAS_ACT, EPI_BIAS_NCHW,
                   EPI_BIAS_TANH_NCHW, EPI_RAW, EPI_RAW_STATS, EPI_SIGN_NCHW, ConvDesc, JpdseError, check)

_LABEL_DTYPES = {torch.float32: 0, torch.uint8: 1, torch.int64: 2}
_INST_DTYPES = {torch.int32: 0, torch.int16: 1, torch.int64: 2, torch.float32: 3}

# counts kernels launched through this module (bench.py reports it as gpu_launches)
launch_count = 0


_cached_stream = None
_cached_dev = None


def _stream():
    if _cached_stream is not None:
        return _cached_stream
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class stream_cached:
    """Query torch's current stream once for a whole kernel sequence (a plan forward / backward issues 60-250 C-ABI
    calls; the per-call lookup is a measurable part of the host time at small batch). Not re-entrant across streams:
    the sequence inside the block must stay on the stream that was current on entry."""

    def __enter__(self):
        global _cached_stream, _cached_dev
        self.prev = (_cached_stream, _cached_dev)
        _cached_dev = torch.cuda.current_device()
        _cached_stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        return self

    def __exit__(self, *exc):
        global _cached_stream, _cached_dev
        _cached_stream, _cached_dev = self.prev
        return False


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _need(t, name, dtype=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise JpdseError("%s must be a CUDA tensor (jpdse_b200 has no CPU path)" % name)
    # the launch goes to the CURRENT device's current stream: a tensor of another GPU would be a foreign pointer there
    cur = _cached_dev if _cached_dev is not None else torch.cuda.current_device()
    if t.device.index != cur:
        raise JpdseError("%s lives on cuda:%d but the current device is cuda:%d; run the call under "
                         "torch.cuda.device(%d)" % (name, t.device.index, cur, t.device.index))
    if not t.is_contiguous():
        raise JpdseError("%s must be contiguous" % name)
    if dtype is not None and t.dtype != dtype:
        raise JpdseError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t


def _count(n=1):
    global launch_count
    launch_count += n


# ------------------------------------------------------------------------------------------------ input build
def build_input(label, instance, image, num_labels, pad=3, c_pad=None, nhwc=True, nchw=False, out_nhwc=None,
                bad_count=None, mean=(0.5, 0.5, 0.5), std=(1.0, 1.0, 1.0)):
    """One-hot + edge + concat (+reflect pad). See jpdse_build_input.

    Returns (nhwc_or_None, nchw_or_None). `nhwc` is bf16 (B, H+2pad, W+2pad, c_pad) carved from a flat
    buffer with 4 KiB of zeroed slack behind it (the 7x7 stem's window view over-reads by < 128 B).
    """
    lib = _lib.load()
    _need(label, "label")
    _need(instance, "instance")
    _need(image, "image")
    if image.dtype not in (torch.float32, torch.uint8):
        raise JpdseError("image must be float32 (normalised) or uint8 (raw; normalised on the device), got %s" % image.dtype)
    if label.dtype not in _LABEL_DTYPES:
        raise JpdseError("label dtype %s not supported (float32, uint8, int64)" % label.dtype)
    if instance.dtype not in _INST_DTYPES:
        raise JpdseError("instance dtype %s not supported (int32, int16, int64, float32)" % instance.dtype)
    B, _, H, W = image.shape
    if tuple(label.shape) != (B, 1, H, W) or tuple(instance.shape) != (B, 1, H, W) or image.shape[1] != 3:
        raise JpdseError("build_input: label/instance must be (B,1,H,W) and image (B,3,H,W)")
    if c_pad is None:
        c_pad = (num_labels + 4 + 7) // 8 * 8
    o_nhwc = None
    if nhwc:
        o_nhwc = out_nhwc if out_nhwc is not None else alloc_nhwc(B, H + 2 * pad, W + 2 * pad, c_pad, image.device)
    o_nchw = torch.empty((B, num_labels + 4, H, W), dtype=torch.float32, device=image.device) if nchw else None
    if image.dtype == torch.uint8:
        m3, s3 = (ctypes.c_float * 3)(*[float(v) for v in mean]), (ctypes.c_float * 3)(*[float(v) for v in std])
        check(lib.jpdse_build_input_u8(_ptr(label), _LABEL_DTYPES[label.dtype], _ptr(instance), _INST_DTYPES[instance.dtype],
                                       _ptr(image), m3, s3, B, H, W, num_labels, _ptr(o_nhwc), pad, c_pad, _ptr(o_nchw),
                                       _ptr(bad_count), _stream()))
    else:
        check(lib.jpdse_build_input(_ptr(label), _LABEL_DTYPES[label.dtype], _ptr(instance), _INST_DTYPES[instance.dtype],
                                    _ptr(image), B, H, W, num_labels, _ptr(o_nhwc), pad, c_pad, _ptr(o_nchw),
                                    _ptr(bad_count), _stream()))
    _count(int(nhwc) + int(nchw))
    return o_nhwc, o_nchw


def alloc_nhwc(B, H, W, C, device, slack_bytes=4096):
    """bf16 (B,H,W,C) view over a zero-initialised flat buffer with trailing slack."""
    n = B * H * W * C
    flat = torch.zeros(n + slack_bytes // 2, dtype=torch.bfloat16, device=device)
    return flat[:n].view(B, H, W, C)


# ------------------------------------------------------------------------------------------------ convolutions
class Conv:
    """One convolution of the generator: descriptor + packed bf16 weights (+ bias for the head)."""

    def __init__(self, kind, epilogue, batch, in_h, in_w, in_pad, cin, cin_real, cout, device, out_pad=0, out_hw=None,
                 slope=0.0, cout_real=0):
        self.lib = _lib.load()
        self.desc = ConvDesc(kind, epilogue, batch, in_h, in_w, in_pad, cin, cin_real, cout, out_pad,
                             0 if out_hw is None else out_hw[0], 0 if out_hw is None else out_hw[1], float(slope), cout_real)
        nbytes = self.lib.jpdse_conv_packed_weight_bytes(ctypes.byref(self.desc))
        if nbytes == 0:
            raise JpdseError("conv descriptor rejected: %s" % self.lib.jpdse_last_error().decode())
        self.w_packed = torch.empty(nbytes // 2, dtype=torch.bfloat16, device=device)
        self.bias = None
        self.flops = self.lib.jpdse_conv_flops(ctypes.byref(self.desc))
        self.launches = self.lib.jpdse_conv_launch_count(ctypes.byref(self.desc))
        self.kind, self.epilogue = kind, epilogue
        self.out_hw = {CONV3X3_S2: (in_h // 2, in_w // 2), CONVT3X3_S2: (in_h * 2, in_w * 2),
                       CONV3X3_FULL: (in_h + 2, in_w + 2), CONV3X3_FULL_SHARED: (in_h + 2, in_w + 2),
                       CONV7X7_FULL: (in_h + 6, in_w + 6),
                       CONV4X4_S2: (in_h // 2 + 1, in_w // 2 + 1), CONV4X4_S1: (in_h + 1, in_w + 1),
                       CONV4X4_S2_DGRAD: out_hw, CONV4X4_S1_FULL: (in_h - 1, in_w - 1)}.get(kind, (in_h, in_w))
        self.out_pad = out_pad
        self.cout = cout
        self.batch = batch

    def pack(self, weight, bias=None):
        """weight: float32 torch layout (Conv2d (Cout,Cin,k,k) / ConvTranspose2d (Cin,Cout,k,k))."""
        w = _need(weight.detach(), "weight", torch.float32)
        check(self.lib.jpdse_conv_pack_weights(ctypes.byref(self.desc), _ptr(w), _ptr(self.w_packed), _stream()))
        _count()
        if bias is not None:
            self.bias = _need(bias.detach(), "bias", torch.float32)

    def forward(self, x, y, stats=None):
        _need(x, "x", torch.bfloat16)
        _need(y, "y")
        check(self.lib.jpdse_conv_forward(ctypes.byref(self.desc), _ptr(x), _ptr(self.w_packed), _ptr(self.bias), _ptr(y),
                                          _ptr(stats), _stream()))
        _count(self.launches)
        return y

    def wgrad_workspace_bytes(self, dy_pad):
        nbytes = self.lib.jpdse_conv_wgrad_workspace_bytes(ctypes.byref(self.desc), dy_pad)
        if nbytes == 0:
            raise JpdseError("conv_wgrad rejected: %s" % self.lib.jpdse_last_error().decode())
        return nbytes

    def wgrad(self, x, dy, dy_pad, dw, accumulate=False, workspace=None):
        """dw (float32, torch weight layout) = / += weight gradient of THIS (forward) conv. See jpdse_conv_wgrad.
        `workspace`: a float32 scratch tensor of at least wgrad_workspace_bytes(dy_pad) owned by the caller (needed when
        weight gradients run on their own stream); default: the per-device scratch buffer shared by one stream."""
        _need(x, "x", torch.bfloat16)
        _need(dy, "dy", torch.bfloat16)
        _need(dw, "dw", torch.float32)
        nbytes = self.lib.jpdse_conv_wgrad_workspace_bytes(ctypes.byref(self.desc), dy_pad)
        if nbytes == 0:
            raise JpdseError("conv_wgrad rejected: %s" % self.lib.jpdse_last_error().decode())
        ws = workspace if workspace is not None else _workspace(nbytes, x.device)
        if ws.numel() * 4 < nbytes:
            raise JpdseError("conv_wgrad: workspace of %d bytes, %d needed" % (ws.numel() * 4, nbytes))
        check(self.lib.jpdse_conv_wgrad(ctypes.byref(self.desc), _ptr(x), _ptr(dy), dy_pad, _ptr(dw), int(bool(accumulate)),
                                        _ptr(ws), ws.numel() * 4, _stream()))
        _count(2)
        return dw


_ws_cache = {}


def _workspace(nbytes, device):
    """One growing float32 scratch buffer per device (kernels on one stream use it one after the other)."""
    key = str(device)
    ws = _ws_cache.get(key)
    if ws is None or ws.numel() * 4 < nbytes:
        ws = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=device)
        _ws_cache[key] = ws
    return ws


def instnorm_apply(raw, stats, out, batch, height, width, channels, pad, relu, residual=None, eps=1e-5):
    lib = _lib.load()
    _need(raw, "raw", torch.bfloat16)
    _need(stats, "stats", torch.float64)
    _need(out, "out", torch.bfloat16)
    if residual is not None:
        _need(residual, "residual", torch.bfloat16)
    check(lib.jpdse_instnorm_apply(_ptr(raw), _ptr(stats), _ptr(residual), _ptr(out), batch, height, width, channels,
                                   pad, int(bool(relu)), eps, _stream()))
    _count()
    return out


def instnorm_backward_reduce(g, g_pad, skip, raw, stats, dy, sums, batch, height, width, channels, relu, eps=1e-5):
    lib = _lib.load()
    _need(g, "g", torch.bfloat16)
    _need(raw, "raw", torch.bfloat16)
    _need(stats, "stats", torch.float64)
    _need(dy, "dy", torch.bfloat16)
    _need(sums, "sums", torch.float64)
    if skip is not None:
        _need(skip, "skip", torch.bfloat16)
    check(lib.jpdse_instnorm_backward_reduce(_ptr(g), g_pad, _ptr(skip), _ptr(raw), _ptr(stats), _ptr(dy), _ptr(sums), batch,
                                             height, width, channels, int(bool(relu)), eps, _stream()))
    _count()
    return dy


def instnorm_backward_apply(dy, raw, stats, sums, dx, dx_pad, batch, height, width, channels, eps=1e-5):
    lib = _lib.load()
    _need(dy, "dy", torch.bfloat16)
    _need(raw, "raw", torch.bfloat16)
    _need(stats, "stats", torch.float64)
    _need(sums, "sums", torch.float64)
    _need(dx, "dx", torch.bfloat16)
    check(lib.jpdse_instnorm_backward_apply(_ptr(dy), _ptr(raw), _ptr(stats), _ptr(sums), _ptr(dx), dx_pad, batch, height,
                                            width, channels, eps, _stream()))
    _count()
    return dx


FUSED_NORM_BACKWARD_MAX_PIXELS = 2048


def instnorm_backward_fused(g, g_pad, skip, raw, stats, dy, dx, dx_pad, batch, height, width, channels, relu, slope=0.0, eps=1e-5):
    """InstanceNorm backward reduce + apply in one launch (height * width <= FUSED_NORM_BACKWARD_MAX_PIXELS); dy optional."""
    lib = _lib.load()
    _need(g, "g", torch.bfloat16)
    _need(raw, "raw", torch.bfloat16)
    _need(stats, "stats", torch.float64)
    _need(dx, "dx", torch.bfloat16)
    if skip is not None:
        _need(skip, "skip", torch.bfloat16)
    if dy is not None:
        _need(dy, "dy", torch.bfloat16)
    check(lib.jpdse_instnorm_backward_fused(_ptr(g), g_pad, _ptr(skip), _ptr(raw), _ptr(stats), _ptr(dy), _ptr(dx), dx_pad, batch,
                                            height, width, channels, int(bool(relu)), float(slope), eps, _stream()))
    _count()
    return dx


def instnorm_backward_reduce_act(g, g_pad, skip, raw, stats, dy, sums, batch, height, width, channels, slope, eps=1e-5):
    """InstanceNorm backward reduce with a LeakyReLU(slope) mask (PatchGAN layers 1-3)."""
    lib = _lib.load()
    _need(g, "g", torch.bfloat16)
    _need(raw, "raw", torch.bfloat16)
    _need(stats, "stats", torch.float64)
    _need(dy, "dy", torch.bfloat16)
    _need(sums, "sums", torch.float64)
    if skip is not None:
        _need(skip, "skip", torch.bfloat16)
    check(lib.jpdse_instnorm_backward_reduce_act(_ptr(g), g_pad, _ptr(skip), _ptr(raw), _ptr(stats), _ptr(dy), _ptr(sums),
                                                 batch, height, width, channels, 1, float(slope), eps, _stream()))
    _count()
    return dy


# ------------------------------------------------------------------------------------------------ discriminator / feature losses
def d_input(a, b, out, c_pad, pool, out_pad=2):
    """cat(a, b) over channels [+ AvgPool2d(3,2,1,count_include_pad=False)] -> interior of the zero-bordered NHWC bf16 `out`."""
    lib = _lib.load()
    _need(a, "a", torch.float32)
    if b is not None:
        _need(b, "b", torch.float32)
    _need(out, "out", torch.bfloat16)
    B, ca, H, W = a.shape
    cb = 0 if b is None else b.shape[1]
    if b is not None and (b.shape[0], b.shape[2], b.shape[3]) != (B, H, W):
        raise JpdseError("d_input: a and b must share batch and spatial size")
    Ho, Wo = ((H - 1) // 2 + 1, (W - 1) // 2 + 1) if pool else (H, W)
    if tuple(out.shape) != (B, Ho + 2 * out_pad, Wo + 2 * out_pad, c_pad):
        raise JpdseError("d_input: out must be %s, got %s" % ((B, Ho + 2 * out_pad, Wo + 2 * out_pad, c_pad), tuple(out.shape)))
    check(lib.jpdse_d_input(_ptr(a), ca, _ptr(b), cb, _ptr(out), B, H, W, c_pad, int(bool(pool)), out_pad, _stream()))
    _count()
    return out


def d_input_ids(label, instance, image_a, out_a, image_b, out_b, num_labels, pool, out_pad=2):
    """Discriminator operand(s) from class ids + instance ids + image(s); see jpdse_d_input_ids."""
    lib = _lib.load()
    _need(label, "label")
    _need(instance, "instance")
    _need(image_a, "image_a", torch.float32)
    _need(out_a, "out_a", torch.bfloat16)
    if (image_b is None) != (out_b is None):
        raise JpdseError("d_input_ids: image_b and out_b go together")
    if image_b is not None:
        _need(image_b, "image_b", torch.float32)
        _need(out_b, "out_b", torch.bfloat16)
    if label.dtype not in _LABEL_DTYPES or instance.dtype not in _INST_DTYPES:
        raise JpdseError("d_input_ids: unsupported label / instance dtype (%s, %s)" % (label.dtype, instance.dtype))
    B, _, H, W = image_a.shape
    if tuple(label.shape) != (B, 1, H, W) or tuple(instance.shape) != (B, 1, H, W) or image_a.shape[1] != 3:
        raise JpdseError("d_input_ids: label/instance must be (B,1,H,W) and the image (B,3,H,W)")
    Ho, Wo = ((H - 1) // 2 + 1, (W - 1) // 2 + 1) if pool else (H, W)
    c_pad = out_a.shape[-1]
    for o in (out_a, out_b):
        if o is not None and tuple(o.shape) != (B, Ho + 2 * out_pad, Wo + 2 * out_pad, c_pad):
            raise JpdseError("d_input_ids: output must be %s, got %s" % ((B, Ho + 2 * out_pad, Wo + 2 * out_pad, c_pad), tuple(o.shape)))
    check(lib.jpdse_d_input_ids(_ptr(label), _LABEL_DTYPES[label.dtype], _ptr(instance), _INST_DTYPES[instance.dtype],
                                _ptr(image_a), _ptr(out_a), _ptr(image_b), _ptr(out_b), B, H, W, num_labels, c_pad,
                                int(bool(pool)), out_pad, _stream()))
    _count()
    return out_a


def d_input_backward(g0, g1, c0, c, out=None):
    """float32 (B,c,H,W) = g0[..., c0:c0+c] + AvgPool backward of g1[..., c0:c0+c] (g1 optional)."""
    lib = _lib.load()
    _need(g0, "g0", torch.bfloat16)
    if g1 is not None:
        _need(g1, "g1", torch.bfloat16)
    B, H, W, cs = g0.shape
    if g1 is not None and tuple(g1.shape) != (B, (H - 1) // 2 + 1, (W - 1) // 2 + 1, cs):
        raise JpdseError("d_input_backward: g1 has the wrong shape")
    if out is None:
        out = torch.empty((B, c, H, W), dtype=torch.float32, device=g0.device)
    check(lib.jpdse_d_input_backward(_ptr(g0), _ptr(g1), _ptr(out), B, H, W, cs, c0, c, _stream()))
    _count()
    return out


def instnorm_apply_act(raw, stats, out, batch, height, width, channels, out_pad, slope, eps=1e-5):
    lib = _lib.load()
    _need(raw, "raw", torch.bfloat16)
    _need(stats, "stats", torch.float64)
    _need(out, "out", torch.bfloat16)
    check(lib.jpdse_instnorm_apply_act(_ptr(raw), _ptr(stats), _ptr(out), batch, height, width, channels, out_pad, float(slope),
                                       eps, _stream()))
    _count()
    return out


def patch_out_gather(z, bias, out, batch, height, width):
    """out (B,1,H+1,W+1) = bias + the 16 tap planes of z (B,16,H+4,W+4) summed at their offsets; see jpdse_patch_out_gather."""
    lib = _lib.load()
    _need(z, "z", torch.float32)
    _need(out, "out", torch.float32)
    if bias is not None:
        _need(bias, "bias", torch.float32)
    if tuple(z.shape) != (batch, 16, height + 4, width + 4) or tuple(out.shape) != (batch, 1, height + 1, width + 1):
        raise JpdseError("patch_out_gather: z %s / out %s do not match (%d, %d, %d)" % (tuple(z.shape), tuple(out.shape), batch, height, width))
    check(lib.jpdse_patch_out_gather(_ptr(z), _ptr(bias), _ptr(out), batch, height, width, _stream()))
    _count()
    return out


def patch_out_scatter(dout, dz, batch, height, width):
    """dz (B,H+4,W+4,c_pad) bf16 = the gradient of the 16 tap planes from dout (B,1,H+1,W+1); see jpdse_patch_out_scatter."""
    lib = _lib.load()
    _need(dout, "dout", torch.float32)
    _need(dz, "dz", torch.bfloat16)
    if tuple(dout.shape) != (batch, 1, height + 1, width + 1) or tuple(dz.shape[:3]) != (batch, height + 4, width + 4):
        raise JpdseError("patch_out_scatter: dout %s / dz %s do not match (%d, %d, %d)" % (tuple(dout.shape), tuple(dz.shape), batch, height, width))
    check(lib.jpdse_patch_out_scatter(_ptr(dout), _ptr(dz), batch, height, width, dz.shape[-1], _stream()))
    _count()
    return dz


def act_backward(g, skip, f, d_pre, dbias, batch, height, width, channels, f_pad, out_pad, slope, g_pad=0):
    lib = _lib.load()
    _need(g, "g", torch.bfloat16)
    _need(f, "f", torch.bfloat16)
    _need(d_pre, "d_pre", torch.bfloat16)
    if skip is not None:
        _need(skip, "skip", torch.bfloat16)
    if dbias is not None:
        _need(dbias, "dbias", torch.float32)
    check(lib.jpdse_act_backward(_ptr(g), g_pad, _ptr(skip), _ptr(f), _ptr(d_pre), _ptr(dbias), batch, height, width, channels,
                                 f_pad, out_pad, float(slope), _stream()))
    _count()
    return d_pre


def l1_pair(a, b, acc):
    """acc (float64, 1 element) += sum |a - b| over two bf16 tensors of one stored shape."""
    lib = _lib.load()
    _need(a, "a", torch.bfloat16)
    _need(b, "b", torch.bfloat16)
    _need(acc, "acc", torch.float64)
    if a.shape != b.shape:
        raise JpdseError("l1_pair: shapes differ")
    check(lib.jpdse_l1_pair(_ptr(a), _ptr(b), a.numel(), _ptr(acc), _stream()))
    _count()
    return acc


def l1_pair_backward(a, b, out, scale_dev, scale_host, batch, height, width, channels, pad):
    lib = _lib.load()
    _need(a, "a", torch.bfloat16)
    _need(b, "b", torch.bfloat16)
    _need(out, "out", torch.bfloat16)
    if scale_dev is not None:
        _need(scale_dev, "scale", torch.float32)
    check(lib.jpdse_l1_pair_backward(_ptr(a), _ptr(b), _ptr(out), _ptr(scale_dev), float(scale_host), batch, height, width,
                                     channels, pad, _stream()))
    _count()
    return out


def maxpool2x2(x, y, batch, height, width, channels, in_pad, out_pad):
    lib = _lib.load()
    _need(x, "x", torch.bfloat16)
    _need(y, "y", torch.bfloat16)
    check(lib.jpdse_maxpool2x2(_ptr(x), _ptr(y), batch, height, width, channels, in_pad, out_pad, _stream()))
    _count()
    return y


def maxpool2x2_backward(x, g, dx, batch, height, width, channels, in_pad, g_pad=0):
    lib = _lib.load()
    _need(x, "x", torch.bfloat16)
    _need(g, "g", torch.bfloat16)
    _need(dx, "dx", torch.bfloat16)
    check(lib.jpdse_maxpool2x2_backward(_ptr(x), _ptr(g), g_pad, _ptr(dx), batch, height, width, channels, in_pad, _stream()))
    _count()
    return dx


def nhwc_pad_to_nchw(x, channels, pad, out=None):
    """stored bf16 (B,H+2pad,W+2pad,Cs) -> float32 (B,channels,H,W)"""
    lib = _lib.load()
    _need(x, "x", torch.bfloat16)
    B, Hs, Ws, cs = x.shape
    H, W = Hs - 2 * pad, Ws - 2 * pad
    if out is None:
        out = torch.empty((B, channels, H, W), dtype=torch.float32, device=x.device)
    check(lib.jpdse_nhwc_pad_to_nchw_f32(_ptr(x), _ptr(out), B, channels, H, W, pad, cs, _stream()))
    _count()
    return out


def tanh_backward_nchw(grad_out, out, d_pre, dbias):
    lib = _lib.load()
    _need(grad_out, "grad_out", torch.float32)
    _need(out, "out", torch.float32)
    _need(d_pre, "d_pre", torch.bfloat16)
    _need(dbias, "dbias", torch.float32)
    B, C, H, W = out.shape
    if tuple(grad_out.shape) != (B, C, H, W) or tuple(d_pre.shape) != (B, H + 12, W + 12, 8):
        raise JpdseError("tanh_backward: shape mismatch")
    check(lib.jpdse_tanh_backward_nchw(_ptr(grad_out), _ptr(out), _ptr(d_pre), _ptr(dbias), B, C, H, W, _stream()))
    _count()
    return d_pre


# ------------------------------------------------------------------------------------------------ layout
def nchw_to_nhwc_bf16(x, pad_reflect=0, c_pad=None, out=None):
    lib = _lib.load()
    _need(x, "x", torch.float32)
    B, C, H, W = x.shape
    c_pad = C if c_pad is None else c_pad
    if out is None:
        out = alloc_nhwc(B, H + 2 * pad_reflect, W + 2 * pad_reflect, c_pad, x.device)
    check(lib.jpdse_nchw_f32_to_nhwc_bf16(_ptr(x), _ptr(out), B, C, H, W, pad_reflect, c_pad, _stream()))
    _count()
    return out


def nhwc_bf16_to_nchw(x, out=None):
    lib = _lib.load()
    _need(x, "x", torch.bfloat16)
    B, H, W, C = x.shape
    if out is None:
        out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    check(lib.jpdse_nhwc_bf16_to_nchw_f32(_ptr(x), _ptr(out), B, C, H, W, _stream()))
    _count()
    return out


# ------------------------------------------------------------------------------------------------ quantisers
def round_f32(x):
    lib = _lib.load()
    _need(x, "x", torch.float32)
    y = torch.empty_like(x)
    check(lib.jpdse_round_f32(_ptr(x), _ptr(y), x.numel(), _stream()))
    _count()
    return y


def sign_f32(x):
    lib = _lib.load()
    _need(x, "x", torch.float32)
    y = torch.empty_like(x)
    check(lib.jpdse_sign_f32(_ptr(x), _ptr(y), x.numel(), _stream()))
    _count()
    return y


def softsign_f32(x, u):
    lib = _lib.load()
    _need(x, "x", torch.float32)
    _need(u, "u", torch.float32)
    if u.shape != x.shape:
        raise JpdseError("softsign: noise must have the input's shape")
    y = torch.empty_like(x)
    check(lib.jpdse_softsign_f32(_ptr(x), _ptr(u), _ptr(y), x.numel(), _stream()))
    _count()
    return y


def binarizer_train_forward(pre_nhwc, noise, y, tanh_out):
    """Binarizer train(): tanh + stochastic sign of the raw 1x1-conv output (see jpdse_binarizer_train_forward)."""
    lib = _lib.load()
    _need(pre_nhwc, "pre", torch.bfloat16)
    _need(noise, "noise", torch.float32)
    _need(y, "y", torch.float32)
    _need(tanh_out, "tanh_out", torch.float32)
    B, C, H, W = y.shape
    if tuple(pre_nhwc.shape) != (B, H, W, C) or noise.shape != y.shape or tanh_out.shape != y.shape:
        raise JpdseError("binarizer_train_forward: shape mismatch")
    check(lib.jpdse_binarizer_train_forward(_ptr(pre_nhwc), _ptr(noise), _ptr(y), _ptr(tanh_out), B, C, H, W, _stream()))
    _count()
    return y


def binarizer_train_backward(grad_y, tanh_out, d_pre_nhwc):
    lib = _lib.load()
    _need(grad_y, "grad_y", torch.float32)
    _need(tanh_out, "tanh_out", torch.float32)
    _need(d_pre_nhwc, "d_pre", torch.bfloat16)
    B, C, H, W = tanh_out.shape
    if grad_y.shape != tanh_out.shape or tuple(d_pre_nhwc.shape) != (B, H, W, C):
        raise JpdseError("binarizer_train_backward: shape mismatch")
    check(lib.jpdse_binarizer_train_backward(_ptr(grad_y), _ptr(tanh_out), _ptr(d_pre_nhwc), B, C, H, W, _stream()))
    _count()
    return d_pre_nhwc


def sign_to_bits(x):
    lib = _lib.load()
    _need(x, "x", torch.float32)
    y = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    check(lib.jpdse_sign_to_bits_u8(_ptr(x), _ptr(y), x.numel(), _stream()))
    _count()
    return y


def s2hvq_encode(x_rows, code_book, sigma, want_scores=False, want_index=False, want_one_hot=False, want_soft=False):
    """x_rows: (rows, center_size) float32; returns dict of the requested outputs."""
    lib = _lib.load()
    _need(x_rows, "x", torch.float32)
    _need(code_book, "code_book", torch.float32)
    rows, d = x_rows.shape
    L = code_book.shape[0]
    if code_book.shape[1] != d:
        raise JpdseError("s2hvq: x rows and code book centers differ in size")
    dev = x_rows.device
    out = {
        "scores": torch.empty((rows, L), dtype=torch.float32, device=dev) if want_scores else None,
        "index": torch.empty((rows,), dtype=torch.int64, device=dev) if want_index else None,
        "one_hot": torch.empty((rows, L), dtype=torch.float32, device=dev) if want_one_hot else None,
        "soft": torch.empty((rows, L), dtype=torch.float32, device=dev) if want_soft else None,
    }
    check(lib.jpdse_s2hvq_encode(_ptr(x_rows), _ptr(code_book), rows, d, L, float(sigma), _ptr(out["scores"]),
                                 _ptr(out["index"]), _ptr(out["one_hot"]), _ptr(out["soft"]), _stream()))
    _count()
    return out


def s2hvq_decode(code_raw_rows, code_book, want_index=False):
    lib = _lib.load()
    _need(code_raw_rows, "code_raw", torch.float32)
    _need(code_book, "code_book", torch.float32)
    rows, L = code_raw_rows.shape
    d = code_book.shape[1]
    if code_book.shape[0] != L:
        raise JpdseError("s2hvq decode: code_raw last dim must equal the number of centers")
    out = torch.empty((rows, d), dtype=torch.float32, device=code_raw_rows.device)
    idx = torch.empty((rows,), dtype=torch.int64, device=out.device) if want_index else None
    check(lib.jpdse_s2hvq_decode(_ptr(code_raw_rows), _ptr(code_book), rows, d, L, _ptr(out), _ptr(idx), _stream()))
    _count()
    return (out, idx) if want_index else out


# ------------------------------------------------------------------------------------------------ eval metric
def _dvec(v, n):
    v = list(v)[:n]
    return (ctypes.c_double * n)(*[float(a) for a in v])


def tensor2im_u8(x, mean=(0.5, 0.5, 0.5), std=(1.0, 1.0, 1.0)):
    """ctu/utils/misc.py tensor2im on-device: float32 (B,C,H,W) -> uint8 (B,H,W,C), bit-exact with numpy."""
    lib = _lib.load()
    _need(x, "x", torch.float32)
    B, C, H, W = x.shape
    out = torch.empty((B, H, W, C), dtype=torch.uint8, device=x.device)
    check(lib.jpdse_tensor2im_u8(_ptr(x), _ptr(out), B, C, H, W, _dvec(mean, C), _dvec(std, C), _stream()))
    _count()
    return out


def distortion_u8(a, b, mode="l1", mean=(0.5, 0.5, 0.5), std=(1.0, 1.0, 1.0)):
    """L1 / MSE between the tensor2im bytes of two float32 (B,C,H,W) images, as a 0-dim float64 tensor (no host sync)."""
    lib = _lib.load()
    _need(a, "a", torch.float32)
    _need(b, "b", torch.float32)
    if a.shape != b.shape:
        raise JpdseError("distortion_u8: shapes differ")
    B, C, H, W = a.shape
    acc = torch.zeros(1, dtype=torch.int64, device=a.device)
    check(lib.jpdse_distortion_u8(_ptr(a), _ptr(b), _ptr(acc), B, C, H, W, {"l1": 0, "mse": 1}[mode], _dvec(mean, C),
                                  _dvec(std, C), _stream()))
    _count(2)
    # tensor / tensor: a true IEEE division (torch turns tensor / python-scalar into a multiply by the reciprocal)
    return acc[0].double() / _const_f64(float(a.numel()), a.device)


_const_cache = {}


def _const_f64(value, device):
    """0-dim float64 device constant, built once per (value, device): torch.tensor(..., device=cuda) blocks the host."""
    key = (value, str(device))
    t = _const_cache.get(key)
    if t is None:
        t = _const_cache[key] = torch.tensor(value, dtype=torch.float64, device=device)
    return t
