"""ORACLE -- TEST INFRASTRUCTURE ONLY. Never imported by the product package (jpd-se_b200/).

CPU restatement of the reference's generator path, written from the behaviour of the reference code
(file:line relative to /root/reference). The arithmetic of the reference lives in PyTorch (pinned
torch==1.1.0 in setup.py:14; semantics of these ops unchanged in torch 2.11), so the restatement uses
torch's CPU fp32 functional ops -- the same third-party arithmetic the reference calls -- for the
convolutions and numpy for the integer work.

Pinning: `oracle/pin_against_reference.py` runs in the build container (where /root/reference exists),
asserts this restatement is bit-identical to the imported reference modules on CPU, and writes the
golden vectors under tests/golden/ that the CPU test-suite re-checks everywhere. The reference ships no
golden vectors of its own for this path, so parity is pinned by those outputs of the
reference itself.
"""
import numpy as np
import torch
import torch.nn.functional as F


# --------------------------------------------------------------------------------------------- preprocessing
def one_hot(label, num_labels):
    """Pix2PixHDModel.preprocess, ctu/models/pix2pixHD_model.py:376-382.

    label: array (B,1,H,W), any numeric dtype; `.long()` truncates toward zero. Returns float32
    (B,num_labels,H,W) with a 1.0 scattered at the label channel. Out-of-range ids raise (scatter_ does).
    """
    lab = np.asarray(label)
    lab = np.trunc(lab).astype(np.int64) if lab.dtype.kind == "f" else lab.astype(np.int64)
    if lab.min() < 0 or lab.max() >= num_labels:
        raise IndexError("label id out of range for scatter_")
    B, _, H, W = lab.shape
    out = np.zeros((B, num_labels, H, W), dtype=np.float32)
    b, h, w = np.meshgrid(np.arange(B), np.arange(H), np.arange(W), indexing="ij")
    out[b, lab[:, 0], h, w] = 1.0
    return out


def get_edges(inst):
    """Pix2PixHDModel.get_edges, ctu/models/pix2pixHD_model.py:774-783: both pixels of every
    horizontally / vertically differing neighbour pair are set; returns float32 (B,1,H,W)."""
    t = np.asarray(inst)
    e = np.zeros(t.shape, dtype=bool)
    dh = t[:, :, :, 1:] != t[:, :, :, :-1]
    dv = t[:, :, 1:, :] != t[:, :, :-1, :]
    e[:, :, :, 1:] |= dh
    e[:, :, :, :-1] |= dh
    e[:, :, 1:, :] |= dv
    e[:, :, :-1, :] |= dv
    return e.astype(np.float32)


def build_input(label, inst, image, num_labels):
    """cat(one-hot, edge) (pix2pixHD_model.py:394) then cat(input_label, image) (:595) -> (B,num_labels+4,H,W)."""
    return np.concatenate([one_hot(label, num_labels), get_edges(inst), np.asarray(image, dtype=np.float32)], axis=1)


def reflect_pad_nhwc(x_nchw, pad, c_pad=None):
    """ReflectionPad2d(pad) (networks.py:210) + NCHW->NHWC (+ zero channel padding): layout of the kernels' x0."""
    t = torch.from_numpy(np.ascontiguousarray(x_nchw))
    if pad:
        t = F.pad(t, (pad, pad, pad, pad), mode="reflect")
    t = t.permute(0, 2, 3, 1).contiguous()
    if c_pad is not None and c_pad > t.shape[-1]:
        t = F.pad(t, (0, c_pad - t.shape[-1]))
    return t.numpy()


# --------------------------------------------------------------------------------------------- generator
def _inorm(x, eps=1e-5):
    # nn.InstanceNorm2d(affine=False, track_running_stats=False): biased variance per (n, c) (networks.py:31)
    return F.instance_norm(x, eps=eps)


def generator_layers(sd, n_downsampling, n_blocks, binarize=False):
    """Yields (kind, prefix) in execution order for a GlobalGenerator state dict (networks.py:210-246); with
    binarize=True a Binarizer sits behind the res blocks (bin_before_res=False, networks.py:231-238)."""
    yield ("stem", "model.1")
    idx = 4
    for _ in range(n_downsampling):
        yield ("down", "model.%d" % idx)
        idx += 3
    for _ in range(n_blocks):
        yield ("res", "model.%d" % idx)
        idx += 1
    if binarize:
        yield ("bin", "model.%d" % idx)
        idx += 1
    for _ in range(n_downsampling):
        yield ("up", "model.%d" % idx)
        idx += 3
    yield ("head", "model.%d" % (idx + 1))


class _SoftSignSTE(torch.autograd.Function):
    """SoftSignFunction (ctu/quantizers/binarize.py:13-28) with the uniform draw supplied: +1 where (1 - x) / 2 <= u, else
    -1; the backward passes the gradient through unchanged."""

    @staticmethod
    def forward(ctx, x, u):
        y = x.clone()
        y[(1 - x) / 2 <= u] = 1
        y[(1 - x) / 2 > u] = -1
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


def generator_forward(sd, x, n_downsampling=4, n_blocks=9, round_fn=None, collect=None, binarize=False, codes_only=False,
                      noise=None):
    """GlobalGenerator.forward(mode='get_continuous_img'), networks.py:249-251, functional form.

    sd: state dict (torch tensors, reference keys); x: torch float32 (B,C,H,W).
    round_fn: optional activation/weight rounding hook -- `lambda t: t.bfloat16().float()` gives the
    "same arithmetic as the kernels" model (bf16 operands, fp32 accumulation) used for tight per-layer
    checks; None is the reference's fp32 path.
    collect: optional dict that receives every intermediate activation by layer prefix.
    """
    r = (lambda t: t) if round_fn is None else round_fn

    emulate = round_fn is not None

    def conv(t, prefix, head=False, **kw):
        # a bias in front of an affine-free InstanceNorm cancels exactly; the kernels skip it, and the
        # bf16-emulation mode must too (it would change the bf16 rounding of the raw conv output)
        bias = None if (emulate and not head) else sd[prefix + ".bias"]
        return F.conv2d(r(t), r(sd[prefix + ".weight"]), bias, **kw)

    for kind, prefix in generator_layers(sd, n_downsampling, n_blocks, binarize):
        if kind == "bin":
            # Binarizer eval (ctu/quantizers/binarize.py:51-54): sign(tanh(conv1x1_nobias(x)))
            # train() mode (noise given): the stochastic SoftSignFunction with the same uniform draw (binarize.py:37-41)
            t = torch.tanh(r(F.conv2d(r(x), r(sd[prefix + ".conv.weight"]))))
            x = torch.sign(t) if noise is None else _SoftSignSTE.apply(t, noise)
            if collect is not None:
                collect[prefix] = x
            if codes_only:
                return x
            continue
        if kind == "stem":
            x = F.relu(_inorm(r(conv(F.pad(x, (3, 3, 3, 3), mode="reflect"), prefix))))
        elif kind == "down":
            x = F.relu(_inorm(r(conv(x, prefix, stride=2, padding=1))))
        elif kind == "res":
            # ResnetBlock, networks.py:271-305: x + IN(conv(pad(ReLU(IN(conv(pad(x)))))))
            t = F.relu(_inorm(r(conv(F.pad(x, (1, 1, 1, 1), mode="reflect"), prefix + ".conv_block.1"))))
            t = r(t)
            t = _inorm(r(conv(F.pad(t, (1, 1, 1, 1), mode="reflect"), prefix + ".conv_block.5")))
            x = x + t
        elif kind == "up":
            t = F.conv_transpose2d(r(x), r(sd[prefix + ".weight"]), None if emulate else sd[prefix + ".bias"],
                                   stride=2, padding=1, output_padding=1)
            x = F.relu(_inorm(r(t)))
        else:
            x = torch.tanh(conv(F.pad(x, (3, 3, 3, 3), mode="reflect"), prefix, head=True))
        if kind != "head":
            x = r(x)
        if collect is not None:
            collect[prefix] = x
    return x


# --------------------------------------------------------------------------------------------- metrics
def psnr(a, b, peak=2.0):
    """PSNR on [-1, 1] images (peak-to-peak 2)."""
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * np.log10(peak * peak / mse)


def tensor2im_uint8(x, mean=(0.5, 0.5, 0.5), std=(1.0, 1.0, 1.0)):
    """ctu/utils/misc.py:64-95 for a (3,H,W) tensor: de-normalise, x255, clip, TRUNCATE to uint8 (HWC)."""
    a = x.detach().cpu().float().numpy()
    a = np.transpose(a, (1, 2, 0))
    a = (a * np.asarray(std) + np.asarray(mean)) * 255.0
    return np.clip(a, 0, 255).astype(np.uint8)


def eval_distortion(recon, real, mean=(0.5, 0.5, 0.5), std=(1.0, 1.0, 1.0), mode="l1"):
    """Pix2PixHDModel.get_eval_loss, ctu/models/pix2pixHD_model.py:636-641: both (B,3,H,W) images go through tensor2im
    (uint8 truncation), back to float, then nn.L1Loss / nn.MSELoss (mean). Restated as the EXACT integer sum of
    |a-b| (or (a-b)^2) over the bytes divided once by the element count, rounded to float32 -- what the device kernel
    (jpdse_distortion_u8) computes. The reference's float32 mean accumulates in floating point, so it can differ from
    this exact value in the last bits on large images; pin_against_reference.py measures that gap (<= 2 float32 ulp
    at the pinned sizes, 0 where every partial sum stays below 2^24)."""
    a = np.stack([tensor2im_uint8(x, mean, std) for x in recon]).astype(np.int64)
    b = np.stack([tensor2im_uint8(x, mean, std) for x in real]).astype(np.int64)
    d = np.abs(a - b) if mode == "l1" else (a - b) ** 2
    return np.float32(np.float64(int(d.sum())) / np.float64(d.size))
