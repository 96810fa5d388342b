"""ORACLE -- TEST INFRASTRUCTURE ONLY. Never imported by the product package (jpd-se_b200/).

CPU restatement of the forward passes of ctu/quantizers (file:line relative to /root/reference).
numpy for the element-wise / index work, torch CPU only where the reference's own arithmetic is a torch
library call whose summation order matters (the S2HVQ score matrix). Pinned against the imported
reference by oracle/pin_against_reference.py -> tests/golden/quantizers.npz.
"""
import numpy as np
import torch


def rounded_identity(x):
    """RoundedIdentity.forward = torch.round, IEEE round-half-to-even (ctu/quantizers/round.py:10-11)."""
    return np.rint(np.asarray(x, dtype=np.float32)).astype(np.float32)


def sign(x):
    """DifferentiableSign.forward in eval mode = x.sign() (ctu/quantizers/binarize.py:41):
    (x > 0) - (x < 0); NaN and -0.0 map to +0.0."""
    x = np.asarray(x, dtype=np.float32)
    return ((x > 0).astype(np.float32) - (x < 0).astype(np.float32)).astype(np.float32)


def soft_sign(x, u):
    """SoftSignFunction.forward with the uniform draw `u` given (ctu/quantizers/binarize.py:20-24):
    x[(1-x)/2 <= u] = 1 ; x[(1-x)/2 > u] = -1, both masks evaluated on the input."""
    x = np.asarray(x, dtype=np.float32)
    u = np.asarray(u, dtype=np.float32)
    t = (np.float32(1) - x) / np.float32(2)
    y = x.copy()
    y[t <= u] = 1
    y[t > u] = -1
    return y


def binarizer_eval(x_nchw, weight):
    """Binarizer.forward in eval mode: sign(tanh(conv1x1_nobias(x))) (ctu/quantizers/binarize.py:51-54).
    x (B,Cin,H,W) float32 torch tensor, weight (Cout,Cin,1,1)."""
    return torch.sign(torch.tanh(torch.nn.functional.conv2d(x_nchw, weight)))


def code_bits(x):
    """(code + 1) / 2 exported as uint8 (pix2pixHD_model.py:614, test.py:103-110)."""
    return ((np.asarray(x, dtype=np.float32) + 1) / 2).astype(np.uint8)


# ------------------------------------------------------------------------------------------ S2HVQ
def s2hvq_scores(x_mtrx, code_book):
    """_get_score_mtrx (ctu/quantizers/s2h_vq.py:85-87): (x.unsqueeze(2) - code_book).pow(2).sum(-1),
    evaluated with torch CPU so the fp32 summation order is the reference's."""
    x = torch.as_tensor(x_mtrx, dtype=torch.float32)
    c = torch.as_tensor(code_book, dtype=torch.float32)
    return (x.unsqueeze(dim=2) - c).pow_(2).sum(dim=-1)


def s2hvq_hard(x_mtrx, code_book):
    """_hard_quantize (:122-129): argmin (first index on ties) -> one-hot. Returns (index, one_hot)."""
    sc = s2hvq_scores(x_mtrx, code_book)
    _, idx = torch.min(sc, dim=-1, keepdim=True)
    one_hot = torch.zeros_like(sc).scatter_(2, idx, 1)
    return idx.squeeze(-1), one_hot


def s2hvq_soft(x_mtrx, code_book, sigma):
    """_soft_quantize (:107-108): softmax(-sigma * scores)."""
    return torch.softmax(s2hvq_scores(x_mtrx, code_book).mul_(-sigma), dim=-1)


def s2hvq_decode(code_raw, code_book):
    """decode (:182-183, :205-207): argmax over centers -> code book gather -> (n, code_len*center_size)."""
    cr = torch.as_tensor(code_raw, dtype=torch.float32)
    cb = torch.as_tensor(code_book, dtype=torch.float32)
    _, idx = torch.max(cr, dim=-1)
    m = cb[idx]
    return m.reshape(-1, m.size(1) * m.size(2)), idx
