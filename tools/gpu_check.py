"""Developer bring-up checks for the sm_100a kernels (run on the GPU box through gpurun).

Usage: python tools/gpu_check.py [stage ...]   (no stage = all, each in its own subprocess with a timeout so a
hung kernel cannot take the whole call down). Compares every kernel with the CPU oracle / torch CPU fp32.
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["elementwise", "conv1x1", "conv3x3", "conv3x3_big", "convs2", "convt", "stem", "head", "generator_small",
          "generator_emul"]


def _bf(t):
    return t.bfloat16().float()


def stage_elementwise():
    import numpy as np
    import torch
    import jpdse_b200  # noqa: F401
    from jpdse_b200 import ops
    from oracle import generator_oracle as orc
    g = torch.Generator().manual_seed(1)
    B, H, W, L = 2, 24, 40, 35
    label = torch.randint(0, L, (B, 1, H, W), generator=g).float()
    inst = torch.randint(0, 5, (B, 1, H // 4, W // 4), generator=g).repeat_interleave(4, 2).repeat_interleave(4, 3).int()
    image = torch.rand(B, 3, H, W, generator=g) - 0.5
    ref = orc.build_input(label.numpy(), inst.numpy(), image.numpy(), L)
    nhwc, nchw = ops.build_input(label.cuda(), inst.cuda(), image.cuda(), L, pad=3, c_pad=40, nhwc=True, nchw=True)
    torch.cuda.synchronize()
    ok1 = np.array_equal(nchw.cpu().numpy(), ref)
    ref_nhwc = _bf(torch.from_numpy(orc.reflect_pad_nhwc(ref, 3, 40))).numpy()
    ok2 = np.array_equal(nhwc.float().cpu().numpy(), ref_nhwc)
    print("build_input nchw exact:", ok1, " nhwc exact:", ok2)
    # layout converts
    x = torch.randn(2, 39, 16, 24, generator=g)
    y = ops.nchw_to_nhwc_bf16(x.cuda(), pad_reflect=3, c_pad=40)
    ok3 = np.array_equal(y.float().cpu().numpy(), _bf(torch.from_numpy(orc.reflect_pad_nhwc(x.numpy(), 3, 40))).numpy())
    z = ops.nhwc_bf16_to_nchw(y)
    ok4 = np.array_equal(z.cpu().numpy(), y.float().permute(0, 3, 1, 2).cpu().numpy())
    print("nchw->nhwc exact:", ok3, " nhwc->nchw exact:", ok4)
    # quantisers
    q = torch.randn(100003, generator=g) * 3
    q[:8] = torch.tensor([0.5, 1.5, 2.5, -0.5, -1.5, float("nan"), 0.0, -0.0])
    ok5 = torch.equal(ops.round_f32(q.cuda()).cpu().nan_to_num(7), torch.round(q).nan_to_num(7))
    ok6 = torch.equal(ops.sign_f32(q.cuda()).cpu(), torch.sign(q))
    print("round exact:", ok5, " sign exact:", ok6)
    return all([ok1, ok2, ok3, ok4, ok5, ok6])


def _run_conv(kind_name, B, H, W, cin, cout, cin_real=None, seed=0):
    import torch
    import torch.nn.functional as F
    import jpdse_b200  # noqa: F401
    from jpdse_b200 import ops
    from jpdse_b200._lib import (CONV1X1, CONV3X3_PAD1, CONV3X3_S2, CONV7X7_PAD3, CONVT3X3_S2, EPI_BIAS_TANH_NCHW,
                                 EPI_RAW_STATS)
    g = torch.Generator().manual_seed(seed)
    cin_real = cin if cin_real is None else cin_real
    dev = torch.device("cuda")
    x = _bf(torch.randn(B, cin_real, H, W, generator=g))
    if kind_name == "conv1x1":
        kind, pad, k = CONV1X1, 0, 1
        w = _bf(torch.randn(cout, cin_real, 1, 1, generator=g) * 0.05)
        ref = F.conv2d(x, w)
    elif kind_name == "conv3x3":
        kind, pad, k = CONV3X3_PAD1, 1, 3
        w = _bf(torch.randn(cout, cin_real, 3, 3, generator=g) * 0.05)
        ref = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w)
    elif kind_name == "convs2":
        kind, pad, k = CONV3X3_S2, 0, 3
        w = _bf(torch.randn(cout, cin_real, 3, 3, generator=g) * 0.05)
        ref = F.conv2d(x, w, stride=2, padding=1)
    elif kind_name == "convt":
        kind, pad, k = CONVT3X3_S2, 0, 3
        w = _bf(torch.randn(cin_real, cout, 3, 3, generator=g) * 0.05)
        ref = F.conv_transpose2d(x, w, stride=2, padding=1, output_padding=1)
    elif kind_name in ("stem", "head"):
        kind, pad, k = CONV7X7_PAD3, 3, 7
        w = _bf(torch.randn(cout, cin_real, 7, 7, generator=g) * 0.05)
        ref = F.conv2d(F.pad(x, (3, 3, 3, 3), mode="reflect"), w)
    head = kind_name == "head"
    epi = EPI_BIAS_TANH_NCHW if head else EPI_RAW_STATS
    bias = torch.randn(cout, generator=g) * 0.1 if head else None
    xd = ops.nchw_to_nhwc_bf16(x.cuda(), pad_reflect=pad, c_pad=cin)
    cv = ops.Conv(kind, epi, B, H, W, pad, cin, cin_real, cout, dev)
    cv.pack(w.cuda(), None if bias is None else bias.cuda())
    oh, ow = cv.out_hw
    if head:
        y = torch.full((B, cout, oh, ow), float("nan"), device=dev)
        cv.forward(xd, y)
        torch.cuda.synchronize()
        ref = torch.tanh(ref + bias.view(1, -1, 1, 1))
        err = (y.cpu() - ref).abs().max().item()
        print("%s B%d %dx%d %d->%d: max abs err %.3e (ref max %.3f)" % (kind_name, B, H, W, cin_real, cout, err,
                                                                     ref.abs().max()))
        return err < 2e-3
    y = torch.full((B, oh, ow, cout), float("nan"), dtype=torch.bfloat16, device=dev)
    stats = torch.zeros(B, cout, 2, dtype=torch.float64, device=dev)
    t0 = time.time()
    cv.forward(xd, y, stats)
    torch.cuda.synchronize()
    dt = time.time() - t0
    got = y.float().cpu().permute(0, 3, 1, 2)
    refb = _bf(ref)
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item()
    s1 = refb.double().sum(dim=(2, 3))
    s2 = (refb.double() ** 2).sum(dim=(2, 3))
    e1 = (stats[:, :, 0].cpu() - s1).abs().max().item()
    e2 = ((stats[:, :, 1].cpu() - s2).abs() / s2.clamp_min(1e-6)).max().item()
    nan = torch.isnan(got).sum().item()
    print("%s B%d %dx%d %d->%d: max abs err %.3e (ref max %.3f) nan %d | stats err sum %.3e sumsq rel %.3e | %.1f ms"
          % (kind_name, B, H, W, cin_real, cout, err, scale, nan, e1, e2, dt * 1e3))
    return nan == 0 and err < 0.02 * max(scale, 1.0) and e1 < 0.5 and e2 < 2e-2


def stage_conv1x1():
    return _run_conv("conv1x1", 2, 8, 16, 64, 64) and _run_conv("conv1x1", 1, 16, 32, 128, 256, seed=1)


def stage_conv3x3():
    return _run_conv("conv3x3", 2, 8, 16, 64, 64) and _run_conv("conv3x3", 2, 16, 16, 128, 256, seed=2)


def stage_conv3x3_big():
    return _run_conv("conv3x3", 3, 32, 64, 1024, 1024, seed=3)


def stage_convs2():
    return _run_conv("convs2", 2, 16, 32, 64, 128) and _run_conv("convs2", 1, 32, 64, 128, 256, seed=4)


def stage_convt():
    return _run_conv("convt", 2, 8, 16, 128, 64) and _run_conv("convt", 1, 16, 32, 256, 128, seed=5)


def stage_stem():
    return (_run_conv("stem", 2, 8, 16, 40, 64, cin_real=39) and _run_conv("stem", 1, 16, 128, 40, 64, cin_real=39, seed=6)
            and _run_conv("stem", 2, 140, 256, 40, 64, cin_real=39, seed=8))


def stage_head():
    return (_run_conv("head", 2, 8, 16, 64, 3) and _run_conv("head", 1, 16, 128, 64, 3, seed=7)
            and _run_conv("head", 2, 140, 256, 64, 3, seed=9))


def _gen_case(B, H, W, n_down, n_blocks, seed=1234):
    import torch
    import jpdse_b200  # noqa: F401
    networks = __import__("importlib").import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
    torch.manual_seed(seed)
    net = networks.define_G(39, 3, 64, "global", n_down, n_blocks, 1, 3, "instance", gpu_ids=[])
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    g = torch.Generator().manual_seed(seed)
    x = torch.zeros(B, 39, H, W)
    lab = torch.randint(0, 35, (B, H // 8, W // 8), generator=g).repeat_interleave(8, 1).repeat_interleave(8, 2)
    x.scatter_(1, lab.unsqueeze(1), 1.0)
    x[:, 36:] = torch.rand(B, 3, H, W, generator=g) - 0.5
    return net, sd, x


def stage_generator_small():
    import torch
    from oracle import generator_oracle as orc
    net, sd, x = _gen_case(2, 128, 256, 4, 9)
    ref = orc.generator_forward(sd, x, 4, 9)
    net = net.cuda().eval()
    with torch.no_grad():
        y = net(x.cuda())
    torch.cuda.synchronize()
    y = y.cpu()
    err = (y - ref).abs()
    print("generator 2x128x256: max abs %.4f mean abs %.5f psnr %.2f dB nan %d" % (
        err.max(), err.mean(), orc.psnr(y, ref), torch.isnan(y).sum()))
    return bool(err.mean() < 0.03 and not torch.isnan(y).any())


def stage_generator_emul():
    import torch
    from oracle import generator_oracle as orc
    net, sd, x = _gen_case(1, 128, 256, 4, 2)
    ref = orc.generator_forward(sd, x, 4, 2, round_fn=_bf)
    ref32 = orc.generator_forward(sd, x, 4, 2)
    net = net.cuda().eval()
    with torch.no_grad():
        y = net(x.cuda()).cpu()
    err = (y - ref).abs()
    e32 = (y - ref32).abs()
    print("generator(2 blocks) vs bf16-emulated oracle: max %.4f mean %.5f | vs fp32 oracle: max %.4f mean %.5f psnr %.2f"
          % (err.max(), err.mean(), e32.max(), e32.mean(), orc.psnr(y, ref32)))
    return bool(err.mean() < 0.01)


def main():
    args = sys.argv[1:]
    if len(args) == 1 and args[0].startswith("stage:"):
        ok = globals()["stage_" + args[0][6:]]()
        print("STAGE %s: %s" % (args[0][6:], "PASS" if ok else "FAIL"))
        sys.exit(0 if ok else 1)
    stages = args or STAGES
    results = {}
    for s in stages:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "stage:" + s], timeout=180)
            results[s] = "ok" if r.returncode == 0 else "FAIL(rc=%d)" % r.returncode
        except subprocess.TimeoutExpired:
            results[s] = "TIMEOUT"
        print("== %s: %s (%.1fs)" % (s, results[s], time.time() - t0), flush=True)
        if results[s] == "TIMEOUT":
            print("stopping after a hang")
            break
    print("SUMMARY", results)
    sys.exit(0 if all(v == "ok" for v in results.values()) else 1)


if __name__ == "__main__":
    main()
