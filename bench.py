#!/usr/bin/env python
"""Benchmark of the JPD-SE generator hot path on B200 (metric and config from BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--height H --width W] [--impl reference]

A step = one forward of the semantic-aware pix2pixHD generator (fused one-hot/edge/concat input build +
182.6 M-parameter GlobalGenerator) over one batch of synthetic Cityscapes-shaped inputs
(BASELINE.json configs[1]: batch 16 at 1024x512, bf16 kernels, random-init weights).

One JSON line on stdout (rank 0):
  value    images/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e      images/s through the ctu-API call (trainer.get_img) with pinned HOST inputs and a device->host
           read of the output image inside the timed region
  roofline the residual-block 3x3 conv kernel (70 % of the FLOPs): algorithmic FLOPs / CUDA-event time
  cpu_baseline  the CPU restatement of the reference path (oracle/) on this box's host cores
`--impl reference` times that CPU path alone, with the same metric/config keys.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "images/sec (1024x512 codec fwd)"
RES_CONV_FLOPS_PER_IMAGE = 38654705664.0   # 2*2048*1024*9216
FWD_FLOPS_PER_IMAGE = 988513566720.0


def synth_inputs(batch, height, width, seed=1234, qf=36):
    """tiled label/instance maps, uniform image, quantise+noise stand-in for BPG."""
    import torch
    g = torch.Generator().manual_seed(seed)
    th, tw = height // 32, width // 32
    label = torch.randint(0, 34, (batch, 1, th, tw), generator=g).repeat_interleave(32, 2).repeat_interleave(32, 3)
    inst = torch.randint(0, 50, (batch, 1, th, tw), generator=g).repeat_interleave(32, 2).repeat_interleave(32, 3)
    image = torch.rand(batch, 3, height, width, generator=g) - 0.5
    q, sigma = {33: (4, 0.01), 36: (6, 0.015), 39: (8, 0.02), 42: (12, 0.03)}[qf]
    deg = torch.round((image + 0.5) * 255.0 / q) * q / 255.0 - 0.5 + torch.randn(image.shape, generator=g) * sigma
    return label.float(), inst.int(), deg.clamp_(-0.5, 0.5)


def synth_inputs_compact(batch, height, width, seed=1234, qf=36):
    """The same synthetic workload in the COMPACT loader format: uint8 class ids, int16 instance ids,
    uint8 RGB of the degraded image (what a BPG/PNG decoder hands over); 3.1 MB per 1024x512 image instead of 10.5 MB.
    Also returns the float32 image the loader's ToTensor + Normalize(.5, 1.) would make of those bytes -- the tensor the
    reference's x_dict['image'] carries (ctu/data/ctu_dataset.py:73-133)."""
    import torch
    label, inst, deg = synth_inputs(batch, height, width, seed, qf)
    img_u8 = torch.round((deg + 0.5) * 255.0).clamp_(0, 255).to(torch.uint8)
    img_f32 = img_u8.float() / 255.0 - 0.5          # ToTensor then Normalize(mean .5, std 1.): exact in float32
    return label.to(torch.uint8), inst.to(torch.int16), img_u8, img_f32


def make_opt():
    return argparse.Namespace(model="pix2pixHD", gpu_ids=[0], is_train=False, num_labels=35,
                              contain_dontcare_label=False, no_label=False, no_instance=False, no_feat=False,
                              no_label_encoding=True, no_feat_encoding=True, no_generator_binarization=True,
                              sem_masking=False, input_nc=3, num_out_channels=3, ngf=64, netG="global",
                              n_downsample_global=4, n_blocks_global=9, n_local_enhancers=1, n_blocks_local=3,
                              norm="instance", use_compressed=False)


class ClockSampler:
    """SM clocks / throttle reasons sampled by one long-lived `nvidia-smi -lms 100` while the timed region runs."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        self.t0 = self.t1 = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def begin(self):
        self.t0 = time.time()

    def end(self):
        self.t1 = time.time()

    def summary(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ""
        import datetime
        rows = []
        for ln in out.splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(f[1]), float(f[2]), f[3:7]))
            except ValueError:
                continue
        inside = [r for r in rows if self.t0 is not None and self.t0 - 0.05 <= r[0] <= self.t1 + 0.05]
        window = "timed region"
        if not inside:  # region shorter than the sampling period: use the samples taken under load (warm-up included)
            inside, window = [r for r in rows if r[1] > 0], "warm-up + timed region"
        mhz = sorted(r[1] for r in inside)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3][i].lower().startswith("active") for r in inside)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None, "sm_max_mhz": max([r[2] for r in inside] or [0]) or None,
                "reasons": reasons, "samples": len(inside), "window": window}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops")), d.get("bf16_tflops"), "measured"
    return 1400.0, 1590.0, "fallback"


def _reference_arm():
    import importlib.util
    spec = importlib.util.spec_from_file_location("reference_arm", os.path.join(ROOT, "tools", "reference_arm.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cpu_reference_run(label, inst, image, steps, warmup, weights=None):
    """Times the reference's CPU path for ONE batch of inputs on all host threads; returns (s/step, output, kind).

    kind "reference": the UNMODIFIED reference installed in baseline/_ref (tools/install_reference.sh), through its own
    public API -- parser -> get_trainer(opt)(opt, 'test') -> trainer.get_img(x_dict) (one-hot scatter_, get_edges, cat,
    GlobalGenerator on ATen/oneDNN). kind "port": the oracle restatement, only when baseline/_ref is absent.
    `weights`: generator state dict to load (reference keys); None = reference define_G under seed 1234."""
    import torch
    torch.set_num_threads(os.cpu_count() or 1)
    ra = _reference_arm()
    times, out = [], None
    if ra.available():
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):  # the reference prints banners; stdout carries ONE JSON line
            trainer, _opt = ra.build_test_trainer(state_dict=weights)
        x_dict = {"label": label.float(), "instance": inst.int(), "image": image.float(), "path": ["synthetic"] * label.shape[0]}
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out = trainer.get_img({k: (v.clone() if hasattr(v, "clone") else v) for k, v in x_dict.items()})
            if i >= warmup:
                times.append(time.perf_counter() - t0)
        return sum(times) / len(times), out, "reference"
    from oracle import generator_oracle as orc
    if weights is None:
        import jpdse_b200  # noqa: F401
        import importlib
        nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
        torch.manual_seed(1234)
        weights = nw.define_G(39, 3, 64, "global", 4, 9, 1, 3, "instance", gpu_ids=[]).state_dict()
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            x = torch.from_numpy(orc.build_input(label.float().numpy(), inst.int().numpy(), image.float().numpy(), 35))
            out = orc.generator_forward(weights, x, 4, 9)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times), out, "port"


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores, one image of the
    workload per step (bounded sample), same metric / config keys as our arm. Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample_b = 1  # bounded sample: one image per step
    _l, _i, _u8, img = synth_inputs_compact(args.batch, args.height, args.width, seed=1234)
    label, inst = _l[:sample_b], _i[:sample_b]
    sec, _out, kind = cpu_reference_run(label, inst, img[:sample_b], args.steps, args.warmup)
    v = sample_b / sec
    what = ("unmodified reference (baseline/_ref) via parser -> get_trainer -> Pix2PixHDTrainer.get_img" if kind == "reference"
            else "oracle port of the reference path (baseline/_ref not installed)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "images/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": v, "unit": "images/s", "cores": cores, "kind": kind,
                             "sample": "image 0 of the batch-%d %dx%d workload per step (batch 1), fp32, torch CPU "
                                       "(oneDNN), %d threads; %s" % (args.batch, args.width, args.height, cores, what)},
            "e2e": {"value": v, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def ncu_traffic(batch, height, width):
    """DRAM bytes per launch of the roofline kernel from the committed `ncu --set full` capture (profiles/), only
    when the capture was taken at this workload; else None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        d = json.load(f)
    for cap in d.get("captures", [d]):
        if (cap.get("batch"), cap.get("height"), cap.get("width")) == (batch, height, width):
            return cap.get("dram_bytes_per_launch")
    return None


def train_measure(args, dev, local, rank, world, barrier):
    """Generator forward+backward on the sm_100a kernels and the whole Pix2PixHDTrainer.step (netD / VGG / losses /
    Adam in stock PyTorch), data-parallel over `world` GPUs with the gradient all-reduce overlapped with backward.
    Not the headline metric: reported under "train"."""
    import importlib
    import torch
    import torch.distributed as dist
    tr = importlib.import_module("jpd-se_b200.ctu.trainers.pix2pixHD_trainer")
    # no pretrained VGG19 checkpoint on an offline box: explicit opt-in to random VGG weights (same FLOPs, recorded below)
    os.environ.setdefault("JPDSE_VGG_RANDOM", "1")
    opt = make_opt()
    opt.gpu_ids, opt.is_train, opt.quiet = [local], True, True
    torch.manual_seed(1234)
    trainer = tr.Pix2PixHDTrainer(opt, mode="train")
    net = trainer.model.netG
    B, H, W = args.train_batch, args.height, args.width
    label, inst, image = synth_inputs(B, H, W, seed=4321 + rank)
    x_dict = {"label": label.to(dev), "instance": inst.to(dev), "image": image.to(dev)}
    steps = max(3, min(args.steps, 10))

    def g_fwd_bwd():
        for p in net.parameters():
            p.grad = None
        y = net.forward_from_maps(x_dict["label"], x_dict["instance"], x_dict["image"], 35)
        ((y - x_dict["image"]).abs().mean() * 10.0).backward()

    def timed(fn):
        for _ in range(2):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    g_ms = timed(g_fwd_bwd)
    step_ms = timed(lambda: trainer.step(x_dict))
    scale = (H * W) / (512.0 * 1024.0)
    # backward = dgrad + wgrad of every conv except the stem's dgrad
    g_flops = (3 * FWD_FLOPS_PER_IMAGE - 128245039104.0) * scale * B
    del trainer
    torch.cuda.empty_cache()
    return {"batch_per_gpu": B, "steps": steps, "generator_fwd_bwd_ms": g_ms,
            "generator_fwd_bwd_images_per_s": world * B / (g_ms * 1e-3),
            "generator_fwd_bwd_tflops_per_gpu": g_flops / (g_ms * 1e-3) / 1e12,
            "trainer_step_ms": step_ms, "trainer_step_images_per_s": world * B / (step_ms * 1e-3),
            "allreduce": "none (1 GPU)" if world == 1 else "730 MB fp32 generator gradients, bucketed, overlapped with backward (NCCL)",
            "vgg_weights": "random (JPDSE_VGG_RANDOM=1: the pretrained checkpoint is not available offline; same FLOPs)",
            "note": "generator, PatchGAN discriminator (GAN + feature-matching losses) and VGG19 loss forward/backward = jpdse_b200 "
                    "kernels; Adam (torch.optim, fused) and the scalar loss arithmetic = PyTorch"}


def cudnn_measure(B, H, W, dev):
    """The existing Blackwell path for scale: the same GlobalGenerator architecture built from stock torch.nn modules
    (ATen / cuDNN) on this GPU, bf16 autocast + channels_last, device-resident, CUDA-event timed. Not on our path."""
    import importlib.util
    import torch
    spec = importlib.util.spec_from_file_location("cudnn_baseline", os.path.join(ROOT, "tools", "cudnn_baseline.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    torch.manual_seed(1234)
    net = mod.generator().to(dev).eval().to(memory_format=torch.channels_last)
    x = torch.randn(B, 39, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        ms = mod.timed(lambda: net(x), 5)
    del net, x
    torch.cuda.empty_cache()
    return {"value": B / ms * 1e3, "unit": "images/s", "ms_per_step": ms,
            "what": "same architecture on stock torch.nn modules (ATen/cuDNN), bf16 autocast + channels_last, batch %d, "
                    "device-resident, this GPU" % B}


def workload_config(args):
    return {"workload": "pix2pixHD-BPG QF36 semantic-aware generator inference, batch %d at %dx%d, 35-class label map "
                        "+ instance edges + RGB, random-init weights" % (args.batch, args.width, args.height),
            "batch_per_gpu": args.batch, "height": args.height, "width": args.width,
            "inputs": "compact loader format: uint8 class ids, int16 instance ids, uint8 RGB (normalised on the device)",
            "execution": "two half-batch plans on two CUDA streams, one captured CUDA graph per step (JPDSE_SPLIT_STREAMS=1 "
                         "for a single stream)" if args.batch >= 8 and args.batch % 2 == 0 else "single stream, one CUDA graph per step",
            "l2": "inputs+activations per step are GBs >> 126 MB L2 (no flush needed)"}


def _max_over_ranks(ms, world, dev):
    if world > 1:
        import torch
        import torch.distributed as dist
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def device_resident(plan, inputs, steps, warmup, barrier, sampler=None):
    """Device-resident loop: inputs already in HBM, K forwards bracketed by barrier + synchronize, CUDA-event timed."""
    import torch
    from jpdse_b200 import ops
    d_label, d_inst, d_image = inputs
    for _ in range(warmup):
        plan.forward_from_maps(d_label, d_inst, d_image, 35)
    barrier()
    ops.launch_count = 0
    if sampler:
        sampler.begin()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(steps):
        plan.forward_from_maps(d_label, d_inst, d_image, 35)
    end.record()
    barrier()
    if sampler:
        sampler.end()
    return start.elapsed_time(end), ops.launch_count


def e2e_measure(trainer, host, B, H, W, steps, barrier, dev):
    """End to end through the ctu API. Every step: H2D of that step's pinned host inputs, trainer.get_img(x_dict) (the
    call test.py makes, ctu/trainers/pix2pixHD_trainer.py:113-116), D2H of the output image. Copies run on their own
    streams with double-buffered device inputs, so step i+1's upload and step i-1's download overlap step i's kernels
    -- all of them inside the timed region. Returns (ms total, h2d bytes/step, d2h bytes/step)."""
    import torch
    pin = {k: v.pin_memory() for k, v in host.items()}
    host_out = [torch.empty((B, 3, H, W), dtype=torch.float32).pin_memory() for _ in range(2)]
    dev_in = [{k: torch.empty_like(v, device=dev) for k, v in pin.items()} for _ in range(2)]
    h2d_stream, d2h_stream = torch.cuda.Stream(), torch.cuda.Stream()
    main = torch.cuda.current_stream()
    in_ready = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]
    out_done = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        s = i % 2
        with torch.cuda.stream(h2d_stream):
            h2d_stream.wait_event(in_free[s])
            for k, v in pin.items():
                dev_in[s][k].copy_(v, non_blocking=True)
            in_ready[s].record(h2d_stream)

    def loop(n):
        for s in range(2):
            in_free[s].record(main)
        upload(0)
        for i in range(n):
            s = i % 2
            if i + 1 < n:
                upload(i + 1)
            main.wait_event(in_ready[s])
            out = trainer.get_img(dict(dev_in[s], path=["synthetic"] * B))
            in_free[s].record(main)
            done = torch.cuda.Event()
            done.record(main)
            out.record_stream(d2h_stream)
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(done)
                host_out[s].copy_(out, non_blocking=True)
                out_done[s].record(d2h_stream)
        for s in range(2):
            main.wait_event(out_done[s])

    with torch.no_grad():
        loop(2)
        barrier()
        s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        loop(steps)
        e2.record()
        barrier()
    trainer.model.check_labels()
    h2d = sum(v.numel() * v.element_size() for v in pin.values())
    d2h = host_out[0].numel() * host_out[0].element_size()
    return s2.elapsed_time(e2), h2d, d2h, host_out[(steps - 1) % 2]


def bandwidth_measure(dev, hbm_peak):
    """Achieved HBM GB/s of the memory-bound kernels of the path (north_star's evidence clause), each timed alone with
    CUDA events over buffers larger than the 126 MB L2 (unless noted), algorithmic bytes / time, vs the measured copy
    peak (MEASURED_PEAKS.json hbm_gbs)."""
    import torch
    from jpdse_b200 import ops

    def timed(fn, reps=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    rows = []

    def add(name, nbytes, fn, note=None):
        ms = timed(fn)
        gbs = nbytes / (ms * 1e-3) / 1e9
        r = {"kernel": name, "bytes": int(nbytes), "ms": ms, "gbs": gbs, "frac_of_hbm_peak": gbs / hbm_peak}
        if note:
            r["note"] = note
        rows.append(r)

    B, H, W = 16, 512, 1024
    px = B * H * W
    lab8, ins16, img8, imgf = synth_inputs_compact(B, H, W)
    x0 = ops.alloc_nhwc(B, H + 6, W + 6, 40, dev)
    out_b = B * (H + 6) * (W + 6) * 80
    lf, ii, fi = lab8.float().to(dev), ins16.int().to(dev), imgf.to(dev)
    add("build_input_nhwc_kernel (float32 ids + int32 instance + float32 image -> padded NHWC bf16 x40)", px * 20 + out_b,
        lambda: ops.build_input(lf, ii, fi, 35, pad=3, c_pad=40, out_nhwc=x0))
    l8, i16, u8 = lab8.to(dev), ins16.to(dev), img8.to(dev)
    add("build_input_nhwc_kernel (compact: uint8 ids + int16 instance + uint8 RGB)", px * 6 + out_b,
        lambda: ops.build_input(l8, i16, u8, 35, pad=3, c_pad=40, out_nhwc=x0))
    del lf, ii, fi, l8, i16, u8, x0
    for (b, h, w, c, pad, relu, res, note) in ((8, 512, 1024, 64, 3, True, False, None), (8, 256, 512, 128, 0, True, False, None),
                                               (8, 64, 128, 512, 0, True, False, None),
                                               (8, 32, 64, 1024, 1, False, True, "67 MB in + 2 x 38 MB: partly L2-resident, as in the real pipeline")):
        raw = torch.randn(b, h, w, c, device=dev).bfloat16()
        st = torch.zeros(b, c, 2, dtype=torch.float64, device=dev)
        st[:, :, 1] = float(h * w)
        out = ops.alloc_nhwc(b, h + 2 * pad, w + 2 * pad, c, dev)
        resid = ops.alloc_nhwc(b, h + 2 * pad, w + 2 * pad, c, dev) if res else None
        nbytes = b * h * w * c * 2 + b * (h + 2 * pad) * (w + 2 * pad) * c * 2 * (2 if res else 1)
        add("instnorm_apply_kernel c%d %dx%d pad %d%s%s" % (c, w, h, pad, " +relu" if relu else "", " +residual" if res else ""),
            nbytes, lambda: ops.instnorm_apply(raw, st, out, b, h, w, c, pad, relu, residual=resid), note)
        del raw, st, out, resid
    xq = torch.randn(1 << 26, device=dev)
    add("quant_elementwise_kernel round (RoundedIdentity fwd)", xq.numel() * 8, lambda: ops.round_f32(xq))
    add("quant_elementwise_kernel sign (DifferentiableSign eval)", xq.numel() * 8, lambda: ops.sign_f32(xq))
    add("sign_to_bits_kernel", xq.numel() * 5, lambda: ops.sign_to_bits(xq))
    rows_n, csz, ncen = 1 << 23, 8, 16
    xv = xq[: rows_n * csz].view(rows_n, csz)
    cb = torch.randn(ncen, csz, device=dev)
    add("s2hvq_encode (index only, %d centers x %d)" % (ncen, csz), rows_n * (csz * 4 + 8),
        lambda: ops.s2hvq_encode(xv, cb, 1.0, want_index=True))
    add("s2hvq_encode (hard one-hot out)", rows_n * (csz * 4 + ncen * 4), lambda: ops.s2hvq_encode(xv, cb, 1.0, want_one_hot=True))
    img = xq[: 16 * 3 * 512 * 1024].view(16, 3, 512, 1024)
    add("tensor2im_u8_kernel", img.numel() * 5, lambda: ops.tensor2im_u8(img))
    del xq
    torch.cuda.empty_cache()
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--height", type=int, default=512)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the (extra, non-headline) training-step measurement")
    ap.add_argument("--no-cudnn-baseline", action="store_true",
                    help="skip timing the same architecture on stock PyTorch/cuDNN on this GPU (extra, rank 0 at N=1)")
    ap.add_argument("--no-fullres", action="store_true", help="skip the 2048x1024 batch-4 measurement (BASELINE configs[2])")
    ap.add_argument("--no-bandwidth", action="store_true", help="skip the memory-bound kernels' GB/s table (rank 0 at N=1)")
    ap.add_argument("--train-batch", type=int, default=2, help="images per GPU of the training-step measurement")
    ap.add_argument("--fullres-batch", type=int, default=4, help="images per GPU of the 2048x1024 measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import jpdse_b200  # noqa: F401  (raises if libjpdse_b200.so is missing)
    import importlib
    trainers = importlib.import_module("jpd-se_b200.ctu.trainers")

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    opt = make_opt()
    opt.gpu_ids = [local]
    torch.manual_seed(1234)
    trainer = trainers.get_trainer(opt)(opt, "test")
    netG = trainer.model.netG
    B, H, W = args.batch, args.height, args.width
    # Image-sharded inference: the job is world*B images, image i -> rank i mod N (jpd-se_b200/sharding.py),
    # every image generated from its own seed (1234 + i), so an image's output does not depend on N or on its batch slot.
    sharding = importlib.import_module("jpd-se_b200.sharding")
    my_images = sharding.shard_indices(world * B, rank, world)
    parts = [synth_inputs_compact(1, H, W, seed=1234 + i) for i in my_images]
    lab8, ins16, img8, img_f32 = (torch.cat([p_[k] for p_ in parts], 0) for k in range(4))
    d_in = (lab8.to(dev), ins16.to(dev), img8.to(dev))
    plan = netG.plan_for(B, H, W, dev)  # batch >= 8: two half-batch plans on two streams, captured into one CUDA graph

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident throughput (the timed region)
    with torch.no_grad():
        sampler = ClockSampler(local) if rank == 0 else None
        elapsed_ms, launches = device_resident(plan, d_in, args.steps, args.warmup, barrier, sampler)
        clocks = sampler.summary() if sampler else None
        out_dev0 = plan.out[0].detach().cpu() if rank == 0 else None  # image 0 of the timed batch, for the parity object
        # the ONLY exchange of the inference path: one gather of a per-image result (here a checksum), outside the timed region
        checks = sharding.gather_results(plan.out.double().mean(dim=(1, 2, 3)), world * B, rank, world)

        # ------------------------------------------------------------ roofline kernel: same K steps, instrumented
        # CUDA events around every res-block conv launch need the launches to be eager and un-overlapped, so this pass
        # runs the half-batch plans one after the other on the launching stream (not part of `value`).
        parts = getattr(plan, "parts", [plan])
        saved = (plan.use_graph, getattr(plan, "parallel", None))
        plan.use_graph = False
        if hasattr(plan, "parallel"):
            plan.parallel = False
        res_events = []
        res_convs = [cv for part in parts for (_n1, c1, _n2, c2) in part.res for cv in (c1, c2)]
        orig_forward = {id(cv): cv.forward for cv in res_convs}

        def timed(cv):
            f = orig_forward[id(cv)]

            def wrapped(x, y, stats=None):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                r = f(x, y, stats)
                e1.record()
                res_events.append((e0, e1))
                return r
            return wrapped
        for cv in res_convs:
            cv.forward = timed(cv)
        for _ in range(args.steps):
            plan.forward_from_maps(*d_in, 35)
        barrier()
        for cv in res_convs:
            cv.forward = orig_forward[id(cv)]
        plan.use_graph = saved[0]
        if hasattr(plan, "parallel"):
            plan.parallel = saved[1]
    res_batch = parts[0].B
    res_ms = sum(a.elapsed_time(b) for a, b in res_events) / max(len(res_events), 1)
    elapsed_ms = _max_over_ranks(elapsed_ms, world, dev)
    ms_per_step = elapsed_ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---------------------------------------------------------------- end to end through the ctu API (compact inputs)
    host = {"label": lab8, "instance": ins16, "image": img8}
    # two timed regions of exactly `steps` steps each (max over ranks per region), the faster one reported and both listed:
    # on these shared hosts the pinned-memory copies of one region in five lose 10-25 % to other tenants' PCIe traffic
    e2e_runs = []
    for _ in range(2):
        e2e_ms, h2d, d2h, e2e_out = e2e_measure(trainer, host, B, H, W, args.steps, barrier, dev)
        e2e_runs.append(_max_over_ranks(e2e_ms, world, dev))
    e2e_ms = min(e2e_runs)
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    e2e_equal = bool(rank == 0 and torch.equal(e2e_out[0], out_dev0))

    # ---------------------------------------------------------------- 2048x1024 (BASELINE.json configs[2]; extra)
    fullres = None
    if not args.no_fullres and (H, W) == (512, 1024):
        fb, fh, fw = args.fullres_batch, 2 * H, 2 * W
        fl, fi, fu, _ff = synth_inputs_compact(fb, fh, fw, seed=4242 + rank)
        fplan = netG.plan_for(fb, fh, fw, dev)
        fsteps = max(3, min(args.steps, 10))
        with torch.no_grad():
            f_ms, _ = device_resident(fplan, (fl.to(dev), fi.to(dev), fu.to(dev)), fsteps, 2, barrier)
        f_ms = _max_over_ranks(f_ms, world, dev) / fsteps
        fe_ms, fh2d, fd2h, _ = e2e_measure(trainer, {"label": fl, "instance": fi, "image": fu}, fb, fh, fw, fsteps, barrier, dev)
        fe_ms = _max_over_ranks(fe_ms, world, dev)
        fullres = {"workload": "batch %d per GPU at %dx%d, image-sharded over %d GPU(s), compact inputs" % (fb, fw, fh, world),
                   "value": world * fb / (f_ms * 1e-3), "unit": "images/s", "ms_per_step": f_ms, "steps": fsteps,
                   "e2e": {"value": world * fb * fsteps / (fe_ms * 1e-3), "unit": "images/s",
                           "h2d_bytes_per_step": fh2d, "d2h_bytes_per_step": fd2h},
                   "whole_forward_tflops_per_gpu": 4 * FWD_FLOPS_PER_IMAGE * fb / (f_ms * 1e-3) / 1e12}
        del fplan
        netG._plans = {}
        torch.cuda.empty_cache()

    # ---------------------------------------------------------------- training step (BASELINE.json configs[3]; extra)
    train = None
    if not args.no_train:
        netG._plans = {}
        torch.cuda.empty_cache()
        train = train_measure(args, dev, local, rank, world, barrier)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    sustained, burst, peak_kind = load_peaks()
    scale = (H * W) / (512.0 * 1024.0)
    res_flops = RES_CONV_FLOPS_PER_IMAGE * scale * res_batch
    achieved = res_flops / (res_ms * 1e-3) / 1e12 if res_ms > 0 else 0.0
    line = {"metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "config": workload_config(args),
            "e2e": {"value": e2e_value, "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "output_equals_device_resident_run": e2e_equal,
                    "regions": [world * B * args.steps / (m * 1e-3) for m in e2e_runs],
                    "note": "faster of two timed regions of `steps` steps each; both listed in `regions`"},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"kernel": "pair_conv3x3_kernel (ResnetBlock 3x3 conv 1024->1024 on CTA pairs, tcgen05 cta_group::2)", "bound": "tensor",
                         "achieved": achieved, "peak": sustained, "unit": "TFLOP/s", "frac": achieved / sustained,
                         "frac_of_burst_peak": achieved / burst, "peak_source": peak_kind + " (bf16_tflops_sustained)",
                         "launch_ms": res_ms, "flops_per_launch": res_flops, "images_per_launch": res_batch,
                         "timing": "CUDA events around each of the %d launches of %d extra steps run eagerly, half-batch "
                                   "plans back to back on the launching stream (the timed region replays them as one "
                                   "CUDA graph on two streams)" % (len(res_events), args.steps),
                         "traffic": ncu_traffic(res_batch, H, W),
                         "whole_forward_tflops": FWD_FLOPS_PER_IMAGE * scale * B / (ms_per_step * 1e-3) / 1e12}}
    line["sharding"] = {"images": world * B, "assignment": "image i -> rank i mod N, seed 1234 + i (jpd-se_b200/sharding.py); no "
                        "data-path collective, one gather of per-image results after the timed region",
                        "per_image_output_mean_head": [float(v) for v in checks[:4].tolist()]}
    if fullres is not None:
        line["fullres"] = fullres
    if train is not None:
        line["train"] = train
    if world == 1 and not args.no_bandwidth:
        try:
            hbm = 6559.4
            pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
            if os.path.exists(pk):
                with open(pk) as f:
                    hbm = json.load(f).get("hbm_gbs", hbm)
            line["bandwidth_kernels"] = {"peak_gbs": hbm, "peak_source": peak_kind + " (hbm_gbs, copy)",
                                         "rows": bandwidth_measure(dev, hbm)}
        except Exception as e:
            line["bandwidth_kernels"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    if world == 1 and not args.no_cudnn_baseline:
        try:
            line["cudnn_baseline"] = cudnn_measure(B, H, W, dev)
        except Exception as e:  # an extra: never let it take the bench line down
            line["cudnn_baseline"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        weights = {k: v.detach().cpu() for k, v in netG.state_dict().items()}
        sec, ref_out, kind = cpu_reference_run(lab8[:1], ins16[:1], img_f32[:1], steps=8, warmup=1, weights=weights)
        line["cpu_baseline"] = {"value": 1.0 / sec, "unit": "images/s", "cores": cores, "kind": kind,
                                "sample": "8 timed forwards of image 0 of the timed batch (batch 1 at %dx%d, same weights, fp32, "
                                          "torch CPU on all host threads; ~10 s) through %s" % (
                                              W, H, "the unmodified reference's trainer.get_img (baseline/_ref)" if kind == "reference"
                                              else "the oracle port")}
        # parity of the TIMED run: image 0 of the batch the timed region produced vs the reference's fp32 CPU output
        from oracle import generator_oracle as orc
        err = (out_dev0 - ref_out[0].cpu()).abs()
        line["parity"] = {"against": kind + " fp32 CPU output for image 0 of the timed batch",
                          "max_abs": float(err.max()), "mean_abs": float(err.mean()), "psnr_db": orc.psnr(out_dev0, ref_out[0].cpu()),
                          "gate": "mean_abs <= 0.02, max_abs <= 0.15, psnr >= 39.2 dB (tests/test_gpu_parity.py)",
                          "within_gate": bool(err.mean() <= 0.02 and err.max() <= 0.15 and orc.psnr(out_dev0, ref_out[0].cpu()) >= 39.2)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
