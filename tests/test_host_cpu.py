"""CPU: the C-ABI library loads and exports every symbol of include/jpdse_b200.h; the ctu-API mirror keeps the
reference's names / keys / error behaviour; nothing computes without a GPU."""
import argparse
import ctypes
import importlib
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "jpdse_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(jpdse_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    import jpdse_b200
    lib = jpdse_b200._lib.load()
    names = _header_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), "libjpdse_b200.so does not export %s" % n
    # every declared function has a ctypes signature and vice versa
    assert sorted(jpdse_b200._lib.SIGNATURES) == names
    assert lib.jpdse_abi_version() == jpdse_b200._lib.ABI_VERSION == 4
    assert lib.jpdse_last_error() is not None


def test_conv_desc_validation_without_gpu():
    import jpdse_b200
    from jpdse_b200._lib import CONV3X3_PAD1, CONV7X7_PAD3, EPI_RAW_STATS, ConvDesc
    lib = jpdse_b200._lib.load()
    assert ctypes.sizeof(ConvDesc) == 14 * 4  # ABI version 2: + out_pad, out_h, out_w, slope, cout_real
    good = ConvDesc(CONV3X3_PAD1, EPI_RAW_STATS, 16, 32, 64, 1, 1024, 1024, 1024)
    assert lib.jpdse_conv_packed_weight_bytes(ctypes.byref(good)) == 1024 * 9 * 1024 * 2
    assert lib.jpdse_conv_flops(ctypes.byref(good)) == 16 * 38654705664.0
    stem = ConvDesc(CONV7X7_PAD3, EPI_RAW_STATS, 1, 512, 1024, 3, 40, 39, 64)
    assert lib.jpdse_conv_flops(ctypes.byref(stem)) == 128245039104.0
    bad = ConvDesc(CONV3X3_PAD1, EPI_RAW_STATS, 1, 32, 64, 1, 100, 100, 64)  # cin not a multiple of 64
    assert lib.jpdse_conv_packed_weight_bytes(ctypes.byref(bad)) == 0
    assert b"multiple of 64" in lib.jpdse_last_error()
    bad2 = ConvDesc(CONV3X3_PAD1, EPI_RAW_STATS, 1, 32, 64, 0, 64, 64, 64)  # PAD1 kind needs the border
    assert lib.jpdse_conv_packed_weight_bytes(ctypes.byref(bad2)) == 0
    # PatchGAN convs (networks.py:430-449): 4x4, zero pad 2; odd sizes; data gradients need the forward input size
    from jpdse_b200._lib import CONV4X4_S1, CONV4X4_S1_FULL, CONV4X4_S2, CONV4X4_S2_DGRAD, EPI_BIAS_ACT, EPI_BIAS_NCHW, EPI_RAW
    d0 = ConvDesc(CONV4X4_S2, EPI_BIAS_ACT, 2, 512, 1024, 2, 64, 39, 64, 2, 0, 0, 0.2, 0)
    assert lib.jpdse_conv_packed_weight_bytes(ctypes.byref(d0)) == 64 * 16 * 64 * 2
    assert lib.jpdse_conv_flops(ctypes.byref(d0)) == 2.0 * 2 * 257 * 513 * 16 * 39 * 64
    d4 = ConvDesc(CONV4X4_S1, EPI_BIAS_NCHW, 2, 66, 130, 2, 512, 512, 1)
    assert lib.jpdse_conv_packed_weight_bytes(ctypes.byref(d4)) == 16 * 16 * 512 * 2  # N tile of 16 for the 1-channel output
    dg = ConvDesc(CONV4X4_S2_DGRAD, EPI_RAW, 2, 129, 257, 2, 128, 128, 64, 0, 257, 513, 0.0, 64)
    assert lib.jpdse_conv_launch_count(ctypes.byref(dg)) == 4
    dg_bad = ConvDesc(CONV4X4_S2_DGRAD, EPI_RAW, 2, 129, 257, 2, 128, 128, 64, 0, 255, 513, 0.0, 64)
    assert lib.jpdse_conv_packed_weight_bytes(ctypes.byref(dg_bad)) == 0 and b"forward input size" in lib.jpdse_last_error()
    df = ConvDesc(CONV4X4_S1_FULL, EPI_RAW, 2, 66, 130, 2, 512, 512, 256)
    assert lib.jpdse_conv_packed_weight_bytes(ctypes.byref(df)) == 256 * 16 * 512 * 2


def _networks():
    return importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")


def test_define_g_matches_reference_layout():
    nw = _networks()
    torch.manual_seed(0)
    net = nw.define_G(39, 3, 64, "global", 4, 9, 1, 3, "instance", gpu_ids=[])
    sd = net.state_dict()
    assert sum(p.numel() for p in net.parameters()) == 182556163
    assert len(sd) == 56
    want = ["model.%d" % i for i in (1, 4, 7, 10, 13)] + \
           ["model.%d.conv_block.%d" % (i, j) for i in range(16, 25) for j in (1, 5)] + \
           ["model.%d" % i for i in (25, 28, 31, 34, 38)]
    assert sorted(sd) == sorted(k + s for k in want for s in (".weight", ".bias"))
    assert tuple(sd["model.25.weight"].shape) == (1024, 512, 3, 3)  # ConvTranspose2d layout (Cin,Cout,3,3)
    assert tuple(sd["model.1.weight"].shape) == (64, 39, 7, 7)
    assert abs(float(sd["model.16.conv_block.1.weight"].std()) - 0.02) < 1e-3  # weights_init N(0, 0.02)


def test_reference_error_behaviour():
    nw = _networks()
    with pytest.raises(TypeError):  # the reference's bare raise('generator not implemented!')
        nw.define_G(39, 3, 64, "nope")
    with pytest.raises(NotImplementedError):
        nw.get_norm_layer("foo")
    net = nw.define_G(39, 3, 64, "global", 1, 0)
    with pytest.raises(ValueError):
        net(torch.zeros(1, 39, 8, 8), mode="bogus")
    with pytest.raises(AttributeError):
        net(torch.zeros(1, 39, 8, 8), mode="get_binary_code")


def test_no_cpu_fallback():
    import jpdse_b200
    nw = _networks()
    net = nw.define_G(39, 3, 64, "global", 1, 0).eval()
    with torch.no_grad(), pytest.raises(jpdse_b200.JpdseError):
        net(torch.zeros(1, 39, 128, 128))
    from jpdse_b200 import ops
    with pytest.raises(jpdse_b200.JpdseError):
        ops.round_f32(torch.zeros(4))


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "jpd-se_b200")
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|import_module\([\"']oracle", re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                text = open(os.path.join(dirpath, f)).read()
                assert not pat.search(text), "%s imports the oracle" % os.path.join(dirpath, f)


def test_trainer_lookup_and_modes():
    trainers = importlib.import_module("jpd-se_b200.ctu.trainers")
    opt = argparse.Namespace(model="pix2pixHD")
    cls = trainers.get_trainer(opt)
    assert cls.__name__ == "Pix2PixHDTrainer"
    with pytest.raises(ValueError):
        trainers.get_trainer(argparse.Namespace(model="pix2pixHD")).__init__(cls.__new__(cls), opt, "bogus")
    with pytest.raises(ModuleNotFoundError):
        trainers.get_trainer(argparse.Namespace(model="toderici2017"))


def test_s2hvq_mirror_argument_errors():
    s2h = importlib.import_module("jpd-se_b200.ctu.quantizers.s2h_vq")
    vq = s2h.S2HVQ(torch.zeros(4, 2), sigma=2.0)
    assert vq.center_size == 2 and vq.sigma == 2.0
    with pytest.raises(ValueError):
        vq.encode(torch.zeros(3, 4), code_len=8)
    with pytest.raises(ValueError):
        vq.encode(torch.zeros(3, 5), code_len=2)
    with pytest.raises(ValueError):
        vq.encode(torch.zeros(3, 8), code_len=2)  # center size 4 != code book's 2
    with pytest.raises(AssertionError):
        s2h.S2HVQ(torch.zeros(4, 2), sigma=0.0)


def test_model_lookup_and_option_setter_match_the_reference_table(golden_dir):
    """ctu/models/__init__.py:10-43 + pix2pixHD_model.py:22-101: `--model pix2pixHD` resolves to our class, and its
    static option setter declares exactly the reference's flags (dest, action, type, default, choices) -- the table in
    tests/golden/model_options.json was dumped from the reference by oracle/pin_against_reference.py."""
    import json
    models = importlib.import_module("jpd-se_b200.ctu.models")
    cls = models.find_model_using_name("pix2pixHD")
    assert cls.__name__ == "Pix2PixHDModel" and issubclass(cls, torch.nn.Module)
    assert models.get_option_setter("pix2pixHD") is cls.modify_commandline_options
    with pytest.raises(ModuleNotFoundError):
        models.find_model_using_name("toderici2017")
    golden = json.load(open(os.path.join(golden_dir, "model_options.json")))
    for mode, is_train in (("train", True), ("test", False)):
        ap = cls.modify_commandline_options(argparse.ArgumentParser(), is_train)
        ours = sorted([a.dest, type(a).__name__, None if a.type is None else a.type.__name__, a.default,
                       None if a.choices is None else list(a.choices)] for a in ap._actions if a.dest != "help")
        assert ours == golden[mode]
    # the shipped training script's model flags (scripts/pix2pixHD_bpg_train.sh) parse
    ap = cls.modify_commandline_options(argparse.ArgumentParser(), True)
    ns, rest = ap.parse_known_args("--no_label_encoding --no_feat_encoding --no_generator_binarization --use_compressed "
                                   "--quality 36 --ext bpg --checkpoints_dir ckpt --no_vgg_loss --dataset cityscapes".split())
    assert ns.no_label_encoding and ns.no_generator_binarization and ns.ext == "bpg" and ns.quality == "36"
    assert rest == ["--dataset", "cityscapes"] and ns.n_downsample_global == 4 and ns.n_blocks_global == 9


def test_bench_reference_arm_prints_one_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the contract's keys on
    rank 0, nothing on the other ranks."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--height", "64", "--width", "128"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="0", WORLD_SIZE="1"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "images/s" and d["higher_is_better"] is True and d["value"] > 0
    # the unmodified reference from baseline/_ref when it is installed (tools/install_reference.sh), else the oracle port
    want_kind = "reference" if os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "ctu", "__init__.py")) else "port"
    assert d["cpu_baseline"]["kind"] == want_kind and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["steps"] == 1 and d["warmup"] == 0 and "workload" in d["config"]
    r1 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r1.returncode == 0 and r1.stdout.strip() == ""


@pytest.mark.skipif(not os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "ctu", "__init__.py")),
                    reason="reference not installed in baseline/_ref (tools/install_reference.sh)")
def test_installed_reference_trainer_equals_oracle_bit_for_bit():
    """The reference arm of bench.py: the UNMODIFIED reference (baseline/_ref) driven through its own parser ->
    get_trainer -> Pix2PixHDTrainer.get_img must give the oracle's output bit for bit on CPU (same torch ops)."""
    import subprocess
    import sys
    code = (
        "import sys, torch; sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import bench, reference_arm as ra\n"
        "from oracle import generator_oracle as orc\n"
        "t, opt = ra.build_test_trainer(extra=['--n_blocks_global', '2', '--n_downsample_global', '3'])\n"
        "l8, i16, u8, img = bench.synth_inputs_compact(2, 64, 128)\n"
        "y = t.get_img({'label': l8.float(), 'instance': i16.int(), 'image': img.clone(), 'path': ['a', 'b']})\n"
        "x = torch.from_numpy(orc.build_input(l8.float().numpy(), i16.int().numpy(), img.numpy(), 35))\n"
        "with torch.no_grad():\n"
        "    ref = orc.generator_forward(t.model.netG.state_dict(), x, 3, 2)\n"
        "assert type(t.model.netG).__module__ == 'ctu.models.pix2pixHD_networks.networks'\n"
        "assert torch.equal(y, ref), float((y - ref).abs().max())\n"
        "print('REFERENCE_EQUALS_ORACLE')\n" % (ROOT, os.path.join(ROOT, "tools")))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    assert r.returncode == 0 and "REFERENCE_EQUALS_ORACLE" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


def test_vgg_refuses_silent_random_weights(monkeypatch):
    """without the pretrained checkpoint in the hub cache Vgg19 must raise, unless JPDSE_VGG_RANDOM=1."""
    nw = importlib.import_module("jpd-se_b200.ctu.models.pix2pixHD_networks.networks")
    cached = os.path.join(torch.hub.get_dir(), "checkpoints", "vgg19-dcbb9e9d.pth")
    if os.path.isfile(cached):
        pytest.skip("pretrained VGG19 present")
    monkeypatch.delenv("JPDSE_VGG_RANDOM", raising=False)
    monkeypatch.delenv("JPDSE_ALLOW_DOWNLOAD", raising=False)
    import jpdse_b200
    with pytest.raises(jpdse_b200.JpdseError):
        nw.Vgg19()
    monkeypatch.setenv("JPDSE_VGG_RANDOM", "1")
    assert nw.Vgg19().pretrained is False


def test_compact_host_loader_on_synthetic_files(tmp_path):
    """jpd-se_b200/ctu/data without the reference: Cityscapes-layout PNGs written here, compact
    tensors out; the float_tensors flavour equals torchvision's ToTensor + Normalize + the reference's id handling."""
    import numpy as np
    from PIL import Image
    import torchvision.transforms as T
    data = importlib.import_module("jpd-se_b200.ctu.data")
    rng = np.random.RandomState(0)
    for city, stem in (("aachen", "aachen_000000_000019"), ("bonn", "bonn_000001_000019")):
        os.makedirs(tmp_path / "leftImg8bit" / "val" / city, exist_ok=True)
        os.makedirs(tmp_path / "gtFine" / "val" / city, exist_ok=True)
        Image.fromarray(rng.randint(0, 256, (64, 128, 3), dtype=np.uint8)).save(tmp_path / "leftImg8bit" / "val" / city / (stem + "_leftImg8bit.png"))
        lab = rng.randint(0, 34, (32, 64), dtype=np.uint8)
        lab[0, 0] = 255  # 'unknown' -> num_labels
        Image.fromarray(lab, mode="L").save(tmp_path / "gtFine" / "val" / city / (stem + "_gtFine_labelIds.png"))
        Image.fromarray(rng.randint(0, 40000, (32, 64)).astype(np.uint16)).save(tmp_path / "gtFine" / "val" / city / (stem + "_gtFine_instanceIds.png"))
    opt = argparse.Namespace(root_dir=str(tmp_path), mode="val", use_gt_semantics=True, no_instance=False, max_dataset_size=10,
                             preprocess_mode="fixed", crop_size=64, aspect_ratio=2.0, is_train=False, no_flip=True, num_labels=35,
                             normalize_mean=[0.5, 0.5, 0.5], normalize_std=[1.0, 1.0, 1.0], batch_size=2, num_workers=0,
                             dataset="cityscapes")
    batch = next(iter(data.create_dataloader(opt)))
    assert batch["label"].dtype == torch.uint8 and tuple(batch["label"].shape) == (2, 1, 32, 64)
    assert batch["image"].dtype == torch.uint8 and tuple(batch["image"].shape) == (2, 3, 32, 64)
    assert batch["instance"].dtype == torch.int16 and int(batch["label"].max()) == 35
    ref = next(iter(data.create_dataloader(opt, float_tensors=True)))
    img0 = Image.open(batch["path"][0]).convert("RGB").resize((64, 32), Image.BICUBIC)
    want = T.Normalize([0.5] * 3, [1.0] * 3)(T.ToTensor()(img0))
    assert torch.equal(ref["image"][0], want)
    assert torch.equal(ref["label"], batch["label"].float()) and torch.equal(ref["instance"], batch["instance"])
    opt.preprocess_mode = "scale_width"
    with pytest.raises(NotImplementedError):
        data.create_dataloader(opt)
