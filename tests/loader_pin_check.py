"""Run by tests/test_reference_pin.py where /root/reference (with its bundled 30 Cityscapes validation triples) exists:
the compact host loader (jpd-se_b200/ctu/data) against the UNMODIFIED reference loader on the same files.
  float_tensors=True  == the reference's x_dict, bit for bit (label float ids, instance ids, normalised float image)
  compact tensors     -> the loader's ToTensor + Normalize formula gives the reference's float image bit for bit, i.e.
                         what jpdse_build_input_u8 computes on the device (tests/test_gpu_parity.py pins that kernel)
"""
import os
import sys
import tempfile
import types

import torch

REF, ROOT = sys.argv[1], sys.argv[2]
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
for m in ("skimage", "skimage.io", "dominate", "dominate.tags"):
    sys.modules.setdefault(m, types.ModuleType(m))
tmp = tempfile.mkdtemp()
data_root = os.path.join(REF, "datasets", "cityscapes_test_CVPR20_1024")
sys.argv = ["test.py", "--model", "pix2pixHD", "--dataset", "cityscapes", "--no_label_encoding", "--no_feat_encoding",
            "--no_generator_binarization", "--normalize_mean", ".5,.5,.5", "--normalize_std", "1.,1.,1.", "--gpu_ids", "-1",
            "--save_dir", tmp, "--checkpoints_dir", tmp, "--root_dir", data_root, "--use_gt_semantics",
            "--load_size", "1024", "--crop_size", "1024", "--test_load_size", "1024", "--test_crop_size", "1024",
            "--test_preprocess_mode", "fixed", "--test_aspect_ratio", "2.0"]
import ctu.parsers  # noqa: E402
import ctu.data  # noqa: E402
opt = ctu.parsers.trainopt2testopt(ctu.parsers.CTUTrainParser().parse(), mode="test")
opt.mode = "val" if os.path.isdir(os.path.join(data_root, "leftImg8bit", "val")) else opt.mode
opt.max_dataset_size = 4
ref_loader = ctu.data.create_dataloader(opt)
import importlib  # noqa: E402
ours = importlib.import_module("jpd-se_b200.ctu.data")
flt = ours.create_dataloader(opt, float_tensors=True)
cmp_ = ours.create_dataloader(opt)
n = 0
for r, f, c in zip(ref_loader, flt, cmp_):
    assert r["path"] == f["path"] == c["path"]
    assert r["image"].shape[-2:] == (512, 1024), r["image"].shape
    assert torch.equal(r["label"], f["label"]) and r["label"].dtype == f["label"].dtype
    assert torch.equal(r["instance"], f["instance"]) and r["instance"].dtype == f["instance"].dtype
    assert torch.equal(r["image"], f["image"]) and r["image"].dtype == torch.float32
    assert c["label"].dtype == torch.uint8 and c["image"].dtype == torch.uint8 and c["instance"].dtype in (torch.int16, torch.int32)
    assert torch.equal(c["label"].float(), r["label"]) and torch.equal(c["instance"].long(), r["instance"].long())
    assert torch.equal((c["image"].float().div(255.0) - 0.5) / 1.0, r["image"])
    n += 1
assert n == 4
small = sum(t.numel() * t.element_size() for t in (c["label"], c["instance"], c["image"]))
big = sum(t.numel() * t.element_size() for t in (r["label"], r["instance"], r["image"]))
print("LOADER_PIN_OK %d samples; bytes per sample %d (compact) vs %d (reference)" % (n, small, big))
