"""Builds the UNMODIFIED reference (`ctu`, installed from /root/reference into baseline/_ref with
`pip install --no-index --no-deps --target baseline/_ref`) through its own public API:
parser -> `trainopt2testopt` -> `get_trainer(opt)(opt, 'test')` -> `trainer.get_img(x_dict)`
(train.py / test.py call sequence: ctu/parsers/base_parser.py:208-249, ctu/parsers/__init__.py:4-34,
ctu/trainers/__init__.py:5-20, ctu/trainers/pix2pixHD_trainer.py:113-116).

Used by `bench.py --impl reference` (the reference arm: stock reference code on the box's host cores), by the
`cpu_baseline` leg and by the GPU test that compares the reference on CUDA fp32 with this repo's drop-in. Nothing of
this repo's kernels, models or engine is on that path. baseline/_ref is git-ignored (it is not product source) but
travels to the GPU box with the gpurun snapshot.
"""
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")

ARGV = ["--model", "pix2pixHD", "--dataset", "cityscapes", "--no_label_encoding", "--no_feat_encoding",
        "--no_generator_binarization", "--normalize_mean", ".5,.5,.5", "--normalize_std", "1.,1.,1."]


def available():
    return os.path.isfile(os.path.join(REF_DIR, "ctu", "trainers", "pix2pixHD_trainer.py"))


def import_reference():
    """Puts baseline/_ref first on sys.path and stubs the two third-party imports the reference makes but never uses
    on this route (skimage in ctu/data/ctu_dataset.py:14, dominate in ctu/utils/html.py:9; both absent offline)."""
    if not available():
        raise RuntimeError("baseline/_ref/ctu is missing: install the reference with `python -m pip install --no-index "
                           "--no-build-isolation --no-deps --target baseline/_ref <copy of /root/reference>`")
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    for m in ("skimage", "skimage.io", "dominate", "dominate.tags"):
        sys.modules.setdefault(m, types.ModuleType(m))
    import ctu  # noqa: F401
    if not os.path.abspath(ctu.__file__).startswith(REF_DIR):
        raise RuntimeError("`import ctu` resolved to %s, not to baseline/_ref" % ctu.__file__)
    return ctu


def parse_options(gpu_id=-1, extra=(), workdir=None):
    """opt Namespace from the reference's own parser for the shipped test configuration (scripts/pix2pixHD_bpg_test.sh
    minus --use_compressed: libbpg is outside the path, the decoded image is supplied as x_dict['image'])."""
    import_reference()
    import ctu.parsers
    workdir = workdir or tempfile.mkdtemp(prefix="jpdse_ref_")
    saved = sys.argv
    sys.argv = ["test.py"] + ARGV + ["--gpu_ids", str(gpu_id), "--save_dir", workdir, "--checkpoints_dir", workdir,
                                      "--root_dir", workdir] + list(extra)
    try:
        opt = ctu.parsers.trainopt2testopt(ctu.parsers.CTUTrainParser().parse(), mode="test")
    finally:
        sys.argv = saved
    return opt, workdir


def build_test_trainer(state_dict=None, gpu_id=-1, seed=1234, extra=()):
    """The reference's Pix2PixHDTrainer in test mode. Its constructor demands checkpoints_dir/net_G.pth
    (ctu/models/pix2pixHD_networks/base_model.py:65-68): `state_dict` (reference keys) or, if None, a random init of the
    reference's own define_G under torch.manual_seed(seed) on CPU is written there first."""
    import torch
    opt, workdir = parse_options(gpu_id, extra)
    from ctu.models.pix2pixHD_networks import networks
    from ctu.trainers import get_trainer
    if state_dict is None:
        torch.manual_seed(seed)
        g = networks.define_G(opt.num_labels + 4, 3, opt.ngf, opt.netG, opt.n_downsample_global, opt.n_blocks_global,
                              opt.n_local_enhancers, opt.n_blocks_local, opt.norm, gpu_ids=[])
        state_dict = g.state_dict()
        del g
    torch.save({k: v.detach().cpu() for k, v in state_dict.items()}, os.path.join(workdir, "net_G.pth"))
    trainer = get_trainer(opt)(opt, mode="test")
    return trainer, opt
