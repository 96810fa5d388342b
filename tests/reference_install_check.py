"""Run by tests/test_reference_pin.py where /root/reference exists: INTEGRATION.md route 1 end to end on the host side.
The UNMODIFIED reference parser, trainer and model are imported; after `jpdse_b200.install_into_reference()` the
reference's own `get_trainer(opt)(opt, 'test')` builds OUR GlobalGenerator through its `networks.define_G` call
(pix2pixHD_model.py:147-150) and loads a checkpoint written by the reference's generator through its own
`load_network` (base_model.py:62-97). Nothing is computed (no GPU here)."""
import os
import sys
import tempfile
import types

import torch

REF, ROOT = sys.argv[1], sys.argv[2]
sys.path.insert(0, REF)
sys.path.insert(0, ROOT)
for m in ("skimage", "skimage.io", "dominate", "dominate.tags"):  # imported, unused on this route
    sys.modules.setdefault(m, types.ModuleType(m))
tmp = tempfile.mkdtemp()
sys.argv = ["test.py", "--model", "pix2pixHD", "--dataset", "cityscapes", "--no_label_encoding", "--no_feat_encoding",
            "--no_generator_binarization", "--normalize_mean", ".5,.5,.5", "--normalize_std", "1.,1.,1.", "--gpu_ids", "-1",
            "--save_dir", tmp, "--checkpoints_dir", tmp, "--n_blocks_global", "1", "--n_downsample_global", "2",
            "--root_dir", tmp]
import ctu.parsers  # noqa: E402
opt = ctu.parsers.trainopt2testopt(ctu.parsers.CTUTrainParser().parse(), mode="test")
from ctu.models.pix2pixHD_networks import networks as ref_networks  # noqa: E402
torch.manual_seed(7)
ref_G = ref_networks.define_G(39, 3, 64, "global", 2, 1, 1, 3, "instance", gpu_ids=[])
torch.save(ref_G.state_dict(), os.path.join(tmp, "net_G.pth"))

import jpdse_b200  # noqa: E402
jpdse_b200.install_into_reference()
from ctu.trainers import get_trainer  # noqa: E402
trainer = get_trainer(opt)(opt, mode="test")
G = trainer.model.netG
assert type(G).__module__.startswith("jpd-se_b200."), type(G).__module__
sd, ref_sd = G.state_dict(), ref_G.state_dict()
assert list(sd.keys()) == list(ref_sd.keys())
assert all(sd[k].dtype == v.dtype and torch.equal(sd[k], v) for k, v in ref_sd.items())
print("install_into_reference: reference trainer built %s.%s and loaded the reference checkpoint" % (
    type(G).__module__, type(G).__name__))
