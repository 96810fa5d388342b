"""Host-side mirror of the reference's ``ctu`` interface for the accelerated path.

Module paths, class / function names, argument meaning and error behaviour follow the reference so the
parity tests read like reference code:
  ctu.models.pix2pixHD_networks.networks   define_G, GlobalGenerator, ResnetBlock, weights_init
  ctu.models.pix2pixHD_model               Pix2PixHDModel (preprocess / get_edges / _get_img / get_img)
  ctu.quantizers.{binarize,round,s2h_vq}   Binarizer, DifferentiableSign, RoundedIdentity, S2HVQ
  ctu.trainers.pix2pixHD_trainer           Pix2PixHDTrainer (get_img / get_code plumbing)
"""


def install():
    """Monkey-patch an importable reference ``ctu`` so its hot path runs on jpdse_b200.

    After ``install()`` the reference's own ``train.py`` / ``test.py`` build their generator through
    our ``define_G`` (same signature, same state-dict keys). Returns the patched reference module.
    """
    import importlib
    ref_networks = importlib.import_module("ctu.models.pix2pixHD_networks.networks")
    from .models.pix2pixHD_networks import networks as ours
    ref_networks.define_G = ours.define_G
    ref_networks.GlobalGenerator = ours.GlobalGenerator
    ref_networks.ResnetBlock = ours.ResnetBlock
    return ref_networks
