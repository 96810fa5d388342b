"""ctypes binding of libjpdse_b200.so (include/jpdse_b200.h).

There is no fallback: if the library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libjpdse_b200.so")

c_void_p, c_int, c_float, c_size_t, c_double = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_double)


class ConvDesc(ctypes.Structure):
    """struct jpdse_conv_desc"""
    _fields_ = [("kind", c_int), ("epilogue", c_int), ("batch", c_int), ("in_h", c_int), ("in_w", c_int),
                ("in_pad", c_int), ("cin", c_int), ("cin_real", c_int), ("cout", c_int),
                # ABI version 2
                ("out_pad", c_int), ("out_h", c_int), ("out_w", c_int), ("slope", c_float), ("cout_real", c_int)]


# enum jpdse_conv_kind / jpdse_conv_epilogue
(CONV3X3_PAD1, CONV3X3_S2, CONVT3X3_S2, CONV7X7_PAD3, CONV1X1, CONV3X3_FULL, CONV7X7_FULL, CONV4X4_S2, CONV4X4_S1,
 CONV4X4_S2_DGRAD, CONV4X4_S1_FULL, CONV3X3_PAD1_NARROW, CONV3X3_FULL_SHARED) = range(13)
PAD_SHARED = 0x100  # JPDSE_PAD_SHARED: shared-border layout flag of the gradient pad arguments
EPI_RAW_STATS, EPI_BIAS_TANH_NCHW, EPI_SIGN_NCHW, EPI_RAW, EPI_BIAS_ACT, EPI_BIAS_NCHW = range(6)
ABI_VERSION = 4

# symbol -> (restype, argtypes); also the list the CPU test checks against the header
SIGNATURES = {
    "jpdse_abi_version": (c_int, []),
    "jpdse_last_error": (ctypes.c_char_p, []),
    "jpdse_build_input": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                  c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "jpdse_build_input_u8": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, ctypes.POINTER(c_float),
                                     ctypes.POINTER(c_float), c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p,
                                     c_void_p, c_void_p]),
    "jpdse_conv_packed_weight_bytes": (c_size_t, [ctypes.POINTER(ConvDesc)]),
    "jpdse_conv_pack_weights": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p]),
    "jpdse_conv_forward": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p]),
    "jpdse_conv_flops": (c_double, [ctypes.POINTER(ConvDesc)]),
    "jpdse_conv_launch_count": (c_int, [ctypes.POINTER(ConvDesc)]),
    "jpdse_conv_wgrad_workspace_bytes": (c_size_t, [ctypes.POINTER(ConvDesc), c_int]),
    "jpdse_conv_wgrad": (c_int, [ctypes.POINTER(ConvDesc), c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                 c_size_t, c_void_p]),
    "jpdse_instnorm_backward_reduce": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_int, c_int, c_int, c_int, c_int, c_float, c_void_p]),
    "jpdse_instnorm_backward_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                              c_int, c_int, c_float, c_void_p]),
    "jpdse_tanh_backward_nchw": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                         c_void_p]),
    "jpdse_instnorm_apply": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_float, c_void_p]),
    "jpdse_nchw_f32_to_nhwc_bf16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "jpdse_nhwc_bf16_to_nchw_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "jpdse_round_f32": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "jpdse_sign_f32": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "jpdse_softsign_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "jpdse_sign_to_bits_u8": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "jpdse_s2hvq_encode": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_float, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p]),
    "jpdse_binarizer_train_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                              c_void_p]),
    "jpdse_binarizer_train_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "jpdse_tensor2im_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, ctypes.POINTER(c_double),
                                   ctypes.POINTER(c_double), c_void_p]),
    "jpdse_distortion_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                    ctypes.POINTER(c_double), ctypes.POINTER(c_double), c_void_p]),
    "jpdse_s2hvq_decode": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "jpdse_instnorm_backward_reduce_act": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                   c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "jpdse_instnorm_backward_fused": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                              c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "jpdse_d_input": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "jpdse_d_input_ids": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                  c_int, c_int, c_int, c_int, c_void_p]),
    "jpdse_d_input_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "jpdse_instnorm_apply_act": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float,
                                         c_void_p]),
    "jpdse_act_backward": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_float, c_void_p]),
    "jpdse_l1_pair": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "jpdse_l1_pair_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int, c_int,
                                       c_void_p]),
    "jpdse_maxpool2x2": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "jpdse_maxpool2x2_backward": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "jpdse_nhwc_pad_to_nchw_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "jpdse_patch_out_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "jpdse_patch_out_scatter": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
}

_lib = None


class JpdseError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise JpdseError(
                "libjpdse_b200.so is missing (%s); build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU or PyTorch fallback for this path." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        raise JpdseError("jpdse_b200 error %d: %s" % (rc, load().jpdse_last_error().decode()))
