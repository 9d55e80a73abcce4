"""ctypes binding of libhcflow_b200.so (the C ABI declared in include/hcflow_b200.h).

There is deliberately no fallback: if the shared library is missing or fails to load the
import raises, and every op raises on a non-zero return code.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhcflow_b200.so")
if os.environ.get("HCFLOW_LIB"):   # e.g. the wait-profiler build (HCF_BUILD_PROF=1), kept beside the product library
    LIB_PATH = os.environ["HCFLOW_LIB"]

c_float_p = C.c_void_p  # device pointers travel as integers


class Seg(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("ld", C.c_int32), ("C", C.c_int32), ("up_shift", C.c_int32),
                ("_pad", C.c_int32)]


class ConvStep(C.Structure):
    _fields_ = [("z", C.c_void_p), ("z_ld", C.c_int32), ("C", C.c_int32), ("n_pass", C.c_int32), ("z16_ld", C.c_int32),
                ("w", C.c_void_p), ("an_scale", C.c_void_p), ("an_bias", C.c_void_p), ("z16_hi", C.c_void_p),
                ("z16_lo", C.c_void_p)]


class ConvArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("nseg", C.c_int32),
        ("seg", Seg * 3),
        ("ks", C.c_int32), ("kpad", C.c_int32), ("cout", C.c_int32), ("npad", C.c_int32),
        ("w", C.c_void_p), ("bias", C.c_void_p), ("scale", C.c_void_p),
        ("act", C.c_int32), ("out_ld", C.c_int32),
        ("out", C.c_void_p), ("out2", C.c_void_p),
        ("out2_ld", C.c_int32), ("res1_ld", C.c_int32),
        ("res1", C.c_void_p), ("res2", C.c_void_p),
        ("res2_ld", C.c_int32), ("alpha1", C.c_float), ("alpha2", C.c_float), ("pre_ld", C.c_int32),
        ("pre", C.c_void_p), ("step", C.POINTER(ConvStep)),
        ("raw2", C.c_void_p), ("raw2_ld", C.c_int32), ("_pad2", C.c_int32),
    ]


class StepArgs(C.Structure):
    _fields_ = [
        ("npix", C.c_int32), ("pix_per_img", C.c_int32),
        ("z", C.c_void_p), ("z_ld", C.c_int32), ("C", C.c_int32),
        ("h", C.c_void_p), ("h_ld", C.c_int32), ("mode", C.c_int32), ("n_pass", C.c_int32), ("_pad", C.c_int32),
        ("w", C.c_void_p), ("an_scale", C.c_void_p), ("an_bias", C.c_void_p), ("logdet", C.c_void_p),
    ]


class PriorArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("Cz", C.c_int32),
        ("h", C.c_void_p), ("h_ld", C.c_int32), ("atan_logscale", C.c_int32),
        ("eps_nchw", C.c_void_p), ("z", C.c_void_p), ("z_ld", C.c_int32), ("_pad", C.c_int32),
        ("logdet", C.c_void_p), ("out_nchw", C.c_void_p),
    ]


class LayoutArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("src", C.c_void_p), ("dst", C.c_void_p), ("ld", C.c_int32), ("post", C.c_int32),
        ("noise", C.c_void_p), ("noise_scale", C.c_float), ("_pad", C.c_int32),
    ]


class SqueezeArgs(C.Structure):
    _fields_ = [
        ("B", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("src", C.c_void_p), ("src_ld", C.c_int32), ("dst_ld", C.c_int32), ("dst", C.c_void_p),
    ]


class Shadow16(C.Structure):
    _fields_ = [("f32", C.c_void_p), ("bytes", C.c_int64), ("hi", C.c_void_p), ("lo", C.c_void_p)]


class Seg16(C.Structure):
    _fields_ = [("hi", C.c_void_p), ("lo", C.c_void_p), ("ld", C.c_int32), ("_pad", C.c_int32)]


class FlowStep(C.Structure):
    _fields_ = [("w1", C.c_void_p), ("w2", C.c_void_p), ("w3", C.c_void_p),
                ("bias1", C.c_void_p), ("scale1", C.c_void_p), ("bias2", C.c_void_p), ("scale2", C.c_void_p),
                ("bias3", C.c_void_p), ("scale3", C.c_void_p),
                ("w", C.c_void_p), ("an_scale", C.c_void_p), ("an_bias", C.c_void_p),
                ("pre", C.c_void_p), ("pre_ld", C.c_int32), ("_pad", C.c_int32)]


class FlowStepChainArgs(C.Structure):
    _fields_ = [("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("n_pass", C.c_int32),
                ("n_steps", C.c_int32), ("split", C.c_int32), ("forward", C.c_int32),
                ("z", C.c_void_p), ("z_ld", C.c_int32), ("_pad", C.c_int32),
                ("z16_a", C.c_void_p), ("z16_b", C.c_void_p), ("done", C.c_void_p), ("logdet", C.c_void_p),
                ("steps", C.POINTER(FlowStep))]


OUT_F32, OUT_HI, OUT_LO = 1, 2, 4
STATUS_F16_OVERFLOW, STATUS_DEP_TIMEOUT = 1, 2
ABI_VERSION = 3   # == HCF_ABI_VERSION (include/hcflow_b200.h)

# every symbol include/hcflow_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "hcf_abi_version": (C.c_int, []),
    "hcf_last_error": (C.c_char_p, []),
    "hcf_launch_count": (C.c_uint64, []),
    "hcf_launch_count_reset": (None, []),
    "hcf_conv_fp32": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p]),
    "hcf_conv_tc_supported": (C.c_int, [C.POINTER(ConvArgs)]),
    "hcf_conv_tc_weight_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "hcf_conv_tc_pack_weights": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "hcf_conv_tc_plan_create": (C.c_int, [C.POINTER(ConvArgs), C.c_void_p, C.c_int32, C.POINTER(C.c_void_p)]),
    "hcf_conv_chain_create": (C.c_int, [C.POINTER(ConvArgs), C.POINTER(C.c_void_p), C.POINTER(C.c_int32), C.c_int32,
                                        C.c_void_p, C.POINTER(C.c_void_p)]),
    "hcf_conv_tc16_supported": (C.c_int, [C.POINTER(ConvArgs)]),
    "hcf_conv_tc16_weight_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "hcf_conv_tc16_pack_weights": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "hcf_conv_chain16_create": (C.c_int, [C.POINTER(ConvArgs), C.POINTER(C.c_void_p), C.POINTER(C.c_int32),
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32, C.c_void_p, C.POINTER(Shadow16), C.c_int32,
                                          C.POINTER(Seg16), C.POINTER(C.c_void_p)]),
    "hcf_split16": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_int32,
                              C.c_void_p]),
    "hcf_conv_tc_plan_layers": (C.c_int32, [C.c_void_p]),
    "hcf_conv_tc_run": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hcf_conv_tc_plan_refresh": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hcf_conv_tc_plan_set_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hcf_conv_tc_plan_destroy": (None, [C.c_void_p]),
    "hcf_flowstep_w1_bytes": (C.c_int64, []),
    "hcf_flowstep_pack_w1": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p]),
    "hcf_flowstep_stage_z1": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "hcf_flowstep_chain_create": (C.c_int, [C.POINTER(FlowStepChainArgs), C.POINTER(C.c_void_p)]),
    "hcf_flowstep_chain_refresh": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hcf_flowstep_chain_set_status": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hcf_flowstep_chain_run": (C.c_int, [C.c_void_p, C.c_void_p]),
    "hcf_flowstep_chain_destroy": (None, [C.c_void_p]),
    "hcf_conv_wgrad": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                 C.c_int32, C.c_void_p, C.c_void_p]),
    "hcf_channel_sum": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "hcf_affine_act_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "hcf_affine_act_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_int64, C.c_int32, C.c_void_p]),
    "hcf_coupling_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                   C.c_void_p]),
    "hcf_coupling_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_int32, C.c_void_p]),
    "hcf_gauss_logp_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    "hcf_gauss_logp_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int64, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]),
    "hcf_gauss_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "hcf_axpby": (C.c_int, [C.c_void_p, C.c_float, C.c_void_p, C.c_float, C.c_void_p, C.c_int64, C.c_void_p]),
    "hcf_quantize8": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    "hcf_downsample_sum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "hcf_step_inverse": (C.c_int, [C.POINTER(StepArgs), C.c_void_p]),
    "hcf_step_forward_head": (C.c_int, [C.POINTER(StepArgs), C.c_void_p]),
    "hcf_step_forward_coupling": (C.c_int, [C.POINTER(StepArgs), C.c_void_p]),
    "hcf_prior_sample": (C.c_int, [C.POINTER(PriorArgs), C.c_void_p]),
    "hcf_prior_logp": (C.c_int, [C.POINTER(PriorArgs), C.c_void_p]),
    "hcf_prior_standardize": (C.c_int, [C.POINTER(PriorArgs), C.c_void_p]),
    "hcf_nchw_to_nhwc": (C.c_int, [C.POINTER(LayoutArgs), C.c_void_p]),
    "hcf_nhwc_to_nchw": (C.c_int, [C.POINTER(LayoutArgs), C.c_void_p]),
    "hcf_squeeze2d": (C.c_int, [C.POINTER(SqueezeArgs), C.c_void_p]),
    "hcf_unsqueeze2d": (C.c_int, [C.POINTER(SqueezeArgs), C.c_void_p]),
    "hcf_haar_forward": (C.c_int, [C.POINTER(SqueezeArgs), C.c_void_p]),
    "hcf_haar_inverse": (C.c_int, [C.POINTER(SqueezeArgs), C.c_void_p]),
    "hcf_copy_view": (C.c_int, [C.POINTER(SqueezeArgs), C.c_void_p]),
    "hcf_upsample_nearest": (C.c_int, [C.POINTER(SqueezeArgs), C.c_int32, C.c_void_p]),
    "hcf_u8_hwc_to_nhwc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_void_p]),
    "hcf_nhwc_to_u8_hwc": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]),
    "hcf_tile_accumulate": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "hcf_tile_normalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    "hcf_image_metrics": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "hcf_gauss_logp_const": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p]),
}

_lib = None


class HcfError(RuntimeError):
    pass


def load():
    """Load the shared library (once) and attach prototypes. Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HcfError("{} not found: run `python -m hcflow_b200.build` (there is no CPU fallback)".format(LIB_PATH))
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.hcf_abi_version() != ABI_VERSION:
        raise HcfError("ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().hcf_last_error().decode(errors="replace")
        raise HcfError("{} failed (rc={}): {}".format(what, rc, msg))
