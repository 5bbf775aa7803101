"""ctypes binding of libegn.so (include/egn.h).  There is no CPU fallback: a missing or
unloadable library, or a non-sm_100 device, raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libegn.so")


class EgnInfo(ctypes.Structure):
    _fields_ = [("device", ctypes.c_int), ("num_sms", ctypes.c_int), ("micro_batch", ctypes.c_int),
                ("products_per_mac", ctypes.c_int), ("tensor_core_path", ctypes.c_int),
                ("workspace_bytes", ctypes.c_longlong), ("activation_bytes_unshared", ctypes.c_longlong),
                ("shared_pool_bytes", ctypes.c_longlong), ("lowered_layers", ctypes.c_int)]


class EgnConfig(ctypes.Structure):
    _fields_ = [("add_edge", ctypes.c_int), ("add_seg", ctypes.c_int), ("seg_detach", ctypes.c_int),
                ("input_concat", ctypes.c_int), ("only_edge", ctypes.c_int), ("style_dim", ctypes.c_int)]


EXPORTS = ["egn_last_error", "egn_version", "egn_create", "egn_destroy", "egn_set_weights", "egn_plan",
           "egn_bdcn_forward", "egn_bdcn_forward_all", "egn_info", "egn_esf_forward", "egn_seg_post", "egn_metrics_accumulate",
           "egn_ellipse_refine", "egn_preprocess_u8", "egn_forward_loss", "egn_launch_count", "egn_flops_per_frame", "egn_debug_read",
           "egn_conv_selfcheck", "egn_profile", "egn_profile_read", "egn_profile_table", "egn_share_workspace"]

_lib = None


class EgnError(RuntimeError):
    pass


def load():
    """Loads libegn.so once and declares the prototypes of include/egn.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise EgnError("libegn.so is not built (%s); run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `make -C <package>/csrc`.  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, ci, cll = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong
    lib.egn_last_error.restype = ctypes.c_char_p
    lib.egn_version.restype = ci
    lib.egn_create.argtypes = [ci, ctypes.POINTER(EgnConfig), ctypes.POINTER(vp)]
    lib.egn_destroy.argtypes = [vp]
    lib.egn_set_weights.argtypes = [vp, ci, vp, ctypes.c_size_t]
    lib.egn_plan.argtypes = [vp, ci]
    lib.egn_share_workspace.argtypes = [ci]
    lib.egn_bdcn_forward.argtypes = [vp, vp, ci, vp, ci, vp]
    lib.egn_bdcn_forward_all.argtypes = [vp, vp, ci, vp, vp, ci, vp]
    lib.egn_info.argtypes = [vp, ctypes.POINTER(EgnInfo)]
    lib.egn_esf_forward.argtypes = [vp, vp, vp, vp, vp, vp, ci, vp]
    lib.egn_seg_post.argtypes = [vp, vp, vp, vp, vp, vp, ci, vp]
    lib.egn_metrics_accumulate.argtypes = [vp, vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, ci, vp]
    lib.egn_ellipse_refine.argtypes = [vp, vp, vp, vp, ci, ci, vp]
    lib.egn_forward_loss.argtypes = [vp, vp, vp, ci, vp, vp, vp, vp, vp, vp, vp, ctypes.c_float, vp, ci, vp]
    lib.egn_preprocess_u8.argtypes = [vp, vp, vp, ci, vp]
    lib.egn_launch_count.argtypes = [vp]
    lib.egn_launch_count.restype = cll
    lib.egn_flops_per_frame.argtypes = [vp, ci]
    lib.egn_flops_per_frame.restype = ctypes.c_double
    lib.egn_debug_read.argtypes = [vp, ctypes.c_char_p, vp, cll, ci, ctypes.POINTER(ci)]
    lib.egn_debug_read.restype = cll
    lib.egn_conv_selfcheck.argtypes = [vp, ctypes.c_char_p, ci, ctypes.POINTER(ctypes.c_double),
                                       ctypes.POINTER(ctypes.c_double)]
    lib.egn_profile.argtypes = [vp, ci]
    lib.egn_profile_read.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                                     ctypes.POINTER(cll), ci]
    lib.egn_profile_table.argtypes = [vp, ctypes.c_char_p, cll]
    lib.egn_profile_table.restype = cll
    for name in EXPORTS:
        if name not in ("egn_last_error", "egn_launch_count", "egn_flops_per_frame", "egn_debug_read",
                        "egn_profile_table"):
            getattr(lib, name).restype = ci
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise EgnError(load().egn_last_error().decode("utf-8", "replace"))
