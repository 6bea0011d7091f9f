"""ctypes binding of include/dmhomo.h (libdmhomo.so, sm_100a).

The library is loaded lazily on first use - never at import time - so importing the
package is safe in forked DataLoader workers and on CPU-only hosts.  There is NO
fallback: if the shared object is missing the first op raises.
"""
import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# DMH_LIB: development knob, an alternative build of the same library (tools/build_variants.sh A/B runs)
LIB_PATH = os.environ.get("DMH_LIB") or os.path.join(_HERE, "libdmhomo.so")
ABI_VERSION = 2

# enums of include/dmhomo.h
S1, S1B, S2_ZEROS, S3_BORDER = 0, 1, 2, 3
PARAM_FLOW, PARAM_COORDS, PARAM_HOMOGRAPHY, PARAM_BASIS8 = 0, 1, 2, 3
LOSS_NONE, LOSS_MASKED_DIFF, LOSS_DIFF_MASKED = 0, 1, 2

_fp = C.c_void_p  # every device pointer crosses the ABI as an opaque address


class WarpDesc(C.Structure):
    """dmh_warp_desc (include/dmhomo.h)."""

    _fields_ = [
        ("struct_size", C.c_uint32),
        ("sampler", C.c_int32),
        ("param_kind", C.c_int32),
        ("loss_form", C.c_int32),
        ("B", C.c_int32),
        ("C", C.c_int32),
        ("Hs", C.c_int32),
        ("Ws", C.c_int32),
        ("h", C.c_int32),
        ("w", C.c_int32),
        ("divide", C.c_int32),
        ("use_border_mask", C.c_int32),
        ("compute_grads", C.c_int32),
        ("reserved0", C.c_int32),
        ("start_x", C.c_float),
        ("start_y", C.c_float),
        ("grad_loss_scale", C.c_float),
        ("reserved1", C.c_float),
        ("src", _fp),
        ("param", _fp),
        ("basis", _fp),
        ("start", _fp),
        ("target", _fp),
        ("soft_mask", _fp),
        ("sample_weight", _fp),
        ("grad_out", _fp),
        ("grad_loss", _fp),
        ("out", _fp),
        ("valid", _fp),
        ("flow_out", _fp),
        ("indices", _fp),
        ("loss_acc", _fp),
        ("grad_src", _fp),
        ("grad_target", _fp),
        ("grad_param", _fp),
        ("grad_soft_mask", _fp),
    ]


_i, _f, _d, _i64 = C.c_int, C.c_float, C.c_double, C.c_int64

# name -> argtypes (return type is int unless listed in _RESTYPES)
SIGNATURES = {
    "dmh_version": [],
    "dmh_last_error_string": [],
    "dmh_launch_count": [],
    "dmh_last_kernel_name": [],
    "dmh_set_tuning": [C.c_char_p, _i],
    "dmh_get_tuning": [C.c_char_p, C.POINTER(_i)],
    "dmh_warp_forward": [C.POINTER(WarpDesc), _i, _fp],
    "dmh_warp_backward": [C.POINTER(WarpDesc), _i, _fp],
    "dmh_loss_finish": [C.POINTER(_fp), C.POINTER(_fp), _i, _i, _f, _fp, _fp],
    "dmh_scale_inplace": [_fp, _i64, _fp, _fp],
    "dmh_dlt4_forward": [_fp, _fp, _fp, _i, _fp],
    "dmh_dlt4_backward": [_fp, _fp, _fp, _fp, _fp, _fp, _i, _fp],
    "dmh_homography_to_flow": [_fp, _fp, _i, _i, _i, _i, _f, _f, _fp, _fp],
    "dmh_homography_to_flow_backward": [_fp, _fp, _fp, _i, _i, _i, _i, _f, _f, _fp, _fp],
    "dmh_homography_to_flow_f64": [_fp, _fp, _i, _i, _i, _d, _i, _i, _fp],
    "dmh_basis_combine": [_fp, _fp, _fp, _i, _i, _i, _fp],
    "dmh_basis_combine_backward": [_fp, _fp, _fp, _i, _i, _i, _fp],
    "dmh_basis_corner_offsets": [_fp, _fp, _fp, _i, _i, _i, _fp],
    "dmh_basis_corner_offsets_backward": [_fp, _fp, _fp, _i, _i, _i, _fp],
    "dmh_basis_homography_forward": [_fp, C.POINTER(_fp), C.POINTER(_fp), _i, _i, _i, _i, _fp],
    "dmh_basis_homography_backward": [_fp, C.POINTER(_fp), C.POINTER(_fp), C.POINTER(_fp), _i, _i, _i, _i, _fp],
    "dmh_border_mask": [_fp, _fp, _fp, _i, _i, _i, _fp],
    "dmh_zero_border_mask": [_fp, _fp, _i, _i, _i, _f, _fp],
    "dmh_l1_sum": [_fp, _fp, _i64, _fp, _fp],
    "dmh_l1_backward": [_fp, _fp, _i64, _fp, _f, _fp, _fp, _fp],
    "dmh_flow_to_rgb": [_fp, _fp, _i, _i, _i, _f, _i, _i, _fp],
    "dmh_warp_perspective": [_fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _fp],
    "dmh_warp_perspective_u8": [_fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _fp],
    "dmh_remap": [_fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _fp],
    "dmh_remap_u8": [_fp, _fp, _fp, _i, _i, _i, _i, _i, _i, _i, _i, _fp],
    "dmh_eval_point_error": [_fp, _fp, _fp, _fp, _i, _i, _i, _i, _fp],
    "dmh_flow_to_homography_ls": [_fp, _fp, _fp, _i, _i, _i, _fp],
    "dmh_pairs_u8_to_gray": [_fp, _fp, _fp, _fp, _fp, C.POINTER(_d), C.POINTER(_d), _i, _i, _i, _i, _i, _i, _fp],
    "dmh_u8_to_f32": [_fp, _fp, _i64, _f, _f, _fp],
    "dmh_grid_normalize": [_fp, _fp, _i, _i, _i, _i, _fp],
    "dmh_flow_upsample": [_fp, _fp, _i, _i, _i, _i, _i, _i, _i, _fp],
    "dmh_flow_upsample_backward": [_fp, _fp, _i, _i, _i, _i, _i, _i, _i, _fp],
}
_RESTYPES = {"dmh_last_error_string": C.c_char_p, "dmh_last_kernel_name": C.c_char_p, "dmh_launch_count": C.c_uint64}

_lib = None
_lock = threading.Lock()


class DmhError(RuntimeError):
    pass


def lib():
    """The loaded library (loads on first call; raises if the extension is not built)."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise DmhError(
                f"dmhomo_b200: CUDA extension not built ({LIB_PATH} missing). Build it with "
                "`make -C dmhomo_b200/csrc` or `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU / PyTorch fallback.")
        l = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(l, name)  # AttributeError here = header / library mismatch
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, C.c_int)
        if l.dmh_version() != ABI_VERSION:
            raise DmhError(f"dmhomo_b200: ABI version {l.dmh_version()} != expected {ABI_VERSION}")
        _lib = l
        # DMH_TUNING="key=value,...": development knob of this binding (the library itself reads no environment)
        for kv in filter(None, os.environ.get("DMH_TUNING", "").split(",")):
            k, _, v = kv.partition("=")
            if l.dmh_set_tuning(k.strip().encode(), int(v)) != 0:
                raise DmhError(f"dmhomo_b200: bad DMH_TUNING entry {kv!r}")
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = lib().dmh_last_error_string()
        raise DmhError(f"libdmhomo {what} failed (status {rc}): {msg.decode() if msg else ''}")


def launch_count():
    return int(lib().dmh_launch_count())


def last_kernel_name():
    n = lib().dmh_last_kernel_name()
    return n.decode() if n else ""


def set_tuning(**knobs):
    """dmh_set_tuning(): development knobs of the library (include/dmhomo.h), e.g. set_tuning(tile=0)."""
    for k, v in knobs.items():
        check(lib().dmh_set_tuning(k.encode(), int(v)), f"set_tuning({k})")


def get_tuning(key):
    v = C.c_int(0)
    check(lib().dmh_get_tuning(key.encode(), C.byref(v)), f"get_tuning({key})")
    return v.value
