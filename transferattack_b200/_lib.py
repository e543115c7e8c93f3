"""ctypes binding of libta_b200.so (the C-ABI declared in include/ta_b200.h).

There is no CPU fallback: if the shared library is missing or cannot be loaded, importing the kernels raises
with the build command. The library is built in-tree by ``transferattack_b200._build`` (nvcc, sm_100a).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libta_b200.so")

TA_OK, TA_EINVAL, TA_ECUDA, TA_EUNSUPPORTED = 0, -1, -2, -3
TA_MEAN_EXACT, TA_MEAN_TORCH = 0, 1
TA_DIR_SIGN, TA_DIR_RAW = 0, 1

_p = ctypes.c_void_p
_f = ctypes.c_float
_i = ctypes.c_int
_l = ctypes.c_int64



class FusedTailArgs(ctypes.Structure):
    """``ta_fused_tail_args`` of include/ta_b200.h, field for field."""
    _fields_ = [("g", _p), ("addend", _p), ("m", _p), ("m_out", _p), ("delta", _p), ("delta_out", _p), ("data", _p),
                ("xadv_out", _p), ("gbar_out", _p), ("scale", _p), ("scale_out", _p), ("mean_mode", _i),
                ("decay", _f), ("alpha", _f), ("eps", _f), ("lo", _f), ("hi", _f), ("B", _i), ("n", _l),
                ("mean_host", _p), ("std_host", _p), ("C", _i), ("plane", _l), ("emit_normalized", _i), ("grad_wrt_xn", _i)]


# name -> (restype, argtypes); mirrors include/ta_b200.h one to one
SIGNATURES = {
    "ta_version": (_i, []),
    "ta_last_error": (ctypes.c_char_p, []),
    "ta_device_info": (_i, [ctypes.POINTER(_i)] * 3),
    "ta_launch_count": (_l, []),
    "ta_tune_set": (_i, [ctypes.c_char_p, _i]),
    "ta_abs_mean_ws_bytes": (_l, [_i, _l]),
    "ta_abs_mean_per_sample": (_i, [_p, _p, _i, _l, _i, _p, _p]),
    "ta_aten_mean_policy": (_i, [_i, _l, _i, _i, ctypes.POINTER(_i), ctypes.POINTER(_i), ctypes.POINTER(_i)]),
    "ta_momentum": (_i, [_p, _p, _p, _f, _p, _i, _l, _p]),
    "ta_update_linf": (_i, [_p, _p, _p, _p, _f, _f, _f, _f, _i, _p, _l, _p]),
    "ta_update_l2_ws_bytes": (_l, [_i]),
    "ta_update_l2": (_i, [_p, _p, _p, _f, _f, _f, _f, _p, _i, _l, _p, _p]),
    "ta_clamp_box": (_i, [_p, _p, _f, _f, _p, _l, _p]),
    "ta_init_l2_scale": (_i, [_p, _p, _p, _f, _f, _f, _p, _i, _l, _p, _p]),
    "ta_fused_update_linf": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _f, _f, _f, _f, _f, _i, _l, _p]),
    "ta_fused_tail": (_i, [ctypes.POINTER(FusedTailArgs), _p]),
    "ta_fused_update_linf_nf": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _i, _f, _f, _f, _f, _f, _i, _l, _p, _p, _i, _l, _i, _p]),
    "ta_fused_allreduce_update_linf": (_i, [ctypes.POINTER(_p), ctypes.POINTER(_p), _i, _p, _p, _p, _p, _p, _p, _p, _i,
                                            _f, _f, _f, _f, _f, _i, _i, _l, _p]),
    "ta_stage_add": (_i, [_p, _p, _p, _f, _p, _l, _p]),
    "ta_normalize_fwd": (_i, [_p, _p, _p, _p, _i, _i, _l, _p]),
    "ta_normalize_bwd": (_i, [_p, _p, _p, _i, _i, _l, _p]),
    "ta_normalize_bwd_colsum": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _l, _p]),
    "ta_abs_mean_from_colsums": (_i, [_p, _p, _i, _l, _p]),
    "ta_sim_fwd": (_i, [_p, _p, _i, _l, _p]),
    "ta_sim_bwd": (_i, [_p, _p, _i, _l, _p]),
    "ta_admix_fwd": (_i, [_p, _p, _f, _p, _i, _i, _i, _l, _p]),
    "ta_admix_bwd": (_i, [_p, _p, _i, _i, _i, _l, _p]),
    "ta_dim_fwd": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ta_dim_bwd": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ta_dim_ws_bytes": (_l, []),
    "ta_dim_fwd_ws": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "ta_dim_bwd_ws": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p]),
    "ta_dim_pack_bytes": (_l, []),
    "ta_dim_pack_build": (_i, [_p, _i, _i, _i, _i, _i, _i]),
    "ta_dim_fwd_dyn": (_i, [_p, _p, _i, _i, _i, _p, _i, _p, _p]),
    "ta_dim_bwd_dyn": (_i, [_p, _p, _i, _i, _i, _p, _i, _p, _p]),
    "ta_counter_add": (_i, [_p, _i, _i, _p]),
    "ta_dwconv2d": (_i, [_p, _p, _i, _p, _i, _i, _i, _i, _p]),
    "ta_dwconv2d_sep": (_i, [_p, _p, _p, _i, _p, _i, _i, _i, _i, _p]),
    "ta_dwconv2d_sep_hw": (_i, [_p, _p, _p, _i, _p, _i, _i, _i, _i, _p]),
    "ta_pi_cut_noise": (_i, [_p, _p, _f, _f, _p, _p, _l, _p]),
    "ta_pi_update_linf": (_i, [_p, _p, _p, _p, _p, _f, _f, _f, _f, _f, _p, _p, _l, _p]),
    "ta_gra_update": (_i, [_p, _p, _p, _f, _f, _p, _p, _f, _f, _f, _p, _p, _l, _p]),
    "ta_adaea_drf": (_i, [ctypes.POINTER(_p), _i, _f, _p, _p, _p, _i, _i, _l, _p]),
    "ta_spectrum_ws_bytes": (_l, [_i, _i]),
    "ta_spectrum_transform": (_i, [_p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p]),
    "ta_lin_sample_fwd": (_i, [_p, _p, ctypes.POINTER(_f), _i, _p, _l, _p]),
    "ta_lin_sample_bwd": (_i, [_p, _p, _i, _l, _p]),
    "ta_neighbor_stage": (_i, [_p, _p, _p, _p, _f, _p, _l, _p]),
    "ta_uniform_fill_policy": (_i, [_l, ctypes.POINTER(_l), ctypes.POINTER(_l)]),
    "ta_neighbor_stage_philox": (_i, [_p, _p, _p, _f, _f, _f, ctypes.c_uint64, ctypes.c_uint64, _p, _p, _l, _p]),
    "ta_accumulate": (_i, [_p, _p, _i, _l, _p]),
    "ta_variance_finalize": (_i, [_p, _p, _i, _p, _l, _p]),
    "ta_add": (_i, [_p, _p, _p, _l, _p]),
    "ta_quantize_u8": (_i, [_p, _p, _p, _i, _i, _l, _i, _p]),
}

_lib = None


class KernelLibraryError(RuntimeError):
    pass


def load():
    """Load libta_b200.so once; raise loudly (no fallback) when it is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise KernelLibraryError(
            "transferattack_b200: %s is missing. Build it with `python -m transferattack_b200._build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the attack hooks." % SO_PATH)
    try:
        lib = ctypes.CDLL(SO_PATH)
    except OSError as e:  # pragma: no cover
        raise KernelLibraryError("transferattack_b200: cannot load %s: %s" % (SO_PATH, e))
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            raise KernelLibraryError("transferattack_b200: %s does not export %s (stale build?)" % (SO_PATH, name))
        fn.restype = res
        fn.argtypes = args
    if lib.ta_version() != 1:
        raise KernelLibraryError("transferattack_b200: ABI version %d, expected 1" % lib.ta_version())
    _lib = lib
    return lib


def last_error():
    return load().ta_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != TA_OK:
        raise RuntimeError("libta_b200 %s failed (%d): %s" % (what, rc, last_error()))


def launch_count():
    return int(load().ta_launch_count())


def tune_set(key, value):
    check(load().ta_tune_set(key.encode(), int(value)), "ta_tune_set")
