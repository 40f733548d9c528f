"""ctypes binding of libboa_b200.so (include/boa_b200.h).  Loads loudly: a missing library is an ImportError-grade
failure, never a fallback."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libboa_b200.so")

BOA_MAX_STAGES = 8
BOA_DT_I16, BOA_DT_F32, BOA_DT_F64 = 0, 1, 2


class BoaArch(C.Structure):
    _fields_ = [
        ("n_stages", C.c_int32), ("in_channels", C.c_int32), ("num_classes", C.c_int32),
        ("features", C.c_int32 * BOA_MAX_STAGES),
        ("n_conv_enc", C.c_int32 * BOA_MAX_STAGES),
        ("n_conv_dec", C.c_int32 * BOA_MAX_STAGES),
        ("strides", (C.c_int32 * 3) * BOA_MAX_STAGES),
        ("kernels", (C.c_int32 * 3) * BOA_MAX_STAGES),
        ("patch", C.c_int32 * 3),
        ("eps", C.c_float), ("leaky_slope", C.c_float),
    ]


class BoaError(RuntimeError):
    pass


_lib = None

_P = C.c_void_p
_SIGS = {
    "boa_last_error": (C.c_char_p, []),
    "boa_abi_version": (C.c_int, []),
    "boa_kernel_launch_count": (C.c_uint64, []),
    "boa_net_create": (C.c_int, [C.POINTER(BoaArch), C.c_int, C.c_int, C.POINTER(_P)]),
    "boa_net_set_tensor": (C.c_int, [_P, C.c_char_p, _P, C.POINTER(C.c_int64), C.c_int]),
    "boa_net_share_workspace": (C.c_int, [_P, _P]),
    "boa_net_finalize": (C.c_int, [_P]),
    "boa_net_set_mode": (C.c_int, [_P, C.c_int]),
    "boa_net_set_graph": (C.c_int, [_P, C.c_int]),
    "boa_net_forward_accumulate": (C.c_int, [_P, _P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, _P, _P, _P]),
    "boa_net_forward_logits": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "boa_net_macs_per_patch": (C.c_int64, [_P]),
    "boa_net_enable_timing": (C.c_int, [_P, C.c_int]),
    "boa_net_read_timing": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int]),
    "boa_net_read_timing_kinds": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double),
                                            C.POINTER(C.c_int64), C.c_int]),
    "boa_net_describe": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_double), C.c_char_p, C.c_int]),
    "boa_net_time_layers": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), _P]),
    "boa_net_destroy": (None, [_P]),
    "boa_ct_normalize": (C.c_int, [_P, C.c_int, C.c_size_t, C.c_float, C.c_float, C.c_float, C.c_float, _P, _P]),
    "boa_accumulate_patch": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P, _P,
                                       C.POINTER(C.c_int32), _P]),
    "boa_accumulate_weights": (C.c_int, [C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32), _P, _P,
                                         C.POINTER(C.c_int32), _P]),
    "boa_finalize_argmax": (C.c_int, [_P, _P, C.c_int, C.c_size_t, C.POINTER(C.c_uint8), C.c_int, _P, _P, _P]),
    "boa_normalize_logits": (C.c_int, [_P, _P, C.c_int, C.c_size_t, C.c_float, _P, _P]),
    "boa_tissue_subclassify": (C.c_int, [_P, C.c_int, _P, C.c_size_t, _P, _P]),
    "boa_slice_label_stats": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_int, C.c_size_t, C.c_int, _P, _P, _P]),
    "boa_label_hu_hist": (C.c_int, [_P, C.c_int, _P, C.c_size_t, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "boa_erode_box": (C.c_int, [_P, C.POINTER(C.c_int32), C.c_int, C.c_int, _P, _P, _P]),
    "boa_mask_label_minus_window": (C.c_int, [_P, C.c_int, _P, C.c_size_t, C.POINTER(C.c_uint8), C.c_int, C.c_int,
                                              C.c_int, _P, _P]),
    "boa_resample_z_cubic": (C.c_int, [_P, C.c_int, C.c_int, C.c_size_t, C.c_int, _P, _P, _P]),
    "boa_resample_z_nearest_u8": (C.c_int, [_P, C.c_int, C.c_size_t, C.c_int, _P, _P]),
    "boa_resample_axis_cubic": (C.c_int, [_P, C.c_int, C.c_size_t, C.c_int, C.c_size_t, C.c_int, _P, _P, C.c_int, _P]),
    "boa_resample_axis_cubic_grid": (C.c_int, [_P, C.c_int, C.c_size_t, C.c_int, C.c_size_t, C.c_int, _P, _P, C.c_int, _P]),
    "boa_clip_slices_f32": (C.c_int, [_P, C.c_size_t, _P, C.c_size_t, C.c_int, _P, _P]),
    "boa_resample_z_nearest_grid_f32": (C.c_int, [_P, C.c_int, C.c_size_t, C.c_int, _P, _P]),
    "boa_finalize_argmax_resampled": (C.c_int, [_P, _P, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int,
                                                C.POINTER(C.c_uint8), C.c_int, _P, _P, _P]),
    "boa_resample_nearest_u8": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _P, _P]),
    "boa_median3x3_slices": (C.c_int, [_P, C.POINTER(C.c_int32), _P, _P]),
    "boa_cc_filter": (C.c_int, [_P, C.POINTER(C.c_int32), C.POINTER(C.c_uint8), C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, _P, _P, _P, _P, _P, _P]),
    "boa_paint_label": (C.c_int, [_P, C.c_size_t, C.c_int, _P, _P]),
    "boa_add_slab": (C.c_int, [_P, _P, C.c_size_t, _P]),
    "boa_add_slab_strided": (C.c_int, [_P, C.c_size_t, _P, C.c_size_t, C.c_int, C.c_size_t, _P]),
    "boa_comm_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P), C.POINTER(C.c_ubyte)]),
    "boa_comm_open": (C.c_int, [C.POINTER(C.c_ubyte), C.POINTER(_P)]),
    "boa_comm_close": (C.c_int, [_P]),
    "boa_comm_free": (C.c_int, [_P]),
    "boa_comm_zero": (C.c_int, [_P, C.c_size_t, _P]),
    "boa_reduce_finalize_peers": (C.c_int, [C.POINTER(_P), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_int, _P, C.c_int, C.POINTER(C.c_uint8), C.c_int, _P,
                                            _P, _P]),
}
EXPORTS = sorted(_SIGS)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BoaError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(boa_b200 has no CPU fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise BoaError(f"libboa_b200 error {rc}: {lib().boa_last_error().decode()}")


def ptr(t) -> C.c_void_p:
    """Device (or host) pointer of a torch tensor / None."""
    return C.c_void_p(0 if t is None else t.data_ptr())


def i32x3(v) -> C.Array:
    return (C.c_int32 * 3)(int(v[0]), int(v[1]), int(v[2]))


def stream_ptr(stream=None) -> C.c_void_p:
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)
