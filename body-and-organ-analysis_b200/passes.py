"""Host wrappers of the HBM-bound passes of libboa_b200 (CT normalisation, tissue rules, per-slice / per-label
statistics, label-set masks, erosion).  Tensors are containers; every byte is touched by our kernels only."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def _ct_dtype(ct: torch.Tensor) -> int:
    if ct.dtype == torch.int16:
        return _lib.BOA_DT_I16
    if ct.dtype == torch.float32:
        return _lib.BOA_DT_F32
    raise TypeError(f"CT volume must be int16 or float32 on the device, got {ct.dtype}")


def _chk(*ts):
    for t in ts:
        if t is not None and not (t.is_cuda and t.is_contiguous()):
            raise ValueError("boa_b200 passes need contiguous CUDA tensors (there is no CPU implementation)")


def ct_normalize(ct: torch.Tensor, lo: float, hi: float, mean: float, std: float) -> torch.Tensor:
    """CTNormalization.run (_external/nnunetv2/preprocessing/normalization/default_normalization_schemes.py:56-67)."""
    _chk(ct)
    out = torch.empty(ct.shape, dtype=torch.float32, device=ct.device)
    with torch.cuda.device(ct.device):
        _lib.check(_lib.lib().boa_ct_normalize(_lib.ptr(ct), _ct_dtype(ct), ct.numel(), lo, hi, mean, std,
                                               _lib.ptr(out), _lib.stream_ptr()))
    return out


def tissue_subclassify(ct: torch.Tensor, regions: torch.Tensor) -> torch.Tensor:
    """subclassify_tissues numerics (_external/body_composition_analysis/tissue/subclassification.py:38-53)."""
    _chk(ct, regions)
    out = torch.empty(regions.shape, dtype=torch.uint8, device=ct.device)
    with torch.cuda.device(ct.device):
        _lib.check(_lib.lib().boa_tissue_subclassify(_lib.ptr(ct), _ct_dtype(ct), _lib.ptr(regions), ct.numel(),
                                                     _lib.ptr(out), _lib.stream_ptr()))
    return out


def slice_label_stats(labels: torch.Tensor, n_labels: int, ct: torch.Tensor | None = None,
                      mask: torch.Tensor | None = None, mask_value: int = 1, want_counts: bool = True):
    """Per-slice (axis 0) per-label voxel counts [Z, L] (uint64 as int64) and HU sums [Z, L] (int64)."""
    _chk(labels, ct, mask)
    Z = labels.shape[0]
    sv = labels.numel() // max(Z, 1)
    counts = torch.zeros((Z, n_labels), dtype=torch.int64, device=labels.device) if want_counts else None
    sums = torch.zeros((Z, n_labels), dtype=torch.int64, device=labels.device) if ct is not None else None
    with torch.cuda.device(labels.device):
        _lib.check(_lib.lib().boa_slice_label_stats(
            _lib.ptr(labels), _lib.ptr(mask), mask_value, _lib.ptr(ct),
            _ct_dtype(ct) if ct is not None else _lib.BOA_DT_I16, Z, sv, n_labels, _lib.ptr(counts), _lib.ptr(sums),
            _lib.stream_ptr()))
    return counts, sums


def label_hu_hist(ct: torch.Tensor, labels: torch.Tensor, n_labels: int, hu_min: int = -1024, n_bins: int = 4096 + 1024):
    """Per-label integer-HU histograms [L, n_bins] (uint32 as int32 bits) + count of out-of-range voxels."""
    _chk(ct, labels)
    hist = torch.zeros((n_labels, n_bins), dtype=torch.int32, device=ct.device)
    oor = torch.zeros(1, dtype=torch.int32, device=ct.device)
    with torch.cuda.device(ct.device):
        _lib.check(_lib.lib().boa_label_hu_hist(_lib.ptr(ct), _ct_dtype(ct), _lib.ptr(labels), ct.numel(), n_labels,
                                                hu_min, n_bins, _lib.ptr(hist), _lib.ptr(oor), _lib.stream_ptr()))
    return hist, oor


def label_set_mask(labels: torch.Tensor, label_ids, ct: torch.Tensor | None = None, lo: int = 0, hi: int = 0,
                   mode: int = 0) -> torch.Tensor:
    """create_mask (+ HU window): mode 0 labels only, 1 inside [lo, hi], 2 outside (strict)."""
    _chk(labels, ct)
    sel = (C.c_uint8 * 256)()
    for i in ([label_ids] if isinstance(label_ids, int) else label_ids):
        sel[int(i)] = 1
    out = torch.empty(labels.shape, dtype=torch.uint8, device=labels.device)
    with torch.cuda.device(labels.device):
        _lib.check(_lib.lib().boa_mask_label_minus_window(
            _lib.ptr(ct), _ct_dtype(ct) if ct is not None else _lib.BOA_DT_I16, _lib.ptr(labels), labels.numel(), sel,
            lo, hi, mode, _lib.ptr(out), _lib.stream_ptr()))
    return out


def median3x3_slices(ct: torch.Tensor) -> torch.Tensor:
    """scipy.ndimage.median_filter(ct, size=[1, 3, 3]) (mode "reflect") of an int16 [z, y, x] volume."""
    _chk(ct)
    if ct.dtype != torch.int16 or ct.dim() != 3:
        raise TypeError("median3x3_slices needs an int16 [z, y, x] volume")
    out = torch.empty_like(ct)
    with torch.cuda.device(ct.device):
        _lib.check(_lib.lib().boa_median3x3_slices(_lib.ptr(ct), _lib.i32x3(ct.shape), _lib.ptr(out), _lib.stream_ptr()))
    return out


def erode_box(mask: torch.Tensor, before: int = 3, after: int = 2) -> torch.Tensor:
    """erode_region (compute/measurements.py:61-71): 6^3 footprint padded at the end => offsets -3..+2."""
    _chk(mask)
    tmp = torch.empty_like(mask)
    out = torch.empty_like(mask)
    with torch.cuda.device(mask.device):
        _lib.check(_lib.lib().boa_erode_box(_lib.ptr(mask), _lib.i32x3(mask.shape), before, after, _lib.ptr(tmp),
                                            _lib.ptr(out), _lib.stream_ptr()))
    return out
