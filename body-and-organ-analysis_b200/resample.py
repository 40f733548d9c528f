"""Slice-thickness resampling on the device (the `resample_only_thickness` branch of nnUNet_predict_image,
_external/totalsegmentator/nnunet.py:457-475,685-687; change_spacing / scipy.ndimage.zoom semantics,
_external/totalsegmentator/resampling.py:24-56,129-222)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def resampled_depth(z_in: int, spacing_z: float, target: float) -> int:
    """ndimage.zoom output size: int(round(z_in * zoom)), zoom = img_spacing / new_spacing as float32 zooms give."""
    zoom = np.float64(np.float32(spacing_z)) / np.float64(target)
    return int(round(z_in * zoom))


def resample_thickness(ct: torch.Tensor, spacing_z: float, target: float = 5.0) -> torch.Tensor:
    """int16 [z,y,x] at slice thickness spacing_z -> int16 [z',y,x] at `target` (order 3, truncated to integers).
    Identity when the spacing already matches (resampling.py:179-181)."""
    if float(np.float32(spacing_z)) == float(np.float32(target)):
        return ct
    if not (ct.is_cuda and ct.is_contiguous()):
        raise ValueError("resample_thickness needs a contiguous CUDA tensor")
    dt = _lib.BOA_DT_I16 if ct.dtype == torch.int16 else _lib.BOA_DT_F32
    if ct.dtype not in (torch.int16, torch.float32):
        raise TypeError(f"CT must be int16 or float32, got {ct.dtype}")
    z_in, plane = ct.shape[0], ct.shape[1] * ct.shape[2]
    z_out = resampled_depth(z_in, spacing_z, target)
    scratch = torch.empty(((z_in + 24) * plane,), dtype=torch.float64, device=ct.device)
    out = torch.empty((z_out, ct.shape[1], ct.shape[2]), dtype=torch.int16, device=ct.device)
    with torch.cuda.device(ct.device):
        _lib.check(_lib.lib().boa_resample_z_cubic(_lib.ptr(ct), dt, z_in, plane, z_out, _lib.ptr(scratch),
                                                   _lib.ptr(out), _lib.stream_ptr()))
    return out


def upsample_labels_nearest(labels: torch.Tensor, z_out: int) -> torch.Tensor:
    """uint8 [z',y,x] -> [z_out,y,x] by order-0 zoom along z."""
    if labels.shape[0] == z_out:
        return labels
    out = torch.empty((z_out, labels.shape[1], labels.shape[2]), dtype=torch.uint8, device=labels.device)
    with torch.cuda.device(labels.device):
        _lib.check(_lib.lib().boa_resample_z_nearest_u8(_lib.ptr(labels), labels.shape[0],
                                                        labels.shape[1] * labels.shape[2], z_out, _lib.ptr(out),
                                                        _lib.stream_ptr()))
    return out
