"""Resampling on the device: to the network spacing before the networks (order 3) and of the label maps back to the
input grid after them (order 0).  Mirrors nnUNet_predict_image's use of change_spacing
(_external/totalsegmentator/nnunet.py:457-475,685-687; scipy.ndimage.zoom semantics,
_external/totalsegmentator/resampling.py:24-56,129-222): slice thickness only for the body-composition networks
(`resample_only_thickness`), all three axes for `total` on inputs that are not at 1.5 mm."""
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


def zoomed_shape(shape, spacing, target) -> tuple:
    """Output shape of change_spacing: per axis round(n * zoom), zoom = float32 header spacing / new spacing
    (resampling.py:171-176, scipy.ndimage.zoom's `round`)."""
    target = [target] * len(shape) if np.isscalar(target) else list(target)
    return tuple(resampled_depth(int(n), float(sp), float(t)) for n, sp, t in zip(shape, spacing, target))


def resample_volume_cubic(ct: torch.Tensor, spacing, target=1.5) -> torch.Tensor:
    """int16 / float32 [z,y,x] at `spacing` (same axis order) -> int16 volume at `target` spacing: order-3 zoom with
    mode="nearest", computed as three 1-D spline passes in fp64 and truncated toward zero once at the end
    (change_spacing(img, [1.5]*3, order=3, dtype=np.int32), nnunet.py:466-467).  Axes already at the target spacing are
    left alone; identity when all three are (resampling.py:179-181)."""
    if not (ct.is_cuda and ct.is_contiguous() and ct.dim() == 3):
        raise ValueError("resample_volume_cubic needs a contiguous 3-D CUDA tensor")
    if ct.dtype not in (torch.int16, torch.float32):
        raise TypeError(f"CT must be int16 or float32, got {ct.dtype}")
    target = [float(target)] * 3 if np.isscalar(target) else [float(t) for t in target]
    axes = [a for a in range(3) if float(np.float32(spacing[a])) != float(np.float32(target[a]))]
    if not axes:
        return ct
    out_shape = zoomed_shape(ct.shape, spacing, target)
    L = _lib.lib()
    cur, cur_dt, shape = ct, (_lib.BOA_DT_I16 if ct.dtype == torch.int16 else _lib.BOA_DT_F32), list(ct.shape)
    with torch.cuda.device(ct.device):
        for k, a in enumerate(axes):
            if shape[a] < 2:
                raise ValueError("resample_volume_cubic: an axis of length 1 cannot be zoomed")
            last = k == len(axes) - 1
            outer = int(np.prod(shape[:a], dtype=np.int64))
            inner = int(np.prod(shape[a + 1:], dtype=np.int64))
            n_in, n_out = shape[a], out_shape[a]
            scratch = torch.empty((outer * (n_in + 24) * inner,), dtype=torch.float64, device=ct.device)
            shape[a] = n_out
            nxt = torch.empty(shape, dtype=torch.int16 if last else torch.float64, device=ct.device)
            _lib.check(L.boa_resample_axis_cubic(_lib.ptr(cur), cur_dt, outer, n_in, inner, n_out, _lib.ptr(scratch),
                                                 _lib.ptr(nxt), 1 if last else 0, _lib.stream_ptr()))
            cur, cur_dt = nxt, _lib.BOA_DT_F64
    return cur


def resample_labels_nearest(labels: torch.Tensor, out_shape) -> torch.Tensor:
    """uint8 [z,y,x] -> out_shape by order-0 zoom (change_spacing(seg, target_shape=..., order=0), nnunet.py:685-687)."""
    out_shape = tuple(int(v) for v in out_shape)
    if tuple(labels.shape) == out_shape:
        return labels
    if not (labels.is_cuda and labels.is_contiguous() and labels.dtype == torch.uint8):
        raise ValueError("resample_labels_nearest needs a contiguous uint8 CUDA tensor")
    out = torch.empty(out_shape, dtype=torch.uint8, device=labels.device)
    with torch.cuda.device(labels.device):
        _lib.check(_lib.lib().boa_resample_nearest_u8(_lib.ptr(labels), _lib.i32x3(labels.shape), _lib.i32x3(out_shape),
                                                      _lib.ptr(out), _lib.stream_ptr()))
    return out
