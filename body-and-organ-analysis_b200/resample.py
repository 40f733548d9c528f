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


# ---------------------------------------------------------------------------------------------------------------------
# nnU-Net's own resampling between the image grid and the plan's spacing
# (_external/nnunetv2/preprocessing/preprocessors/default_preprocessor.py:57-90 on the way in,
#  _external/nnunetv2/inference/export_prediction.py:25-38 on the way out;
#  resample_data_or_seg_to_shape / determine_do_sep_z_and_axis / compute_new_shape:
#  _external/nnunetv2/preprocessing/resampling/default_resampling.py:14-203, ANISO_THRESHOLD = 3 in configuration.py)
ANISO_THRESHOLD = 3.0


def nnunet_new_shape(shape, old_spacing, new_spacing) -> tuple:
    """compute_new_shape (default_resampling.py:25-31)."""
    return tuple(int(round(float(i) / float(j) * int(k))) for i, j, k in zip(old_spacing, new_spacing, shape))


def nnunet_separate_z(current_spacing, new_spacing):
    """determine_do_sep_z_and_axis(force_separate_z=None, ...) (default_resampling.py:34-67): (do_separate_z, axis)."""
    def aniso(sp):
        return (max(sp) / min(sp)) > ANISO_THRESHOLD

    def lowres_axis(sp):
        sp = np.asarray(sp, dtype=np.float64)
        return np.where(sp.max() / sp == 1)[0]

    if aniso(current_spacing):
        axis = lowres_axis(current_spacing)
    elif aniso(new_spacing):
        axis = lowres_axis(new_spacing)
    else:
        return False, None
    if len(axis) != 1:  # (0.24, 1.25, 1.25)-like spacings: no separate treatment of the out-of-plane axis
        return False, None
    return True, int(axis[0])


def _cubic_grid_axis(cur: torch.Tensor, axis: int, n_out: int, last: bool) -> torch.Tensor:
    shape = list(cur.shape)
    outer = int(np.prod(shape[:axis], dtype=np.int64))
    inner = int(np.prod(shape[axis + 1:], dtype=np.int64))
    n_in = shape[axis]
    if n_in < 2:
        raise ValueError("an axis of length 1 cannot be resampled")
    scratch = torch.empty((outer * (n_in + 24) * inner,), dtype=torch.float64, device=cur.device)
    shape[axis] = n_out
    nxt = torch.empty(shape, dtype=torch.float32 if last else torch.float64, device=cur.device)
    dt = _lib.BOA_DT_F32 if cur.dtype == torch.float32 else _lib.BOA_DT_F64
    _lib.check(_lib.lib().boa_resample_axis_cubic_grid(_lib.ptr(cur), dt, outer, n_in, inner, n_out, _lib.ptr(scratch),
                                                       _lib.ptr(nxt), 3 if last else 0, _lib.stream_ptr()))
    return nxt


def resample_to_plan_spacing(data: torch.Tensor, current_spacing, new_spacing) -> torch.Tensor:
    """resampling_fn_data of the plans = resample_data_or_seg_to_shape(data, new_shape, current, new, is_seg=False,
    order=3, order_z=0, force_separate_z=None): the NORMALISED fp32 volume [z,y,x] -> fp32 volume on the plan's grid.
    skimage.transform.resize(order=3, mode="edge", anti_aliasing=False, clip=True) is scipy's cubic zoom with
    grid_mode coordinates, clipped to the value range of its input - per slice when the anisotropic axis is resampled
    separately (order 0 along it), over the whole volume otherwise."""
    if not (data.is_cuda and data.is_contiguous() and data.dim() == 3 and data.dtype == torch.float32):
        raise ValueError("resample_to_plan_spacing needs a contiguous fp32 [z,y,x] CUDA tensor")
    new_shape = nnunet_new_shape(data.shape, current_spacing, new_spacing)
    if tuple(new_shape) == tuple(data.shape):
        return data  # default_resampling.py:139: "no resampling necessary"
    sep, axis = nnunet_separate_z(current_spacing, new_spacing)
    if sep and axis != 0:
        raise NotImplementedError("separate-z resampling along an in-plane axis is not implemented (slices are dim 0)")
    L = _lib.lib()
    with torch.cuda.device(data.device):
        axes = [a for a in ((1, 2) if sep else (0, 1, 2)) if new_shape[a] != data.shape[a]]
        cur = data
        for k, a in enumerate(axes):
            cur = _cubic_grid_axis(cur, a, new_shape[a], last=k == len(axes) - 1)
        if axes:
            n_slices = int(data.shape[0]) if sep else 1
            mm = torch.empty(2 * n_slices, dtype=torch.int32, device=data.device)
            _lib.check(L.boa_clip_slices_f32(_lib.ptr(data), data.numel() // n_slices, _lib.ptr(cur),
                                             cur.numel() // n_slices, n_slices, _lib.ptr(mm), _lib.stream_ptr()))
        if sep and new_shape[0] != cur.shape[0]:
            out = torch.empty((new_shape[0], cur.shape[1], cur.shape[2]), dtype=torch.float32, device=data.device)
            _lib.check(L.boa_resample_z_nearest_grid_f32(_lib.ptr(cur), int(cur.shape[0]),
                                                         int(cur.shape[1] * cur.shape[2]), int(new_shape[0]),
                                                         _lib.ptr(out), _lib.stream_ptr()))
            cur = out
    return cur
