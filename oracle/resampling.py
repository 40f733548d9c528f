"""Oracle: nnU-Net's own resampling between the image grid and the plan's spacing (TEST INFRASTRUCTURE).

Restates, on numpy / scipy:
  compute_new_shape, determine_do_sep_z_and_axis, resample_data_or_seg
      _external/nnunetv2/preprocessing/resampling/default_resampling.py:14-203
  the two call sites: DefaultPreprocessor.run_case_npy (default_preprocessor.py:57-90, order 3 / order_z 0 on the image
  data) and convert_predicted_logits_to_segmentation_with_correct_shape (export_prediction.py:25-38, order 1 / order_z 0 on
  the logits, then argmax).
`skimage.transform.resize(image, shape, order, mode="edge", anti_aliasing=False)` is not installable here; per its
source (scikit-image 0.26, transform/_warps.py: resize -> ndi.zoom(image, zoom, order=order, mode="nearest",
grid_mode=True) followed by _clip_warp_output: np.clip to the input's [min, max]) it is restated on scipy.
Parity of this file: compute_new_shape, determine_do_sep_z_and_axis and resample_data are PINNED to outputs of the
reference's own default_resampling.py run in the development container (tests/golden/make_golden_resampling.py ->
tests/golden/resampling.{json,npz}, bit-equal); only `resize` itself is unpinned by skimage (not importable) and
rests on its source as quoted above.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

ANISO_THRESHOLD = 3


def resize(image: np.ndarray, output_shape, order: int) -> np.ndarray:
    image = image.astype(np.float64)
    zoom = [o / i for o, i in zip(output_shape, image.shape)]
    out = ndimage.zoom(image, zoom, order=order, mode="nearest", grid_mode=True)
    assert out.shape == tuple(output_shape)
    return np.clip(out, image.min(), image.max())


def compute_new_shape(old_shape, old_spacing, new_spacing):
    return tuple(int(round(i / j * k)) for i, j, k in zip(old_spacing, new_spacing, old_shape))


def determine_do_sep_z_and_axis(current_spacing, new_spacing):
    def aniso(sp):
        return (np.max(sp) / np.min(sp)) > ANISO_THRESHOLD

    def lowres(sp):
        return np.where(max(sp) / np.array(sp) == 1)[0]

    if aniso(current_spacing):
        axis = lowres(current_spacing)
    elif aniso(new_spacing):
        axis = lowres(new_spacing)
    else:
        return False, None
    if len(axis) != 1:
        return False, None
    return True, int(axis[0])


def resample_data(data: np.ndarray, new_shape, current_spacing, new_spacing, order: int, order_z: int = 0) -> np.ndarray:
    """data [c, z, y, x] -> [c, *new_shape] (resample_data_or_seg with is_seg=False)."""
    new_shape = tuple(int(v) for v in new_shape)
    if tuple(data.shape[1:]) == new_shape:
        return data
    sep, axis = determine_do_sep_z_and_axis(current_spacing, new_spacing)
    out = np.zeros((data.shape[0], *new_shape), dtype=data.dtype)
    for c in range(data.shape[0]):
        if not sep:
            out[c] = resize(data[c], new_shape, order)
            continue
        assert axis == 0
        here = np.stack([resize(data[c, i], new_shape[1:], order) for i in range(data.shape[1])])
        if here.shape[0] != new_shape[0]:
            scale = here.shape[0] / new_shape[0]
            zz, yy, xx = np.mgrid[:new_shape[0], :new_shape[1], :new_shape[2]]
            coords = np.array([scale * (zz + 0.5) - 0.5, yy.astype(float), xx.astype(float)])
            here = ndimage.map_coordinates(here, coords, order=order_z, mode="nearest")
        out[c] = here
    return out


def logits_to_segmentation(logits: np.ndarray, shape_before_resampling, plan_spacing, original_spacing) -> np.ndarray:
    """export_prediction.py:25-38: resample every channel (order 1, order_z 0), argmax with the first maximum."""
    r = resample_data(logits.astype(np.float64), shape_before_resampling, plan_spacing, original_spacing, order=1)
    return r.argmax(0).astype(np.uint8)
