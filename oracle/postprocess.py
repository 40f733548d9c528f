"""CPU restatement (test infrastructure) of the connected-component post-processing of the body-composition label maps:

  postprocess_region_segmentation   body_composition_analysis/body_regions/postprocess.py:8-40
  remove_small_labeled_objects      body_composition_analysis/body_parts/postprocess.py:7-52

skimage.measure.label / regionprops and skimage.morphology.remove_small_objects are restated on scipy.ndimage.label
(full connectivity, components numbered in raster order); the slice-wise contour fill uses OpenCV like the reference
when cv2 is importable and a scipy restatement (4-connected background not reachable from the slice border) otherwise.
Pinned against vectors produced by the reference's own functions: tests/golden/postprocess.npz.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

FULL3 = np.ones((3, 3, 3), dtype=bool)
THORACIC_CAVITY, MEDIASTINUM, PERICARDIUM, ABDOMINAL_CAVITY = 4, 9, 7, 3  # body_regions/definition.py:4-15


def _filter_largest_unique_segment(seg: np.ndarray, mask: np.ndarray, weights=None) -> None:
    """postprocess.py:8-16: every component of `mask` but the largest becomes 255 (ties: the first in raster order,
    because sorted(..., reverse=True) is stable and regionprops lists components by label)."""
    lab, n = ndimage.label(mask, structure=FULL3)
    if n <= 1:
        return
    w = None if weights is None else np.broadcast_to(np.asarray(weights)[:, None, None], mask.shape).ravel()
    area = np.bincount(lab.ravel(), weights=w, minlength=n + 1)[1:]
    keep = int(np.argmax(area)) + 1  # argmax returns the first maximum
    seg[(lab != 0) & (lab != keep)] = 255


def postprocess_region_segmentation(seg: np.ndarray, weights=None) -> np.ndarray:
    """postprocess.py:19-40.  weights: optional per-slice voxel weights (see boa_b200.postprocess)."""
    seg = np.array(seg, copy=True)
    _filter_largest_unique_segment(seg, seg > 0, weights)
    _filter_largest_unique_segment(seg, (seg == THORACIC_CAVITY) | (seg == MEDIASTINUM) | (seg == PERICARDIUM), weights)
    for region in (PERICARDIUM, ABDOMINAL_CAVITY):
        _filter_largest_unique_segment(seg, seg == region, weights)
    return seg


def fill_external_contours(label_mask: np.ndarray) -> np.ndarray:
    """body_parts/postprocess.py:33-40: per slice cv2.findContours(RETR_EXTERNAL) + drawContours(FILLED)."""
    filled = np.zeros(label_mask.shape, dtype=np.uint8)
    try:
        import cv2
    except Exception:  # pragma: no cover - scipy restatement of the same fill
        for i in range(label_mask.shape[0]):
            bg, n = ndimage.label(~label_mask[i])  # 4-connected background
            border = np.unique(np.concatenate([bg[0], bg[-1], bg[:, 0], bg[:, -1]]))
            filled[i] = ~np.isin(bg, border[border != 0])
        return filled.astype(bool)
    for i in range(label_mask.shape[0]):
        contours, _ = cv2.findContours(label_mask[i].astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
        cv2.drawContours(filled[i], contours, -1, color=1, thickness=cv2.FILLED)
    return filled.astype(bool)


def _remove_small(mask: np.ndarray, max_size: int, weights=None) -> None:
    """skimage.morphology.remove_small_objects(mask, max_size=max_size, connectivity=3, out=mask)."""
    lab, n = ndimage.label(mask, structure=FULL3)
    w = None if weights is None else np.broadcast_to(np.asarray(weights)[:, None, None], mask.shape).ravel()
    sizes = np.bincount(lab.ravel(), weights=w, minlength=n + 1)
    small = sizes <= max_size
    small[0] = False
    mask[small[lab]] = False


def remove_small_labeled_objects(seg: np.ndarray, threshold: int = 3000, weights=None) -> np.ndarray:
    """body_parts/postprocess.py:7-52."""
    out = np.zeros(seg.shape, dtype=seg.dtype)
    for label in [v for v in np.unique(seg) if v > 0]:
        filled = fill_external_contours(seg == label)
        _remove_small(filled, threshold - 1, weights)
        np.invert(filled, out=filled)
        _remove_small(filled, threshold - 1, weights)
        np.invert(filled, out=filled)
        out[filled] = label
    return out
