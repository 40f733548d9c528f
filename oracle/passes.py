"""Oracle: memory-bound passes of the hot path, restated in numpy (TEST INFRASTRUCTURE, see __init__.py).

  ct_normalize            _external/nnunetv2/preprocessing/normalization/default_normalization_schemes.py:56-67
  subclassify_tissues     _external/body_composition_analysis/tissue/subclassification.py:38-53, tissue/definition.py:6-30
  slice_label_stats       _external/body_composition_analysis/report/builder.py:403-444 (per-slice counts, torso mask),
                          :284-305 (mean HU = sum / count), :56-99 and commands.py:34-44 (slice presence)
  create_mask             compute/util.py:25-31
  region_minus_fat        compute/measurements.py:29-39
  erode_region            compute/measurements.py:61-71 (skimage binary_erosion, even footprint padded at the end)
  metrics_for_region      compute/measurements.py:74-123
  metrics_from_hist       the same statistics computed from an integer-HU histogram (what the CUDA path does on the
                          host after boa_label_hu_hist); tested equal to metrics_for_region
Each is pinned against the reference's own function by tests/golden/make_golden.py where that function is importable.
"""
from __future__ import annotations

import numpy as np
from scipy import ndimage

HU_RANGES = {"ALL": (-1000, 3000), "ADIPOSE_TISSUE": (-190, -30), "MUSCLE_TISSUE": (-29, 150)}
# tissue id -> (HU range, body region id)   (tissue/definition.py:22-30, body_regions/definition.py:4-15)
TISSUE_RULES = {1: ("MUSCLE_TISSUE", 2), 2: ("ALL", 5), 3: ("ADIPOSE_TISSUE", 1), 4: ("ADIPOSE_TISSUE", 3),
                5: ("ADIPOSE_TISSUE", 2), 6: ("ADIPOSE_TISSUE", 9), 7: ("ADIPOSE_TISSUE", 7)}
ADIPOSE_TISSUE = (-200, -40)  # compute/measurements.py:22


def ct_normalize(image: np.ndarray, props: dict) -> np.ndarray:
    image = image.astype(np.float32, copy=True)
    np.clip(image, props["percentile_00_5"], props["percentile_99_5"], out=image)
    image -= props["mean"]
    image /= max(props["std"], 1e-8)
    return image


def subclassify_tissues(image: np.ndarray, regions: np.ndarray) -> np.ndarray:
    masks = {k: np.logical_and(image >= lo, image <= hi) for k, (lo, hi) in HU_RANGES.items()}
    out = np.zeros_like(regions)
    for tissue, (hu, region) in TISSUE_RULES.items():
        out[masks[hu] & (regions == region)] = tissue
    return out


def slice_label_stats(labels: np.ndarray, n_labels: int, ct: np.ndarray | None = None, mask: np.ndarray | None = None,
                      mask_value: int = 1):
    Z = labels.shape[0]
    counts = np.zeros((Z, n_labels), dtype=np.int64)
    sums = np.zeros((Z, n_labels), dtype=np.int64)
    for k in range(n_labels):
        m = labels == k
        if mask is not None:
            m = np.logical_and(m, mask == mask_value)
        counts[:, k] = m.sum(axis=(1, 2))
        if ct is not None:
            sums[:, k] = np.where(m, ct.astype(np.int64), 0).sum(axis=(1, 2))
    return counts, (sums if ct is not None else None)


def create_mask(region_data: np.ndarray, labels) -> np.ndarray:
    mask = np.zeros(region_data.shape, dtype=bool)
    if isinstance(labels, (int, np.integer)):
        mask[region_data == labels] = True
    else:
        mask[np.isin(region_data, labels)] = True
    return mask


def region_minus_fat(ct: np.ndarray, mask: np.ndarray) -> np.ndarray:
    return np.logical_and(mask, np.logical_or(ct < ADIPOSE_TISSUE[0], ct > ADIPOSE_TISSUE[1]))


def erode_region(mask: np.ndarray, kernel_value: int = 6) -> np.ndarray:
    """skimage 0.26 binary_erosion(mask, pad_footprint(ones(6,6,6), pad_end=True)) ==
    scipy binary_erosion(structure=that footprint, border_value=True): window offsets -3..+2 per axis."""
    fp = np.ones([kernel_value] * 3, dtype=bool)
    if kernel_value % 2 == 0:
        fp = np.pad(fp, [(0, 1)] * 3)
    return ndimage.binary_erosion(mask, structure=fp, border_value=True)


def metrics_for_region(ct: np.ndarray, mask: np.ndarray, aut_mean, aut_std, spacing) -> dict:
    m = {}
    if np.sum(mask) == 0:
        return {"present": False}
    ml_per_voxel = np.prod(spacing) / 1000.0
    m["present"] = True
    hu = ct[mask]
    m["volume_ml"] = np.sum(mask) * ml_per_voxel
    m["mean_hu"], m["std_hu"] = float(np.mean(hu)), float(np.std(hu))
    m["min_hu"], m["median_hu"], m["max_hu"] = float(np.min(hu)), float(np.median(hu)), float(np.max(hu))
    for p in (25, 75):
        m[f"{p}th_percentile_hu"] = float(np.percentile(hu, p))
    m["cnr"] = (np.mean(hu) - aut_mean) / aut_std if aut_mean is not None and aut_std is not None else None
    return m


def _order_stat(cum: np.ndarray, values: np.ndarray, k: int) -> float:
    return float(values[np.searchsorted(cum, k, side="right")])


def _percentile(cum, values, n, q):
    pos = (n - 1) * q / 100.0  # numpy 'linear' interpolation
    lo, hi = int(np.floor(pos)), int(np.ceil(pos))
    a, b = _order_stat(cum, values, lo), _order_stat(cum, values, hi)
    t = pos - lo
    # numpy's _lerp: a + (b - a) * t, switched to b - (b - a) * (1 - t) for t >= 0.5
    return float(b - (b - a) * (1 - t)) if t >= 0.5 else float(a + (b - a) * t)


def metrics_from_hist(hist: np.ndarray, hu_min: int, aut_mean, aut_std, spacing) -> dict:
    """hist[i] = number of voxels with HU == hu_min + i."""
    hist = hist.astype(np.int64)
    n = int(hist.sum())
    if n == 0:
        return {"present": False}
    values = np.arange(hu_min, hu_min + hist.size, dtype=np.int64)
    nz = np.nonzero(hist)[0]
    cum = np.cumsum(hist)
    s1 = int((hist * values).sum())
    mean = s1 / n
    var = float((hist * (values - mean) ** 2).sum() / n)
    m = {"present": True, "volume_ml": n * (np.prod(spacing) / 1000.0), "mean_hu": float(mean),
         "std_hu": float(np.sqrt(var)), "min_hu": float(values[nz[0]]),
         "median_hu": _percentile(cum, values, n, 50), "max_hu": float(values[nz[-1]])}
    for p in (25, 75):
        m[f"{p}th_percentile_hu"] = _percentile(cum, values, n, p)
    m["cnr"] = (mean - aut_mean) / aut_std if aut_mean is not None and aut_std is not None else None
    return m
