"""Sliding-window geometry: host-side mirror of
  compute_gaussian / compute_steps_for_sliding_window  (_external/nnunetv2/inference/sliding_window_prediction.py:10-54)
  nnUNetPredictor._internal_get_sliding_window_slicers (_external/nnunetv2/inference/predict_from_raw_data.py:506-538)
  acvl_utils.cropping_and_padding.padding.pad_nd_image (acvl-utils==0.2.5, call site predict_from_raw_data.py:657)
Pure integer / numpy host logic - the only floating-point object, the Gaussian map, is tiny and input independent.
"""
from __future__ import annotations

from functools import lru_cache
from itertools import product

import numpy as np


def compute_steps_for_sliding_window(image_size, tile_size, tile_step_size: float) -> list[list[int]]:
    if not all(i >= j for i, j in zip(image_size, tile_size)):
        raise ValueError("image size must be as large or larger than patch_size")
    if not 0 < tile_step_size <= 1:
        raise ValueError("step_size must be larger than 0 and smaller or equal to 1")
    target = [i * tile_step_size for i in tile_size]
    num_steps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image_size, target, tile_size)]
    steps = []
    for dim in range(len(tile_size)):
        max_step = image_size[dim] - tile_size[dim]
        actual = max_step / (num_steps[dim] - 1) if num_steps[dim] > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(num_steps[dim])])
    return steps


def sliding_window_origins(image_size, tile_size, tile_step_size: float) -> np.ndarray:
    """Patch origins in the reference's slicer order: array dim 0 outermost, dim 2 innermost. int32 [n, 3]."""
    steps = compute_steps_for_sliding_window(image_size, tile_size, tile_step_size)
    return np.array(list(product(*steps)), dtype=np.int32).reshape(-1, len(tile_size))


@lru_cache(maxsize=8)
def _gaussian_cached(tile_size: tuple, sigma_scale: float, value_scaling_factor: float) -> np.ndarray:
    from scipy.ndimage import gaussian_filter

    tmp = np.zeros(tile_size)
    tmp[tuple(i // 2 for i in tile_size)] = 1
    g = gaussian_filter(tmp, [i * sigma_scale for i in tile_size], 0, mode="constant", cval=0)
    g = g / (g.max() / value_scaling_factor)
    g = g.astype(np.float16)            # the reference keeps the map in half precision
    g[g == 0] = g[g != 0].min()
    return g


def compute_gaussian(tile_size, sigma_scale: float = 1.0 / 8, value_scaling_factor: float = 10.0) -> np.ndarray:
    """fp16 importance map exactly as the reference builds it (predict_from_raw_data.py:593 passes
    value_scaling_factor=10); callers widen it to fp32 for the fp32 accumulators."""
    return _gaussian_cached(tuple(int(t) for t in tile_size), float(sigma_scale), float(value_scaling_factor))


def pad_to_patch(shape, patch) -> tuple[list[tuple[int, int]], tuple[slice, ...]]:
    """pad_nd_image(..., 'constant', value 0): pad only where shape < patch, below = diff // 2,
    above = diff // 2 + diff % 2.  Returns the (below, above) pads and the slicer that undoes them."""
    pads, slicer = [], []
    for s, p in zip(shape, patch):
        diff = max(p - s, 0)
        below, above = diff // 2, diff // 2 + diff % 2
        pads.append((below, above))
        slicer.append(slice(below, below + s))
    return pads, tuple(slicer)


def shard_patches(n_patches: int, world_size: int, rank: int) -> tuple[int, int]:
    """Count-balanced contiguous run [begin, end) of the slicer list owned by `rank`."""
    base, rem = divmod(n_patches, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


TRIPLE_SPLIT_VOXELS = 512 * 512 * 900  # totalsegmentator/nnunet.py:483
TRIPLE_SPLIT_MARGIN = 20               # :495


def needs_triple_split(shape, multimodel: bool, force_split: bool = False) -> bool:
    """nnUNet_predict_image splits a volume into three overlapping z-parts when it is very large and several models
    run on it (`total`), or when the caller forces it (the body-composition networks on more than 400 slices at 5 mm,
    compute/inference.py:109-128): totalsegmentator/nnunet.py:483-492.  shape = array shape [z, y, x]."""
    z = int(shape[0])
    return bool(force_split or (int(np.prod([int(v) for v in shape], dtype=np.int64)) > TRIPLE_SPLIT_VOXELS and z > 200
                                and multimodel))


def triple_split_ranges(z: int, margin: int = TRIPLE_SPLIT_MARGIN):
    """The three parts of nnunet.py:493-505 and how :582-586 stitches their predictions, along the slice axis:
    [(part_lo, part_hi, keep_lo, keep_hi, dst_lo, dst_hi)] - predict volume[part_lo:part_hi], write its slices
    [keep_lo:keep_hi] to out[dst_lo:dst_hi]."""
    third = z // 3
    parts = [(0, third + margin), (third + 1 - margin, 2 * third + margin), (2 * third + 1 - margin, z)]
    keeps = [(0, third), (margin - 1, margin - 1 + third), (margin - 1, None)]
    dsts = [(0, third), (third, 2 * third), (2 * third, z)]
    out = []
    for (plo, phi), (klo, khi), (dlo, dhi) in zip(parts, keeps, dsts):
        if plo < 0 or phi > z or phi - plo <= 0:
            raise ValueError(f"a volume of {z} slices is too short for the triple split (margin {margin})")
        khi = (phi - plo) if khi is None else khi
        if khi - klo != dhi - dlo:
            raise ValueError(f"a volume of {z} slices is too short for the triple split (margin {margin})")
        out.append((plo, phi, klo, khi, dlo, dhi))
    return out
