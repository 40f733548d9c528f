"""Oracle: sliding-window inference, Gaussian aggregation, argmax, part merge (TEST INFRASTRUCTURE, see __init__.py).

Restates, in numpy / torch fp32 on the CPU:
  compute_gaussian, compute_steps_for_sliding_window   _external/nnunetv2/inference/sliding_window_prediction.py:10-54
  _internal_get_sliding_window_slicers                  _external/nnunetv2/inference/predict_from_raw_data.py:506-538
  _internal_predict_sliding_window_return_logits        predict_from_raw_data.py:560-631
  predict_sliding_window_return_logits (pad / un-pad)   predict_from_raw_data.py:634-680  (+ acvl_utils 0.2.5 pad_nd_image)
  predict_logits_from_preprocessed_data (fold mean)     predict_from_raw_data.py:471-504
  convert_logits_to_segmentation (argmax, first max)    _external/nnunetv2/utilities/label_handling/label_handling.py:174-180
  part -> global label merge                            _external/totalsegmentator/nnunet.py:553-556
The geometry / Gaussian functions are pinned against the reference's own functions by tests/golden/make_golden.py.
"""
from __future__ import annotations

from itertools import product

import numpy as np
import torch
from scipy.ndimage import gaussian_filter

from .network import unet_forward


def compute_gaussian(tile_size, sigma_scale=1.0 / 8, value_scaling_factor=1.0) -> np.ndarray:
    """sliding_window_prediction.py:10-27 - returns the fp16 map as a numpy float16 array."""
    tmp = np.zeros(tile_size)
    center = [i // 2 for i in tile_size]
    tmp[tuple(center)] = 1
    g = gaussian_filter(tmp, [i * sigma_scale for i in tile_size], 0, mode="constant", cval=0)
    g = torch.from_numpy(g)
    g /= (torch.max(g) / value_scaling_factor)
    g = g.to(dtype=torch.float16)
    mask = g == 0
    g[mask] = torch.min(g[~mask])
    return g.numpy()


def compute_steps_for_sliding_window(image_size, tile_size, tile_step_size):
    """sliding_window_prediction.py:30-54."""
    target = [i * tile_step_size for i in tile_size]
    num_steps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image_size, target, tile_size)]
    steps = []
    for dim in range(len(tile_size)):
        max_step_value = image_size[dim] - tile_size[dim]
        actual = max_step_value / (num_steps[dim] - 1) if num_steps[dim] > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(num_steps[dim])])
    return steps


def sliding_window_slicers(image_size, patch, step):
    """predict_from_raw_data.py:524-537 (3-D branch): dim 0 outermost."""
    steps = compute_steps_for_sliding_window(image_size, patch, step)
    out = []
    for sx, sy, sz in product(*steps):
        out.append(tuple(slice(si, si + ti) for si, ti in zip((sx, sy, sz), patch)))
    return out


def pad_nd_image(image: np.ndarray, new_shape):
    """acvl_utils 0.2.5 pad_nd_image(image, new_shape, 'constant', {'value': 0}, True, None) on the trailing dims."""
    old = np.array(image.shape[-len(new_shape):])
    new = np.array([max(n, o) for n, o in zip(new_shape, old)])
    diff = new - old
    below = diff // 2
    above = diff // 2 + diff % 2
    pad = [[0, 0]] * (image.ndim - len(new_shape)) + [[int(b), int(a)] for b, a in zip(below, above)]
    res = np.pad(image, pad, mode="constant", constant_values=0) if any(diff) else image
    slicer = tuple([slice(None)] * (image.ndim - len(new_shape)) +
                   [slice(int(b), int(b) + int(o)) for b, o in zip(below, old)])
    return res, slicer


@torch.inference_mode()
def predict_sliding_window_return_logits(arch, fold_state_dicts, data: np.ndarray, step: float, use_gaussian=True,
                                         emulate_fp16=False, accumulator_dtype=torch.float32,
                                         patch_range=None) -> np.ndarray:
    """data [1, x, y, z] fp32 -> logits [C, x, y, z] (fp32), mean over folds.

    accumulator_dtype=torch.float16 reproduces the reference's half-precision accumulators
    (predict_from_raw_data.py:587-590); float32 is what the product uses."""
    patch = arch["patch_size"]
    padded, unpad = pad_nd_image(data, patch)
    slicers = sliding_window_slicers(padded.shape[1:], patch, step)
    if patch_range is not None:
        slicers = slicers[patch_range[0]:patch_range[1]]
    g = torch.from_numpy(compute_gaussian(tuple(patch), 1.0 / 8, 10).astype(np.float32)) if use_gaussian else \
        torch.ones(tuple(patch))
    x = torch.from_numpy(np.ascontiguousarray(padded, dtype=np.float32))
    total = None
    for sd in fold_state_dicts:
        logits = torch.zeros((arch["num_classes"], *padded.shape[1:]), dtype=accumulator_dtype)
        n_pred = torch.zeros(padded.shape[1:], dtype=accumulator_dtype)
        for sl in slicers:
            pred = unet_forward(arch, sd, x[(slice(None), *sl)][None], emulate_fp16)[0]
            pred = (pred * g).to(accumulator_dtype)
            logits[(slice(None), *sl)] += pred
            n_pred[sl] += g.to(accumulator_dtype)
        logits = (logits.float() / n_pred.float())
        if torch.any(torch.isinf(logits)):
            raise RuntimeError("Encountered inf in predicted array. Aborting...")
        total = logits if total is None else total + logits
    total /= len(fold_state_dicts)
    return total[unpad].numpy() if patch_range is None else total.numpy()


def convert_logits_to_segmentation(logits: np.ndarray) -> np.ndarray:
    """label_handling.py:178 (non-region labels): argmax over channel 0; uint8 (export_prediction.py:46)."""
    return logits.argmax(0).astype(np.uint8)


def merge_parts(part_segs, part_luts, shape) -> np.ndarray:
    """totalsegmentator/nnunet.py:534-556: seg_combined[seg == jdx] = global id, later parts overwrite earlier ones,
    background never overwrites."""
    combined = np.zeros(shape, dtype=np.uint8)
    for seg, lut in zip(part_segs, part_luts):
        for jdx, gid in enumerate(lut):
            if jdx == 0:
                continue
            combined[seg == jdx] = gid
    return combined
