"""Connected-component post-processing of the body-composition label maps on the device.  Mirrors

  postprocess_region_segmentation  (_external/body_composition_analysis/body_regions/postprocess.py:19-40)
  postprocess_part_segmentation    (_external/body_composition_analysis/body_parts/postprocess.py:55-60,
                                    remove_small_labeled_objects :7-52)

which the reference applies to `body_regions.nii.gz` / `body_parts.nii.gz` right after the networks
(infer/infer.py:81-89), i.e. before the tissue rules and every measurement.

The reference runs them on the label map at the INPUT grid, after the 5 mm prediction has been replicated along z
(order-0 zoom).  Replicating slices maps 26-connected components (and the 4-connected background of every slice) one to
one, so the same result comes from labelling the 5 mm map with every slice weighted by the number of output slices it
becomes (`slice_weights`) - 3.3 x less work at 1.5 mm; the functions take the weights as an optional argument and
count plain voxels without it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

THORACIC_CAVITY, MEDIASTINUM, PERICARDIUM, ABDOMINAL_CAVITY = 4, 9, 7, 3  # body_regions/definition.py:4-15
SMALL_OBJECT_THRESHOLD = 3000                                             # body_parts/postprocess.py:7
OP_KEEP_LARGEST, OP_REMOVE_SMALL, OP_FILL_ENCLOSED = 0, 1, 2
MODE_26, MODE_SLICE_4 = 0, 1


def slice_weights(z_in: int, z_out: int, device) -> torch.Tensor:
    """int32 [z_in]: how many slices of the order-0 zoom to z_out come from each source slice
    (index = floor(o * (z_in - 1) / (z_out - 1) + 0.5), as upsample_labels_nearest / scipy order 0)."""
    if z_in == z_out:
        return torch.ones(z_in, dtype=torch.int32, device=device)
    zoom = (z_in - 1) / (z_out - 1) if z_out > 1 else 1.0
    src = np.clip(np.floor(np.arange(z_out, dtype=np.float64) * zoom + 0.5).astype(np.int64), 0, z_in - 1)
    return torch.from_numpy(np.bincount(src, minlength=z_in).astype(np.int32)).to(device)


class _Scratch:
    def __init__(self, seg: torch.Tensor, need_border: bool):
        n = seg.numel()
        if n >= 2 ** 31:
            raise ValueError("connected-component labelling is limited to 2^31 voxels")
        self.labels = torch.empty(n, dtype=torch.int32, device=seg.device)
        self.sizes = torch.empty(n, dtype=torch.int32, device=seg.device)
        self.border = torch.empty(n, dtype=torch.int32, device=seg.device) if need_border else None
        self.best = torch.empty(2, dtype=torch.int64, device=seg.device)


def _label_set(ids) -> C.Array:
    sel = (C.c_uint8 * 256)()
    for i in ids:
        sel[int(i)] = 1
    return sel


def _cc_filter(seg, ids, invert, mode, op, threshold, fill_value, weights, scratch) -> None:
    if not (seg.is_cuda and seg.is_contiguous() and seg.dtype == torch.uint8 and seg.dim() == 3):
        raise ValueError("label maps must be contiguous uint8 CUDA tensors [z, y, x]")
    if weights is not None and not (weights.dtype == torch.int32 and weights.numel() == seg.shape[0] and weights.is_cuda):
        raise ValueError("slice weights must be an int32 CUDA tensor of length z")
    with torch.cuda.device(seg.device):
        _lib.check(_lib.lib().boa_cc_filter(
            _lib.ptr(seg), _lib.i32x3(seg.shape), _label_set(ids), int(invert), mode, op, int(threshold), int(fill_value),
            _lib.ptr(weights), _lib.ptr(scratch.labels), _lib.ptr(scratch.sizes), _lib.ptr(scratch.border),
            _lib.ptr(scratch.best), _lib.stream_ptr()))


def postprocess_region_segmentation(body_regions: torch.Tensor, weights: torch.Tensor | None = None) -> torch.Tensor:
    """Every region that can only be one piece keeps its largest 26-connected component, the rest becomes 255
    (body_regions/postprocess.py:19-40): all non-zero voxels; thoracic cavity + mediastinum + pericardium; pericardium;
    abdominal cavity - in this order, each on the result of the previous."""
    seg = body_regions.clone()
    scratch = _Scratch(seg, need_border=False)
    for ids in (range(1, 256), (THORACIC_CAVITY, MEDIASTINUM, PERICARDIUM), (PERICARDIUM,), (ABDOMINAL_CAVITY,)):
        _cc_filter(seg, ids, False, MODE_26, OP_KEEP_LARGEST, 0, 255, weights, scratch)
    return seg


def postprocess_pair_distributed(body_parts: torch.Tensor, body_regions: torch.Tensor, weights, dist_ctx):
    """Both post-processings at once on >= 2 ranks: ranks 0 .. W-2 share the body_parts labels, rank W-1 runs the four
    dependent body_regions passes; a MAX all-reduce combines the parts, a broadcast distributes the regions.  Same
    results as running the two functions one after the other on every rank."""
    import torch.distributed as dist
    W, rank = dist_ctx.world_size, dist_ctx.rank
    last = W - 1
    if rank == last:
        regions = postprocess_region_segmentation(body_regions, weights)
        parts = torch.zeros_like(body_parts)
    else:
        regions = torch.empty_like(body_regions)
        parts = postprocess_part_segmentation(body_parts, weights, label_share=(rank, W - 1))
    dist.all_reduce(parts, op=dist.ReduceOp.MAX, group=dist_ctx.group)
    src = last if dist_ctx.group is None else dist.get_global_rank(dist_ctx.group, last)
    dist.broadcast(regions, src=src, group=dist_ctx.group)
    return parts, regions


def postprocess_part_segmentation(body_parts: torch.Tensor, weights: torch.Tensor | None = None,
                                  threshold: int = SMALL_OBJECT_THRESHOLD, labels=None, dist_ctx=None,
                                  label_share=None) -> torch.Tensor:
    """remove_small_labeled_objects (body_parts/postprocess.py:7-52): per label (ascending) fill the external contours
    of every slice, drop 26-connected objects of fewer than `threshold` voxels, close 26-connected holes of fewer than
    `threshold` voxels, paint the label (later labels overwrite earlier ones).

    Every label is processed from the ORIGINAL map and painting in ascending order means "the largest label that
    claims a voxel wins", so with several GPUs (dist_ctx) the labels are dealt out to the ranks round robin and the
    per-rank results are combined with an element-wise MAX all-reduce - the same map on every rank."""
    from . import passes

    if labels is None:
        present = torch.bincount(body_parts.flatten().to(torch.int64), minlength=256).cpu().numpy()
        labels = [int(v) for v in np.nonzero(present)[0] if v > 0]  # np.unique(mask), labels > 0
    labels = sorted(labels)
    world = dist_ctx.world_size if dist_ctx is not None else 1
    if label_share is not None:      # (index, count): this caller's share, the caller combines the results
        labels = labels[label_share[0]::label_share[1]]
    elif world > 1:
        labels = labels[dist_ctx.rank::world]
    out = torch.zeros_like(body_parts)
    scratch = _Scratch(body_parts, need_border=True) if labels else None
    for label in labels:
        filled = passes.label_set_mask(body_parts, [label])  # uint8 0 / 1
        # slice-wise external-contour fill == background that no 4-connected path links to the slice border
        _cc_filter(filled, (1,), True, MODE_SLICE_4, OP_FILL_ENCLOSED, 0, 1, None, scratch)
        _cc_filter(filled, (1,), False, MODE_26, OP_REMOVE_SMALL, threshold - 1, 0, weights, scratch)
        _cc_filter(filled, (1,), True, MODE_26, OP_REMOVE_SMALL, threshold - 1, 1, weights, scratch)
        with torch.cuda.device(out.device):
            _lib.check(_lib.lib().boa_paint_label(_lib.ptr(filled), filled.numel(), label, _lib.ptr(out), _lib.stream_ptr()))
    if world > 1 and label_share is None:
        import torch.distributed as dist
        dist.all_reduce(out, op=dist.ReduceOp.MAX, group=dist_ctx.group)
    return out
