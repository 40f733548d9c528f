"""Body cross-section axes at L3: drop-in for `major_minor_axis` (body_organ_analysis/compute/ts_metrics.py:32-61) and
the geometry it rests on, `find_axes` (compute/geometry.py:49-85).

The device side is one pass the pipeline already has - the per-slice presence of the vertebrae_L3 label
(boa_slice_label_stats) - plus a single 2-D slice of the body mask copied to the host (a few hundred KB); the geometry
of that one slice (convex hull, farthest pair, the perpendicular through its midpoint intersected with the drawn
contour) is host work on OpenCV / scipy, exactly like the reference's, because its result is defined by OpenCV's
rasterisation of the contour and of the probing line.
"""
from __future__ import annotations

import math

import numpy as np


def find_axes(middle_slice: np.ndarray):
    """((x, y) major end 1, major end 2, minor end 1, minor end 2) of a binary slice [rows, cols]."""
    import cv2
    from scipy import spatial

    mask = np.asarray(middle_slice).astype(bool)
    # (x, y) points in the order the reference feeds the hull (reversed raster order): ties of the farthest pair are
    # broken by that order
    pts = np.argwhere(mask)[::-1, ::-1]
    hull = pts[spatial.ConvexHull(pts).vertices]
    dist = spatial.distance.cdist(hull, hull, metric="euclidean")
    i, j = np.unravel_index(dist.argmax(), dist.shape)
    a, b = hull[i], hull[j]
    mid = (int((a[0] + b[0]) // 2), int((a[1] + b[1]) // 2))
    reach = int(sum(mask.shape))  # longer than any chord of the slice
    dx, dy = float(a[0] - b[0]), float(a[1] - b[1])
    norm = math.sqrt(dx * dx + dy * dy)
    dx, dy = dx / norm, dy / norm
    contours, _ = cv2.findContours(mask.astype(np.uint8), cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_SIMPLE)
    outline = cv2.drawContours(np.zeros(mask.shape), contours, contourIdx=-1, color=1, thickness=2)

    def probe(px: float, py: float):
        far = (int(mid[0] + px * reach), int(mid[1] + py * reach))
        ray = cv2.line(np.zeros(mask.shape), [far[0], far[1]], [mid[0], mid[1]], 1, 2)
        rows, cols = np.logical_and(outline, ray).nonzero()
        return (int(cols[0]), int(rows[0]))  # first hit in raster order, as the reference takes it

    return (int(a[0]), int(a[1])), (int(b[0]), int(b[1])), probe(-dy, dx), probe(dy, -dx)


def axes_of_slice(body_slice: np.ndarray, spacing_xy):
    """(major, minor) in mm of one body-mask slice, or (None, None) when it is empty."""
    if not np.any(body_slice):
        return None, None
    p1, p2, q1, q2 = find_axes(body_slice)
    avg = float(np.mean(spacing_xy))
    return math.dist(p1, p2) * avg, math.dist(q1, q2) * avg


def major_minor_axis(total, body_parts, spacing_xy, l3_label: int | None = None):
    """total / body_parts: uint8 [z, y, x] label maps (device tensors or numpy).  The middle one (median index) of the
    slices that contain vertebrae_L3 -> body mask (body_parts == 1) of that slice -> (major, minor) axis in mm."""
    from .labels import class_map
    if l3_label is None:
        l3_label = {v: k for k, v in class_map("total").items()}["vertebrae_L3"]
    if hasattr(total, "is_cuda") and total.is_cuda:
        from . import passes
        counts, _ = passes.slice_label_stats(total, l3_label + 1)
        present = np.nonzero(counts[:, l3_label].cpu().numpy() > 0)[0]
        if present.size == 0:
            return None, None
        z = int(np.median(present))
        sl = (body_parts[z] == 1).cpu().numpy()
    else:
        total, body_parts = np.asarray(total), np.asarray(body_parts)
        present = np.nonzero((total == l3_label).any(axis=(1, 2)))[0]
        if present.size == 0:
            return None, None
        z = int(np.median(present))
        sl = body_parts[z] == 1
    return axes_of_slice(sl, spacing_xy)


_PINNED: dict = {}


class PendingL3Axes:
    """major_minor_axis for device label maps without a stall of the launch queue and off the critical path: the middle
    L3 slice index is computed on the device (median of the slice indices that contain the label, int() as the
    reference: floor of the mean of the two middle ones), the body mask of that slice is copied to pinned host memory
    asynchronously, and a host thread evaluates the geometry (OpenCV / qhull release the GIL) while the caller keeps
    enqueuing the networks that follow.  finish() joins it."""

    def __init__(self, total, body_parts, spacing_xy, l3_label: int | None = None):
        import threading

        import torch
        from .labels import class_map
        if l3_label is None:
            l3_label = {v: k for k, v in class_map("total").items()}["vertebrae_L3"]
        self.spacing_xy, self.result = spacing_xy, (None, None)
        present = (total == l3_label).flatten(1).any(dim=1)
        rank = present.cumsum(0)
        n = rank[-1]
        lo = torch.argmax(((rank == (n - 1) // 2 + 1) & present).to(torch.uint8))
        hi = torch.argmax(((rank == n // 2 + 1) & present).to(torch.uint8))
        z = (lo + hi) // 2
        mask = body_parts.index_select(0, z.view(1))[0] == 1
        key = (str(total.device), tuple(mask.shape))
        if key not in _PINNED:
            _PINNED[key] = (torch.empty(mask.shape, dtype=torch.bool, pin_memory=True),
                            torch.empty(1, dtype=torch.int64, pin_memory=True))
        self._mask, self._n = _PINNED[key]
        self._mask.copy_(mask, non_blocking=True)
        self._n.copy_(n.view(1), non_blocking=True)
        self._event = torch.cuda.Event()
        self._event.record(torch.cuda.current_stream(total.device))
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def _run(self):
        self._event.synchronize()
        if int(self._n[0]) > 0:
            self.result = axes_of_slice(self._mask.numpy(), self.spacing_xy)

    def finish(self):
        self._thread.join()
        return self.result
