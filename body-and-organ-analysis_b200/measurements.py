"""`total-measurements.json` numerics on the GPU: drop-in for compute_measurements / metrics_for_each_region / ct_pfav /
autochthon_reference (compute/measurements.py:22-343).

One pass over (CT, label map) builds a per-label integer-HU histogram (boa_label_hu_hist); every statistic the
reference computes per label with boolean masks, gathers and sorts - count, mean, std, min, median, max, 25th / 75th
percentile, the lung-fat-window subsets and the label unions - is an exact function of those histograms, evaluated on
the host in float64 (a few KB).  The autochthon reference and the optional CNR adjustment need an eroded mask:
label-set mask -> 6^3 box erosion -> histogram, all on the device.
"""
from __future__ import annotations

from typing import Any

import numpy as np
import torch

from . import passes
from .labels import measurement_label_map

ADIPOSE_TISSUE = (-200, -40)
CNR_ADJUSTED_REGIONS = {"total": {"aorta", "autochthon_left", "autochthon_right"},
                        "heartchambers_highres": {"pulmonary_artery"}}
LUNG_MASKS = ["lung_upper_lobe_left", "lung_lower_lobe_left", "lung_upper_lobe_right", "lung_middle_lobe_right",
              "lung_lower_lobe_right"]
HU_MIN, N_BINS = -32768, 65536  # the whole int16 range: no voxel can fall outside


def _order_stat(cum, values, k):
    return float(values[np.searchsorted(cum, k, side="right")])


def _percentile(cum, values, n, q):
    pos = (n - 1) * q / 100.0  # numpy method="linear"
    lo, hi = int(np.floor(pos)), int(np.ceil(pos))
    a, b = _order_stat(cum, values, lo), _order_stat(cum, values, hi)
    t = pos - lo
    return float(b - (b - a) * (1 - t)) if t >= 0.5 else float(a + (b - a) * t)


def metrics_from_hist(hist: np.ndarray, hu_min: int, autochthon_mean, autochthon_std, img_spacing,
                      cnr_none: bool = False) -> dict[str, Any]:
    """metrics_for_region (measurements.py:74-123) from hist[i] = #voxels with HU == hu_min + i."""
    hist = hist.astype(np.int64)
    n = int(hist.sum())
    if n == 0:
        return {"present": False}
    nz = np.nonzero(hist)[0]
    lo, hi = int(nz[0]), int(nz[-1]) + 1
    h = hist[lo:hi]
    values = np.arange(hu_min + lo, hu_min + hi, dtype=np.int64)
    cum = np.cumsum(h)
    mean = float(int((h * values).sum()) / n)
    var = float((h * (values - mean) ** 2).sum() / n)
    m: dict[str, Any] = {"present": True}
    m["volume_ml"] = n * (np.prod(img_spacing) / 1000.0)
    m["mean_hu"] = mean
    m["std_hu"] = float(np.sqrt(var))
    m["min_hu"] = float(values[0])
    m["median_hu"] = _percentile(cum, values, n, 50)
    m["max_hu"] = float(values[-1])
    for p in (25, 75):
        m[f"{p}th_percentile_hu"] = _percentile(cum, values, n, p)
    if autochthon_mean is not None and autochthon_std is not None and not cnr_none:
        m["cnr"] = (mean - autochthon_mean) / autochthon_std
    else:
        m["cnr"] = None
    return m


def _window(hist_row: np.ndarray, lo: int, hi: int, hu_min: int = HU_MIN) -> np.ndarray:
    """hist_row restricted to lo <= HU <= hi (hist_row[i] counts HU == hu_min + i)."""
    out = np.zeros_like(hist_row)
    a, b = max(lo - hu_min, 0), min(hi - hu_min + 1, hist_row.shape[0])
    if b > a:
        out[a:b] = hist_row[a:b]
    return out


def _to_host_async(t: torch.Tensor) -> torch.Tensor:
    """Device -> pinned host copy on the current stream (no host synchronisation; see PendingMeasurements)."""
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t, non_blocking=True)
    return h


def _masked_hist_dev(ct: torch.Tensor, mask: torch.Tensor, hu_lo: int, n_bins: int) -> torch.Tensor:
    hist, _ = passes.label_hu_hist(ct, mask, 2, hu_lo, n_bins)
    return hist[1]


def _eroded_region_hist_dev(ct, labels, ids, minus_fat: bool, hu_lo: int, n_bins: int, slab=None) -> torch.Tensor:
    """Histogram of the CT under erode_6(label set [minus fat]).  slab = (z0, z1): only the slices z0..z1 of the volume
    contribute (multi-GPU: every rank takes its dim-0 slab and the tables are all-reduced); the erosion window reaches 3
    slices back and 2 forward, so the mask is built on the slab plus that halo and the halo rows are dropped."""
    if slab is None:
        mask = passes.label_set_mask(labels, ids, ct, ADIPOSE_TISSUE[0], ADIPOSE_TISSUE[1], 2 if minus_fat else 0)
        return _masked_hist_dev(ct, passes.erode_box(mask, 3, 2), hu_lo, n_bins)
    z0, z1 = slab
    lo, hi = max(0, z0 - 3), min(int(ct.shape[0]), z1 + 2)
    cts, labs = ct[lo:hi], labels[lo:hi]
    mask = passes.label_set_mask(labs, ids, cts, ADIPOSE_TISSUE[0], ADIPOSE_TISSUE[1], 2 if minus_fat else 0)
    # kept rows z0..z1 look at z-3..z+2, i.e. never beyond [lo, hi) unless that is the image border itself, where
    # erode_region's "outside counts as foreground" is exactly what erode_box does
    er = passes.erode_box(mask, 3, 2)
    return _masked_hist_dev(ct[z0:z1], er[z0 - lo:z1 - lo].contiguous(), hu_lo, n_bins)


def _all_reduce_sum(t: torch.Tensor, dist_ctx) -> torch.Tensor:
    import torch.distributed as dist
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=dist_ctx.group)
    return t


class PendingMeasurements:
    """Device work of compute_measurements is enqueued (histograms, eroded-region histograms, the ct_pfav mask) and the
    small tables are on their way to pinned host memory; finish() waits for them and does the host arithmetic.  The
    pipeline calls finish() after the following networks are enqueued, so the ~300 per-name statistics (tens of ms of
    numpy) run while the GPU is busy instead of between two networks."""

    def __init__(self, spacing, cnr_adjustment: bool):
        self.spacing, self.cnr_adjustment = spacing, cnr_adjustment
        self.models: list = []      # (model_name, label_map, hu_lo, hist_host, aut_hist_host | None, {region: hist_host})
        self.pfav_mask = None
        self.event = None

    def finish(self):
        if self.event is not None:
            self.event.synchronize()
        spacing = self.spacing
        measurements: dict[str, Any] = {"segmentations": {}, "info": {}}
        aut_mean = aut_std = None
        for model_name, label_map, hu0, hist_h, aut_h, adj_h in self.models:
            hist = hist_h.numpy().view(np.uint32)
            if aut_h is not None:
                h = aut_h.numpy().view(np.uint32)
                if h.sum() > 0:
                    ref = metrics_from_hist(h, hu0, None, None, spacing)
                    aut_mean, aut_std = ref["mean_hu"], ref["std_hu"]
            res, per_label = {}, {}
            for region, label in label_map.items():
                # ~60 % of the names are aliases of a label id that was evaluated already (total_v1_*, total_mr_*, ...)
                if label not in per_label:
                    per_label[label] = metrics_from_hist(hist[label], hu0, aut_mean, aut_std, spacing)
                res[region] = dict(per_label[label])
            if "autochthon_left" in label_map and "autochthon_right" in label_map:
                union = hist[label_map["autochthon_left"]].astype(np.int64) + hist[label_map["autochthon_right"]]
                res["autochthon"] = metrics_from_hist(union, hu0, aut_mean, aut_std, spacing)
            if model_name == "total":
                def lung(names):
                    u = np.zeros(hist.shape[1], dtype=np.int64)
                    for nme in names:
                        u += _window(hist[label_map[nme]], *ADIPOSE_TISSUE, hu_min=hu0)
                    return metrics_from_hist(u, hu0, aut_mean, aut_std, spacing)
                for nme in LUNG_MASKS:
                    res["ct_pfav_" + nme] = lung([nme])
                for side in ("left", "right"):
                    res[f"ct_pfav_lobe_{side}"] = lung([ll for ll in LUNG_MASKS if ll.endswith(side)])
                res["ct_pfav_lungs"] = lung(LUNG_MASKS)
            measurements["segmentations"][model_name] = res
            if (self.cnr_adjustment and model_name in CNR_ADJUSTED_REGIONS and aut_mean is not None
                    and aut_std is not None):
                adj = {}
                sel = {r: v for r, v in label_map.items() if r in CNR_ADJUSTED_REGIONS[model_name]}
                for region, label in sel.items():
                    if hist[label].sum() == 0:
                        adj[region] = {"present": False}
                        continue
                    adj[region] = metrics_from_hist(adj_h[region].numpy().view(np.uint32), hu0, aut_mean, aut_std,
                                                    spacing, cnr_none=region.partition("_")[0] == "autochthon")
                if "autochthon_left" in sel and "autochthon_right" in sel:
                    if sum(int(hist[i].sum()) for i in (sel["autochthon_left"], sel["autochthon_right"])) == 0:
                        adj["autochthon"] = {"present": False}
                    else:
                        adj["autochthon"] = metrics_from_hist(aut_h.numpy().view(np.uint32), hu0, aut_mean, aut_std,
                                                              spacing, cnr_none=True)
                measurements.setdefault("cnr_adjusted", {}).update(adj)
        measurements["info"]["autochthon_mean"] = aut_mean
        measurements["info"]["autochthon_std"] = aut_std
        return measurements, self.pfav_mask


def enqueue_measurements(ct: torch.Tensor, segmentations: dict[str, torch.Tensor], spacing,
                         cnr_adjustment: bool = False, return_ct_pfav_mask: bool = False,
                         hu_range: tuple[int, int] | None = None, dist_ctx=None) -> PendingMeasurements:
    """Device half of compute_measurements (compute/measurements.py:244-343).  hu_range = (min, max) HU of the CT if
    the caller knows it (the histograms then cover exactly that range); None: read it here (one host sync).
    dist_ctx (every rank holds the whole CT and label map): each rank reduces its dim-0 slab and the integer tables are
    all-reduced - exact, and every rank ends up with the same tables."""
    from .geometry import shard_patches
    pend = PendingMeasurements(spacing, cnr_adjustment)
    if not segmentations:
        return pend
    if ct.dtype != torch.int16:
        raise TypeError("compute_measurements_on_device needs an int16 CT (exact integer-HU histograms)")
    if hu_range is None:
        lo_t, hi_t = torch.aminmax(ct)
        hu_range = (int(lo_t), int(hi_t))
    hu_lo, n_bins = int(hu_range[0]), int(hu_range[1]) - int(hu_range[0]) + 1
    sharded = dist_ctx is not None and dist_ctx.world_size > 1
    slab = shard_patches(int(ct.shape[0]), dist_ctx.world_size, dist_ctx.rank) if sharded else None
    red = (lambda t: _all_reduce_sum(t, dist_ctx)) if sharded else (lambda t: t)
    for model_name in sorted(segmentations, key=lambda m: m != "total"):
        labels = segmentations[model_name]
        if labels.shape != ct.shape:
            raise ValueError("The spacing of the image and of the segmentation should be the same")
        label_map = measurement_label_map(model_name)
        n_labels = max(label_map.values()) + 1
        if sharded:
            hist_dev, _ = passes.label_hu_hist(ct[slab[0]:slab[1]], labels[slab[0]:slab[1]], n_labels, hu_lo, n_bins)
        else:
            hist_dev, _ = passes.label_hu_hist(ct, labels, n_labels, hu_lo, n_bins)
        hist_h = _to_host_async(red(hist_dev))
        aut_h, adj_h = None, {}
        if model_name == "total":
            ids = [label_map["autochthon_right"], label_map["autochthon_left"]]
            # the same histogram serves the autochthon reference (:42-58) and the CNR-adjusted "autochthon" entry
            aut_h = _to_host_async(red(_eroded_region_hist_dev(ct, labels, ids, True, hu_lo, n_bins, slab).contiguous()))
            if return_ct_pfav_mask:
                pend.pfav_mask = passes.label_set_mask(labels, [label_map[ll] for ll in LUNG_MASKS], ct,
                                                       ADIPOSE_TISSUE[0], ADIPOSE_TISSUE[1], 1)
        if cnr_adjustment and model_name in CNR_ADJUSTED_REGIONS:
            for region, label in label_map.items():
                if region in CNR_ADJUSTED_REGIONS[model_name]:
                    adj_h[region] = _to_host_async(red(_eroded_region_hist_dev(
                        ct, labels, [label], "autochthon" in region, hu_lo, n_bins, slab).contiguous()))
        pend.models.append((model_name, label_map, hu_lo, hist_h, aut_h, adj_h))
    pend.event = torch.cuda.Event()
    pend.event.record()
    return pend


def compute_measurements_on_device(ct: torch.Tensor, segmentations: dict[str, torch.Tensor], spacing,
                                   cnr_adjustment: bool = False, return_ct_pfav_mask: bool = False,
                                   hu_range: tuple[int, int] | None = None):
    """ct int16 [z,y,x] on the device; segmentations: model name -> uint8 label map (same shape); spacing as
    SimpleITK's GetSpacing().  Returns the dict compute_measurements returns (+ the ct_pfav mask tensor on request)."""
    m, pfav = enqueue_measurements(ct, segmentations, spacing, cnr_adjustment, return_ct_pfav_mask, hu_range).finish()
    return (m, pfav) if return_ct_pfav_mask else m
