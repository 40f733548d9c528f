"""Body-composition numerics on the GPU: drop-in for
  subclassify_tissues                     (_external/body_composition_analysis/tissue/subclassification.py:10-63)
  AggregatableBodyPart.from_body_regions  (_external/body_composition_analysis/report/builder.py:45-112)
  create_vertebrae_info                   (_external/body_composition_analysis/commands.py:24-45)
  Builder.prepare / generate_aggregated_measurements / _descriptive_statistics_from_measurements /
  generate_secondary_findings (volumes) / create_json   (builder.py:163-361,397-444,520-598)
The volumes are touched once by boa_tissue_subclassify and four boa_slice_label_stats passes; everything the report
needs is then a function of the small per-slice tables [Z, L] (integer counts and HU sums), evaluated on the host.
Plots / PDF are out of scope; the breast-implant finding labels the (small) implant mask on the host.
"""
from __future__ import annotations

import enum
from typing import Any

import numpy as np
import torch

from . import passes
from .labels import BODY_PART_TORSO, BODY_REGION, TISSUES, class_map

TISSUE_COLUMNS = ["Bone", "Muscle", "TAT", "IMAT", "SAT", "VAT", "PAT", "EAT"]
_COL_TO_TISSUE = {"Bone": "BONE", "Muscle": "MUSCLE", "IMAT": "IMAT", "SAT": "SAT", "VAT": "VAT", "PAT": "PAT",
                  "EAT": "EAT"}
_ADIPOSE = ["IMAT", "SAT", "VAT", "PAT", "EAT"]


class AggregatableBodyPart(enum.IntFlag):
    NONE = 0
    ABDOMEN = 1
    THORAX = 2
    NECK = 4


def subclassify_tissues(ct: torch.Tensor, body_regions: torch.Tensor, median_filtering: bool = False) -> torch.Tensor:
    """subclassify_tissues (tissue/subclassification.py:10-63); median_filtering: the HU rules see the CT after an
    in-plane 3x3 median (:20-36; the arrays here are [z, y, x] with z the slice axis), the report keeps the original."""
    if median_filtering:
        ct = passes.median3x3_slices(ct)
    return passes.tissue_subclassify(ct, body_regions)


class SliceTables:
    """Everything the BCA report reads from the four volumes, reduced to per-slice tables."""

    def __init__(self, ct, tissues, body_parts, body_regions, total=None):
        self.Z = int(tissues.shape[0])
        c, s = passes.slice_label_stats(tissues, 8, ct=ct)
        self.tissue_counts, self.tissue_hu = c.cpu().numpy(), s.cpu().numpy()
        c, s = passes.slice_label_stats(tissues, 8, ct=ct, mask=body_parts, mask_value=BODY_PART_TORSO)
        self.tissue_counts_torso, self.tissue_hu_torso = c.cpu().numpy(), s.cpu().numpy()
        c, _ = passes.slice_label_stats(body_regions, 12)
        self.region_counts = c.cpu().numpy()
        self.total_counts = None
        if total is not None:
            c, _ = passes.slice_label_stats(total, 118)
            self.total_counts = c.cpu().numpy()


def slice_tables_from_arrays(Z, tissue_counts, tissue_hu, tissue_counts_torso, tissue_hu_torso, region_counts,
                             total_counts=None) -> SliceTables:
    """Build the tables from host arrays (multi-GPU merge of per-rank tables, tests)."""
    t = SliceTables.__new__(SliceTables)
    t.Z = int(Z)
    t.tissue_counts, t.tissue_hu = np.asarray(tissue_counts), np.asarray(tissue_hu)
    t.tissue_counts_torso, t.tissue_hu_torso = np.asarray(tissue_counts_torso), np.asarray(tissue_hu_torso)
    t.region_counts = np.asarray(region_counts)
    t.total_counts = None if total_counts is None else np.asarray(total_counts)
    return t


def _presence(counts: np.ndarray, ids) -> np.ndarray:
    return np.where(counts[:, list(ids)].sum(axis=1) > 0)[0]


def body_part_from_regions(t: SliceTables, slice_thickness: float, min_abdomen_length=200.0, min_neck_length=100.0,
                           min_thorax_length=200.0) -> AggregatableBodyPart:
    R = BODY_REGION
    result = AggregatableBodyPart.NONE
    abd = _presence(t.region_counts, [R["ABDOMINAL_CAVITY"]])
    n_abd = abd.max() - abd.min() + 1 if abd.size else 0
    if n_abd * slice_thickness >= min_abdomen_length:
        result |= AggregatableBodyPart.ABDOMEN
    med = _presence(t.region_counts, [R["MEDIASTINUM"]])
    above = t.Z - med.max() if med.size else 0
    if above * slice_thickness >= min_neck_length:
        result |= AggregatableBodyPart.NECK
    thorax_ids = [R["THORACIC_CAVITY"], R["MEDIASTINUM"], R["PERICARDIUM"]]
    tho = _presence(t.region_counts, thorax_ids)
    inter = np.intersect1d(abd, tho).size > 0
    n_tho = tho.max() - tho.min() + 1 if tho.size else 0
    if inter and n_tho * slice_thickness >= min_thorax_length:
        result |= AggregatableBodyPart.THORAX
    return result


def create_vertebrae_info(t: SliceTables, detected: AggregatableBodyPart) -> dict[str, tuple[int, int]]:
    vmap = {v.removeprefix("vertebrae_"): k for k, v in class_map("total").items() if v.startswith("vertebrae_")}
    info = {}
    for vid, label in vmap.items():
        sl = _presence(t.total_counts, [label])
        if len(sl) == 0:
            continue
        if (("C" in vid and AggregatableBodyPart.NECK not in detected)
                or ("T" in vid and AggregatableBodyPart.THORAX not in detected)
                or ("L" in vid and AggregatableBodyPart.ABDOMEN not in detected)):
            continue
        info[vid] = (int(sl.min()), int(sl.max() + 1))
    return info


def _slice_frame(counts: np.ndarray, ml_per_voxel: float) -> dict[str, np.ndarray]:
    cols = {c: counts[:, TISSUES[_COL_TO_TISSUE[c]]] * ml_per_voxel for c in TISSUE_COLUMNS if c != "TAT"}
    cols["TAT"] = cols["SAT"] + cols["VAT"] + cols["IMAT"] + cols["PAT"] + cols["EAT"]
    return {c: cols[c] for c in TISSUE_COLUMNS}


def _none(v):
    return None if v is None or (isinstance(v, float) and np.isnan(v)) else v


def _describe(frame, counts, hu, lo, hi) -> dict[str, dict[str, Any]]:
    """pandas describe() + Total + MeanHU of builder.py:257-307, as {column: {row: value}} with create_json's names.
    All eight columns at once (one numpy call per statistic: a report has ~30 groups x 2 tables)."""
    x = np.stack([frame[col][lo:hi] for col in TISSUE_COLUMNS], axis=1).astype(np.float64)  # [n, 8]
    n = x.shape[0]
    stats: dict[str, Any] = {}
    if n:
        stats["mean"] = x.mean(axis=0)
        stats["std"] = x.std(axis=0, ddof=1) if n > 1 else None
        stats["min"] = x.min(axis=0)
        q = np.percentile(x, (25, 50, 75), axis=0)
        stats["q1"], stats["q2"], stats["q3"] = q[0], q[1], q[2]
        stats["max"] = x.max(axis=0)
    total = x.sum(axis=0)
    csum, hsum = counts[lo:hi].sum(axis=0), hu[lo:hi].sum(axis=0)  # integer column sums per tissue id
    out = {}
    for j, col in enumerate(TISSUE_COLUMNS):
        d: dict[str, Any] = {}
        for k in ("mean", "std", "min", "q1", "q2", "q3", "max"):
            v = stats.get(k) if n else None
            d[k] = None if v is None else float(v[j])
        d["sum"] = float(total[j])
        ids = [TISSUES[a] for a in _ADIPOSE] if col == "TAT" else [TISSUES[_COL_TO_TISSUE[col]]]
        cnt = int(csum[ids].sum())
        d["mean_hu"] = float(int(hsum[ids].sum()) / cnt) if cnt else None
        out[col.lower()] = {k: _none(v) for k, v in d.items()}
    return out


def aggregation_groups(t: SliceTables, examined: AggregatableBodyPart, vertebrae) -> list[tuple[str, int, int]]:
    R = BODY_REGION
    groups = [("Whole Scan", 0, t.Z)]

    def span(ids):
        s = _presence(t.region_counts, ids)
        return int(s.min()), int(s.max() + 1)

    if AggregatableBodyPart.ABDOMEN in examined:
        groups.append(("Abdominal Cavity", *span([R["ABDOMINAL_CAVITY"]])))
    if AggregatableBodyPart.THORAX in examined:
        groups.append(("Thoracic Cavity", *span([R["THORACIC_CAVITY"], R["MEDIASTINUM"], R["PERICARDIUM"]])))
        groups.append(("Mediastinum", *span([R["MEDIASTINUM"]])))
        groups.append(("Pericardium", *span([R["PERICARDIUM"]])))
    if AggregatableBodyPart.ABDOMEN in examined and AggregatableBodyPart.THORAX in examined:
        groups.insert(1, ("Ventral Cavity", groups[1][1], groups[2][2]))
    if vertebrae:
        for name, g in vertebrae.items():
            groups.append((name, g[0], g[1]))
    return groups


def _pretty_volume(value: float) -> str:
    return f"{value / 1000:.3f} L" if value >= 1000 else f"{value:.2f} mL"


def breast_implant_finding(body_regions, ml_per_voxel: float, slice_weights=None) -> str | None:
    """builder.py:363-395: 26-connected components of the BREAST_IMPLANT label larger than 10 ml, ordered by the integer
    part of their centroid along the last array axis; one or two of them make a sentence (side = centroid against the
    middle of array axis 1, as the reference compares them), more are an error.
    On the device nothing but the few numbers of the sentence goes to the host: boa_cc_filter labels the components
    (root = first voxel in raster order, the order skimage numbers them in) and counts their voxels; the components
    above 10 ml - normally none, at most two - get their centroid from a per-column count of their voxels.
    slice_weights (device path): the map is the 5 mm-grid map and slice z stands for slice_weights[z] replicated slices
    of the input grid (postprocess.slice_weights) - components map one to one, voxel counts and column counts are
    weighted, 3.3x less labelling work.  The host path (numpy label map) is scipy.ndimage.label + two bincounts."""
    R = BODY_REGION["BREAST_IMPLANT"]
    # largest voxel count whose volume is NOT above 10 ml, with the reference's float comparison
    small = int(10.0 / ml_per_voxel)
    while (small + 1) * ml_per_voxel <= 10:
        small += 1
    while small > 0 and small * ml_per_voxel > 10:
        small -= 1
    mid_index = int(body_regions.shape[1]) // 2
    if getattr(body_regions, "is_cuda", False):
        from . import passes
        from .postprocess import MODE_26, OP_REMOVE_SMALL, _cc_filter, _Scratch
        mask_d = passes.label_set_mask(body_regions, [R])
        scratch = _Scratch(mask_d, need_border=False)
        _cc_filter(mask_d, (1,), False, MODE_26, OP_REMOVE_SMALL, small, 0, slice_weights, scratch)
        roots = torch.nonzero(scratch.sizes > small).flatten()[:3].tolist()  # sizes are non-zero at roots only
        props = []
        if len(roots) <= 2:
            cols = torch.arange(mask_d.shape[2], dtype=torch.float64, device=mask_d.device)
            for r in roots:
                per_col = (scratch.labels.view(mask_d.shape) == r).sum(dim=1).to(torch.float64)  # [z, x]
                if slice_weights is not None:
                    per_col = per_col * slice_weights.view(-1, 1)
                per_col = per_col.sum(dim=0)
                area = float(per_col.sum())
                props.append((float((per_col * cols).sum()) / area, area * ml_per_voxel))
        else:
            props = [(0.0, 0.0)] * 3
    else:
        from scipy import ndimage
        mask = np.asarray(body_regions) == R
        if not mask.any():
            return None
        lab, n = ndimage.label(mask, structure=np.ones((3, 3, 3)))
        flat = lab.ravel()
        area = np.bincount(flat, minlength=n + 1)
        xs = np.broadcast_to(np.arange(mask.shape[2], dtype=np.float64), mask.shape).ravel()
        xsum = np.bincount(flat, weights=xs, minlength=n + 1)
        props = [(xsum[i] / area[i], area[i] * ml_per_voxel) for i in range(1, n + 1) if area[i] * ml_per_voxel > 10]
    props.sort(key=lambda p: int(p[0]))
    found = [("right" if x < mid_index else "left", v) for x, v in props]
    if len(found) == 1:
        return (f"Patient has a single breast implant on the {found[0][0]} side with volume of "
                f"{_pretty_volume(found[0][1])}")
    if len(found) == 2:
        return (f"Patient has two breast implants with volume of {_pretty_volume(found[0][1])} ({found[0][0]}) and "
                f"{_pretty_volume(found[1][1])} ({found[1][0]})")
    if len(found) > 2:
        import logging
        logging.getLogger(__name__).error("More than two breast implant segments found")
    return None


def secondary_findings(t: SliceTables, examined: AggregatableBodyPart, ml_per_voxel: float,
                       body_regions=None, slice_weights=None) -> list[str]:
    """generate_secondary_findings (builder.py:309-395).  body_regions (the label map, or the 5 mm-grid map with its
    slice weights): also look for breast implants."""
    R = BODY_REGION
    tot = t.region_counts.sum(axis=0)
    out = []
    if AggregatableBodyPart.ABDOMEN in examined:
        out.append(f"Total volume of the abdominal cavity is {_pretty_volume(tot[R['ABDOMINAL_CAVITY']] * ml_per_voxel)}")
    if AggregatableBodyPart.THORAX in examined:
        v = (tot[R["THORACIC_CAVITY"]] + tot[R["MEDIASTINUM"]] + tot[R["PERICARDIUM"]]) * ml_per_voxel
        out.append(f"Volume of thoracic cavity is {_pretty_volume(v)}")
        v = (tot[R["MEDIASTINUM"]] + tot[R["PERICARDIUM"]]) * ml_per_voxel
        out.append(f"Volume of mediastinum is {_pretty_volume(v)}")
        out.append(f"Volume enclosed by the pericardial sack is {_pretty_volume(tot[R['PERICARDIUM']] * ml_per_voxel)}")
        if body_regions is not None and tot[R["BREAST_IMPLANT"]] * ml_per_voxel > 10:  # else no component can be
            sentence = breast_implant_finding(body_regions, ml_per_voxel, slice_weights)
            if sentence:
                out.append(sentence)
    return out


def build_bca_measurements(ct, tissues, body_parts, body_regions, total, spacing,
                           examined_body_region: str | None = None, tables: SliceTables | None = None):
    """-> (bca-measurements.json dict, vertebrae.json dict, SliceTables).  Volumes: [z,y,x] tensors on the device
    (already re-oriented as the reference's process_image does); spacing = (sx, sy, sz)."""
    t = tables if tables is not None else SliceTables(ct, tissues, body_parts, body_regions, total)
    ml_per_voxel = float(np.prod(spacing) / 1000.0)
    examined = (AggregatableBodyPart[examined_body_region.upper()] if examined_body_region
                else body_part_from_regions(t, float(spacing[2])))
    vertebrae = create_vertebrae_info(t, examined) if t.total_counts is not None else {}
    frame = _slice_frame(t.tissue_counts, ml_per_voxel)
    frame_nl = _slice_frame(t.tissue_counts_torso, ml_per_voxel)

    def records(fr):
        return [{c.lower(): float(fr[c][z]) for c in TISSUE_COLUMNS} for z in range(t.Z)]

    aggregated = {}
    for name, lo, hi in aggregation_groups(t, examined, vertebrae):
        aggregated[name.lower().replace(" ", "_").replace("-", "_")] = {
            "num_slices": int(hi - lo), "min_slice_idx": int(lo), "max_slice_idx": int(hi),
            "measurements": _describe(frame, t.tissue_counts, t.tissue_hu, lo, hi),
            "measurements_no_extremities": _describe(frame_nl, t.tissue_counts_torso, t.tissue_hu_torso, lo, hi),
        }
    data = {
        "slices": records(frame), "slices_no_extremities": records(frame_nl), "aggregated": aggregated,
        "body_parts": {"abdomen": AggregatableBodyPart.ABDOMEN in examined,
                       "neck": AggregatableBodyPart.NECK in examined,
                       "thorax": AggregatableBodyPart.THORAX in examined},
    }
    return data, vertebrae, t
