"""Oracle: `total-measurements.json` and `bca-measurements.json` numerics, restated with numpy / pandas on whole
volumes exactly as the reference computes them (TEST INFRASTRUCTURE, see __init__.py).

  compute_measurements / metrics_for_each_region / ct_pfav / autochthon_reference   compute/measurements.py:42-343
  AggregatableBodyPart.from_body_regions, Builder.prepare / generate_aggregated_measurements /
  _descriptive_statistics_from_measurements / create_json        _external/body_composition_analysis/report/builder.py
  create_vertebrae_info                                           _external/body_composition_analysis/commands.py:24-45
Pinned against the reference's own functions (run with a fake SimpleITK) by tests/golden/make_golden.py.
"""
from __future__ import annotations

import json
import os

import numpy as np
import pandas as pd

from . import passes as P

_TABLES = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "body-and-organ-analysis_b200",
                       "data", "class_maps.json")
LUNG_MASKS = ["lung_upper_lobe_left", "lung_lower_lobe_left", "lung_upper_lobe_right", "lung_middle_lobe_right",
              "lung_lower_lobe_right"]
REGION = {"SUBCUTANEOUS_TISSUE": 1, "MUSCLE": 2, "ABDOMINAL_CAVITY": 3, "THORACIC_CAVITY": 4, "BONE": 5, "GLANDS": 6,
          "PERICARDIUM": 7, "BREAST_IMPLANT": 8, "MEDIASTINUM": 9, "BRAIN": 10, "NERVOUS_SYSTEM": 11}
TISSUE = {"MUSCLE": 1, "BONE": 2, "SAT": 3, "VAT": 4, "IMAT": 5, "PAT": 6, "EAT": 7}


def label_map_for(model_name: str) -> dict:
    with open(_TABLES) as f:
        t = json.load(f)
    out = {}
    for task in t["class_map_order"]:
        for idx, name in t["class_map_all_keys"][task].items():
            key = f"{task}_{name}"
            if key.startswith(model_name) and not key.startswith(model_name + "_v2"):
                out[key[len(model_name) + 1:]] = int(idx)
    return out


def _metrics(ct, mask, am, asd, spacing, cnr_adjustment=False, region_name=""):
    if np.sum(mask) == 0:
        return {"present": False}
    if cnr_adjustment:
        if "autochthon" in region_name:
            mask = P.region_minus_fat(ct, mask)
        mask = P.erode_region(mask)
    if np.sum(mask) == 0:
        return {"present": False}
    m = P.metrics_for_region(ct, mask, am, asd, spacing)
    if cnr_adjustment and region_name.partition("_")[0] == "autochthon" and am is not None and asd is not None:
        m["cnr"] = None
    return m


def compute_measurements(ct: np.ndarray, total: np.ndarray, spacing, cnr_adjustment=False) -> dict:
    out = {"segmentations": {}, "info": {}}
    lm = label_map_for("total")
    aut = P.region_minus_fat(ct, np.logical_or(P.create_mask(total, lm["autochthon_right"]),
                                                P.create_mask(total, lm["autochthon_left"])))
    aut = P.erode_region(aut)
    am = asd = None
    if aut.sum() > 0:
        am, asd = float(np.mean(ct[aut])), float(np.std(ct[aut]))
    res = {}
    for region, label in lm.items():
        res[region] = _metrics(ct, P.create_mask(total, label), am, asd, spacing)
    res["autochthon"] = _metrics(ct, P.create_mask(total, [lm["autochthon_left"], lm["autochthon_right"]]), am, asd, spacing)

    def lung(names):
        mask = P.create_mask(total, [lm[n] for n in names])
        fat = np.logical_and(mask, np.logical_and(ct >= P.ADIPOSE_TISSUE[0], ct <= P.ADIPOSE_TISSUE[1]))
        return fat, _metrics(ct, fat, am, asd, spacing)

    for n in LUNG_MASKS:
        _, res["ct_pfav_" + n] = lung([n])
    for side in ("left", "right"):
        _, res[f"ct_pfav_lobe_{side}"] = lung([ll for ll in LUNG_MASKS if ll.endswith(side)])
    pfav, res["ct_pfav_lungs"] = lung(LUNG_MASKS)
    out["segmentations"]["total"] = res
    if cnr_adjustment and am is not None:
        sel = {r: v for r, v in lm.items() if r in {"aorta", "autochthon_left", "autochthon_right"}}
        adj = {r: _metrics(ct, P.create_mask(total, v), am, asd, spacing, True, r) for r, v in sel.items()}
        adj["autochthon"] = _metrics(ct, P.create_mask(total, [sel["autochthon_left"], sel["autochthon_right"]]), am, asd,
                                     spacing, True, "autochthon")
        out["cnr_adjusted"] = adj
    out["info"] = {"autochthon_mean": am, "autochthon_std": asd}
    out["_ct_pfav_mask"] = pfav.astype(np.uint8)
    return out


def body_part_flags(regions: np.ndarray, slice_thickness: float) -> dict:
    abd_mask = regions == REGION["ABDOMINAL_CAVITY"]
    abd = np.where(abd_mask.any(axis=(1, 2)))[0]
    n_abd = abd.max() - abd.min() + 1 if abd.size else 0
    med = np.where((regions == REGION["MEDIASTINUM"]).any(axis=(1, 2)))[0]
    above = regions.shape[0] - med.max() if med.size else 0
    tho_mask = np.isin(regions, [REGION["THORACIC_CAVITY"], REGION["MEDIASTINUM"], REGION["PERICARDIUM"]])
    tho = np.where(tho_mask.any(axis=(1, 2)))[0]
    inter = np.logical_and(abd_mask.any(axis=(1, 2)), tho_mask.any(axis=(1, 2))).any()
    n_tho = tho.max() - tho.min() + 1 if tho.size else 0
    return {"abdomen": bool(n_abd * slice_thickness >= 200), "neck": bool(above * slice_thickness >= 100),
            "thorax": bool(inter and n_tho * slice_thickness >= 200)}


def vertebrae_info(total: np.ndarray, flags: dict) -> dict:
    with open(_TABLES) as f:
        cm = json.load(f)["class_map"]["total"]
    vmap = {v.removeprefix("vertebrae_"): int(k) for k, v in cm.items() if v.startswith("vertebrae_")}
    info = {}
    for vid, label in vmap.items():
        sl = np.where((total == label).any(axis=(1, 2)))[0]
        if len(sl) == 0:
            continue
        if ("C" in vid and not flags["neck"]) or ("T" in vid and not flags["thorax"]) or ("L" in vid and not flags["abdomen"]):
            continue
        info[vid] = (int(sl.min()), int(sl.max() + 1))
    return info


def _frame(tissues, ml, mask=None):
    names = {"MUSCLE": "Muscle", "BONE": "Bone"}
    data = {}
    for t, v in TISSUE.items():
        m = tissues == v
        if mask is not None:
            m = np.logical_and(mask, m)
        data[names.get(t, t)] = m.sum(axis=(1, 2)) * ml
    df = pd.DataFrame(data)
    df["TAT"] = df.SAT + df.VAT + df.IMAT + df.PAT + df.EAT
    df["slice_idx"] = range(len(df))
    return df[["slice_idx", "Bone", "Muscle", "TAT", "IMAT", "SAT", "VAT", "PAT", "EAT"]]


def _describe(df, image, tissue):
    df = df.drop("slice_idx", axis=1)
    m = df.describe()
    m.drop("count", inplace=True)
    m.index = ["Mean", "StdDev", "Minimum", "25%", "Median", "75%", "Maximum"]
    m.loc["Total"] = df.sum()
    names = {"MUSCLE": "Muscle", "BONE": "Bone"}
    for t, v in TISSUE.items():
        d = image[tissue == v]
        m.loc["MeanHU", names.get(t, t)] = np.mean(d) if d.size else None
    d = image[np.isin(tissue, [TISSUE[k] for k in ("IMAT", "SAT", "VAT", "PAT", "EAT")])]
    m.loc["MeanHU", "TAT"] = np.mean(d) if d.size else None
    return m.replace({np.nan: None})


def bca_json(ct, tissues, parts, regions, total, spacing) -> tuple[dict, dict]:
    ml = np.prod(spacing) / 1000.0
    flags = body_part_flags(regions, spacing[2])
    vert = vertebrae_info(total, flags) if total is not None else {}
    torso = parts == 1
    df, dfn = _frame(tissues, ml), _frame(tissues, ml, torso)
    groups = [("Whole Scan", 0, ct.shape[0])]

    def span(ids):
        s = np.where(np.isin(regions, ids).any(axis=(1, 2)))[0]
        return s.min(), s.max() + 1

    if flags["abdomen"]:
        groups.append(("Abdominal Cavity", *span([3])))
    if flags["thorax"]:
        groups.append(("Thoracic Cavity", *span([4, 9, 7])))
        groups.append(("Mediastinum", *span([9])))
        groups.append(("Pericardium", *span([7])))
    if flags["abdomen"] and flags["thorax"]:
        groups.insert(1, ("Ventral Cavity", groups[1][1], groups[2][2]))
    for name, g in vert.items():
        groups.append((name, g[0], g[1]))
    ren = {"Mean": "mean", "StdDev": "std", "Minimum": "min", "25%": "q1", "Median": "q2", "75%": "q3",
           "Maximum": "max", "Total": "sum", "MeanHU": "mean_hu"}
    agg = {}
    for name, lo, hi in groups:
        a = _describe(df[(df.slice_idx >= lo) & (df.slice_idx < hi)], ct[lo:hi], tissues[lo:hi])
        b = _describe(dfn[(dfn.slice_idx >= lo) & (dfn.slice_idx < hi)], ct[lo:hi],
                      np.where(torso[lo:hi], tissues[lo:hi], 0))
        agg[name.lower().replace(" ", "_").replace("-", "_")] = {
            "num_slices": int(hi - lo), "min_slice_idx": int(lo), "max_slice_idx": int(hi),
            "measurements": a.rename(index=ren, columns={x: x.lower() for x in a.columns}).to_dict(),
            "measurements_no_extremities": b.rename(index=ren, columns={x: x.lower() for x in b.columns}).to_dict()}

    def rec(d):
        return d.rename(columns={x: x.lower() for x in d.columns}).drop("slice_idx", axis=1).astype(float).to_dict("records")

    return {"slices": rec(df), "slices_no_extremities": rec(dfn), "aggregated": agg, "body_parts": flags}, vert
