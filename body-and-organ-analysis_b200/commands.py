"""Orchestrator: mirror of analyze_ct (body_organ_analysis/commands.py:73-288) and compute_all_models
(body_organ_analysis/compute/inference.py:50-143) for the NIfTI fast path and the models on the hot path.
Output-directory contract (README.md:239-257, tests/test_generated_files.py:42-57): total.nii.gz, body_parts.nii.gz,
body_regions.nii.gz, tissues.nii.gz, ct_pfav.nii.gz, total-measurements.json, bca-measurements.json, vertebrae.json,
debug_information.txt.  DICOM conversion, contrast prediction, Excel and PDF reporting are out of scope."""
from __future__ import annotations

import json
import logging
import time
from pathlib import Path

import numpy as np
import torch

from . import nifti
from .config import IMPLEMENTED_MODELS
from .labels import class_map
from .pipeline import ModelZoo, analyze_volume

logger = logging.getLogger(__name__)


def range_warning(ct: np.ndarray) -> None:
    """compute/inference.py:21-30."""
    lo, hi = int(ct.min()), int(ct.max())
    if lo < -1024 or hi > 3071:
        logger.warning("The CT has HU values outside of the usual range [-1024, 3071]: [%s, %s]", lo, hi)


def to_int16_hu(data: np.ndarray) -> np.ndarray:
    """The device pipeline keeps the CT as int16 HU (integer HU make every statistic exact).  Float data are truncated
    like the reference's `.astype(np.int32)` after resampling (totalsegmentator/resampling.py:54); values outside the
    int16 range would wrap silently, so they are clipped with a warning instead."""
    if data.dtype == np.int16:
        return np.ascontiguousarray(data)
    if data.dtype.kind == "f":
        if not np.isfinite(data).all():
            raise ValueError("the CT contains NaN / inf values")
        data = np.trunc(data)
    lo, hi = float(data.min()), float(data.max())
    if lo < -32768 or hi > 32767:
        logger.warning("HU values outside the int16 range [%s, %s] are clipped to [-32768, 32767]", lo, hi)
        data = np.clip(data, -32768, 32767)
    return np.ascontiguousarray(data.astype(np.int16))


def analyze_ct(input_folder: Path, processed_output_folder: Path, excel_output_folder: Path | None = None,
               models=("total", "bca"), fast_bca: bool = False, fast_total: bool = False,
               bca_median_filtering: bool = False, cnr_adjustment: bool = True,
               device: str = "gpu",
               recompute: bool = True, weights_root: str | None = None, zoo: ModelZoo | None = None, **_ignored):
    """NIfTI in -> segmentations + measurement JSONs out.  Returns (output folder, stats dict) like the reference."""
    start = time.time()
    models = set(models)
    todo = models - IMPLEMENTED_MODELS
    if todo:
        raise NotImplementedError(f"models {sorted(todo)} are not implemented (licence-only TotalSegmentator tasks)")
    dev_str, _, gpu_id = device.partition(":")
    if dev_str != "gpu":
        raise RuntimeError(f"device '{device}': boa_b200 has no CPU/MPS implementation, use -d gpu[:id]")
    dev = torch.device("cuda", int(gpu_id) if gpu_id else 0)
    out_dir = Path(processed_output_folder)
    out_dir.mkdir(parents=True, exist_ok=True)
    img = nifti.load(input_folder)
    data, zooms, order = nifti.to_canonical(img.data, img.affine)
    range_warning(data)  # on the data as loaded, before the conversion below (compute/inference.py:21-30,46)
    ct_np = to_int16_hu(data)
    stats = {"num_voxels": int(ct_np.size), "num_slices": int(ct_np.shape[0])}
    zoo = zoo or ModelZoo(weights_root, device=dev)
    t0 = time.time()
    ct = torch.from_numpy(np.array(ct_np, copy=True)).pin_memory().to(dev, non_blocking=True)
    # recompute=False: label maps that already exist in the output folder are loaded instead of computed, and an
    # existing total-measurements.json is kept (compute/inference.py:82-84,95-105; infer/infer.py:59-61)
    precomputed, keep_total_json = {}, False
    if not recompute:
        from .labels import CROP_TASKS
        for name in ("total", "body_parts", "body_regions", *[m for m in CROP_TASKS if m in models]):
            f = out_dir / f"{name}.nii.gz"
            if f.is_file():
                logger.info("Loading already computed %s...", name)
                prev = nifti.load(f)
                arr, _, _ = nifti.to_canonical(prev.data, prev.affine)
                if arr.shape != ct_np.shape:
                    raise ValueError(f"{f} has shape {arr.shape}, the CT {ct_np.shape}: use --force-recompute")
                precomputed[name] = torch.from_numpy(np.ascontiguousarray(arr.astype(np.uint8)))
        keep_total_json = (out_dir / "total-measurements.json").is_file() and (out_dir / "ct_pfav.nii.gz").is_file()
        if keep_total_json:
            logger.info("The total measurements were already computed, skipping...")
    res = analyze_volume(ct, (zooms[2], zooms[1], zooms[0]), zoo, models=tuple(models), fast_bca=fast_bca,
                         fast_total=fast_total, cnr_adjustment=cnr_adjustment,
                         median_filtering=bca_median_filtering, precomputed=precomputed,
                         total_measurements=not keep_total_json)
    stats["inference_time"] = time.time() - t0

    def write(name, tensor, labels=None):
        arr = nifti.from_canonical(tensor.cpu().numpy(), order)
        nifti.save(out_dir / f"{name}.nii.gz", arr, img.affine, labels)

    if res.total is not None:
        if "total" not in precomputed:
            write("total", res.total, class_map("total"))
        if res.total_measurements is not None:
            write("ct_pfav", res.ct_pfav)
            with (out_dir / "total-measurements.json").open("w") as f:
                json.dump(res.total_measurements, f, indent=2)
    for name, lab in res.extra.items():
        if name not in precomputed:
            write(name, lab, class_map(name))
    if res.body_parts is not None and "body_parts" not in precomputed:
        write("body_parts", res.body_parts, class_map("body_parts"))
    if res.body_regions is not None and "body_regions" not in precomputed:
        write("body_regions", res.body_regions, class_map("body_regions"))
    if res.tissues is not None:
        write("tissues", res.tissues)
        with (out_dir / "bca-measurements.json").open("w") as f:
            json.dump(res.bca_measurements, f, indent=2)
        if res.vertebrae:
            with (out_dir / "vertebrae.json").open("w") as f:
                json.dump(res.vertebrae, f, indent=2)
    if res.l3_axes_mm is not None and res.l3_axes_mm[0] is not None:
        stats["l3_major_axis_cm"], stats["l3_minor_axis_cm"] = res.l3_axes_mm[0] / 10, res.l3_axes_mm[1] / 10
    if res.other_findings:
        stats["other_findings"] = " | ".join(res.other_findings)
    stats["total_time"] = time.time() - start
    stats.update({f"gpu_{k}_seconds": v for k, v in res.timings.items()})
    with (out_dir / "debug_information.txt").open("w") as f:
        for k, v in stats.items():
            f.write(f"{k}: {v}\n")
    return out_dir, stats
