"""Volume-level pipeline on the device: preprocessing -> sliding-window networks -> label maps -> tissue map ->
measurement tables.  Host mirror of the per-volume work of

  nnUNet_predict_image            (_external/totalsegmentator/nnunet.py:326-829; `total` = 5 part models merged by LUT)
  DefaultPreprocessor.run_case_npy (_external/nnunetv2/preprocessing/preprocessors/default_preprocessor.py:45-118:
                                    crop_to_nonzero -> CTNormalization -> resample)
  export_prediction (un-crop)      (_external/nnunetv2/inference/export_prediction.py:45-47)
  body_composition_analysis.inference / run_pipeline (infer/infer.py:39-89, commands.py:84-170)
  compute_all_models               (compute/inference.py:50-143)

The CT goes to the device once (int16, 2 bytes / voxel); label maps come back as uint8 and the measurement tables as a
few KB.  Logits never leave HBM.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass, field

import numpy as np
import torch
from torch.cuda import nvtx  # NVTX ranges around the stages (visible in nsys / ncu --nvtx)

from . import bca, passes
from .dist import DistContext, exchange_slabs, gather_label_slabs, plan_shards
from .labels import (BODY_PARTS_TASK_ID, BODY_REGIONS_TASK_ID, CROP_ADDON_MM, CROP_PREPASS_TASK_ID, CROP_TASKS,
                     TOTAL_FAST_TASK_ID, TOTAL_TASK_IDS, class_map, part_luts)
from .measurements import compute_measurements_on_device, enqueue_measurements
from .plans import find_model_folder, load_model_folder
from .predictor import finalize_argmax, nnUNetPredictor, raise_if_nonfinite, weight_sum

TRAINERS = {**{t: "nnUNetTrainerNoMirroring" for t in TOTAL_TASK_IDS},
            TOTAL_FAST_TASK_ID: "nnUNetTrainer_4000epochs_NoMirroring",
            BODY_REGIONS_TASK_ID: "nnUNetTrainerNoMirroring",
            BODY_PARTS_TASK_ID: "nnUNetTrainer_1500epochs_NoMirroring",
            CROP_PREPASS_TASK_ID: "nnUNetTrainer_4000epochs_NoMirroring",
            **{tid: "nnUNetTrainer" for tid, _, _ in CROP_TASKS.values()}}


class ModelZoo:
    """Networks stay resident across tasks and volumes (the reference reloads every checkpoint for every volume,
    predict_from_raw_data.py:86-118); networks of one geometry share one activation workspace."""

    def __init__(self, weights_root: str | None = None, device=None, max_batch: int | None = None):
        self._specs = None
        self.root = weights_root or os.environ.get("TOTALSEG_WEIGHTS_PATH") or os.environ.get("nnUNet_results")
        if not self.root:
            raise RuntimeError("no weights directory: pass weights_root or set TOTALSEG_WEIGHTS_PATH")
        self.device, self.max_batch = device, max_batch
        self._cache: dict = {}
        self._donor: nnUNetPredictor | None = None

    @classmethod
    def from_specs(cls, specs: dict, device=None, max_batch: int | None = None) -> "ModelZoo":
        """In-memory zoo: task id -> ModelSpec holding ALL folds of that task (synthetic weights, tests, bench)."""
        zoo = cls.__new__(cls)
        zoo.root, zoo.device, zoo.max_batch = None, device, max_batch
        zoo._cache, zoo._donor, zoo._specs = {}, None, specs
        return zoo

    def get(self, task_id: int, folds, step_size: float) -> nnUNetPredictor:
        """folds None = every fold of the model (nnU-Net's auto-detection, predict_from_raw_data.py:131-140)."""
        key = (task_id, None if folds is None else tuple(folds), step_size)
        if key not in self._cache:
            p = nnUNetPredictor(tile_step_size=step_size, use_gaussian=True, use_mirroring=False,
                                perform_everything_on_device=True, device=self.device, max_batch=self.max_batch,
                                workspace_donor=self._donor)
            if getattr(self, "_specs", None) is not None:
                import copy
                spec = copy.copy(self._specs[task_id])
                if folds is not None:
                    spec.fold_weights = [self._specs[task_id].fold_weights[int(f)] for f in folds]
                p.manual_initialization(spec)
            else:
                p.initialize_from_trained_model_folder(find_model_folder(self.root, task_id, TRAINERS[task_id]), folds)
            self._donor = self._donor or p
            self._cache[key] = p
        return self._cache[key]


def nonzero_bbox(vol: torch.Tensor):
    """crop_to_nonzero (nnunetv2/preprocessing/cropping/cropping.py:6-39): bounding box of data != 0 (filling holes
    does not change a bounding box)."""
    nz = vol != 0
    box = []
    for ax in range(3):
        other = tuple(a for a in range(3) if a != ax)
        idx = torch.nonzero(nz.any(dim=other)).flatten()
        if idx.numel() == 0:
            return [(0, s) for s in vol.shape]
        box.append((int(idx[0]), int(idx[-1]) + 1))
    return box


def predict_labels_sharded(pred: nnUNetPredictor, data: torch.Tensor, lut=None, label_inout=None,
                           overwrite_nonzero_only=False, dist_ctx: DistContext | None = None,
                           defer: list | None = None) -> torch.Tensor:
    """predict_labels with the patches of the volume sharded over the ranks of dist_ctx (dist.py, NCCL path: one
    exchange, one finalize and one label all-gather per network)."""
    if dist_ctx is None or dist_ctx.world_size == 1:
        return pred.predict_labels(data, lut, label_inout, overwrite_nonzero_only, defer=defer)
    with torch.cuda.device(pred.device_index):
        vol, origins, unpad = pred._prepare(data)
        plan = plan_shards(origins, pred.patch_size[0], vol.shape[0], dist_ctx.world_size, dist_ctx.rank)
        acc = torch.zeros((pred.num_classes, plan.zhi - plan.zlo, *vol.shape[1:]), dtype=torch.float32,
                          device=pred.device)
        if plan.end > plan.begin:
            local = origins[plan.begin:plan.end].copy()
            local[:, 0] -= plan.zlo
            pred.accumulate(vol[plan.zlo:plan.zhi], local, acc)
        slab = exchange_slabs(acc, plan, dist_ctx)
        del acc
        w = weight_sum(vol.shape, pred.patch_size, origins, pred.gaussian(), pred.gaussian_kind)
        lo, hi = plan.slabs[dist_ctx.rank]
        lab_slab = finalize_argmax(slab, w[lo:hi], lut, defer=defer) if hi > lo else torch.zeros(
            (0, *vol.shape[1:]), dtype=torch.uint8, device=pred.device)
        lab = gather_label_slabs(lab_slab, plan, dist_ctx)[unpad].contiguous()
        if label_inout is None:
            return lab
        if overwrite_nonzero_only:
            label_inout.copy_(torch.where(lab != 0, lab, label_inout))
        else:
            label_inout.copy_(lab)
        return label_inout


def check_plan_geometry(spec, cropped_shape, spacing_zyx) -> None:
    """DefaultPreprocessor.run_case_npy applies plans.transpose_forward (only the identity is implemented: anything
    else raises instead of silently producing a transposed result) and resamples the cropped, normalised volume to the
    configuration's spacing (default_preprocessor.py:57-90).  Returns that target spacing (None: no spacing given, the
    caller vouches for the grid)."""
    if list(spec.transpose_forward) != [0, 1, 2] or list(spec.transpose_backward) != [0, 1, 2]:
        raise NotImplementedError(
            f"plans.json asks for transpose_forward={list(spec.transpose_forward)} / transpose_backward="
            f"{list(spec.transpose_backward)}: only the identity is implemented (every BOA model is trained with it)")
    if spacing_zyx is None:
        return None
    target = list(spec.spacing)
    if len(target) < 3:  # 2d configurations keep the slice spacing (default_preprocessor.py:72-75)
        target = [spacing_zyx[0]] + target
    # compute_new_shape (default_resampling.py:25-31); the reference resamples iff the shape changes (:139)
    return target


def _preprocess(ct: torch.Tensor, spec, spacing_zyx=None, box=None) -> tuple[torch.Tensor, list]:
    """crop_to_nonzero -> CTNormalization (fp32).  Returns the [1, z, y, x] network input and the crop box (pass `box`
    when it is known already: it depends on the volume only, and finding it costs three device -> host reads)."""
    if box is None:
        box = nonzero_bbox(ct)
    crop = ct[box[0][0]:box[0][1], box[1][0]:box[1][1], box[2][0]:box[2][1]].contiguous()
    target = check_plan_geometry(spec, crop.shape, spacing_zyx)
    p = spec.intensity
    norm = passes.ct_normalize(crop, float(p["percentile_00_5"]), float(p["percentile_99_5"]), float(p["mean"]),
                               float(p["std"]))
    if target is not None:
        # normalise first, then resample to the plan's spacing (default_preprocessor.py:76-90); identity when the
        # shape does not change (`total`: TotalSegmentator resampled to 1.5 mm already)
        from .resample import resample_to_plan_spacing
        norm = resample_to_plan_spacing(norm, spacing_zyx, target)
    return norm[None], box


def segment_task(ct: torch.Tensor, zoo: ModelZoo, task_ids, folds, step_size: float, luts=None,
                 dist_ctx: DistContext | None = None, force_split: bool = False, spacing_zyx=None) -> torch.Tensor:
    """One nnUNet_predict_image call on a volume that is already at the task's spacing: uint8 label map [z,y,x].
    Very large volumes (or force_split) are predicted as three overlapping z-parts and stitched, exactly where the
    reference does it (totalsegmentator/nnunet.py:483-505,582-586) - each part is preprocessed on its own."""
    from .geometry import needs_triple_split, triple_split_ranges
    if needs_triple_split(ct.shape, len(task_ids) > 1, force_split):
        out = torch.zeros(ct.shape, dtype=torch.uint8, device=ct.device)
        for plo, phi, klo, khi, dlo, dhi in triple_split_ranges(int(ct.shape[0])):
            part = _segment_task_whole(ct[plo:phi].contiguous(), zoo, task_ids, folds, step_size, luts, dist_ctx,
                                       spacing_zyx)
            out[dlo:dhi] = part[klo:khi]
        return out
    return _segment_task_whole(ct, zoo, task_ids, folds, step_size, luts, dist_ctx, spacing_zyx)


def _segment_task_peers(ct, zoo, task_ids, folds, step_size, luts, dist_ctx, spacing_zyx, ex) -> torch.Tensor | None:
    """Multi-GPU, peer-memory path (dist.PeerExchange): every network's patches go into this rank's private buffer; one
    fused kernel per network reduces this rank's dim-0 slab over NVLink, finalises it and merges the part labels into
    the rank's label slab; ONE all-gather of the uint8 slabs per task.  Returns None when the volume needs the
    general path (cropped by crop_to_nonzero, or smaller than the patch) - decided from the volume alone, so every rank
    decides the same."""
    box = nonzero_bbox(ct)
    if not all(b == 0 and e == s for (b, e), s in zip(box, ct.shape)):
        return None
    preds = [zoo.get(tid, folds, step_size) for tid in task_ids]
    if any(int(s) < int(p) for pr in preds for s, p in zip(ct.shape, pr.patch_size)):
        return None
    if any(_plan_shape(pr.spec, ct.shape, spacing_zyx) != tuple(ct.shape) for pr in preds):
        return None  # nnU-Net's own resampling to the plan's spacing: general path
    rank, world = dist_ctx.rank, dist_ctx.world_size
    Y, X = int(ct.shape[1]), int(ct.shape[2])
    plans = []
    need = 0
    for pred in preds:
        from .geometry import sliding_window_origins
        origins = sliding_window_origins(ct.shape, pred.patch_size, pred.tile_step_size)
        plan = plan_shards(origins, pred.patch_size[0], int(ct.shape[0]), world, rank)
        plans.append((origins, plan))
        need = max(need, max((hi - lo) for lo, hi in plan.touched) * pred.num_classes * Y * X * 4)
    ex.ensure(need)  # collective; `need` is the same number on every rank
    multi = len(task_ids) > 1
    lo, hi = plans[0][1].slabs[rank]
    lab_slab = torch.zeros((hi - lo, Y, X), dtype=torch.uint8, device=ct.device)
    bad = torch.zeros(1, dtype=torch.int32, device=ct.device)
    for i, (tid, pred) in enumerate(zip(task_ids, preds)):
        nvtx.range_push(f"boa/network/{tid}")
        data, _ = _preprocess(ct, pred.spec, spacing_zyx, box)
        vol = data[0]
        origins, plan = plans[i]
        b = ex.next_buffer()
        if plan.end > plan.begin:
            ex.zero(b, pred.num_classes * (plan.zhi - plan.zlo) * Y * X * 4)
            local = origins[plan.begin:plan.end].copy()
            local[:, 0] -= plan.zlo
            pred.accumulate(vol[plan.zlo:plan.zhi], local, ex.local[b])
        ex.barrier()
        if hi > lo:
            w = weight_sum(vol.shape, pred.patch_size, origins, pred.gaussian(), pred.gaussian_kind)
            ex.reduce_finalize(b, plan, Y, X, w[lo:hi], pred.num_classes, luts[i] if luts is not None else None,
                               multi, lab_slab, bad)
        nvtx.range_pop()
    out = gather_label_slabs(lab_slab, plans[0][1], dist_ctx)
    raise_if_nonfinite([bad])
    return out


def _plan_shape(spec, cropped_shape, spacing_zyx) -> tuple:
    """Shape of the network's grid for a cropped volume of this shape (compute_new_shape, default_resampling.py:25-31)."""
    from .resample import nnunet_new_shape
    target = check_plan_geometry(spec, cropped_shape, spacing_zyx)
    if target is None:
        return tuple(int(v) for v in cropped_shape)
    return nnunet_new_shape(cropped_shape, spacing_zyx, target)


def _segment_task_whole(ct, zoo, task_ids, folds, step_size, luts, dist_ctx, spacing_zyx=None) -> torch.Tensor:
    if dist_ctx is not None and dist_ctx.world_size > 1 and ct.is_cuda:
        ex = dist_ctx.peer_exchange(ct.device)
        if ex is not None:
            out = _segment_task_peers(ct, zoo, task_ids, folds, step_size, luts, dist_ctx, spacing_zyx, ex)
            if out is not None:
                return out
    out = torch.zeros(ct.shape, dtype=torch.uint8, device=ct.device)
    multi = len(task_ids) > 1
    flags: list = []
    box0 = nonzero_bbox(ct)
    for i, tid in enumerate(task_ids):
        pred = zoo.get(tid, folds, step_size)
        nvtx.range_push(f"boa/network/{tid}")
        data, box = _preprocess(ct, pred.spec, spacing_zyx, box0)
        sl = tuple(slice(b, e) for b, e in box)
        full = all(b == 0 and e == s for (b, e), s in zip(box, ct.shape))
        lut = luts[i] if luts is not None else None
        cropped_shape = tuple(e - b for b, e in box)
        if tuple(data.shape[1:]) != cropped_shape:
            # the network runs on the plan's grid; the logits come back to the cropped grid (order 1) inside the argmax
            # pass (export_prediction.py:25-38)
            from .resample import nnunet_separate_z
            if dist_ctx is not None and dist_ctx.world_size > 1:
                raise NotImplementedError("multi-GPU prediction of a volume that nnU-Net resamples to the plan's "
                                          "spacing is not implemented")
            target = check_plan_geometry(pred.spec, cropped_shape, spacing_zyx)
            sep, axis = nnunet_separate_z(target, spacing_zyx)  # current = plan spacing, new = original spacing
            if sep and axis != 0:
                raise NotImplementedError("separate-z resampling along an in-plane axis is not implemented")
            lab = pred.predict_labels(data, lut, defer=flags, resample_to=cropped_shape, separate_z=sep)
            view = out[sl]
            if multi:
                nz = lab != 0
                view[nz] = lab[nz]
            else:
                view.copy_(lab)
        elif full:
            predict_labels_sharded(pred, data, lut, out, overwrite_nonzero_only=multi, dist_ctx=dist_ctx, defer=flags)
        else:
            lab = predict_labels_sharded(pred, data, lut, dist_ctx=dist_ctx, defer=flags)
            if multi:
                view = out[sl]
                nz = lab != 0
                view[nz] = lab[nz]
            else:
                out[sl] = lab
        nvtx.range_pop()
    raise_if_nonfinite(flags)  # one device -> host read per task (predict_from_raw_data.py:622-625)
    return out


def segment_total(ct: torch.Tensor, zoo: ModelZoo, dist_ctx: DistContext | None = None, spacing_zyx=None) -> torch.Tensor:
    """task `total`, 1.5 mm: 5 part models, fold 0, step 0.8, merged through the part -> global LUTs
    (totalsegmentator/python_api.py:182-189, nnunet.py:507-559)."""
    return segment_task(ct, zoo, TOTAL_TASK_IDS, [0], 0.8, part_luts(), dist_ctx, spacing_zyx=spacing_zyx)


def segment_total_fast(ct_3mm: torch.Tensor, zoo: ModelZoo, dist_ctx: DistContext | None = None,
                       spacing_zyx=None) -> torch.Tensor:
    """task `total` with --fast-total: the single 3 mm model 297, fold 0, step 0.5 (the 0.8 step applies only below
    3 mm, totalsegmentator/nnunet.py:507-514); its labels are the global class ids already."""
    return segment_task(ct_3mm, zoo, [TOTAL_FAST_TASK_ID], [0], 0.5, None, dist_ctx, spacing_zyx=spacing_zyx)


def rough_total_6mm(ct: torch.Tensor, spacing_zyx, zoo: ModelZoo) -> torch.Tensor:
    """The rough organ segmentation of the crop pre-pass (totalsegmentator/python_api.py:673-712): the volume at 6 mm,
    model 298 (`total` classes, fold 0, step 0.5), labels back on the input grid (order 0)."""
    from .resample import resample_labels_nearest, resample_volume_cubic
    ct6 = resample_volume_cubic(ct, spacing_zyx, 6.0)
    seg6 = segment_task(ct6, zoo, [CROP_PREPASS_TASK_ID], [0], 0.5, None, None, spacing_zyx=(6.0, 6.0, 6.0))
    return resample_labels_nearest(seg6, ct.shape)


def crop_box_from_rois(rough: torch.Tensor, roi_names, spacing_zyx):
    """crop_to_mask / get_bbox_from_mask (totalsegmentator/cropping.py:11-37,75-103): bounding box of the listed `total`
    structures in the rough segmentation, grown by 20 mm (python_api.py:726) per axis, clipped to the volume.  None: the
    structures are absent (the reference then returns an empty segmentation, nnunet.py:428-445)."""
    inv = {v: k for k, v in class_map("total").items()}
    mask = passes.label_set_mask(rough, [inv[r] for r in roi_names])
    extents = []
    for ax in range(3):
        other = tuple(a for a in range(3) if a != ax)
        idx = torch.nonzero(mask.any(dim=other)).flatten()
        if idx.numel() == 0:
            return None
        extents.append((int(idx[0]), int(idx[-1])))
    return grow_crop_box(extents, rough.shape, spacing_zyx)


def grow_crop_box(extents, shape, spacing_zyx, addon_mm: float = CROP_ADDON_MM):
    """[(first, last)] index of the mask per axis -> [(lo, hi)] per axis: get_bbox_from_mask with the addon of crop_to_mask
    (cropping.py:11-37,97-99): mm -> voxels by truncation of `addon / zoom` with the zoom as the float32 the NIfTI header
    holds (20 mm / 0.8 mm is 24 voxels there, not 25), clipped to the volume."""
    box = []
    for ax, (first, last) in enumerate(extents):
        addon = int(np.float64(addon_mm) / np.float32(spacing_zyx[ax]))
        box.append((max(0, first - addon), min(int(shape[ax]), last + 1 + addon)))
    return box


def segment_crop_task(ct: torch.Tensor, spacing_zyx, zoo: ModelZoo, task: str, rough: torch.Tensor) -> torch.Tensor:
    """One task behind the crop pre-pass (lung_vessels, cerebral_bleed, hip_implant, pleural_pericard_effusion,
    liver_vessels; python_api.py:236-330): its network runs on the cropped volume at the volume's NATIVE spacing
    (`resample=None`), i.e. nnU-Net's own preprocessing resamples to the plan's spacing and the logits come back with
    order 1 (_segment_task_whole); the labels are put back into a zero volume (undo_crop, cropping.py:127-133)."""
    task_id, folds, rois = CROP_TASKS[task]
    out = torch.zeros(ct.shape, dtype=torch.uint8, device=ct.device)
    box = crop_box_from_rois(rough, rois, spacing_zyx)
    if box is None:
        return out
    sl = tuple(slice(b, e) for b, e in box)
    crop = ct[sl].contiguous()
    out[sl] = segment_task(crop, zoo, [task_id], folds, 0.5, None, None, spacing_zyx=tuple(float(v) for v in spacing_zyx))
    return out


FORCE_SPLIT_THRESHOLD = 400  # slices at 5 mm (commands.py:155-170 -> compute/inference.py:109-128)


def segment_bca_net(ct_5mm: torch.Tensor, zoo: ModelZoo, task: str, fast: bool,
                    dist_ctx: DistContext | None = None, spacing_zyx=None) -> torch.Tensor:
    tid = BODY_PARTS_TASK_ID if task == "body_parts" else BODY_REGIONS_TASK_ID
    folds = [0] if fast else [0, 1, 2, 3, 4]  # body_composition_analysis/tasks.py:15-48
    return segment_task(ct_5mm, zoo, [tid], folds, 0.5, None, dist_ctx,
                        force_split=int(ct_5mm.shape[0]) > FORCE_SPLIT_THRESHOLD, spacing_zyx=spacing_zyx)


@dataclass
class VolumeResult:
    total: torch.Tensor | None = None
    body_parts: torch.Tensor | None = None
    body_regions: torch.Tensor | None = None
    tissues: torch.Tensor | None = None
    ct_pfav: torch.Tensor | None = None
    extra: dict = field(default_factory=dict)   # crop-pre-pass tasks: name -> uint8 label map
    other_findings: list | None = None          # generate_secondary_findings sentences (builder.py:309-395)
    l3_axes_mm: tuple | None = None             # (major, minor) body axis at L3 (compute/ts_metrics.py:32-61)
    total_measurements: dict | None = None
    bca_measurements: dict | None = None
    vertebrae: dict | None = None
    timings: dict = field(default_factory=dict)


class HostStager:
    """Device -> host staging of the label maps for callers that hold host buffers: each map is copied into a pinned
    buffer on a side stream as soon as it is final, so the transfer overlaps the networks that follow (the reference
    writes every map to disk between tasks, totalsegmentator/nnunet.py:553-559,723-726).  Buffers are reused across
    volumes; `collect()` waits for the copies and hands out the host tensors (valid until the next volume)."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._buf: dict = {}
        self._pending: dict = {}

    def stage(self, name: str, t: torch.Tensor | None) -> None:
        if t is None:
            return
        key = (name, tuple(t.shape), t.dtype)
        if key not in self._buf:
            self._buf[key] = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        self.stream.wait_event(ready)
        with torch.cuda.stream(self.stream):
            self._buf[key].copy_(t, non_blocking=True)
        t.record_stream(self.stream)
        self._pending[name] = self._buf[key]

    def collect(self) -> dict:
        self.stream.synchronize()
        out, self._pending = self._pending, {}
        return out


def analyze_volume(ct: torch.Tensor, spacing_zyx, zoo: ModelZoo, models=("total", "bca"), fast_bca: bool = False,
                   cnr_adjustment: bool = False, dist_ctx: DistContext | None = None,
                   stager: HostStager | None = None, fast_total: bool = False,
                   postprocess: bool = True, median_filtering: bool = False, precomputed: dict | None = None,
                   total_measurements: bool = True) -> VolumeResult:
    """compute_all_models + run_pipeline numerics for one CT already on the device (int16 [z,y,x]).

    spacing_zyx: voxel spacing of the array axes.  `total` expects 1.5 mm (resampling is identity there,
    totalsegmentator/resampling.py:179-181); the BCA nets run at 5 mm slice thickness (resample_only_thickness).
    precomputed: label maps on the input grid (uint8 [z,y,x], keys total / body_parts / body_regions) that are reused
    instead of running their networks, total_measurements=False keeps an existing total-measurements.json - the
    reference's `recompute=False` (compute/inference.py:82-84,95-105; infer/infer.py:59-61)."""
    precomputed = precomputed or {}
    from .resample import resample_labels_nearest, resample_thickness, resample_volume_cubic, upsample_labels_nearest

    res = VolumeResult()
    models = set(models)
    if "bca" in models:
        models.add("total")  # compute/config.py:54-55
    if models & set(CROP_TASKS):
        models.add("total")  # compute_measurements needs it (autochthon reference) and the CLI always runs it first
    pending_total = pending_l3 = None
    # HU range of the CT: sizes the per-label histograms exactly; read once here, before any network is enqueued (the
    # only device -> host read ahead of the label maps)
    lo_t, hi_t = torch.aminmax(ct)
    hu_range = (int(lo_t), int(hi_t))
    ev = lambda: torch.cuda.Event(enable_timing=True)
    marks = [("start", ev())]
    marks[-1][1].record()

    def mark(name):
        e = ev()
        e.record()
        marks.append((name, e))

    if "total" in models:
        # `total` runs at 1.5 mm: other inputs are resampled (order 3) and the label map goes back to the input grid
        # (order 0), totalsegmentator/nnunet.py:466-470,685-687
        net_spacing = 3.0 if fast_total else 1.5  # python_api.py:169-189
        if "total" in precomputed:
            res.total = precomputed["total"].to(ct.device, torch.uint8).contiguous()
        else:
            ct_net = resample_volume_cubic(ct, spacing_zyx, net_spacing)
            if ct_net is not ct:
                mark("resample_total")
            sp_net = (net_spacing,) * 3
            seg = (segment_total_fast(ct_net, zoo, dist_ctx, sp_net) if fast_total
                   else segment_total(ct_net, zoo, dist_ctx, sp_net))
            res.total = resample_labels_nearest(seg, ct.shape)
            del ct_net, seg
        mark("total_nets")
        if stager is not None:
            stager.stage("total", res.total)
        sx_sy_sz = (spacing_zyx[2], spacing_zyx[1], spacing_zyx[0])

        # tasks behind the crop pre-pass (`--models all`): one rough 6 mm segmentation serves all of them
        crop_tasks = [m for m in sorted(models) if m in CROP_TASKS]
        if crop_tasks:
            rough = rough_total_6mm(ct, spacing_zyx, zoo)
            for m in crop_tasks:
                res.extra[m] = precomputed[m].to(ct.device, torch.uint8).contiguous() if m in precomputed else \
                    segment_crop_task(ct, spacing_zyx, zoo, m, rough)
                if stager is not None:
                    stager.stage(m, res.extra[m])
            del rough
            mark("crop_tasks")
        # `total-measurements.json` needs only the CT and the `total` label map: its device passes are enqueued here and
        # the small tables travel to pinned host memory; the ~300 per-name statistics (host numpy) are evaluated at the
        # end, after the body-composition networks are enqueued, i.e. while the GPU is busy.
        if total_measurements:
            with nvtx.range("boa/total_measurements"):
                pending_total = enqueue_measurements(ct, {"total": res.total, **res.extra}, sx_sy_sz, cnr_adjustment,
                                                     return_ct_pfav_mask=True, hu_range=hu_range, dist_ctx=dist_ctx)
            res.ct_pfav = pending_total.pfav_mask
            mark("total_measurements")
            if stager is not None:
                stager.stage("ct_pfav", res.ct_pfav)
    if "bca" in models or "body_parts" in models or "body_regions" in models:
        ct5 = resample_thickness(ct, spacing_zyx[0], 5.0)
        # the spacing the 5 mm volume carries into nnU-Net (totalsegmentator/nnunet.py:457-459: only the thickness changes)
        sp5 = (5.0 if ct5 is not ct else float(spacing_zyx[0]), float(spacing_zyx[1]), float(spacing_zyx[2]))
        mark("resample")
        want_parts = "bca" in models or "body_parts" in models
        want_regions = "bca" in models or "body_regions" in models
        # connected-component post-processing of both maps (infer/infer.py:81-89).  The reference applies it after the
        # maps are back on the input grid; replicated slices map components one to one, so it runs here on the 5 mm
        # grid with every slice weighted by the number of output slices it becomes (postprocess.py)
        from .postprocess import postprocess_part_segmentation, postprocess_region_segmentation, slice_weights
        # slices are REPLICATED on the way back (input thinner than 5 mm, the usual case): weighted labelling on the 5 mm
        # grid is exact.  Input thicker than 5 mm: slices are dropped on the way back, which can split or join
        # components, so the post-processing runs on the input grid like the reference's.
        on_5mm = postprocess and ct5.shape[0] <= ct.shape[0]
        weights = slice_weights(ct5.shape[0], ct.shape[0], ct.device) if on_5mm else None

        grid5 = {}  # post-processed maps on the 5 mm grid (each slice stands for weights[z] slices of the input grid)

        def l3_axes():
            # body cross-section at L3 (compute/ts_metrics.py:32-61): needs `total` and the final body_parts map
            if "bca" not in models or res.total is None:
                return None
            from .ts_metrics import PendingL3Axes
            return PendingL3Axes(res.total, res.body_parts, (spacing_zyx[2], spacing_zyx[1]))

        def finish(net_out, fn, name):
            mark(f"{name}_net")
            with nvtx.range(f"boa/postprocess/{name}"):
                kw_d = {"dist_ctx": dist_ctx} if fn is postprocess_part_segmentation else {}
                if on_5mm:
                    net_out = fn(net_out, weights, **kw_d)
                    grid5[name] = net_out
                out = upsample_labels_nearest(net_out, ct.shape[0])
                if postprocess and not on_5mm:
                    out = fn(out, None, **kw_d)
            if postprocess:
                mark(f"{name}_postprocess")
            return out

        # >= 4 GPUs and both maps wanted: both networks first, then the two post-processings CONCURRENTLY - the labels of
        # body_parts are dealt out to ranks 0 .. W-2 and the last rank runs the (sequential) body_regions passes
        pair = (want_parts and want_regions and on_5mm and dist_ctx is not None
                and dist_ctx.world_size >= int(os.environ.get("BOA_B200_PAIR_MIN_RANKS", "4"))
                and "body_parts" not in precomputed and "body_regions" not in precomputed)
        if pair:
            from .postprocess import postprocess_pair_distributed
            parts_raw = segment_bca_net(ct5, zoo, "body_parts", fast_bca, dist_ctx, sp5)
            mark("body_parts_net")
            regions_raw = segment_bca_net(ct5, zoo, "body_regions", fast_bca, dist_ctx, sp5)
            mark("body_regions_net")
            with nvtx.range("boa/postprocess/pair"):
                parts_pp, regions_pp = postprocess_pair_distributed(parts_raw, regions_raw, weights, dist_ctx)
                res.body_parts = upsample_labels_nearest(parts_pp, ct.shape[0])
                res.body_regions = upsample_labels_nearest(regions_pp, ct.shape[0])
                grid5["body_regions"] = regions_pp
            mark("bca_postprocess")
            if stager is not None:
                stager.stage("body_parts", res.body_parts)
            pending_l3 = l3_axes()
        elif want_parts and "body_parts" in precomputed:
            res.body_parts = precomputed["body_parts"].to(ct.device, torch.uint8).contiguous()
            pending_l3 = l3_axes()
        elif want_parts:
            res.body_parts = finish(segment_bca_net(ct5, zoo, "body_parts", fast_bca, dist_ctx, sp5),
                                    postprocess_part_segmentation, "body_parts")
            if stager is not None:
                stager.stage("body_parts", res.body_parts)
            pending_l3 = l3_axes()  # its host part runs while the body_regions networks are on the device
        if pair:
            pass
        elif want_regions and "body_regions" in precomputed:
            res.body_regions = precomputed["body_regions"].to(ct.device, torch.uint8).contiguous()
        elif want_regions:
            res.body_regions = finish(segment_bca_net(ct5, zoo, "body_regions", fast_bca, dist_ctx, sp5),
                                      postprocess_region_segmentation, "body_regions")
        if stager is not None:
            stager.stage("body_regions", res.body_regions)
        mark("bca_nets")
    if pending_total is not None:
        t_host = time.perf_counter()
        res.total_measurements, _ = pending_total.finish()
        res.timings["total_measurements_host"] = time.perf_counter() - t_host
    if "bca" in models:
        nvtx.range_push("boa/bca_measurements")
        res.tissues = bca.subclassify_tissues(ct, res.body_regions, median_filtering)
        if stager is not None:
            stager.stage("tissues", res.tissues)
        sx_sy_sz = (spacing_zyx[2], spacing_zyx[1], spacing_zyx[0])
        res.bca_measurements, res.vertebrae, tables = bca.build_bca_measurements(
            ct, res.tissues, res.body_parts, res.body_regions, res.total, sx_sy_sz)
        examined = bca.body_part_from_regions(tables, float(sx_sy_sz[2]))
        # breast implants: components of the region map, labelled on the 5 mm grid when that map exists
        regions5 = grid5.get("body_regions")
        res.other_findings = bca.secondary_findings(
            tables, examined, float(np.prod(sx_sy_sz) / 1000.0),
            body_regions=regions5 if regions5 is not None else res.body_regions,
            slice_weights=weights if regions5 is not None else None)
        if pending_l3 is not None:
            res.l3_axes_mm = pending_l3.finish()
        nvtx.range_pop()
        mark("bca_measurements")
    torch.cuda.synchronize()
    for (n0, e0), (n1, e1) in zip(marks, marks[1:]):
        res.timings[n1] = e0.elapsed_time(e1) / 1e3
    return res


def analyze_from_host(ct_host: torch.Tensor, spacing_zyx, zoo: ModelZoo, device=None, maps_on_all_ranks: bool = False,
                      **kw) -> dict:
    """The call a user of the Python API makes for one CT held in (pinned) host memory: H2D of the int16 volume,
    all networks and passes on the device, D2H of the uint8 label maps (pinned staging buffers owned by the zoo,
    copied on a side stream while later networks run); returns host tensors + measurement dicts.  The host tensors
    are views of the staging buffers: copy them if they must outlive the next call on the same zoo.
    With a dist_ctx (one volume on N GPUs) the job has ONE caller: rank 0 receives the label maps, the other ranks
    only the measurement dicts (their maps stay None unless maps_on_all_ranks) - N copies of 0.67 GB into the same
    host memory cost 49 ms of a 375 ms step at 8 GPUs."""
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    dist_ctx = kw.get("dist_ctx")
    deliver = maps_on_all_ranks or dist_ctx is None or dist_ctx.rank == 0
    stager = getattr(zoo, "_stager", None)
    if deliver and (stager is None or stager.device != dev):
        stager = zoo._stager = HostStager(dev)
    ct = ct_host.to(dev, non_blocking=True)
    res = analyze_volume(ct, spacing_zyx, zoo, stager=stager if deliver else None, **kw)
    out = {"total_measurements": res.total_measurements, "bca_measurements": res.bca_measurements,
           "vertebrae": res.vertebrae, "timings": res.timings}
    host = stager.collect() if deliver else {}
    for name in ("total", "body_parts", "body_regions", "tissues", "ct_pfav"):
        out[name] = host.get(name)
    if not deliver:
        torch.cuda.current_stream(dev).synchronize()  # the call returns when this rank's share of the volume is done
    return out
