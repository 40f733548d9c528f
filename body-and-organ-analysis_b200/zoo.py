"""Synthetic model zoo and synthetic CT phantoms (there are no real weights or scans offline).

Writes checkpoints in the real on-disk contract read by the reference
(`Dataset{ID:03d}_<name>/<Trainer>__nnUNetPlans__3d_fullres/{plans.json,dataset.json,fold_k/checkpoint_final.pth}`,
predict_from_raw_data.py:76-94; old-format plans as shipped by TotalSegmentator, plans_handler.py:36-97) with seeded
random weights of the TotalSegmentator 3d_fullres PlainConvUNet geometry, and builds seeded body phantoms with integer
HU values.
"""
from __future__ import annotations

import json
import math
import os

import numpy as np
import torch

# dataset id -> (folder name, trainer, number of classes incl. background)   (totalsegmentator/python_api.py:168-189,
# totalsegmentator/map_to_binary.py:808-958, body_composition_analysis/tasks.py:15-48, infer/infer.py:26-31)
DATASETS = {
    291: ("Dataset291_TotalSegmentator_part1_organs_1559subj", "nnUNetTrainerNoMirroring", 25),
    292: ("Dataset292_TotalSegmentator_part2_vertebrae_1532subj", "nnUNetTrainerNoMirroring", 27),
    293: ("Dataset293_TotalSegmentator_part3_cardiac_1559subj", "nnUNetTrainerNoMirroring", 19),
    294: ("Dataset294_TotalSegmentator_part4_muscles_1559subj", "nnUNetTrainerNoMirroring", 24),
    295: ("Dataset295_TotalSegmentator_part5_ribs_1559subj", "nnUNetTrainerNoMirroring", 27),
    297: ("Dataset297_TotalSegmentator_total_3mm_1559subj", "nnUNetTrainer_4000epochs_NoMirroring", 118),
    298: ("Dataset298_TotalSegmentator_total_6mm_1559subj", "nnUNetTrainer_4000epochs_NoMirroring", 118),
    258: ("Dataset258_lung_vessels_248subj", "nnUNetTrainer", 3),
    150: ("Dataset150_icb_v0", "nnUNetTrainer", 2),
    260: ("Dataset260_hip_implant_71subj", "nnUNetTrainer", 2),
    315: ("Dataset315_thoraxCT", "nnUNetTrainer", 4),
    8: ("Dataset008_HepaticVessel", "nnUNetTrainer", 3),
    542: ("Dataset542_BodyRegions", "nnUNetTrainerNoMirroring", 12),
    543: ("Dataset543_BodyParts", "nnUNetTrainer_1500epochs_NoMirroring", 7),
}


# plan spacing of the synthetic models: the 3 mm / 6 mm `total` models, and the crop-task models at a spacing that is
# NOT the test volumes' (their networks run at native resolution, so nnU-Net's own resampling is exercised)
_SPACING = {297: (3.0, 3.0, 3.0), 298: (6.0, 6.0, 6.0), 258: (1.0, 1.0, 1.0), 150: (1.0, 1.0, 1.0), 260: (1.25, 1.25, 1.25),
            315: (2.0, 2.0, 2.0), 8: (1.0, 1.0, 1.0)}


def default_plans(patch=(128, 128, 128), base=32, max_features=320, n_stages=6, spacing=(1.5, 1.5, 1.5),
                  name="Dataset000") -> dict:
    return {
        "dataset_name": name, "plans_name": "nnUNetPlans", "image_reader_writer": "SimpleITKIO",
        "transpose_forward": [0, 1, 2], "transpose_backward": [0, 1, 2],
        "experiment_planner_used": "ExperimentPlanner",
        "foreground_intensity_properties_per_channel": {
            "0": {"mean": -370.0, "std": 436.6, "percentile_00_5": -1024.0, "percentile_99_5": 276.0,
                  "min": -1024.0, "max": 3071.0, "median": -249.0}},
        "configurations": {"3d_fullres": {
            "data_identifier": "nnUNetPlans_3d_fullres", "preprocessor_name": "DefaultPreprocessor",
            "batch_size": 2, "patch_size": list(patch), "spacing": list(spacing),
            "normalization_schemes": ["CTNormalization"], "use_mask_for_norm": [False],
            "UNet_class_name": "PlainConvUNet", "UNet_base_num_features": base, "unet_max_num_features": max_features,
            "n_conv_per_stage_encoder": [2] * n_stages, "n_conv_per_stage_decoder": [2] * (n_stages - 1),
            "num_pool_per_axis": [n_stages - 1] * 3,
            "pool_op_kernel_sizes": [[1, 1, 1]] + [[2, 2, 2]] * (n_stages - 1),
            "conv_kernel_sizes": [[3, 3, 3]] * n_stages,
            "resampling_fn_data": "resample_data_or_seg_to_shape", "resampling_fn_seg": "resample_data_or_seg_to_shape",
            "resampling_fn_probabilities": "resample_data_or_seg_to_shape",
            "resampling_fn_data_kwargs": {"is_seg": False, "order": 3, "order_z": 0, "force_separate_z": None},
            "resampling_fn_seg_kwargs": {"is_seg": True, "order": 1, "order_z": 0, "force_separate_z": None},
            "resampling_fn_probabilities_kwargs": {"is_seg": False, "order": 1, "order_z": 0, "force_separate_z": None},
            "batch_dice": False}},
    }


def random_state_dict(arch: dict, seed: int, head_gain: float = 4.0) -> dict:
    """Seeded weights under dynamic_network_architectures' key names (incl. the alias keys a real checkpoint has)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(prefix, cin, cout, ks):
        fan_in = cin * ks[0] * ks[1] * ks[2]
        std = math.sqrt(2.0 / ((1 + 0.01 ** 2) * fan_in))  # kaiming_normal_(a=1e-2), InitWeights_He
        sd[prefix + ".conv.weight"] = torch.randn((cout, cin, *ks), generator=g) * std
        sd[prefix + ".conv.bias"] = torch.randn((cout,), generator=g) * 0.05
        sd[prefix + ".norm.weight"] = 1.0 + 0.1 * torch.randn((cout,), generator=g)
        sd[prefix + ".norm.bias"] = 0.1 * torch.randn((cout,), generator=g)
        # nn.Sequential alias registered by ConvDropoutNormReLU.all_modules
        sd[prefix + ".all_modules.0.weight"] = sd[prefix + ".conv.weight"]
        sd[prefix + ".all_modules.0.bias"] = sd[prefix + ".conv.bias"]
        sd[prefix + ".all_modules.1.weight"] = sd[prefix + ".norm.weight"]
        sd[prefix + ".all_modules.1.bias"] = sd[prefix + ".norm.bias"]

    feats, n = arch["features"], len(arch["features"])
    cin = arch["in_channels"]
    for s in range(n):
        for i in range(arch["n_conv_enc"][s]):
            conv(f"encoder.stages.{s}.0.convs.{i}", cin, feats[s], arch["kernels"][s])
            cin = feats[s]
    for j in range(n - 1):
        below, skip = feats[-(j + 1)], feats[-(j + 2)]
        st = arch["strides"][-(j + 1)]
        fan = skip * st[0] * st[1] * st[2]
        sd[f"decoder.transpconvs.{j}.weight"] = torch.randn((below, skip, *st), generator=g) * math.sqrt(2.0 / fan)
        sd[f"decoder.transpconvs.{j}.bias"] = torch.randn((skip,), generator=g) * 0.05
        cin = 2 * skip
        for i in range(arch["n_conv_dec"][j]):
            conv(f"decoder.stages.{j}.convs.{i}", cin, skip, arch["kernels"][-(j + 2)])
            cin = skip
        # deep-supervision heads exist in the checkpoint for every level; only the last is used at inference
        sd[f"decoder.seg_layers.{j}.weight"] = torch.randn((arch["num_classes"], skip, 1, 1, 1), generator=g) * (
            head_gain / math.sqrt(skip))
        sd[f"decoder.seg_layers.{j}.bias"] = torch.randn((arch["num_classes"],), generator=g) * 0.5
    for k in [k for k in sd if k.startswith("encoder.")]:
        sd["decoder." + k] = sd[k]  # UNetDecoder keeps a reference to the encoder
    return sd


def write_model(root: str, dataset_id: int, plans: dict, num_classes: int, folds=(0,), seed: int = 0,
                folder_name: str | None = None, trainer: str | None = None) -> str:
    from .plans import arch_from_plans

    name, default_trainer, _ = DATASETS.get(dataset_id, (f"Dataset{dataset_id:03d}_Synthetic", "nnUNetTrainerNoMirroring", 0))
    folder_name, trainer = folder_name or name, trainer or default_trainer
    out = os.path.join(root, folder_name, f"{trainer}__nnUNetPlans__3d_fullres")
    os.makedirs(out, exist_ok=True)
    plans = dict(plans, dataset_name=folder_name)
    with open(os.path.join(out, "plans.json"), "w") as f:
        json.dump(plans, f)
    labels = {"background": 0, **{f"class_{i}": i for i in range(1, num_classes)}}
    with open(os.path.join(out, "dataset.json"), "w") as f:
        json.dump({"channel_names": {"0": "CT"}, "labels": labels, "numTraining": 1, "file_ending": ".nii.gz"}, f)
    arch = arch_from_plans(plans, "3d_fullres", 1, num_classes)
    for fold in folds:
        os.makedirs(os.path.join(out, f"fold_{fold}"), exist_ok=True)
        sd = random_state_dict(arch, seed * 1000 + dataset_id * 10 + int(fold))
        torch.save({"network_weights": sd, "trainer_name": trainer,
                    "init_args": {"plans": plans, "configuration": "3d_fullres", "fold": fold},
                    "inference_allowed_mirroring_axes": None},
                   os.path.join(out, f"fold_{fold}", "checkpoint_final.pth"))
    return out


def write_zoo(root: str, patch=(128, 128, 128), base=32, max_features=320, n_stages=6, bca_folds=(0, 1, 2, 3, 4),
              seed: int = 0, datasets=None) -> dict:
    """All seven networks of `--models total+bca`; returns dataset id -> model folder."""
    out = {}
    for did, (name, trainer, ncls) in DATASETS.items():
        if datasets is not None and did not in datasets:
            continue
        spacing = _SPACING.get(did, (1.5, 1.5, 1.5) if did < 500 else (5.0, 1.5, 1.5))
        plans = default_plans(patch, base, max_features, n_stages, spacing, name)
        folds = (0,) if did < 500 else tuple(bca_folds)
        out[did] = write_model(root, did, plans, ncls, folds, seed)
    return out


def synthetic_ct(shape, seed: int = 0) -> np.ndarray:
    """Seeded body phantom, integer HU in [-1024, 2047], int16 [d0, d1, d2] (axis 0 = cranio-caudal)."""
    rng = np.random.default_rng(seed)
    d0, d1, d2 = shape
    z = np.linspace(-1, 1, d0, dtype=np.float32)[:, None, None]
    y = np.linspace(-1, 1, d1, dtype=np.float32)[None, :, None]
    x = np.linspace(-1, 1, d2, dtype=np.float32)[None, None, :]
    vol = np.full(shape, -1000.0, dtype=np.float32)
    body = (x / 0.8) ** 2 + (y / 0.6) ** 2 < 1.0
    vol = np.where(body, 40.0, vol)
    fat_ring = body & ((x / 0.68) ** 2 + (y / 0.48) ** 2 >= 1.0)
    vol = np.where(fat_ring, -100.0, vol)
    lungs = (((x - 0.3) / 0.22) ** 2 + (y / 0.3) ** 2 + ((z + 0.35) / 0.35) ** 2 < 1.0) | (
        ((x + 0.3) / 0.22) ** 2 + (y / 0.3) ** 2 + ((z + 0.35) / 0.35) ** 2 < 1.0)
    vol = np.where(lungs, -800.0, vol)
    spine = (x / 0.07) ** 2 + ((y - 0.35) / 0.07) ** 2 < 1.0
    vol = np.where(spine & np.ones_like(z, dtype=bool), 700.0, vol)
    for _ in range(6):  # organ blobs
        c = rng.uniform(-0.5, 0.5, 3)
        r = rng.uniform(0.08, 0.2, 3)
        hu = rng.uniform(20, 120)
        blob = ((z - c[0]) / r[0]) ** 2 + ((y - c[1]) / r[1]) ** 2 + ((x - c[2]) / r[2]) ** 2 < 1.0
        vol = np.where(blob & body, hu, vol)
    sigma = np.where(vol > 500, 150.0, np.where(vol < -500, 50.0, 18.0)).astype(np.float32)
    vol = vol + sigma * rng.standard_normal(shape, dtype=np.float32)
    return np.clip(np.rint(vol), -1024, 2047).astype(np.int16)


def synthetic_specs(patch=(128, 128, 128), base=32, max_features=320, n_stages=6, bca_folds=5, seed: int = 0,
                    datasets=None) -> dict:
    """In-memory ModelSpecs (no disk round trip) for ModelZoo.from_specs: same geometry and seeds as write_zoo."""
    from .plans import ModelSpec, arch_from_plans

    out = {}
    for did, (name, trainer, ncls) in DATASETS.items():
        if datasets is not None and did not in datasets:
            continue
        spacing = _SPACING.get(did, (1.5, 1.5, 1.5) if did < 500 else (5.0, 1.5, 1.5))
        plans = default_plans(patch, base, max_features, n_stages, spacing, name)
        arch = arch_from_plans(plans, "3d_fullres", 1, ncls)
        nf = 1 if did < 500 else bca_folds
        weights = [random_state_dict(arch, seed * 1000 + did * 10 + f) for f in range(nf)]
        out[did] = ModelSpec(arch=arch, intensity=plans["foreground_intensity_properties_per_channel"]["0"],
                             labels={"background": 0, **{f"class_{i}": i for i in range(1, ncls)}},
                             transpose_forward=[0, 1, 2], transpose_backward=[0, 1, 2], spacing=list(spacing),
                             configuration="3d_fullres", fold_weights=weights, folder="<synthetic>")
    return out
