"""plans.json / dataset.json / checkpoint loading: host mirror of
  PlansManager / ConfigurationManager incl. the old-plans reconstruction
      (_external/nnunetv2/utilities/plans_handling/plans_handler.py:36-97,142-152,264-325)
  nnUNetPredictor.initialize_from_trained_model_folder (_external/nnunetv2/inference/predict_from_raw_data.py:67-129)
  LabelManager channel counts (_external/nnunetv2/utilities/label_handling/label_handling.py:241-245,294-311)
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field


@dataclass
class ModelSpec:
    arch: dict
    intensity: dict            # mean, std, percentile_00_5, percentile_99_5 of channel 0
    labels: dict               # name -> int
    transpose_forward: list
    transpose_backward: list
    spacing: list
    configuration: str
    fold_weights: list = field(default_factory=list)   # list of state dicts (name -> tensor)
    folder: str = ""
    allowed_mirroring_axes: object = None              # `inference_allowed_mirroring_axes` of the checkpoint


def arch_from_plans(plans: dict, configuration: str, num_input_channels: int, num_classes: int) -> dict:
    cfg = dict(plans["configurations"][configuration])
    while "inherits_from" in cfg:  # plans_handler.py:231-253
        parent = dict(plans["configurations"][cfg.pop("inherits_from")])
        parent.update(cfg)
        cfg = parent
    if "architecture" in cfg:
        kw = cfg["architecture"]["arch_kwargs"]
        cls = cfg["architecture"]["network_class_name"].rsplit(".", 1)[-1]
        features = list(kw["features_per_stage"])
        kernels, strides = kw["kernel_sizes"], kw["strides"]
        n_enc, n_dec = kw["n_conv_per_stage"], kw["n_conv_per_stage_decoder"]
        eps = (kw.get("norm_op_kwargs") or {}).get("eps", 1e-5)
    else:
        cls = cfg["UNet_class_name"]
        n_enc, n_dec = cfg["n_conv_per_stage_encoder"], cfg["n_conv_per_stage_decoder"]
        features = [min(cfg["UNet_base_num_features"] * 2 ** i, cfg["unet_max_num_features"])
                    for i in range(len(n_enc))]
        kernels, strides = cfg["conv_kernel_sizes"], cfg["pool_op_kernel_sizes"]
        eps = 1e-5
    if cls != "PlainConvUNet":
        raise NotImplementedError(f"network class {cls} is not implemented (PlainConvUNet only)")
    if isinstance(n_enc, int):
        n_enc = [n_enc] * len(features)
    if isinstance(n_dec, int):
        n_dec = [n_dec] * (len(features) - 1)
    return {
        "in_channels": num_input_channels, "num_classes": num_classes, "features": features,
        "kernels": [list(k) for k in kernels], "strides": [list(s) for s in strides],
        "n_conv_enc": list(n_enc), "n_conv_dec": list(n_dec), "eps": float(eps), "leaky_slope": 0.01,
        "patch_size": list(cfg["patch_size"]),
    }


def macs_per_patch(arch: dict) -> int:
    """Algorithmic multiply-accumulates of one PlainConvUNet forward over one patch (roofline accounting; equals
    boa_net_macs_per_patch of the built network)."""
    feats, n = arch["features"], len(arch["features"])
    dims, shape = [], list(arch["patch_size"])
    for s in range(n):
        shape = [d // st for d, st in zip(shape, arch["strides"][s])]
        dims.append(list(shape))
    vox = lambda s: dims[s][0] * dims[s][1] * dims[s][2]
    k3 = lambda s: arch["kernels"][s][0] * arch["kernels"][s][1] * arch["kernels"][s][2]
    total, cin = 0, arch["in_channels"]
    for s in range(n):
        for _ in range(arch["n_conv_enc"][s]):
            total += k3(s) * cin * feats[s] * vox(s)
            cin = feats[s]
    for j in range(n - 1):
        below, s = n - 1 - j, n - 2 - j
        st = arch["strides"][below]
        total += st[0] * st[1] * st[2] * feats[below] * feats[s] * vox(below)
        cin = 2 * feats[s]
        for _ in range(arch["n_conv_dec"][j]):
            total += k3(s) * cin * feats[s] * vox(s)
            cin = feats[s]
    return total + feats[0] * arch["num_classes"] * vox(0)


def load_model_folder(model_training_output_dir: str, use_folds, checkpoint_name: str = "checkpoint_final.pth") -> ModelSpec:
    import torch

    with open(os.path.join(model_training_output_dir, "dataset.json")) as f:
        dataset_json = json.load(f)
    with open(os.path.join(model_training_output_dir, "plans.json")) as f:
        plans = json.load(f)
    if use_folds is None:  # auto-detect (predict_from_raw_data.py:75,131-140): every fold_k with the checkpoint
        use_folds = sorted(int(d[5:]) for d in os.listdir(model_training_output_dir)
                           if d.startswith("fold_") and d[5:].isdigit()
                           and os.path.isfile(os.path.join(model_training_output_dir, d, checkpoint_name)))
    if isinstance(use_folds, (str, int)):
        use_folds = [use_folds]
    weights, configuration, mirror = [], None, None
    for fold in use_folds:
        fold = int(fold) if fold != "all" else fold
        ckpt = torch.load(os.path.join(model_training_output_dir, f"fold_{fold}", checkpoint_name),
                          map_location="cpu", weights_only=False)
        configuration = ckpt["init_args"]["configuration"]
        # checkpoints of the plain nnUNetTrainer allow mirroring; the reference predicts with tta=False everywhere on
        # this path (totalsegmentator/python_api.py:708,746,752 -> disable_tta), so the axes are only recorded: the
        # predictor refuses use_mirroring=True instead of silently skipping it
        mirror = ckpt.get("inference_allowed_mirroring_axes")
        weights.append({k: v for k, v in ckpt["network_weights"].items()})
    labels = dataset_json["labels"]
    if any(isinstance(v, (list, tuple)) for v in labels.values()):
        raise NotImplementedError("region-based label sets are not implemented")
    num_classes = len([k for k in labels if k != "ignore"])
    channels = dataset_json.get("channel_names", dataset_json.get("modality"))
    arch = arch_from_plans(plans, configuration, len(channels), num_classes)
    props = plans.get("foreground_intensity_properties_per_channel",
                      plans.get("foreground_intensity_properties_by_modality"))["0"]
    return ModelSpec(arch=arch, intensity=props, labels=labels, transpose_forward=plans["transpose_forward"],
                     transpose_backward=plans["transpose_backward"],
                     spacing=plans["configurations"][configuration].get("spacing", [1, 1, 1]),
                     configuration=configuration, fold_weights=weights, folder=model_training_output_dir,
                     allowed_mirroring_axes=mirror)


def find_model_folder(results_root: str, dataset_id: int, trainer: str, plans: str = "nnUNetPlans",
                      configuration: str = "3d_fullres") -> str:
    """Dataset%03d_* lookup (nnunetv2/utilities/dataset_name_id_conversion.py:21-39, file_path_utilities.py:11-26)."""
    prefix = f"Dataset{dataset_id:03d}"
    cands = [d for d in sorted(os.listdir(results_root)) if d.startswith(prefix)]
    if len(cands) != 1:
        raise RuntimeError(f"expected exactly one folder starting with {prefix} under {results_root}, found {cands}")
    return os.path.join(results_root, cands[0], f"{trainer}__{plans}__{configuration}")
