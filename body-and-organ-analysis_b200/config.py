"""Model / device resolution: mirror of body_organ_analysis/compute/config.py:13-69 and constants.py:16-36 (same names,
argument meaning and error behaviour; pinned by the reference's tests/test_config.py truth tables)."""
from __future__ import annotations

import logging
import os

logger = logging.getLogger(__name__)

BASE_MODELS = {"bca", "body_regions", "body_parts"}
ALL_MODELS = {"bca", "body_parts", "body_regions", "cerebral_bleed", "hip_implant", "liver_vessels", "lung_vessels",
              "pleural_pericard_effusion", "total"}
LICENSE_MODELS = {"heartchambers_highres"}
AVAILABLE_MODELS = ALL_MODELS | LICENSE_MODELS
# what this framework computes (the hot path of BASELINE.json); the other tasks are further nnU-Net models with
# crop pre-passes and raise NotImplementedError when requested
IMPLEMENTED_MODELS = {"total", "bca", "body_regions", "body_parts"}


def env_bool(name: str, default: bool = False) -> bool:
    raw = os.getenv(name)
    if raw is None:
        return default
    return raw.strip().lower() in {"1", "true"}


def env_str(name: str, default: str | None = None) -> str | None:
    raw = os.getenv(name)
    if raw is None or raw.strip().lower() in {"", "todo"}:
        return default
    return raw.strip()


def resolve_models(spec: str | None, strict: bool = False, license_number: str | None = None) -> set[str]:
    if not spec or spec.lower() == "all":
        models = set(ALL_MODELS)
        # license-only models need TotalSegmentator's online license check: never added offline
    else:
        models = {s.replace("-", "_") for s in spec.split("+")}
        invalid = models - AVAILABLE_MODELS
        if invalid:
            if strict:
                raise ValueError(f"Unknown model(s): {', '.join(sorted(invalid))}. "
                                 f"Available: {', '.join(sorted(AVAILABLE_MODELS))}")
            logger.error("Ignoring invalid model entries: %s. Available models are: %s.", invalid,
                         sorted(AVAILABLE_MODELS))
            models -= invalid
    if "bca" in models:
        models = (models | {"total"}) - {"body_regions", "body_parts"}
    return models


def resolve_device(device: str | None = None) -> str:
    device_str = device or os.environ.get("DEVICE", "gpu")
    device_str, _, gpu_id = device_str.partition(":")
    if device_str == "cuda":
        device_str = "gpu"
    gpu_id = gpu_id or os.environ.get("NVIDIA_ID", "")
    if gpu_id and device_str == "gpu":
        os.environ.setdefault("NVIDIA_VISIBLE_DEVICES", gpu_id)
        device_str = f"gpu:{gpu_id}"
    return device_str
