"""Which models to run and on which device.

Behavioural mirror of the reference's selection rules (body_organ_analysis/compute/config.py:13-69 with the model sets
of compute/constants.py:16-36); the truth tables of the reference's tests/test_config.py are restated in
tests/test_cli_config.py.  Rules:

  * models: "a+b+c", dashes and underscores interchangeable; empty or "all" selects every model that needs no licence;
    unknown names are an error (strict) or are logged and dropped; `bca` implies `total` and absorbs the two
    body-composition networks it is made of;
  * device: "<kind>[:<gpu id>]" from the argument, else $DEVICE, else "gpu"; "cuda" is an alias of "gpu"; a missing id
    falls back to $NVIDIA_ID and is exported as NVIDIA_VISIBLE_DEVICES unless that is already set;
  * environment flags: "1" / "true" (any case, surrounding blanks ignored) are true; the placeholders "" and "todo" of
    the reference's .env template count as "not set" for string settings.
"""
from __future__ import annotations

import logging
import os

logger = logging.getLogger(__name__)

BASE_MODELS = frozenset({"bca", "body_regions", "body_parts"})
ALL_MODELS = frozenset({"bca", "body_parts", "body_regions", "cerebral_bleed", "hip_implant", "liver_vessels",
                        "lung_vessels", "pleural_pericard_effusion", "total"})
LICENSE_MODELS = frozenset({"heartchambers_highres"})
AVAILABLE_MODELS = ALL_MODELS | LICENSE_MODELS
# What this framework computes: the hot path of BASELINE.json and the tasks behind a crop pre-pass (SURVEY.md 8f rank 3,
# labels.CROP_TASKS) - i.e. everything `--models all` selects.  The licence-only model raises NotImplementedError in
# commands.analyze_ct.
IMPLEMENTED_MODELS = frozenset(ALL_MODELS)

_TRUE_WORDS = ("1", "true")
_PLACEHOLDERS = ("", "todo")
_BCA_PARTS = ("body_regions", "body_parts")


def _env(name: str) -> str | None:
    value = os.environ.get(name)
    return None if value is None else value.strip()


def env_bool(name: str, default: bool = False) -> bool:
    value = _env(name)
    return default if value is None else value.lower() in _TRUE_WORDS


def env_str(name: str, default: str | None = None) -> str | None:
    value = _env(name)
    return default if value is None or value.lower() in _PLACEHOLDERS else value


def _parse_model_list(spec: str) -> tuple[set[str], set[str]]:
    """'total+lung-vessels+foo' -> ({'total', 'lung_vessels'}, {'foo'})."""
    wanted = {token.replace("-", "_") for token in spec.split("+")}
    unknown = {m for m in wanted if m not in AVAILABLE_MODELS}
    return wanted - unknown, unknown


def resolve_models(spec: str | None, strict: bool = False, license_number: str | None = None) -> set[str]:
    """The set of models a `--models` / $MODELS string selects.  `license_number` is accepted for signature
    compatibility: the licence-only model needs TotalSegmentator's online check and is never added offline."""
    everything = spec is None or spec == "" or spec.lower() == "all"
    if everything:
        chosen = set(ALL_MODELS)
    else:
        chosen, unknown = _parse_model_list(spec)
        if unknown:
            known = ", ".join(sorted(AVAILABLE_MODELS))
            if strict:
                raise ValueError(f"Unknown model(s): {', '.join(sorted(unknown))}. Available: {known}")
            logger.error("Ignoring invalid model entries: %s. Available models are: %s.", sorted(unknown), known)
    if "bca" in chosen:
        chosen.add("total")
        chosen.difference_update(_BCA_PARTS)
    return chosen


def resolve_device(device: str | None = None) -> str:
    """'gpu', 'gpu:<id>' or whatever other kind was asked for (commands.analyze_ct refuses everything but gpu)."""
    requested = device if device else os.environ.get("DEVICE", "gpu")
    kind, _, index = requested.partition(":")
    kind = "gpu" if kind == "cuda" else kind
    index = index if index else os.environ.get("NVIDIA_ID", "")
    if kind != "gpu" or not index:
        return kind
    os.environ.setdefault("NVIDIA_VISIBLE_DEVICES", index)
    return f"gpu:{index}"
