"""Label tables of the reference (exported by tools/export_class_maps.py from
_external/totalsegmentator/map_to_binary.py) and the LUTs derived from them."""
from __future__ import annotations

import json
import os
from functools import lru_cache

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "class_maps.json")

TOTAL_TASK_IDS = [291, 292, 293, 294, 295]           # totalsegmentator/python_api.py:182-189
TOTAL_FAST_TASK_ID = 297                              # --fast-total: one 3 mm model, all 117 classes (python_api.py:169-175)
BODY_REGIONS_TASK_ID, BODY_PARTS_TASK_ID = 542, 543  # body_composition_analysis/tasks.py:15-48

# Tasks behind a crop pre-pass (totalsegmentator/python_api.py:236-330,673-736): a rough 6 mm `total` segmentation
# (model 298) gives the bounding box of the listed structures (+ 20 mm, :726), the task's own network then runs on the
# cropped volume at its native spacing.  task -> (dataset id, folds (None = every fold in the folder), crop structures)
CROP_PREPASS_TASK_ID = 298
_LUNG_LOBES = ["lung_upper_lobe_left", "lung_lower_lobe_left", "lung_upper_lobe_right", "lung_middle_lobe_right",
               "lung_lower_lobe_right"]
CROP_TASKS = {
    "lung_vessels": (258, [0], _LUNG_LOBES),
    "cerebral_bleed": (150, [0], ["brain"]),
    "hip_implant": (260, [0], ["femur_left", "femur_right", "hip_left", "hip_right"]),
    "pleural_pericard_effusion": (315, None, _LUNG_LOBES),
    "liver_vessels": (8, [0], ["liver"]),
}
CROP_ADDON_MM = 20.0

# body_composition_analysis/body_regions/definition.py, body_parts/definition.py, tissue/definition.py
BODY_REGION = {"SUBCUTANEOUS_TISSUE": 1, "MUSCLE": 2, "ABDOMINAL_CAVITY": 3, "THORACIC_CAVITY": 4, "BONE": 5,
               "GLANDS": 6, "PERICARDIUM": 7, "BREAST_IMPLANT": 8, "MEDIASTINUM": 9, "BRAIN": 10, "NERVOUS_SYSTEM": 11}
BODY_PART_TORSO = 1
TISSUES = {"MUSCLE": 1, "BONE": 2, "SAT": 3, "VAT": 4, "IMAT": 5, "PAT": 6, "EAT": 7}


@lru_cache(maxsize=1)
def tables() -> dict:
    with open(_DATA) as f:
        return json.load(f)


def class_map(task: str) -> dict[int, str]:
    return {int(k): v for k, v in tables()["class_map_all_keys"][task].items()}


def part_luts() -> list[list[int]]:
    """For each of the five `total` part models: LUT part-class index -> global `total` label
    (totalsegmentator/nnunet.py:534-556: class_map_inv[class_map_5_parts[part][jdx]])."""
    t = tables()
    inv = {v: int(k) for k, v in t["class_map"]["total"].items()}
    luts = []
    for tid in TOTAL_TASK_IDS:
        part = t["class_map_5_parts"][t["map_taskid_to_partname_ct"][str(tid)]]
        lut = [0] * (len(part) + 1)
        for jdx, name in part.items():
            lut[int(jdx)] = inv[name]
        luts.append(lut)
    return luts


def measurement_label_map(model_name: str) -> dict[str, int]:
    """compute/measurements.py:290-294: every `<task>_<name>` key of the complete reverse class map that starts with
    the model name (and not `<model>_v2`), stripped of `<model>_` - in class_map definition order."""
    t = tables()
    out = {}
    for task in t["class_map_order"]:
        for idx, name in t["class_map_all_keys"][task].items():
            key = f"{task}_{name}"
            if key.startswith(model_name) and not key.startswith(model_name + "_v2"):
                out[key[len(model_name) + 1:]] = int(idx)
    return out
