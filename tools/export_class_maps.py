"""Export the label tables of the reference (data, not code) into the package, so that output JSON keys and the
part -> global label LUTs match the reference exactly.  Source: _external/totalsegmentator/map_to_binary.py
(class_map, class_map_5_parts, map_taskid_to_partname_ct) - imported from /root/reference; run in the dev container."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden"))
import _ref_stubs as S

S.install()
from totalsegmentator.map_to_binary import class_map, class_map_5_parts, map_taskid_to_partname_ct  # noqa: E402

keep = {k: {str(i): n for i, n in v.items()} for k, v in class_map.items()
        if k.startswith("total") or k in ("body_regions", "body_parts", "bca")}
out = {
    "class_map": keep,
    "class_map_order": list(class_map.keys()),  # measurements.py iterates class_map in definition order
    "class_map_all_keys": {k: {str(i): n for i, n in v.items()} for k, v in class_map.items()},
    "class_map_5_parts": {k: {str(i): n for i, n in v.items()} for k, v in class_map_5_parts.items()},
    "map_taskid_to_partname_ct": {str(k): v for k, v in map_taskid_to_partname_ct.items()},
}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "body-and-organ-analysis_b200", "data")
os.makedirs(dst, exist_ok=True)
with open(os.path.join(dst, "class_maps.json"), "w") as f:
    json.dump(out, f, indent=0, sort_keys=False)
print({k: len(v) for k, v in keep.items()}, "all keys:", len(out["class_map_all_keys"]))
