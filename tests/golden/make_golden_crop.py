"""Golden vectors for the crop pre-pass geometry, from the REFERENCE's own
_external/totalsegmentator/cropping.py (get_bbox_from_mask with the mm -> voxel addon of crop_to_mask, :11-37,97-99):
random ROI masks in volumes of several spacings (zooms as the float32 values a NIfTI header returns).

    python tests/golden/make_golden_crop.py      # needs /root/reference
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_stubs as S  # noqa: E402


def main():
    S.install()
    sys.modules["nibabel.processing"] = types.ModuleType("nibabel.processing")
    sys.modules["nibabel"].processing = sys.modules["nibabel.processing"]
    sys.modules["nibabel"].Nifti1Image = object
    m = types.ModuleType("totalsegmentator")
    m.__path__ = [S.EXT + "/totalsegmentator"]
    sys.modules["totalsegmentator"] = m
    from totalsegmentator.cropping import get_bbox_from_mask
    rng = np.random.default_rng(4)
    cases = []
    for zooms in ((0.8, 0.8, 2.5), (1.5, 1.5, 1.5), (0.7, 0.7, 5.0), (0.976562, 0.976562, 3.0), (2.0, 1.0, 0.5)):
        z32 = tuple(np.float32(v) for v in zooms)           # nibabel's header.get_zooms()
        for _ in range(4):
            shape = tuple(int(v) for v in rng.integers(20, 90, size=3))
            lo = [int(rng.integers(0, s - 3)) for s in shape]
            hi = [int(rng.integers(l + 1, s)) for l, s in zip(lo, shape)]
            mask = np.zeros(shape, np.float64)
            mask[lo[0]:hi[0] + 1, lo[1]:hi[1] + 1, lo[2]:hi[2] + 1] = rng.random((hi[0] - lo[0] + 1, hi[1] - lo[1] + 1, hi[2] - lo[2] + 1)) > 0.7
            mask[lo[0], lo[1], lo[2]] = mask[hi[0], hi[1], hi[2]] = 1
            addon = (np.array([20, 20, 20]) / z32).astype(int)   # crop_to_mask :99
            bbox = get_bbox_from_mask(mask, outside_value=0, addon=addon)
            cases.append({"zooms": [float(v) for v in zooms], "shape": shape, "first_last": [[l, h] for l, h in zip(lo, hi)],
                          "bbox": [[int(a), int(b)] for a, b in bbox]})
    with open(os.path.join(HERE, "crop.json"), "w") as f:
        json.dump(cases, f)
    print(len(cases), "cases; addon voxels at 0.8 mm:", int((np.array([20]) / np.float32(0.8)).astype(int)[0]))


if __name__ == "__main__":
    main()
