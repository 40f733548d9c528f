#!/bin/bash
set -u
timeout 600 python -m pytest tests/test_gpu_postprocess.py -x -q 2>&1 | tail -15
timeout 300 python - <<'PY'
import sys, time; sys.path.insert(0, '.')
import numpy as np, torch
from scipy import ndimage
from boa_b200 import postprocess as pp
rng = np.random.default_rng(0)
shape = (154, 512, 512)
field = ndimage.gaussian_filter(rng.standard_normal((39, 128, 128)), 2.0)
field = np.repeat(np.repeat(np.repeat(field, 4, 0), 4, 1), 4, 2)[:154]
lab = (np.digitize(ndimage.gaussian_filter(rng.standard_normal((39, 128, 128)), 6.0), [-0.01, 0.0, 0.01]) + 1)
lab = np.repeat(np.repeat(np.repeat(lab, 4, 0), 4, 1), 4, 2)[:154]
regions = torch.from_numpy(np.where(field > -0.02, lab * 3 - 2, 0).astype(np.uint8)).cuda()
parts = torch.from_numpy(np.where(field > -0.02, lab, 0).astype(np.uint8)).cuda()
w = pp.slice_weights(154, 512, "cuda")
for name, fn, arg in (("regions", pp.postprocess_region_segmentation, regions), ("parts", pp.postprocess_part_segmentation, parts)):
    fn(arg, weights=w); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3): out = fn(arg, weights=w)
    torch.cuda.synchronize()
    print(f"{name} post-processing on 154x512x512 (fg {float((arg != 0).float().mean()):.2f}): {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms, labels {torch.unique(out).tolist()}")
PY
