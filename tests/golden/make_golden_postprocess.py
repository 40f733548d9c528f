"""Golden vectors for the connected-component post-processing of the body-composition label maps, produced by the
reference's OWN functions (run in the development container, where /root/reference exists):

  body_composition_analysis/body_regions/postprocess.py::postprocess_region_segmentation
  body_composition_analysis/body_parts/postprocess.py::remove_small_labeled_objects   (real OpenCV for the contours)

scikit-image is not installed here: `skimage.measure.label / regionprops` and
`skimage.morphology.remove_small_objects(max_size=, connectivity=)` are restated on scipy.ndimage.label (what
scikit-image itself calls for boolean input) - the only arithmetic in these vectors that is not the reference's.

    python tests/golden/make_golden_postprocess.py        # writes tests/golden/postprocess.npz
"""
import os
import sys
import types

import numpy as np
from scipy import ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_stubs as S  # noqa: E402


def install_skimage():
    measure = sys.modules["skimage.measure"]

    class Prop:
        def __init__(self, label, area):
            self.label, self.area = label, area

    def regionprops(lab):
        counts = np.bincount(lab.ravel())
        return [Prop(i, int(counts[i])) for i in range(1, len(counts)) if counts[i] > 0]

    measure.regionprops = regionprops  # label(): full connectivity, numbered in raster order (already stubbed)
    morph = sys.modules["skimage.morphology"]

    def remove_small_objects(ar, min_size=64, connectivity=1, *, max_size=None, out=None):
        assert ar.dtype == bool and max_size is not None
        lab, _ = ndimage.label(ar, structure=ndimage.generate_binary_structure(ar.ndim, connectivity))
        sizes = np.bincount(lab.ravel())
        small = sizes <= max_size          # scikit-image 0.26: "this number of pixels or fewer"
        small[0] = False
        if out is None:
            out = ar.copy()
        out[small[lab]] = False
        return out

    morph.remove_small_objects = remove_small_objects


def blobs(shape, n_labels, seed, sigma=2.0, density=0.45):
    rng = np.random.default_rng(seed)
    field = ndimage.gaussian_filter(rng.standard_normal(shape), sigma)
    lab_field = ndimage.gaussian_filter(rng.standard_normal(shape), 3 * sigma)
    fg = field > np.quantile(field, 1 - density)
    edges = np.quantile(lab_field, np.linspace(0, 1, n_labels + 1)[1:-1])
    lab = (np.digitize(lab_field, edges) + 1).astype(np.uint8)
    return np.where(fg, lab, 0).astype(np.uint8)


def main():
    S.install()
    install_skimage()
    reg = S.load_by_path("ref_regions_post", S.EXT + "/body_composition_analysis/body_regions/postprocess.py")
    par = S.load_by_path("ref_parts_post", S.EXT + "/body_composition_analysis/body_parts/postprocess.py")
    out = {}
    # ---- regions: labels 0..11 in blobs; one case with two components of exactly the same size (tie -> first in
    #      raster order survives), one with a single component, one empty
    cases = [blobs((18, 40, 44), 11, 1, 1.6, 0.35), blobs((12, 36, 30), 11, 2, 1.2, 0.25)]
    tie = np.zeros((6, 20, 20), np.uint8)
    tie[1:3, 2:6, 2:6] = 3
    tie[3:5, 12:16, 10:14] = 3          # same size, later in raster order -> 255
    tie[1:3, 12:14, 2:4] = 7
    tie[4, 4:6, 15:19] = 7              # pericardium: 8 vs 8 voxels
    cases += [tie, np.where(np.ones((5, 9, 9), bool), 4, 0).astype(np.uint8), np.zeros((4, 8, 8), np.uint8)]
    for i, seg in enumerate(cases):
        res = reg.postprocess_region_segmentation(S.FakeImage(seg)).arr
        out[f"regions_in_{i}"], out[f"regions_out_{i}"] = seg, res.astype(np.uint8)
    # ---- parts: holes inside slices, 3-D cavities, small objects; small thresholds for the small volumes
    pcases = [(blobs((16, 44, 40), 6, 3, 1.8, 0.5), 40), (blobs((10, 30, 34), 6, 4, 1.1, 0.4), 25)]
    ring = np.zeros((7, 24, 24), np.uint8)
    ring[1:6, 3:21, 3:21] = 2
    ring[1:6, 7:17, 7:17] = 0           # a tube: every slice is a ring -> the contour fill closes it
    ring[2:5, 10:14, 10:14] = 5         # another label inside the tube (painted after label 2 -> stays)
    ring[3, 0:2, 0:2] = 1               # 4 voxels: below the threshold -> removed
    diag = np.zeros((3, 12, 12), np.uint8)
    for k in range(1, 6):               # a diamond outline: 8-connected ring with diagonal steps only
        diag[1, 6 - k + 0, 6 + (5 - k) * 0 + k - 5 + 5] = 0
    yy, xx = np.mgrid[0:12, 0:12]
    diag[1][np.abs(yy - 6) + np.abs(xx - 6) == 4] = 3
    pcases += [(ring, 10), (diag, 1)]
    for i, (seg, thr) in enumerate(pcases):
        res = par.remove_small_labeled_objects(seg.copy(), threshold=thr)
        out[f"parts_in_{i}"], out[f"parts_out_{i}"], out[f"parts_thr_{i}"] = seg, res.astype(np.uint8), np.int32(thr)
    np.savez_compressed(os.path.join(HERE, "postprocess.npz"), **out)
    for k in sorted(out):
        if "_out_" in k:
            src = out[k.replace("_out_", "_in_")]
            print(k, out[k].shape, "changed voxels:", int((out[k] != src).sum()), "labels:", np.unique(out[k]).tolist())


if __name__ == "__main__":
    main()
