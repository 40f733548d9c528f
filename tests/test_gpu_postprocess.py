"""GPU: connected-component post-processing of the body-composition label maps against the reference's own outputs
(golden vectors) and the oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from boa_b200 import postprocess as pp
from oracle import postprocess as opp

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_regions_golden(cuda):
    z = np.load(os.path.join(G, "postprocess.npz"))
    i = 0
    while f"regions_in_{i}" in z:
        got = pp.postprocess_region_segmentation(_dev(z[f"regions_in_{i}"])).cpu().numpy()
        assert np.array_equal(got, z[f"regions_out_{i}"]), i
        i += 1


def test_parts_golden(cuda):
    z = np.load(os.path.join(G, "postprocess.npz"))
    i = 0
    while f"parts_in_{i}" in z:
        got = pp.postprocess_part_segmentation(_dev(z[f"parts_in_{i}"]), threshold=int(z[f"parts_thr_{i}"])).cpu().numpy()
        assert np.array_equal(got, z[f"parts_out_{i}"]), i
        i += 1


@pytest.mark.parametrize("shape,seed", [((37, 96, 88), 0), ((64, 128, 120), 1), ((9, 33, 257), 2)])
def test_random_volumes_against_oracle(cuda, shape, seed):
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    field = ndimage.gaussian_filter(rng.standard_normal(shape), 1.7)
    lab = (np.digitize(ndimage.gaussian_filter(rng.standard_normal(shape), 5.0), [-0.02, 0.0, 0.02]) + 1) * 3 - 2
    regions = np.where(field > 0.02, lab, 0).astype(np.uint8)          # labels 1, 4, 7, 10 in blobs
    got = pp.postprocess_region_segmentation(_dev(regions)).cpu().numpy()
    assert np.array_equal(got, opp.postprocess_region_segmentation(regions))
    parts = np.where(field > -0.01, (lab + 2) // 3, 0).astype(np.uint8)  # labels 1..4
    got = pp.postprocess_part_segmentation(_dev(parts), threshold=300).cpu().numpy()
    assert np.array_equal(got, opp.remove_small_labeled_objects(parts, 300))


def test_slice_weights_reproduce_full_grid(cuda):
    """Post-processing the 5 mm map with slice weights, then replicating == the reference's order."""
    from scipy import ndimage
    from boa_b200.resample import upsample_labels_nearest
    rng = np.random.default_rng(5)
    shape, z_out = (31, 72, 80), 103
    field = ndimage.gaussian_filter(rng.standard_normal(shape), 1.5)
    lab = np.digitize(ndimage.gaussian_filter(rng.standard_normal(shape), 4.0), [-0.03, 0.0, 0.03]) + 1
    low = np.where(field > 0.0, lab, 0).astype(np.uint8)
    w = pp.slice_weights(shape[0], z_out, "cuda")
    full = upsample_labels_nearest(_dev(low), z_out)
    assert int(w.sum()) == z_out
    for fn, ofn, kw, okw in ((pp.postprocess_region_segmentation, opp.postprocess_region_segmentation, {}, {}),
                             (pp.postprocess_part_segmentation, opp.remove_small_labeled_objects, {"threshold": 500},
                              {"threshold": 500})):
        ref = ofn(full.cpu().numpy(), **okw)                               # the reference's order, on the CPU
        got = upsample_labels_nearest(fn(_dev(low), weights=w, **kw), z_out).cpu().numpy()
        assert np.array_equal(got, ref)
        assert np.array_equal(fn(full, **kw).cpu().numpy(), ref)           # and the plain full-grid device path


def test_errors(cuda):
    with pytest.raises(ValueError):
        pp.postprocess_region_segmentation(torch.zeros((4, 4, 4), dtype=torch.int32, device="cuda"))
    with pytest.raises(ValueError):
        pp.postprocess_region_segmentation(torch.zeros((4, 4, 4), dtype=torch.uint8, device="cuda"),
                                           weights=torch.ones(3, dtype=torch.int32, device="cuda"))


def test_repeated_runs_are_identical_on_a_solid_volume(cuda):
    """A solid region (one root for millions of voxels) with speckle around it: the concurrent union-find must give the
    same, correct answer every time (regression test for a late path-halving store overwriting a final root)."""
    from scipy import ndimage
    rng = np.random.default_rng(3)
    shape = (40, 256, 256)
    solid = ndimage.gaussian_filter(rng.standard_normal(shape), 6.0) > -0.01
    speckle = rng.random(shape) < 0.08
    seg = np.where(solid | speckle, 3, 0).astype(np.uint8)
    seg[rng.random(shape) < 0.02] = 7
    ref = opp.postprocess_region_segmentation(seg)
    dev = _dev(seg)
    for _ in range(5):
        assert np.array_equal(pp.postprocess_region_segmentation(dev).cpu().numpy(), ref)
