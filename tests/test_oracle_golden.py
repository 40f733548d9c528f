"""CPU: the oracle restatement against golden vectors produced by the REFERENCE's own functions
(tests/golden/make_golden.py), and the product's host logic against both."""
import hashlib
import json
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _close(a, b, path=""):
    if isinstance(a, dict):
        assert isinstance(b, dict) and list(a.keys()) == list(b.keys()), f"{path}: keys differ {list(a)[:5]} vs {list(b)[:5]}"
        for k in a:
            _close(a[k], b[k], f"{path}/{k}")
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            _close(x, y, f"{path}[{i}]")
    elif a is None or isinstance(a, (bool, str)):
        assert a == b, f"{path}: {a!r} != {b!r}"
    else:
        assert b is not None, f"{path}: {a} vs None"
        assert np.isclose(float(a), float(b), rtol=1e-11, atol=1e-11), f"{path}: {a} != {b}"


def test_sliding_window_steps_match_reference():
    from boa_b200.geometry import compute_steps_for_sliding_window as prod_steps, sliding_window_origins
    from oracle.sliding_window import compute_steps_for_sliding_window as orc_steps, sliding_window_slicers
    doc = json.load(open(os.path.join(G, "geometry.json")))
    for c in doc["steps"]:
        assert orc_steps(c["image"], c["patch"], c["step"]) == c["steps"]
        assert prod_steps(c["image"], c["patch"], c["step"]) == c["steps"]
        origins = sliding_window_origins(c["image"], c["patch"], c["step"])
        sl = sliding_window_slicers(c["image"], c["patch"], c["step"])
        assert [tuple(int(s.start) for s in t) for t in sl] == [tuple(map(int, o)) for o in origins]
    # the counts BASELINE.md quotes
    assert len(sliding_window_origins((512, 512, 512), (128,) * 3, 0.8)) == 125
    assert len(sliding_window_origins((300, 512, 512), (128,) * 3, 0.8)) == 75
    assert len(sliding_window_origins((154, 512, 512), (128,) * 3, 0.5)) == 98


def test_gaussian_matches_reference():
    from boa_b200.geometry import compute_gaussian as prod_g
    from oracle.sliding_window import compute_gaussian as orc_g
    z = np.load(os.path.join(G, "gaussian.npz"))
    for key in z.files:
        tile = tuple(int(t) for t in key.split("x"))
        assert np.array_equal(orc_g(tile, 1.0 / 8, 10), z[key])
        assert np.array_equal(prod_g(tile, 1.0 / 8, 10.0), z[key])
    doc = json.load(open(os.path.join(G, "geometry.json")))["gaussian128"]
    g = prod_g((128, 128, 128))
    assert hashlib.sha256(g.tobytes()).hexdigest() == doc["sha256"]
    assert float(g.max()) == doc["max"] == 10.0 and float(g.min()) == doc["min"]


def test_ct_normalization_matches_reference():
    from oracle.passes import ct_normalize
    z = np.load(os.path.join(G, "ct_norm.npz"))
    props = json.loads(str(z["props"]))
    assert np.array_equal(ct_normalize(z["x"], props), z["y"])


def test_tissue_rules_match_reference():
    from oracle.passes import subclassify_tissues
    z = np.load(os.path.join(G, "tissue.npz"))
    assert np.array_equal(subclassify_tissues(z["ct"], z["regions"]), z["tissues"])


def test_total_measurements_match_reference():
    from oracle.report import compute_measurements
    z = np.load(os.path.join(G, "phantom.npz"))
    gold = json.load(open(os.path.join(G, "measurements.json")))
    got = compute_measurements(z["ct"], z["total"], tuple(z["spacing"]), cnr_adjustment=True)
    assert np.array_equal(got.pop("_ct_pfav_mask"), z["ct_pfav"])
    assert len(gold["segmentations"]["total"]) == 304
    _close(gold, got)


def test_bca_report_matches_reference():
    from oracle.report import bca_json
    z = np.load(os.path.join(G, "phantom.npz"))
    zb = np.load(os.path.join(G, "phantom_bca.npz"))
    gold = json.load(open(os.path.join(G, "bca.json")))
    js, vert = bca_json(z["ct"], zb["tissues"], z["parts"], z["regions"], z["total"], tuple(zb["spacing"]))
    _close(gold["json"], js)
    assert {k: list(v) for k, v in vert.items()} == gold["vertebrae"]


def test_product_host_statistics_match_reference():
    """The product's host side (histogram -> metrics, slice tables -> report) fed with oracle-computed tables."""
    from boa_b200 import bca as pbca
    from boa_b200.measurements import HU_MIN, N_BINS, metrics_from_hist
    from oracle import passes as op
    z = np.load(os.path.join(G, "phantom.npz"))
    zb = np.load(os.path.join(G, "phantom_bca.npz"))
    gold = json.load(open(os.path.join(G, "measurements.json")))
    ct, total = z["ct"], z["total"]
    am, asd = gold["info"]["autochthon_mean"], gold["info"]["autochthon_std"]
    from boa_b200.labels import measurement_label_map
    lm = measurement_label_map("total")
    assert list(lm.keys()) == [k for k in gold["segmentations"]["total"] if not k.startswith("ct_pfav") and k != "autochthon"]
    for region, label in list(lm.items())[::7]:
        h = np.bincount(ct[total == label].astype(np.int64) - HU_MIN, minlength=N_BINS)
        _close(gold["segmentations"]["total"][region], metrics_from_hist(h, HU_MIN, am, asd, tuple(z["spacing"])), region)
    # report from tables
    sp = tuple(zb["spacing"])
    tc, th = op.slice_label_stats(zb["tissues"], 8, ct)
    tct, tht = op.slice_label_stats(zb["tissues"], 8, ct, z["parts"], 1)
    rc, _ = op.slice_label_stats(z["regions"], 12)
    oc, _ = op.slice_label_stats(total, 118)
    t = pbca.slice_tables_from_arrays(ct.shape[0], tc, th, tct, tht, rc, oc)
    js, vert, _ = pbca.build_bca_measurements(None, None, None, None, None, sp, tables=t)
    goldb = json.load(open(os.path.join(G, "bca.json")))
    _close(goldb["json"], js)
    assert {k: list(v) for k, v in vert.items()} == goldb["vertebrae"]
    ex = pbca.AggregatableBodyPart(goldb["body_part"])
    assert pbca.secondary_findings(t, ex, float(np.prod(sp) / 1000.0)) == goldb["other_findings"]


def test_postprocess_oracle_matches_reference_vectors():
    """oracle/postprocess.py against vectors produced by the reference's own postprocess_region_segmentation /
    remove_small_labeled_objects (tests/golden/make_golden_postprocess.py)."""
    from oracle import postprocess as opp
    z = np.load(os.path.join(G, "postprocess.npz"))
    i = 0
    while f"regions_in_{i}" in z:
        assert np.array_equal(opp.postprocess_region_segmentation(z[f"regions_in_{i}"]), z[f"regions_out_{i}"]), i
        i += 1
    assert i == 5
    i = 0
    while f"parts_in_{i}" in z:
        got = opp.remove_small_labeled_objects(z[f"parts_in_{i}"], int(z[f"parts_thr_{i}"]))
        assert np.array_equal(got, z[f"parts_out_{i}"]), i
        i += 1
    assert i == 4


def test_postprocess_on_the_5mm_grid_with_slice_weights_equals_the_full_grid():
    """The claim behind boa_b200.postprocess: replicating slices (order-0 zoom along z) maps components one to one, so
    post-processing the 5 mm map with per-slice weights and replicating afterwards equals the reference's order
    (replicate, then post-process)."""
    from scipy import ndimage
    from oracle import postprocess as opp
    z = np.load(os.path.join(G, "postprocess.npz"))
    for key, fn, kw in (("regions_in_0", opp.postprocess_region_segmentation, {}),
                        ("regions_in_2", opp.postprocess_region_segmentation, {}),
                        ("parts_in_0", opp.remove_small_labeled_objects, {"threshold": 130}),
                        ("parts_in_2", opp.remove_small_labeled_objects, {"threshold": 33})):
        low = z[key]
        z_out = int(round(low.shape[0] * 3.3))
        zoom = (z_out / low.shape[0], 1, 1)
        full = ndimage.zoom(low, zoom, order=0, mode="nearest")
        scale = (low.shape[0] - 1) / (z_out - 1)  # scipy's order-0 coordinate: floor(o * scale + 0.5)
        src = np.clip(np.floor(np.arange(z_out, dtype=np.float64) * scale + 0.5).astype(int), 0, low.shape[0] - 1)
        assert np.array_equal(full, low[src])
        w = np.bincount(src, minlength=low.shape[0])
        from boa_b200.postprocess import slice_weights
        assert np.array_equal(slice_weights(low.shape[0], z_out, "cpu").numpy(), w)
        ref = fn(full, **kw)
        got = fn(low, weights=w, **kw)[src]
        assert np.array_equal(got, ref), key


def test_breast_implant_finding_and_l3_axes_against_the_reference():
    """Golden vectors from the reference's own generate_secondary_findings / find_axes
    (tests/golden/make_golden_findings.py)."""
    from boa_b200 import bca as pbca
    from boa_b200 import ts_metrics
    gold = json.load(open(os.path.join(G, "findings.json")))
    for case in gold["implants"]:
        regions = np.load(os.path.join(G, f"implants_{case['n']}.npz"))["regions"]
        t = pbca.slice_tables_from_arrays(regions.shape[0], np.zeros((regions.shape[0], 8), dtype=np.int64),
                                          np.zeros((regions.shape[0], 8), dtype=np.int64),
                                          np.zeros((regions.shape[0], 8), dtype=np.int64),
                                          np.zeros((regions.shape[0], 8), dtype=np.int64),
                                          np.stack([(regions == r).sum(axis=(1, 2)) for r in range(12)], axis=1))
        got = pbca.secondary_findings(t, pbca.AggregatableBodyPart.THORAX, float(np.prod(case["spacing"]) / 1000.0),
                                      body_regions=regions)
        assert got == case["findings"], (case["n"], got)
    for k, ax in enumerate(gold["axes"]):
        sl = np.load(os.path.join(G, f"axes_{k}.npz"))["slice"]
        pts = ts_metrics.find_axes(sl)
        assert [list(map(float, p)) for p in pts] == ax["points"]
        major, minor = ts_metrics.axes_of_slice(sl, ax["spacing"])
        assert abs(major - ax["major_mm"]) < 1e-9 and abs(minor - ax["minor_mm"]) < 1e-9
    # the wrapper: middle slice of the L3 range, body mask = body_parts == 1
    sl = np.load(os.path.join(G, "axes_0.npz"))["slice"]
    total = np.zeros((9, *sl.shape), dtype=np.uint8)
    parts = np.zeros_like(total)
    total[3:6, 10:20, 10:20] = 29   # any label standing in for vertebrae_L3
    parts[4] = sl
    parts[3] = 1
    assert ts_metrics.major_minor_axis(total, parts, gold["axes"][0]["spacing"], l3_label=29) == \
        (gold["axes"][0]["major_mm"], gold["axes"][0]["minor_mm"])
    assert ts_metrics.major_minor_axis(np.zeros_like(total), parts, (1, 1), l3_label=29) == (None, None)


def test_nnunet_resampling_decisions_and_volumes_match_the_reference():
    """compute_new_shape / determine_do_sep_z_and_axis / resample_data_or_seg_to_shape of the reference
    (tests/golden/make_golden_resampling.py ran default_resampling.py itself) against the oracle restatement and the
    product's host logic."""
    from boa_b200.resample import nnunet_new_shape, nnunet_separate_z
    from oracle import resampling as orr
    with open(os.path.join(G, "resampling.json")) as f:
        g = json.load(f)
    assert g["aniso_threshold"] == orr.ANISO_THRESHOLD
    n_sep = 0
    for d in g["decisions"]:
        want = (d["separate_z"], d["axis"])
        assert orr.determine_do_sep_z_and_axis(d["current"], d["new"]) == want, d
        assert tuple(nnunet_separate_z(d["current"], d["new"])) == want, d
        n_sep += d["separate_z"]
        for shape, new_shape in zip(g["shapes"], d["new_shapes"]):
            assert list(orr.compute_new_shape(shape, d["current"], d["new"])) == new_shape
            assert list(nnunet_new_shape(shape, d["current"], d["new"])) == new_shape
    assert 0 < n_sep < len(g["decisions"])
    z = np.load(os.path.join(G, "resampling.npz"))
    for name in sorted(k[:-3] for k in z.files if k.endswith("_in")):
        meta = z[name + "_meta"]
        cur, new, order = tuple(meta[:3]), tuple(meta[3:6]), int(meta[6])
        ref = z[name + "_out"]
        got = orr.resample_data(z[name + "_in"], ref.shape[1:], cur, new, order=order)
        assert got.shape == ref.shape and got.dtype == ref.dtype, name
        assert np.array_equal(got, ref), (name, np.abs(got - ref).max())


def test_crop_box_matches_the_reference_bbox():
    """The crop pre-pass box (bounding box of the ROI mask + 20 mm, cropping.py) against the reference's
    get_bbox_from_mask, incl. the float32 zoom that makes 20 mm / 0.8 mm 24 voxels."""
    from boa_b200.pipeline import grow_crop_box
    with open(os.path.join(G, "crop.json")) as f:
        cases = json.load(f)
    assert len(cases) >= 20
    for c in cases:
        got = grow_crop_box([tuple(fl) for fl in c["first_last"]], c["shape"], c["zooms"])
        assert [list(b) for b in got] == c["bbox"], c
    assert grow_crop_box([(30, 40)], (100,), (0.8,)) == [(6, 65)]


def _change_spacing_cases():
    z = np.load(os.path.join(G, "change_spacing.npz"))
    for name in sorted(k[:-3] for k in z.files if k.endswith("_ct")):
        yield name, {k[len(name) + 1:]: z[k] for k in z.files if k.startswith(name + "_")}


def test_change_spacing_shapes_and_scipy_restatement_match_the_reference():
    """Outputs of the reference's change_spacing (tests/golden/make_golden_change_spacing.py): the product's shape
    arithmetic (float32 header zooms) and the scipy call the GPU tests compare the kernels with reproduce them exactly."""
    from scipy import ndimage
    from boa_b200.resample import resampled_depth, zoomed_shape
    n = 0
    for name, c in _change_spacing_cases():
        ct, sp, tg, ref = c["ct"], c["spacing_zyx"], c["target_zyx"], c["resampled"]
        if name.startswith("thickness"):
            assert resampled_depth(ct.shape[0], sp[0], tg[0]) == ref.shape[0] and ref.shape[1:] == ct.shape[1:]
        else:
            assert zoomed_shape(ct.shape, sp, float(tg[0])) == ref.shape, name
        zoom = [np.float64(np.float32(s)) / np.float64(np.float32(t)) for s, t in zip(sp, tg)]
        # in the reference's own memory order ([x, y, z]) the restated call is the reference's call: bit-equal
        xyz = ndimage.zoom(ct.transpose(2, 1, 0).astype(np.float64), zoom[::-1], order=3, mode="nearest")
        assert np.array_equal(xyz.astype(np.int32).transpose(2, 1, 0), ref), name
        # evaluated on the [z, y, x] array (what the GPU tests do) the separable passes run in another order: the spline
        # values agree to ~1e-11, so the truncated integers differ (by one) only where the value is numerically an integer
        zyx = ndimage.zoom(ct.astype(np.float64), zoom, order=3, mode="nearest")
        diff = np.abs(zyx.astype(np.int32) - ref)
        assert diff.max() <= 1 and np.all(np.abs(zyx - np.rint(zyx))[diff != 0] < 1e-6), name
        back = ndimage.zoom(c["labels"], np.array(ct.shape) / np.array(ref.shape), order=0, mode="nearest")
        assert np.array_equal(back, c["labels_back"]), name
        n += 1
    assert n == 4
