"""GPU parity (bit-exact) of the HBM-bound passes against the oracle and the reference-generated golden vectors,
through the C ABI."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from boa_b200 import passes
from boa_b200.predictor import finalize_argmax, weight_sum
from oracle import passes as op

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_ct_normalize_golden(cuda):
    z = np.load(os.path.join(G, "ct_norm.npz"))
    p = json.loads(str(z["props"]))
    for x in (z["x"], z["x"].astype(np.float32)):
        y = passes.ct_normalize(_dev(x), p["percentile_00_5"], p["percentile_99_5"], p["mean"], p["std"])
        assert np.array_equal(y.cpu().numpy(), z["y"])


@pytest.mark.parametrize("n", [0, 1, 7, 16, 1000, 4099, 1 << 20])
def test_ct_normalize_ragged(cuda, n):
    rng = np.random.default_rng(n)
    x = rng.integers(-2000, 4000, size=n).astype(np.int16)
    props = {"mean": 12.5, "std": 300.25, "percentile_00_5": -900.0, "percentile_99_5": 1500.0}
    y = passes.ct_normalize(_dev(x), -900.0, 1500.0, 12.5, 300.25) if n else torch.empty(0)
    assert np.array_equal(y.cpu().numpy(), op.ct_normalize(x, props))


def test_tissue_golden_and_random(cuda):
    z = np.load(os.path.join(G, "tissue.npz"))
    t = passes.tissue_subclassify(_dev(z["ct"]), _dev(z["regions"]))
    assert np.array_equal(t.cpu().numpy(), z["tissues"])
    rng = np.random.default_rng(1)
    for shape in [(3, 5, 7), (16, 64, 64), (9, 33, 31)]:
        ct = rng.integers(-1100, 3100, size=shape).astype(np.int16)
        reg = rng.integers(0, 12, size=shape).astype(np.uint8)
        for c in (ct, ct.astype(np.float32)):
            t = passes.tissue_subclassify(_dev(c), _dev(reg))
            assert np.array_equal(t.cpu().numpy(), op.subclassify_tissues(ct, reg))


@pytest.mark.parametrize("shape,L", [((5, 16, 16), 8), ((7, 33, 29), 12), ((3, 300, 300), 118), ((2, 512, 512), 8)])
def test_slice_label_stats(cuda, shape, L):
    rng = np.random.default_rng(2)
    lab = rng.integers(0, L + 3, size=shape).astype(np.uint8)   # labels >= L are ignored
    ct = rng.integers(-1024, 3071, size=shape).astype(np.int16)
    mask = rng.integers(0, 3, size=shape).astype(np.uint8)
    for m in (None, mask):
        c, s = passes.slice_label_stats(_dev(lab), L, ct=_dev(ct), mask=None if m is None else _dev(m), mask_value=1)
        oc, osum = op.slice_label_stats(lab, L, ct, m, 1)
        assert np.array_equal(c.cpu().numpy(), oc) and np.array_equal(s.cpu().numpy(), osum)
    c, s = passes.slice_label_stats(_dev(lab), L)
    assert s is None and np.array_equal(c.cpu().numpy(), op.slice_label_stats(lab, L)[0])


def test_label_hist_and_measurements_golden(cuda):
    from boa_b200.measurements import compute_measurements_on_device
    z = np.load(os.path.join(G, "phantom.npz"))
    gold = json.load(open(os.path.join(G, "measurements.json")))
    ct, total = _dev(z["ct"]), _dev(z["total"])
    got, pfav = compute_measurements_on_device(ct, {"total": total}, tuple(z["spacing"]), cnr_adjustment=True,
                                               return_ct_pfav_mask=True)
    assert np.array_equal(pfav.cpu().numpy(), z["ct_pfav"])
    from test_oracle_golden import _close
    _close(gold, got)


def test_erode_matches_oracle(cuda):
    rng = np.random.default_rng(3)
    for shape in [(20, 20, 20), (9, 31, 17), (6, 6, 6), (40, 12, 50)]:
        m = (rng.random(shape) < 0.97).astype(np.uint8)
        got = passes.erode_box(_dev(m), 3, 2).cpu().numpy()
        assert np.array_equal(got.astype(bool), op.erode_region(m.astype(bool)))


def test_bca_report_golden(cuda):
    from boa_b200 import bca
    z = np.load(os.path.join(G, "phantom.npz"))
    zb = np.load(os.path.join(G, "phantom_bca.npz"))
    gold = json.load(open(os.path.join(G, "bca.json")))
    ct, regions, parts, total = _dev(z["ct"]), _dev(z["regions"]), _dev(z["parts"]), _dev(z["total"])
    tissues = bca.subclassify_tissues(ct, regions)
    assert np.array_equal(tissues.cpu().numpy(), zb["tissues"])
    js, vert, _ = bca.build_bca_measurements(ct, tissues, parts, regions, total, tuple(zb["spacing"]))
    from test_oracle_golden import _close
    _close(gold["json"], js)
    assert {k: list(v) for k, v in vert.items()} == gold["vertebrae"]


def test_finalize_argmax_ties_lut_merge_and_inf(cuda):
    rng = np.random.default_rng(4)
    C, shape = 6, (8, 12, 16)
    acc = rng.standard_normal((C, *shape)).astype(np.float32)
    acc[3][acc[1] > 0.5] = acc[1][acc[1] > 0.5]          # exact ties: first maximum must win
    acc[:, 0, 0, :] = 0.0                                  # all equal -> class 0
    w = rng.uniform(0.1, 10.0, size=shape).astype(np.float32)
    ref = (acc / w).argmax(0).astype(np.uint8)
    lab = finalize_argmax(_dev(acc), _dev(w)).cpu().numpy()
    assert np.array_equal(lab, ref)
    lut = [0, 10, 0, 30, 40, 50]
    prev = rng.integers(0, 5, size=shape).astype(np.uint8)
    merged = prev.copy()
    mapped = np.array(lut, dtype=np.uint8)[ref]
    merged[mapped != 0] = mapped[mapped != 0]
    got = finalize_argmax(_dev(acc), _dev(w), lut, _dev(prev), True).cpu().numpy()
    assert np.array_equal(got, merged)
    bad = acc.copy()
    bad[2, 1, 1, 1] = np.inf
    with pytest.raises(RuntimeError, match="inf"):
        finalize_argmax(_dev(bad), _dev(w))


@pytest.mark.parametrize("shape,C", [((31, 33, 35), 25), ((13, 7, 5), 3), ((16, 24, 32), 27), ((9, 10, 11), 1)])
def test_finalize_argmax_odd_shapes_near_ties_and_tiny_values(cuda, shape, C):
    """V % 4 != 0 takes the scalar kernel; the vector kernel's division-free fast path must agree with numpy's
    divide-then-argmax on near ties (quotients that round to the same float), denormal quotients and zeros."""
    rng = np.random.default_rng(11)
    acc = rng.standard_normal((C, *shape)).astype(np.float32) * 5
    w = rng.uniform(5.96e-8, 80.0, size=shape).astype(np.float32)
    if C > 2:
        m = rng.random(shape) < 0.3                       # channel 2 one ulp above / below channel 0
        acc[2][m] = np.nextafter(acc[0][m], np.float32(np.inf))
        m = rng.random(shape) < 0.3
        acc[1][m] = np.nextafter(acc[2][m], np.float32(-np.inf))
        acc[:, 0] *= np.float32(1e-42)                    # quotients underflow into ties
        acc[:, 1] = 0.0
        acc[:, 2] = -np.abs(acc[:, 2])                    # all negative
    ref = (acc / w).argmax(0).astype(np.uint8)
    lab = finalize_argmax(_dev(acc), _dev(w)).cpu().numpy()
    assert np.array_equal(lab, ref)
    # a slab that starts at a misaligned element offset (what a sharded volume hands in)
    if shape[0] > 4:
        a = _dev(acc)[:, 1:-1].contiguous()
        wv = _dev(w)[1:-1]                                # storage offset = Y * X elements: not 16-byte aligned when odd
        lab2 = finalize_argmax(a, wv).cpu().numpy()
        assert np.array_equal(lab2, ref[1:-1])


def test_weight_sum_and_accumulate_patch(cuda):
    from boa_b200 import _lib
    from boa_b200.geometry import compute_gaussian, sliding_window_origins
    import ctypes as C
    patch, shape = (16, 16, 16), (40, 24, 33)
    origins = sliding_window_origins(shape, patch, 0.5)
    g = compute_gaussian(patch).astype(np.float32)
    gd = _dev(g)
    w = weight_sum(shape, patch, origins, gd).cpu().numpy()
    ref = np.zeros(shape, dtype=np.float32)
    for o in origins:
        ref[o[0]:o[0] + 16, o[1]:o[1] + 16, o[2]:o[2] + 16] += g
    assert np.array_equal(w, ref)
    rng = np.random.default_rng(5)
    Cn = 3
    acc = torch.zeros((Cn, *shape), device="cuda")
    racc = np.zeros((Cn, *shape), dtype=np.float32)
    for o in origins[:5]:
        lg = rng.standard_normal((Cn, *patch)).astype(np.float32)
        _lib.check(_lib.lib().boa_accumulate_patch(_lib.ptr(_dev(lg)), Cn, _lib.i32x3(patch), _lib.i32x3(o), _lib.ptr(gd),
                                                   _lib.ptr(acc), _lib.i32x3(shape), _lib.stream_ptr()))
        racc[:, o[0]:o[0] + 16, o[1]:o[1] + 16, o[2]:o[2] + 16] += lg * g
    assert np.array_equal(acc.cpu().numpy(), racc)


def test_resample_thickness_vs_scipy(cuda):
    from scipy import ndimage
    from boa_b200.resample import resample_thickness, upsample_labels_nearest
    rng = np.random.default_rng(6)
    ct = rng.integers(-1000, 2000, size=(47, 20, 24)).astype(np.int16)
    ref = ndimage.zoom(ct.astype(np.float64), (np.float64(np.float32(1.5)) / 5.0, 1, 1), order=3, mode="nearest")
    got = resample_thickness(_dev(ct), 1.5, 5.0).cpu().numpy()
    assert got.shape == ref.shape
    # fp64 spline as scipy; values agree to ~1e-12 before truncation, so integers differ by at most 1 and only where
    # the spline value is (numerically) an integer
    diff = np.abs(got.astype(np.int64) - ref.astype(np.int32))
    assert diff.max() <= 1 and (diff != 0).mean() < 0.08
    inner = (np.abs(ref - np.rint(ref)) > 1e-6)
    assert np.array_equal(got[inner], ref.astype(np.int32)[inner])
    lab = rng.integers(0, 7, size=ref.shape).astype(np.uint8)
    up = upsample_labels_nearest(_dev(lab), 47).cpu().numpy()
    assert np.array_equal(up, ndimage.zoom(lab, (47 / lab.shape[0], 1, 1), order=0, mode="nearest"))


@pytest.mark.parametrize("shape,spacing", [((37, 44, 52), (2.0, 0.9765625, 0.9765625)),   # thick slices, fine in-plane
                                           ((40, 33, 29), (1.0, 1.5, 0.8)),               # one axis already at 1.5 mm
                                           ((21, 64, 48), (3.0, 0.7, 0.7))])
def test_resample_3d_vs_scipy(cuda, shape, spacing):
    """change_spacing(img, [1.5]*3, order=3) and back with order=0 against scipy.ndimage.zoom, the function the
    reference calls (totalsegmentator/resampling.py:24-56)."""
    from scipy import ndimage
    from boa_b200.resample import resample_labels_nearest, resample_volume_cubic, zoomed_shape
    rng = np.random.default_rng(11)
    # smooth-ish field + noise so that neighbouring voxels are correlated like a CT
    ct = (rng.integers(-1000, 2000, size=shape) * 0.3 + 400 * np.sin(np.arange(shape[2]) / 5.0)[None, None, :]).astype(np.int16)
    zoom = [np.float64(np.float32(s)) / 1.5 for s in spacing]
    # scipy evaluates the axes that are already at 1.5 mm at integer coordinates, where the spline reproduces the
    # samples; the product skips those axes (like the thickness-only path)
    ref = ndimage.zoom(ct.astype(np.float64), zoom, order=3, mode="nearest")
    got = resample_volume_cubic(_dev(ct), spacing, 1.5).cpu().numpy()
    assert got.shape == ref.shape == zoomed_shape(shape, spacing, 1.5)
    assert got.dtype == np.int16
    # fp64 splines on both sides: the values agree to ~1e-11 before truncation, so the integers can differ (by one)
    # only where the spline value is numerically an integer - output points that coincide with input samples, which
    # the (1.0, 1.5, 0.8) case has on purpose (zoom scales 1.5 and 2.0); there scipy's own rounding noise decides
    diff = np.abs(got.astype(np.int64) - ref.astype(np.int32))
    assert diff.max() <= 1, diff.max()
    inner = np.abs(ref - np.rint(ref)) > 1e-6
    assert inner.mean() > 0.4
    assert np.array_equal(got[inner], ref.astype(np.int32)[inner])
    lab = rng.integers(0, 118, size=ref.shape).astype(np.uint8)
    back = resample_labels_nearest(_dev(lab), shape).cpu().numpy()
    ref_back = ndimage.zoom(lab, np.array(shape) / np.array(lab.shape), order=0, mode="nearest")
    assert back.shape == tuple(shape) and np.array_equal(back, ref_back)


def test_resample_identity_and_errors(cuda):
    from boa_b200.resample import resample_labels_nearest, resample_volume_cubic
    ct = _dev(np.zeros((8, 9, 10), dtype=np.int16))
    assert resample_volume_cubic(ct, (1.5, 1.5, 1.5), 1.5) is ct
    lab = _dev(np.zeros((8, 9, 10), dtype=np.uint8))
    assert resample_labels_nearest(lab, (8, 9, 10)) is lab
    with pytest.raises(TypeError):
        resample_volume_cubic(ct.to(torch.float64), (1.0, 1.0, 1.0), 1.5)


@pytest.mark.parametrize("shape", [(5, 33, 47), (3, 1, 9), (2, 16, 1), (4, 64, 64)])
def test_median3x3_slices_vs_scipy(cuda, shape):
    """--bca-median-filtering: scipy.ndimage.median_filter(image, size=[1, 3, 3]) (subclassification.py:33-36)."""
    from scipy import ndimage
    rng = np.random.default_rng(21)
    ct = rng.integers(-1100, 3100, size=shape).astype(np.int16)
    got = passes.median3x3_slices(_dev(ct)).cpu().numpy()
    assert np.array_equal(got, ndimage.median_filter(ct, size=[1, 3, 3]))
    regions = rng.integers(0, 12, size=shape).astype(np.uint8)
    from boa_b200 import bca
    tis = bca.subclassify_tissues(_dev(ct), _dev(regions), median_filtering=True).cpu().numpy()
    assert np.array_equal(tis, op.subclassify_tissues(ndimage.median_filter(ct, size=[1, 3, 3]), regions))


@pytest.mark.parametrize("shape,cur,new", [((12, 40, 36), (5.0, 0.9, 0.9), (5.0, 1.5, 1.5)),     # separate z, 2-D cubic per slice
                                           ((10, 30, 44), (6.0, 1.0, 1.0), (5.0, 0.8, 0.8)),     # separate z + order-0 pick along z
                                           ((20, 24, 28), (1.0, 0.8, 0.8), (1.5, 1.5, 1.5)),     # 3-D cubic
                                           ((9, 33, 31), (5.0, 1.5, 1.5), (5.0, 1.5, 1.5))])      # identity
def test_resample_to_plan_spacing_vs_oracle(cuda, shape, cur, new):
    """nnU-Net's own resampling of the normalised volume to the plan's spacing (default_preprocessor.py:57-90)."""
    from boa_b200.resample import nnunet_new_shape, resample_to_plan_spacing
    from oracle import resampling as orr
    rng = np.random.default_rng(6)
    data = rng.standard_normal(shape).astype(np.float32)
    data[:, :5] = -2.2  # a flat border: cubic overshoot at its edge is what the clip acts on
    ref = orr.resample_data(data[None], orr.compute_new_shape(shape, cur, new), cur, new, order=3)[0]
    got = resample_to_plan_spacing(_dev(data), cur, new).cpu().numpy()
    assert got.shape == ref.shape == nnunet_new_shape(shape, cur, new)
    assert np.allclose(got, ref, rtol=0, atol=2e-6), np.abs(got - ref).max()


@pytest.mark.parametrize("net_shape,out_shape,cur,new", [((8, 20, 24), (8, 33, 40), (5.0, 1.5, 1.5), (5.0, 0.9, 0.9)),
                                                         ((10, 18, 22), (12, 30, 37), (6.0, 1.5, 1.5), (5.0, 0.9, 0.9)),
                                                         ((12, 14, 16), (18, 26, 30), (1.5, 1.5, 1.5), (1.0, 0.8, 0.8))])
def test_finalize_argmax_resampled_vs_oracle(cuda, net_shape, out_shape, cur, new):
    """Logits / n resampled with order 1 to the pre-resampling shape, then argmax (export_prediction.py:25-38)."""
    from boa_b200.predictor import finalize_argmax_resampled
    from boa_b200.resample import nnunet_separate_z
    from oracle import resampling as orr
    rng = np.random.default_rng(8)
    C = 5
    w = rng.uniform(0.5, 10.0, size=net_shape).astype(np.float32)
    logits = (rng.standard_normal((C, *net_shape)) * 3).astype(np.float32)
    acc = (logits * w).astype(np.float32)
    q = acc / w  # what the reference resamples
    ref = orr.logits_to_segmentation(q, out_shape, cur, new)
    sep, axis = nnunet_separate_z(cur, new)
    assert (sep, axis) == orr.determine_do_sep_z_and_axis(cur, new)
    got = finalize_argmax_resampled(_dev(acc), _dev(w), out_shape, sep).cpu().numpy()
    r = orr.resample_data(q.astype(np.float64), out_shape, cur, new, order=1)
    top2 = np.sort(r, axis=0)[-2:]
    safe = (top2[1] - top2[0]) > 1e-4  # fp32 quotients vs the oracle's fp64 interpolation
    assert got.shape == ref.shape
    assert np.array_equal(got[safe], ref[safe]) and (got == ref).mean() > 0.9999


def test_nnunet_resampling_against_the_reference_vectors(cuda):
    """The device resampling to the plan's spacing and the resampled argmax against outputs of the reference's own
    resample_data_or_seg_to_shape (tests/golden/make_golden_resampling.py)."""
    import os
    from boa_b200.predictor import finalize_argmax_resampled
    from boa_b200.resample import nnunet_separate_z, resample_to_plan_spacing
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resampling.npz"))
    for name in sorted(k[:-3] for k in z.files if k.endswith("_in")):
        meta = z[name + "_meta"]
        cur, new, order = tuple(float(v) for v in meta[:3]), tuple(float(v) for v in meta[3:6]), int(meta[6])
        data, ref = z[name + "_in"], z[name + "_out"]
        if order == 3:   # DefaultPreprocessor: every channel of the normalised image
            for c in range(data.shape[0]):
                got = resample_to_plan_spacing(_dev(data[c]), cur, new).cpu().numpy()
                assert got.shape == ref[c].shape, name
                assert np.allclose(got, ref[c], rtol=0, atol=2e-6), (name, np.abs(got - ref[c]).max())
        else:            # export: logits back to the pre-resampling shape, argmax
            sep, _ = nnunet_separate_z(cur, new)
            w = np.ones(data.shape[1:], np.float32)
            got = finalize_argmax_resampled(_dev(data), _dev(w), ref.shape[1:], sep).cpu().numpy()
            want = ref.argmax(0).astype(np.uint8)
            safe = np.abs(ref[0].astype(np.float64) - ref[1]) > 1e-4
            assert np.array_equal(got[safe], want[safe]) and (got == want).mean() > 0.999, name


def test_change_spacing_against_the_reference_vectors(cuda):
    """resample_volume_cubic / resample_thickness / resample_labels_nearest against outputs of the reference's own
    change_spacing (tests/golden/make_golden_change_spacing.py).  Order 0 is bit-exact; order 3 agrees except (by one)
    where the fp64 spline value is numerically an integer and the truncation to int is decided by rounding noise."""
    from scipy import ndimage
    from boa_b200.resample import resample_labels_nearest, resample_thickness, resample_volume_cubic
    z = np.load(os.path.join(G, "change_spacing.npz"))
    names = sorted(k[:-3] for k in z.files if k.endswith("_ct"))
    assert len(names) == 4
    for name in names:
        ct, sp, tg, ref = z[name + "_ct"], z[name + "_spacing_zyx"], z[name + "_target_zyx"], z[name + "_resampled"]
        if name.startswith("thickness"):
            got = resample_thickness(_dev(ct), float(sp[0]), float(tg[0])).cpu().numpy()
        else:
            got = resample_volume_cubic(_dev(ct), tuple(float(v) for v in sp), float(tg[0])).cpu().numpy()
        assert got.shape == ref.shape and got.dtype == np.int16, name
        zoom = [np.float64(np.float32(s)) / np.float64(np.float32(t)) for s, t in zip(sp, tg)]
        spline = ndimage.zoom(ct.astype(np.float64), zoom, order=3, mode="nearest")
        diff = np.abs(got.astype(np.int64) - ref)
        assert diff.max() <= 1, (name, diff.max())
        inner = np.abs(spline - np.rint(spline)) > 1e-6
        assert inner.mean() > 0.4 and np.array_equal(got[inner], ref[inner]), name
        back = resample_labels_nearest(_dev(z[name + "_labels"]), ct.shape).cpu().numpy()
        assert np.array_equal(back, z[name + "_labels_back"]), name


def test_normalize_logits_equals_ieee_division_and_flags_nonfinite(cuda):
    """boa_normalize_logits: acc[c][v] / (w[v] * folds) bit-equal to the fp32 division torch does, counter set by inf."""
    import ctypes as C
    from boa_b200 import _lib
    rng = np.random.default_rng(2)
    for V, Cn, folds in ((1000, 5, 1), (4099, 3, 5)):
        acc = (rng.standard_normal((Cn, V)) * 50).astype(np.float32)
        w = rng.uniform(1e-3, 40.0, size=V).astype(np.float32)
        a, wd = _dev(acc), _dev(w)
        want = (a / (wd * float(folds))).cpu().numpy()
        bad = torch.zeros(1, dtype=torch.int32, device="cuda")
        _lib.check(_lib.lib().boa_normalize_logits(_lib.ptr(a), _lib.ptr(wd), Cn, V, float(folds), _lib.ptr(bad),
                                                   _lib.stream_ptr()))
        assert np.array_equal(a.cpu().numpy(), want) and int(bad) == 0
        acc[1, 7] = np.inf
        a = _dev(acc)
        _lib.check(_lib.lib().boa_normalize_logits(_lib.ptr(a), _lib.ptr(wd), Cn, V, float(folds), _lib.ptr(bad),
                                                   _lib.stream_ptr()))
        assert int(bad) > 0
