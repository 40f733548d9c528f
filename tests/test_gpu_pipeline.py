"""GPU: the whole per-volume path (`total+bca`) on a small synthetic CT against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from boa_b200 import zoo
from boa_b200.labels import part_luts
from boa_b200.pipeline import ModelZoo, analyze_volume
from oracle import passes as op
from oracle.report import bca_json, compute_measurements
from oracle.sliding_window import convert_logits_to_segmentation, merge_parts, predict_sliding_window_return_logits


@pytest.fixture(scope="module")
def small_zoo():
    specs = zoo.synthetic_specs((32, 32, 32), 32, 64, 3, bca_folds=5, seed=1)
    return specs, ModelZoo.from_specs(specs, device=torch.device("cuda", 0), max_batch=3)


def _oracle_labels(spec, ct, step, folds):
    data = op.ct_normalize(ct, spec.intensity)[None]
    sds = [spec.fold_weights[f] for f in folds]
    logits = predict_sliding_window_return_logits(spec.arch, sds, data, step, emulate_fp16=True)
    return convert_logits_to_segmentation(logits), logits


def test_total_and_bca_against_oracle(cuda, small_zoo):
    specs, mz = small_zoo
    ct = zoo.synthetic_ct((48, 64, 56), seed=2)
    res = analyze_volume(torch.from_numpy(ct).cuda(), (1.5, 1.5, 1.5), mz, models=("total", "bca"), fast_bca=False,
                         cnr_adjustment=True)
    # ---- total: 5 part networks, merged
    segs, margins = [], []
    for tid in (291, 292, 293, 294, 295):
        seg, logits = _oracle_labels(specs[tid], ct, 0.8, [0])
        segs.append(seg)
        top2 = np.sort(logits, axis=0)[-2:]
        margins.append(top2[1] - top2[0])
    ref_total = merge_parts(segs, part_luts(), ct.shape)
    total = res.total.cpu().numpy()
    agree = (total == ref_total).mean()
    # voxels whose top-2 margin exceeds the fp16 noise in EVERY part model must agree exactly
    safe = np.all(np.stack(margins) > 0.05, axis=0)
    print(f"total: agreement {agree:.5f}, safe voxels {safe.mean():.3f}")
    assert agree > 0.99
    assert np.array_equal(total[safe], ref_total[safe])
    # ---- everything downstream is integer / exact: recompute with the oracle FROM OUR label maps
    ref_meas = compute_measurements(ct, total, (1.5, 1.5, 1.5), cnr_adjustment=True)
    assert np.array_equal(res.ct_pfav.cpu().numpy(), ref_meas.pop("_ct_pfav_mask"))
    from test_oracle_golden import _close
    _close(ref_meas, res.total_measurements)
    regions, parts = res.body_regions.cpu().numpy(), res.body_parts.cpu().numpy()
    assert regions.shape == ct.shape and parts.shape == ct.shape
    # connected-component post-processing: the pipeline runs it on the 5 mm grid with slice weights; the reference's
    # order is replicate to the input grid first, then post-process (infer/infer.py:67-89) - same result
    from boa_b200.pipeline import segment_bca_net
    from boa_b200.resample import resample_thickness, upsample_labels_nearest
    from oracle import postprocess as opp
    ct5 = resample_thickness(torch.from_numpy(ct).cuda(), 1.5, 5.0)
    for task, got, fn in (("body_regions", regions, opp.postprocess_region_segmentation),
                          ("body_parts", parts, opp.remove_small_labeled_objects)):
        raw = upsample_labels_nearest(segment_bca_net(ct5, mz, task, fast=False), ct.shape[0]).cpu().numpy()
        assert np.array_equal(got, fn(raw)), task
    tissues = res.tissues.cpu().numpy()
    assert np.array_equal(tissues, op.subclassify_tissues(ct, regions))
    js, vert = bca_json(ct, tissues, parts, regions, total, (1.5, 1.5, 1.5))
    _close(js, res.bca_measurements)
    assert {k: tuple(v) for k, v in vert.items()} == {k: tuple(v) for k, v in res.vertebrae.items()}


def test_bca_nets_against_oracle(cuda, small_zoo):
    """5 mm path: thickness resampling (scipy restated on the device) -> fold ensemble -> nearest up-sampling."""
    from scipy import ndimage
    from boa_b200.pipeline import segment_bca_net
    from boa_b200.resample import resample_thickness, upsample_labels_nearest
    specs, mz = small_zoo
    ct = zoo.synthetic_ct((120, 40, 48), seed=4)
    ct5 = resample_thickness(torch.from_numpy(ct).cuda(), 1.5, 5.0)
    ct5_np = ct5.cpu().numpy()
    assert ct5_np.shape[0] == 36
    for task, tid in (("body_regions", 542), ("body_parts", 543)):
        for fast, folds in ((True, [0]), (False, [0, 1, 2, 3, 4])):  # --fast-bca / 5-fold logit ensemble
            lab = segment_bca_net(ct5, mz, task, fast=fast).cpu().numpy()
            ref, _ = _oracle_labels(specs[tid], ct5_np, 0.5, folds)
            agree = (lab == ref).mean()
            print(f"{task} folds {folds}: agreement {agree:.5f}")
            assert agree > 0.99
        up = upsample_labels_nearest(torch.from_numpy(lab).cuda(), 120).cpu().numpy()
        assert np.array_equal(up, ndimage.zoom(lab, (120 / 36, 1, 1), order=0, mode="nearest"))


def test_volume_smaller_than_patch_is_padded(cuda, small_zoo):
    specs, mz = small_zoo
    from boa_b200.pipeline import segment_task
    ct = zoo.synthetic_ct((20, 40, 24), seed=5)
    lab = segment_task(torch.from_numpy(ct).cuda(), mz, [291], [0], 0.5).cpu().numpy()
    ref, _ = _oracle_labels(specs[291], ct, 0.5, [0])
    assert lab.shape == ct.shape
    assert (lab == ref).mean() > 0.99


def test_crop_to_nonzero(cuda, small_zoo):
    specs, mz = small_zoo
    from boa_b200.pipeline import nonzero_bbox, segment_task
    ct = zoo.synthetic_ct((40, 40, 40), seed=6)
    ct[:4] = 0; ct[:, :3] = 0; ct[:, :, 37:] = 0
    assert nonzero_bbox(torch.from_numpy(ct).cuda()) == [(4, 40), (3, 40), (0, 37)]
    lab = segment_task(torch.from_numpy(ct).cuda(), mz, [291], [0], 0.5).cpu().numpy()
    assert (lab[:4] == 0).all() and (lab[:, :3] == 0).all() and (lab[:, :, 37:] == 0).all()
    ref, _ = _oracle_labels(specs[291], ct[4:, 3:, :37], 0.5, [0])
    assert (lab[4:, 3:, :37] == ref).mean() > 0.99


def test_host_api_matches_device_api(cuda, small_zoo):
    """analyze_from_host (pinned staging buffers, copies on a side stream) returns exactly what analyze_volume
    leaves on the device, twice in a row on the same zoo (buffer reuse)."""
    from boa_b200.pipeline import analyze_from_host

    specs, mz = small_zoo
    for seed in (5, 6):
        ct = zoo.synthetic_ct((40, 48, 40), seed=seed)
        ct_host = torch.from_numpy(ct).pin_memory()
        res = analyze_volume(ct_host.cuda(), (1.5, 1.5, 1.5), mz, models=("total", "bca"), fast_bca=True)
        out = analyze_from_host(ct_host, (1.5, 1.5, 1.5), mz, models=("total", "bca"), fast_bca=True)
        for name in ("total", "body_parts", "body_regions", "tissues", "ct_pfav"):
            assert out[name].device.type == "cpu" and out[name].is_pinned()
            assert np.array_equal(out[name].numpy(), getattr(res, name).cpu().numpy()), name
        from test_oracle_golden import _close
        _close(res.total_measurements, out["total_measurements"])
        _close(res.bca_measurements, out["bca_measurements"])


def test_total_on_a_volume_that_is_not_at_1_5_mm(cuda, small_zoo):
    """Non-1.5 mm input: order-3 resampling to 1.5 mm, networks, order-0 resampling of the label map back to the input
    grid (totalsegmentator/nnunet.py:466-470,685-687) - checked against scipy.ndimage.zoom + the oracle networks."""
    from scipy import ndimage
    specs, mz = small_zoo
    ct = zoo.synthetic_ct((36, 80, 72), seed=9)
    spacing = (2.5, 0.9, 0.9)
    res = analyze_volume(torch.from_numpy(ct).cuda(), spacing, mz, models=("total",))
    total = res.total.cpu().numpy()
    assert total.shape == ct.shape and total.dtype == np.uint8
    zoomf = [np.float64(np.float32(s)) / 1.5 for s in spacing]
    ct15 = ndimage.zoom(ct.astype(np.float64), zoomf, order=3, mode="nearest").astype(np.int32).astype(np.int16)
    segs = [_oracle_labels(specs[tid], ct15, 0.8, [0])[0] for tid in (291, 292, 293, 294, 295)]
    ref15 = merge_parts(segs, part_luts(), ct15.shape)
    ref = ndimage.zoom(ref15, np.array(ct.shape) / np.array(ref15.shape), order=0, mode="nearest")
    agree = (total == ref).mean()
    print(f"total on a {spacing} mm volume: agreement {agree:.5f}")
    assert agree > 0.99
    # the measurements are taken on the ORIGINAL grid with the ORIGINAL spacing
    ref_meas = compute_measurements(ct, total, (spacing[2], spacing[1], spacing[0]), cnr_adjustment=False)
    ref_meas.pop("_ct_pfav_mask")
    from test_oracle_golden import _close
    _close(ref_meas, res.total_measurements)


def test_volume_off_the_plans_grid_is_resampled_like_nnunet(cuda, small_zoo):
    """A 5 mm volume whose in-plane spacing is not the plan's (every real CT for the body-composition nets, which are
    only resampled in thickness before nnU-Net sees them): nnU-Net's preprocessing resamples the NORMALISED volume to
    the plan's spacing (order 3, slice by slice - separate z) and the export resamples the logits back (order 1) before
    the argmax (default_preprocessor.py:57-90, export_prediction.py:25-38).  Against the oracle's restatement."""
    from boa_b200.pipeline import segment_bca_net
    from oracle import resampling as orr
    specs, mz = small_zoo
    spec = specs[543]
    ct5 = zoo.synthetic_ct((36, 72, 80), seed=8)
    sp = (5.0, 0.9, 0.9)                      # plan: (5.0, 1.5, 1.5)
    lab = segment_bca_net(torch.from_numpy(ct5).cuda(), mz, "body_parts", fast=True, spacing_zyx=sp).cpu().numpy()
    assert lab.shape == ct5.shape
    norm = op.ct_normalize(ct5, spec.intensity)[None]
    new_shape = orr.compute_new_shape(ct5.shape, sp, spec.spacing)
    assert new_shape == (36, 43, 48)
    data = orr.resample_data(norm, new_shape, sp, spec.spacing, order=3).astype(np.float32)
    logits = predict_sliding_window_return_logits(spec.arch, [spec.fold_weights[0]], data, 0.5, emulate_fp16=True)
    ref = orr.logits_to_segmentation(logits, ct5.shape, spec.spacing, sp)
    agree = (lab == ref).mean()
    print(f"body_parts on a (5.0, 0.9, 0.9) mm volume: agreement with the oracle {agree:.5f}")
    assert agree > 0.99


@pytest.mark.gpu
def test_breast_implant_finding_device_path_equals_host_path(cuda):
    """The device pre-filter (components <= 10 ml removed by boa_cc_filter, bounding box to the host) gives the sentence
    of the plain host evaluation, also on a label map full of speckle (where a per-component host loop never ends)."""
    import time
    from boa_b200 import bca
    R = bca.BODY_REGION["BREAST_IMPLANT"]
    rng = np.random.default_rng(3)
    ml = 0.9 * 0.9 * 1.5 / 1000.0
    for n_blobs in (0, 1, 2, 3):
        reg = rng.integers(0, 12, size=(96, 160, 192)).astype(np.uint8)
        reg[reg == R] = 0
        reg[rng.random(reg.shape) < 0.02] = R            # speckle: tens of thousands of tiny components
        for c in [(40, 50, 40), (44, 52, 150), (70, 120, 96)][:n_blobs]:
            z, y, x = np.ogrid[:96, :160, :192]
            reg[(z - c[0]) ** 2 + (y - c[1]) ** 2 + (x - c[2]) ** 2 < 17 ** 2] = R
        t0 = time.perf_counter()
        dev = bca.breast_implant_finding(torch.from_numpy(reg).to(cuda), ml)
        assert time.perf_counter() - t0 < 20
        assert dev == bca.breast_implant_finding(reg, ml)
        assert (dev is not None) == (n_blobs in (1, 2))
        # the same on the 5 mm grid: every slice stands for weights[z] replicated slices of the input grid
        from boa_b200.postprocess import slice_weights
        from boa_b200.resample import upsample_labels_nearest
        reg5 = torch.from_numpy(reg[::3].copy()).to(cuda)
        full = upsample_labels_nearest(reg5, 96)
        w = slice_weights(reg5.shape[0], 96, cuda)
        assert bca.breast_implant_finding(reg5, ml, w) == bca.breast_implant_finding(full.cpu().numpy(), ml)


@pytest.mark.gpu
def test_pending_l3_axes_equals_the_synchronous_evaluation(cuda):
    """PendingL3Axes (slice index on the device, host geometry in a thread) == ts_metrics.major_minor_axis on numpy maps,
    for an odd and an even number of L3 slices and for a volume without L3."""
    from boa_b200.labels import class_map
    from boa_b200.ts_metrics import PendingL3Axes, major_minor_axis
    l3 = {v: k for k, v in class_map("total").items()}["vertebrae_L3"]
    z, y, x = np.ogrid[:40, :96, :128]
    for slices in ((11, 12, 13, 17, 20), (9, 14, 15, 30), ()):
        total = np.zeros((40, 96, 128), np.uint8)
        for s in slices:
            total[s, 40:50, 60:70] = l3
        parts = (((y - 48) / (30.0 + 0.3 * z)) ** 2 + ((x - 64) / (50.0 - 0.5 * z)) ** 2 < 1).astype(np.uint8)
        want = major_minor_axis(total, parts, (0.8, 0.7))
        got = PendingL3Axes(torch.from_numpy(total).to(cuda), torch.from_numpy(parts).to(cuda), (0.8, 0.7)).finish()
        assert got == want, (slices, got, want)
        assert (want[0] is None) == (len(slices) == 0)
