"""CPU: host-side logic of the drop-in (plans / checkpoint contract, label tables, padding, config) and the C ABI
surface (the library loads without a GPU and exports every symbol include/boa_b200.h declares)."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_abi_exports_every_declared_symbol():
    from boa_b200 import _lib
    header = open(os.path.join(ROOT, "include", "boa_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(boa_[a-z0-9_]+)\s*\(", header))
    declared -= {"boa_arch", "boa_net"}
    assert len(declared) >= 25
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/boa_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), (declared ^ set(_lib.EXPORTS))
    assert _lib.lib().boa_abi_version() == 1


def test_compute_entries_fail_loudly_without_a_gpu():
    from boa_b200 import _lib
    from boa_b200.predictor import nnUNetPredictor
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        nnUNetPredictor(device=torch.device("cpu"))
    if not torch.cuda.is_available():
        a = _lib.BoaArch()
        a.n_stages, a.in_channels, a.num_classes = 2, 1, 2
        for s in range(2):
            a.features[s] = 8
            a.n_conv_enc[s] = 1
            for k in range(3):
                a.kernels[s][k] = 3
                a.strides[s][k] = 1 if s == 0 else 2
        a.n_conv_dec[0] = 1
        for k in range(3):
            a.patch[k] = 8
        h = ctypes.c_void_p()
        rc = _lib.lib().boa_net_create(ctypes.byref(a), 0, 1, ctypes.byref(h))
        assert rc == -2, "without a CUDA device boa_net_create must return BOA_ERR_CUDA, not fall back"
        assert b"cuda" in _lib.lib().boa_last_error().lower()


def test_pad_to_patch_matches_oracle_pad_nd_image():
    from boa_b200.geometry import pad_to_patch
    from oracle.sliding_window import pad_nd_image
    rng = np.random.default_rng(0)
    for shape, patch in [((5, 9, 7), (8, 8, 8)), ((64, 128, 128), (128, 128, 128)), ((20, 20, 20), (16, 16, 16)),
                         ((3, 40, 17), (16, 32, 16))]:
        x = rng.standard_normal((1, *shape)).astype(np.float32)
        ref, sl = pad_nd_image(x, patch)
        pads, unpad = pad_to_patch(shape, patch)
        got = np.pad(x[0], pads)
        assert np.array_equal(got, ref[0]) and sl[1:] == unpad


def test_plans_and_checkpoint_contract_roundtrip(tmp_path):
    from boa_b200 import zoo
    from boa_b200.plans import find_model_folder, load_model_folder
    zoo.write_zoo(str(tmp_path), patch=(16, 16, 16), base=8, max_features=16, n_stages=3, bca_folds=(0, 1), datasets=[291, 543])
    folder = find_model_folder(str(tmp_path), 543, "nnUNetTrainer_1500epochs_NoMirroring")
    assert os.path.exists(os.path.join(folder, "plans.json")) and os.path.exists(os.path.join(folder, "fold_1", "checkpoint_final.pth"))
    spec = load_model_folder(folder, [0, 1])
    assert spec.arch["features"] == [8, 16, 16] and spec.arch["num_classes"] == 7 and len(spec.fold_weights) == 2
    assert spec.arch["strides"][0] == [1, 1, 1] and spec.arch["n_conv_dec"] == [2, 2]
    sd = spec.fold_weights[0]
    assert "encoder.stages.0.0.convs.0.conv.weight" in sd and "decoder.seg_layers.1.weight" in sd
    assert "decoder.encoder.stages.0.0.convs.0.conv.weight" in sd  # alias keys of a real checkpoint are present
    with pytest.raises(RuntimeError):
        find_model_folder(str(tmp_path), 999, "x")


def test_new_format_plans_are_parsed():
    from boa_b200.plans import arch_from_plans
    plans = {"configurations": {"3d_fullres": {"patch_size": [32, 64, 64], "architecture": {
        "network_class_name": "dynamic_network_architectures.architectures.unet.PlainConvUNet",
        "arch_kwargs": {"n_stages": 3, "features_per_stage": [16, 32, 64], "kernel_sizes": [[1, 3, 3], [3, 3, 3], [3, 3, 3]],
                        "strides": [[1, 1, 1], [1, 2, 2], [2, 2, 2]], "n_conv_per_stage": [2, 2, 2],
                        "n_conv_per_stage_decoder": [2, 2], "norm_op_kwargs": {"eps": 1e-5, "affine": True}}}}}}
    a = arch_from_plans(plans, "3d_fullres", 1, 5)
    assert a["features"] == [16, 32, 64] and a["kernels"][0] == [1, 3, 3] and a["strides"][1] == [1, 2, 2]
    plans["configurations"]["3d_fullres"]["architecture"]["network_class_name"] = "x.ResidualEncoderUNet"
    with pytest.raises(NotImplementedError):
        arch_from_plans(plans, "3d_fullres", 1, 5)


def test_label_tables():
    from boa_b200.labels import class_map, measurement_label_map, part_luts
    total = class_map("total")
    assert len(total) == 117 and total[1] == "spleen"
    luts = part_luts()
    assert [len(l) for l in luts] == [25, 27, 19, 24, 27]  # classes incl. background
    covered = sorted(v for l in luts for v in l[1:])
    assert covered == list(range(1, 118)), "the five part models together cover the 117 labels exactly once"
    lm = measurement_label_map("total")
    assert len(lm) == 295 and max(lm.values()) <= 117


def test_macs_per_patch_matches_baseline_table():
    from boa_b200 import zoo
    from boa_b200.plans import arch_from_plans
    from oracle.network import count_macs
    arch = arch_from_plans(zoo.default_plans((128,) * 3, 32, 320, 6), "3d_fullres", 1, 25)
    assert abs(count_macs(arch) / 1e9 - 478.8) < 0.5  # 478.8 GMAC per 128^3 patch


def test_resampled_depth_and_weight_keys():
    from boa_b200.resample import resampled_depth
    assert resampled_depth(512, 1.5, 5.0) == 154 and resampled_depth(300, 1.5, 5.0) == 90


def test_resampling_shapes_and_slice_weights_follow_scipy():
    """zoomed_shape == scipy.ndimage.zoom's output shape for the float32 header spacings change_spacing uses
    (totalsegmentator/resampling.py:171-176); slice_weights == multiplicities of scipy's order-0 source indices."""
    from scipy import ndimage
    from boa_b200.postprocess import slice_weights
    from boa_b200.resample import resampled_depth, zoomed_shape
    for shape, spacing, target in [((37, 44, 52), (2.0, 0.9765625, 0.9765625), 1.5), ((40, 33, 29), (1.0, 1.5, 0.8), 1.5),
                                   ((61, 30, 30), (0.7, 0.7, 0.7), 3.0), ((9, 9, 9), (1.5, 1.5, 1.5), 1.5)]:
        zoom = [np.float64(np.float32(s)) / target for s in spacing]
        ref = ndimage.zoom(np.zeros(shape, np.uint8), zoom, order=0).shape
        assert zoomed_shape(shape, spacing, target) == ref
    for z_in, z_out in [(154, 512), (36, 120), (31, 103), (7, 7), (2, 9)]:
        lab = np.arange(z_in, dtype=np.int32).reshape(z_in, 1, 1)
        up = ndimage.zoom(lab, (z_out / z_in, 1, 1), order=0, mode="nearest")[:, 0, 0]
        assert up.shape[0] == z_out
        w = slice_weights(z_in, z_out, "cpu").numpy()
        assert np.array_equal(w, np.bincount(up, minlength=z_in)) and w.sum() == z_out
    assert resampled_depth(512, 1.5, 5.0) == 154


def test_new_compute_entries_refuse_cpu_tensors():
    """No CPU fallback for the passes added with resampling / post-processing / median filtering either."""
    from boa_b200 import passes, postprocess, resample
    ct = torch.zeros((4, 5, 6), dtype=torch.int16)
    with pytest.raises(ValueError):
        resample.resample_volume_cubic(ct, (1.0, 1.0, 1.0), 1.5)
    with pytest.raises(ValueError):
        resample.resample_labels_nearest(ct.to(torch.uint8), (8, 10, 12))
    with pytest.raises(ValueError):
        postprocess.postprocess_region_segmentation(ct.to(torch.uint8))
    with pytest.raises(ValueError):
        passes.median3x3_slices(ct)


def test_triple_split_ranges_stitch_like_the_reference():
    """Very large volumes are predicted in three overlapping z-parts (totalsegmentator/nnunet.py:483-505) and stitched
    (:582-586).  With `prediction == input` the reference's index arithmetic (transcribed below on the slice axis) and
    triple_split_ranges must rebuild the volume from the same source slices."""
    from boa_b200.geometry import needs_triple_split, triple_split_ranges
    for z in (201, 300, 512, 800, 1201):
        vol = np.arange(z)
        third, margin = z // 3, 20
        p1, p2, p3 = vol[:third + margin], vol[third + 1 - margin:third * 2 + margin], vol[third * 2 + 1 - margin:]
        ref = np.zeros(z, dtype=vol.dtype)
        ref[:third] = p1[:-margin]
        ref[third:third * 2] = p2[margin - 1:-margin]
        ref[third * 2:] = p3[margin - 1:]
        got = np.zeros(z, dtype=vol.dtype)
        parts = triple_split_ranges(z)
        for (plo, phi, klo, khi, dlo, dhi), p in zip(parts, (p1, p2, p3)):
            assert np.array_equal(vol[plo:phi], p)
            got[dlo:dhi] = vol[plo:phi][klo:khi]
        assert np.array_equal(got, ref) and np.array_equal(got, vol)
    with pytest.raises(ValueError):
        triple_split_ranges(30)
    assert not needs_triple_split((512, 512, 512), True)            # config 3: 1.3e8 voxels < 512*512*900
    assert needs_triple_split((800, 1024, 1024), True)              # config 5
    assert not needs_triple_split((800, 1024, 1024), False)         # single-model tasks only split when forced
    assert not needs_triple_split((150, 2048, 2048), True)          # z <= 200
    assert needs_triple_split((154, 512, 512), False, force_split=True)


def test_plan_geometry_transpose_raises_and_resampling_decisions_match_the_oracle():
    from boa_b200.pipeline import check_plan_geometry
    from boa_b200.plans import ModelSpec
    from boa_b200.resample import nnunet_new_shape, nnunet_separate_z
    from oracle import resampling as orr
    spec = ModelSpec(arch={}, intensity={}, labels={}, transpose_forward=[0, 2, 1], transpose_backward=[0, 2, 1],
                     spacing=[1.5] * 3, configuration="3d_fullres")
    with pytest.raises(NotImplementedError, match="transpose_forward"):
        check_plan_geometry(spec, (64, 64, 64), None)
    spec.transpose_forward = spec.transpose_backward = [0, 1, 2]
    assert check_plan_geometry(spec, (64, 64, 64), None) is None
    assert check_plan_geometry(spec, (64, 64, 64), (1.0, 1.0, 1.0)) == [1.5, 1.5, 1.5]
    # compute_new_shape / determine_do_sep_z_and_axis (default_resampling.py:14-67) against the oracle's restatement
    for shape, cur, new in (((64, 512, 512), (5.0, 0.7, 0.7), (5.0, 0.8, 0.8)), ((30, 100, 90), (6.0, 1.0, 1.0), (5.0, 0.8, 0.8)),
                            ((80, 90, 100), (1.0, 0.8, 0.8), (1.5, 1.5, 1.5)), ((40, 50, 60), (0.24, 1.25, 1.25), (1.0, 1.0, 1.0)),
                            ((40, 50, 60), (1.0, 1.0, 4.0), (1.0, 1.0, 1.0))):
        assert nnunet_new_shape(shape, cur, new) == orr.compute_new_shape(shape, cur, new)
        assert nnunet_separate_z(cur, new) == orr.determine_do_sep_z_and_axis(cur, new)
