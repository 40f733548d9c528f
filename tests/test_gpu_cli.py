"""GPU: the CLI end to end on a synthetic NIfTI with a synthetic model zoo on disk (real on-disk contract)."""
import json

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cli_total_bca_outputs(cuda, tmp_path):
    from boa_b200 import nifti, zoo
    from boa_b200.cli import run
    weights = tmp_path / "weights"
    zoo.write_zoo(str(weights), patch=(32, 32, 32), base=32, max_features=64, n_stages=3, bca_folds=(0,), seed=1)
    ct = zoo.synthetic_ct((40, 48, 44), seed=9)
    # LPS file orientation (the usual DICOM-derived NIfTI): exercises canonicalisation and its undo
    aff = np.array([[-1.5, 0, 0, 30.0], [0, -1.5, 0, 40.0], [0, 0, 1.5, -10.0], [0, 0, 0, 1.0]])
    nifti.save(tmp_path / "ct.nii.gz", ct, aff)
    out = tmp_path / "out"
    run(["--input-image", str(tmp_path / "ct.nii.gz"), "--models", "total+bca", "--fast-bca", "--bca-no-pdf",
         "--cnr-adjustment", "-d", "gpu", "-o", str(out), "--weights", str(weights)])
    for f in ("total.nii.gz", "body_parts.nii.gz", "body_regions.nii.gz", "tissues.nii.gz", "ct_pfav.nii.gz",
              "total-measurements.json", "bca-measurements.json", "debug_information.txt"):
        assert (out / f).is_file() and (out / f).stat().st_size > 0, f
    total = nifti.load(out / "total.nii.gz")
    assert total.data.shape == ct.shape and total.data.dtype == np.uint8 and np.allclose(total.affine, aff)
    assert total.data.max() <= 117 and (total.data != 0).any()
    # file-space consistency: statistics recomputed from the written files equal the JSON
    from oracle.report import compute_measurements
    tm = json.load(open(out / "total-measurements.json"))
    ref = compute_measurements(ct, total.data, (1.5, 1.5, 1.5), cnr_adjustment=True)
    ref.pop("_ct_pfav_mask")
    from test_oracle_golden import _close
    _close(ref, tm)
    tissues, regions = nifti.load(out / "tissues.nii.gz").data, nifti.load(out / "body_regions.nii.gz").data
    from oracle.passes import subclassify_tissues
    assert np.array_equal(tissues, subclassify_tissues(ct, regions))
    bj = json.load(open(out / "bca-measurements.json"))
    assert len(bj["slices"]) == ct.shape[0] and "whole_scan" in bj["aggregated"]


def test_cli_fast_total_on_anisotropic_input(cuda, tmp_path):
    """--fast-total: 3-D resampling to 3 mm, the single model 297, labels back on the input grid
    (totalsegmentator/python_api.py:169-175, nnunet.py:466-470,685-687)."""
    from boa_b200 import nifti, zoo
    from boa_b200.cli import run
    weights = tmp_path / "weights"
    zoo.write_zoo(str(weights), patch=(32, 32, 32), base=32, max_features=64, n_stages=3, bca_folds=(0,), seed=1,
                  datasets=[297])
    ct = zoo.synthetic_ct((44, 96, 88), seed=4)          # array axes z, y, x
    aff = np.diag([1.0, 1.0, 2.5, 1.0])                  # file axes x, y, z: 1 x 1 x 2.5 mm
    nifti.save(tmp_path / "ct.nii.gz", ct, aff)
    out = tmp_path / "out"
    run(["--input-image", str(tmp_path / "ct.nii.gz"), "--models", "total", "--fast-total", "-d", "gpu", "-o", str(out),
         "--weights", str(weights)])
    total = nifti.load(out / "total.nii.gz")
    assert total.data.shape == ct.shape and total.data.dtype == np.uint8 and np.allclose(total.affine, aff)
    assert total.data.max() <= 117 and (total.data != 0).any()
    from oracle.report import compute_measurements
    tm = json.load(open(out / "total-measurements.json"))
    ref = compute_measurements(ct, total.data, (1.0, 1.0, 2.5), cnr_adjustment=False)
    ref.pop("_ct_pfav_mask")
    from test_oracle_golden import _close
    _close(ref, tm)


def test_cli_crop_prepass_tasks(cuda, tmp_path):
    """Tasks behind the crop pre-pass (`--models all` territory; totalsegmentator/python_api.py:236-330,673-736): rough
    6 mm `total` segmentation (model 298) -> bounding box of the task's structures + 20 mm -> the task's network on the
    crop at the volume's native spacing (nnU-Net resamples to ITS plan's spacing and the logits back) -> un-crop.
    Checked against the same chain run step by step, and the measurements against the written files."""
    import torch
    from boa_b200 import nifti, zoo
    from boa_b200.cli import run
    from boa_b200.pipeline import ModelZoo, crop_box_from_rois, rough_total_6mm, segment_task
    from boa_b200.labels import CROP_TASKS
    weights = tmp_path / "weights"
    zoo.write_zoo(str(weights), patch=(32, 32, 32), base=32, max_features=64, n_stages=3, bca_folds=(0,), seed=1,
                  datasets=[291, 292, 293, 294, 295, 298, 258, 8, 315])
    # pleural_pericard_effusion uses every fold in its folder (folds=None): give it two
    zoo.write_zoo(str(weights), patch=(32, 32, 32), base=32, max_features=64, n_stages=3, bca_folds=(0,), seed=1, datasets=[315])
    import shutil
    f315 = [d for d in (weights).iterdir() if d.name.startswith("Dataset315")][0] / "nnUNetTrainer__nnUNetPlans__3d_fullres"
    shutil.copytree(f315 / "fold_0", f315 / "fold_1")
    ct = zoo.synthetic_ct((56, 64, 60), seed=11)
    aff = np.diag([1.5, 1.5, 1.5, 1.0])
    nifti.save(tmp_path / "ct.nii.gz", ct, aff)
    out = tmp_path / "out"
    run(["--input-image", str(tmp_path / "ct.nii.gz"), "--models", "total+lung_vessels+liver_vessels+pleural_pericard_effusion",
         "-d", "gpu", "-o", str(out), "--weights", str(weights)])
    tm = json.load(open(out / "total-measurements.json"))
    assert set(tm["segmentations"]) == {"total", "lung_vessels", "liver_vessels", "pleural_pericard_effusion"}
    mz = ModelZoo(str(weights), device=torch.device("cuda", 0))
    ctd = torch.from_numpy(ct).cuda()
    rough = rough_total_6mm(ctd, (1.5, 1.5, 1.5), mz)
    for task in ("lung_vessels", "liver_vessels", "pleural_pericard_effusion"):
        img = nifti.load(out / f"{task}.nii.gz")
        assert img.data.shape == ct.shape and img.data.dtype == np.uint8
        tid, folds, rois = CROP_TASKS[task]
        box = crop_box_from_rois(rough, rois, (1.5, 1.5, 1.5))
        ref = np.zeros(ct.shape, dtype=np.uint8)
        if box is not None:
            sl = tuple(slice(b, e) for b, e in box)
            ref[sl] = segment_task(ctd[sl].contiguous(), mz, [tid], folds, 0.5, None, None,
                                   spacing_zyx=(1.5, 1.5, 1.5)).cpu().numpy()
            assert (ref != 0).any()
        assert np.array_equal(img.data, ref), task
    assert len(mz.get(315, None, 0.5).networks) == 2
