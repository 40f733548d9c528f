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
