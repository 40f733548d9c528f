"""CPU: the CLI / config contract, mirroring the truth tables of the reference's tests/test_config.py:16-144, and the
NIfTI reader / writer."""
import os
from unittest import mock

import numpy as np
import pytest

from boa_b200 import nifti
from boa_b200.cli import get_parser
from boa_b200.config import ALL_MODELS, env_bool, env_str, resolve_device, resolve_models


def test_resolve_models_truth_table():
    assert resolve_models("total") == {"total"}
    assert resolve_models("bca") == {"bca", "total"}
    assert resolve_models("bca+body_regions+body_parts") == {"bca", "total"}
    assert resolve_models("total+body-parts") == {"total", "body_parts"}
    assert resolve_models(None) == set(ALL_MODELS - {"body_regions", "body_parts"}) | {"total"}
    assert resolve_models("all") == resolve_models(None)
    assert resolve_models("total+nonsense") == {"total"}
    with pytest.raises(ValueError, match="Unknown model"):
        resolve_models("total+nonsense", strict=True)


def test_resolve_device():
    with mock.patch.dict(os.environ, {}, clear=True):
        assert resolve_device(None) == "gpu"
        assert resolve_device("cuda") == "gpu"
        assert resolve_device("cpu") == "cpu"
    with mock.patch.dict(os.environ, {}, clear=True):
        assert resolve_device("cuda:3") == "gpu:3"
        assert os.environ["NVIDIA_VISIBLE_DEVICES"] == "3"
    with mock.patch.dict(os.environ, {"DEVICE": "gpu", "NVIDIA_ID": "1"}, clear=True):
        assert resolve_device(None) == "gpu:1"
    with mock.patch.dict(os.environ, {"DEVICE": "cpu", "NVIDIA_ID": "1"}, clear=True):
        assert resolve_device(None) == "cpu"


def test_env_helpers():
    with mock.patch.dict(os.environ, {"A": "true", "B": " 1 ", "C": "no", "D": "TODO", "E": " x "}, clear=True):
        assert env_bool("A") and env_bool("B") and not env_bool("C") and not env_bool("MISSING")
        assert env_bool("MISSING", True)
        assert env_str("D") is None and env_str("E") == "x" and env_str("MISSING", "d") == "d"


def test_cli_parser_flags():
    a = get_parser().parse_args(["--input-image", "x.nii.gz", "--models", "total+bca", "-d", "gpu", "--fast-bca",
                                 "--bca-no-pdf", "-o", "out"])
    assert str(a.input_image) == "x.nii.gz" and a.models == "total+bca" and a.fast_bca and a.bca_no_pdf
    assert not a.bca_median_filtering and a.fast_total is None
    b = get_parser().parse_args(["-i", "x.nii.gz", "-m", "total", "--fast-total", "--bca-median-filtering"])
    assert b.fast_total and b.bca_median_filtering  # body_organ_analysis/cli.py:155-163,175-187
    with pytest.raises(SystemExit):
        get_parser().parse_args(["--models", "foo"])
    with pytest.raises(SystemExit):
        get_parser().parse_args(["-i", "x.nii.gz"])  # --models is required


def test_cpu_device_is_refused(tmp_path):
    from boa_b200.commands import analyze_ct
    with pytest.raises(RuntimeError, match="no CPU"):
        analyze_ct(tmp_path / "x.nii.gz", tmp_path, models={"total"}, device="cpu")
    with pytest.raises(NotImplementedError):
        analyze_ct(tmp_path / "x.nii.gz", tmp_path, models={"heartchambers_highres"}, device="gpu")  # licence-only


def test_nifti_roundtrip_and_orientation(tmp_path):
    rng = np.random.default_rng(0)
    d = rng.integers(-1000, 1000, size=(5, 6, 7)).astype(np.int16)
    for aff in (np.diag([1.5, 1.5, 1.5, 1.0]),
                np.array([[-1.5, 0, 0, 10], [0, -1.5, 0, 20], [0, 0, 5, 30], [0, 0, 0, 1.0]]),
                np.array([[0, 0, 2.0, 0], [1.0, 0, 0, 0], [0, -3.0, 0, 0], [0, 0, 0, 1.0]])):
        p = tmp_path / "a.nii.gz"
        nifti.save(p, d, aff, {1: "spleen", 2: "kidney_right"})
        im = nifti.load(p)
        assert np.array_equal(im.data, d) and np.allclose(im.affine, aff)
        c, zooms, order = nifti.to_canonical(im.data, im.affine)
        assert np.array_equal(nifti.from_canonical(c, order), d)
    c, zooms, _ = nifti.to_canonical(d, np.array([[-1.5, 0, 0, 0], [0, -1.5, 0, 0], [0, 0, 5, 0], [0, 0, 0, 1.0]]))
    assert zooms == (1.5, 1.5, 5.0) and np.array_equal(c, d[:, ::-1, ::-1])
    u8 = (d % 7).astype(np.uint8)
    nifti.save(tmp_path / "b.nii", u8, np.eye(4))
    assert np.array_equal(nifti.load(tmp_path / "b.nii").data, u8)


def test_nifti_parallel_deflate_is_one_valid_gzip_stream(tmp_path, monkeypatch):
    """Volumes above one chunk are written as concatenated gzip members deflated on several threads: the decompressed
    stream is byte-identical to the single-stream file, and the label extension survives."""
    import gzip
    rng = np.random.default_rng(1)
    d = rng.integers(0, 30, size=(40, 300, 400)).astype(np.int16)  # 9.6 MB: three chunks
    aff = np.diag([0.8, 0.8, 2.5, 1.0])
    monkeypatch.setenv("BOA_B200_GZIP_THREADS", "1")
    nifti.save(tmp_path / "one.nii.gz", d, aff, {1: "spleen"})
    monkeypatch.setenv("BOA_B200_GZIP_THREADS", "4")
    nifti.save(tmp_path / "many.nii.gz", d, aff, {1: "spleen"})
    one, many = (tmp_path / "one.nii.gz").read_bytes(), (tmp_path / "many.nii.gz").read_bytes()
    assert many.count(b"\x1f\x8b\x08") >= 3 and one != many
    assert gzip.decompress(one) == gzip.decompress(many)
    im = nifti.load(tmp_path / "many.nii.gz")
    assert np.array_equal(im.data, d) and np.allclose(im.affine, aff)
