"""Golden vectors for TotalSegmentator's resampling either side of the networks, from the REFERENCE's own
_external/totalsegmentator/resampling.py::change_spacing (-> resample_img -> scipy.ndimage.zoom) called the way
nnunet.py:457-475,685-687 calls it: CT to 1.5 mm / 3 mm / 5 mm thickness (order 3, dtype int32), label map back to the
input shape (order 0, target_shape, dtype uint8).  nibabel is stood in by a minimal image class (array in nibabel's
[x, y, z] order, float32 zooms like a NIfTI header).

    python tests/golden/make_golden_change_spacing.py      # needs /root/reference
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_stubs as S  # noqa: E402


class FakeNifti:
    def __init__(self, data, affine):
        self.data, self.affine = np.asarray(data), np.asarray(affine, dtype=np.float64)
        zooms = tuple(np.float32(np.sqrt((self.affine[:3, i] ** 2).sum())) for i in range(3))
        self.header = types.SimpleNamespace(get_zooms=lambda: zooms)
        self.shape = self.data.shape

    def get_fdata(self):
        return self.data.astype(np.float64)


def main():
    S.install()
    sys.modules["nibabel"].Nifti1Image = FakeNifti
    m = types.ModuleType("totalsegmentator")
    m.__path__ = [S.EXT + "/totalsegmentator"]
    sys.modules["totalsegmentator"] = m
    from totalsegmentator.resampling import change_spacing
    rng = np.random.default_rng(12)
    out = {}
    for name, shape_xyz, zooms, target in (
            ("to_1p5", (40, 44, 22), (0.8, 0.8, 2.5), [1.5, 1.5, 1.5]),
            ("fast_3mm", (36, 30, 41), (0.9765625, 0.9765625, 1.0), [3.0, 3.0, 3.0]),
            ("one_axis_at_target", (29, 33, 40), (0.8, 1.5, 1.0), [1.5, 1.5, 1.5]),
            ("thickness_5mm", (24, 20, 47), (0.7, 0.7, 1.5), None)):
        x = np.arange(shape_xyz[0])[:, None, None]
        ct = (rng.integers(-1000, 2000, size=shape_xyz) * 0.3 + 400 * np.sin(x / 5.0)).astype(np.int16)
        img = FakeNifti(ct, np.diag([*zooms, 1.0]))
        resample = target if target is not None else list(img.header.get_zooms()[:2]) + [5.0]   # nnunet.py:457-459
        rsp = change_spacing(img, resample, order=3, dtype=np.int32, nr_cpus=1)
        lab = rng.integers(0, 118, size=rsp.data.shape).astype(np.uint8)
        back = change_spacing(FakeNifti(lab, rsp.affine), resample, img.shape, order=0, dtype=np.uint8, nr_cpus=1,
                              force_affine=img.affine)
        assert back.data.shape == ct.shape
        # stored in the product's [z, y, x] order
        out[name + "_ct"] = np.ascontiguousarray(ct.transpose(2, 1, 0))
        out[name + "_spacing_zyx"] = np.array(zooms[::-1], dtype=np.float64)
        out[name + "_target_zyx"] = np.array([float(v) for v in resample][::-1], dtype=np.float64)
        out[name + "_resampled"] = np.ascontiguousarray(rsp.data.transpose(2, 1, 0))
        out[name + "_labels"] = np.ascontiguousarray(lab.transpose(2, 1, 0))
        out[name + "_labels_back"] = np.ascontiguousarray(back.data.transpose(2, 1, 0))
        print(name, ct.shape, "->", rsp.data.shape, rsp.data.dtype, back.data.dtype)
    np.savez_compressed(os.path.join(HERE, "change_spacing.npz"), **out)


if __name__ == "__main__":
    main()
