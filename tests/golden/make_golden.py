"""Generate the golden fixtures of tests/golden/ by running the REFERENCE's own functions (imported from
/root/reference with the stubs of _ref_stubs.py) on small seeded inputs.  Run in the development container only:

    python tests/golden/make_golden.py

Outputs (committed): geometry.json, gaussian.npz, ct_norm.npz, tissue.npz, measurements.json (+ inputs .npz),
bca.json (+ inputs .npz), class_maps_check.json.
"""
import hashlib
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import _ref_stubs as S  # noqa: E402

S.install()
for pkg, path in [("body_organ_analysis", S.BOA), ("body_organ_analysis.compute", S.BOA + "/compute")]:
    m = types.ModuleType(pkg)
    m.__path__ = [path]
    sys.modules[pkg] = m


def jsonable(o):
    if isinstance(o, dict):
        return {str(k): jsonable(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [jsonable(v) for v in o]
    if isinstance(o, (np.floating, np.integer)):
        return o.item()
    if isinstance(o, np.bool_):
        return bool(o)
    return o


def phantom(shape, seed):
    """Small synthetic (ct, total, regions, parts) with every structure the numerics look at."""
    rng = np.random.default_rng(seed)
    Z, Y, X = shape
    ct = rng.integers(-300, 400, size=shape).astype(np.int16)
    total = np.zeros(shape, dtype=np.uint8)
    from totalsegmentator.map_to_binary import class_map
    inv = {v: k for k, v in class_map["total"].items()}
    zz, yy, xx = np.meshgrid(np.arange(Z), np.arange(Y), np.arange(X), indexing="ij")
    # big autochthon blocks (survive the 6^3 erosion), lungs, aorta, a few vertebrae, random small organs
    total[4:Z - 4, 4:16, 2:14] = inv["autochthon_left"]
    total[4:Z - 4, 4:16, X - 14:X - 2] = inv["autochthon_right"]
    total[2:Z // 2, 18:30, 2:12] = inv["lung_upper_lobe_left"]
    total[Z // 2:Z - 2, 18:30, 2:12] = inv["lung_lower_lobe_left"]
    total[2:Z // 3, 18:30, X - 12:X - 2] = inv["lung_upper_lobe_right"]
    total[Z // 3:2 * Z // 3, 18:30, X - 12:X - 2] = inv["lung_middle_lobe_right"]
    total[2 * Z // 3:Z - 2, 18:30, X - 12:X - 2] = inv["lung_lower_lobe_right"]
    total[1:Z - 1, 18:28, 14:24] = inv["aorta"]
    for i, v in enumerate(["vertebrae_L5", "vertebrae_L4", "vertebrae_L3", "vertebrae_T12", "vertebrae_T11", "vertebrae_C7"]):
        total[i * (Z // 6):(i + 1) * (Z // 6), 0:3, 14:20] = inv[v]
    sm = rng.integers(1, 90, size=shape).astype(np.uint8)
    sm[np.isin(sm, [inv[k] for k in inv if k.startswith("vertebrae_") or k.startswith("autochthon") or k == "aorta"])] = inv["liver"]
    total = np.where((total == 0) & (rng.random(shape) < 0.15), sm, total)
    # autochthon: muscle-like HU with a fat streak, so that "minus fat" + 6^3 erosion leaves a non-empty core
    aut = (total == inv["autochthon_left"]) | (total == inv["autochthon_right"])
    ct[aut] = rng.integers(10, 90, size=int(aut.sum()))
    ct[4:Z - 4, 4:6, 2:14] = rng.integers(-180, -60, size=ct[4:Z - 4, 4:6, 2:14].shape)
    ao = total == inv["aorta"]
    ct[ao] = rng.integers(100, 300, size=int(ao.sum()))
    # lung fat window voxels
    ct[total == inv["lung_upper_lobe_left"]] = rng.integers(-260, -20, size=int((total == inv["lung_upper_lobe_left"]).sum()))
    regions = rng.integers(0, 12, size=shape).astype(np.uint8)
    regions[: Z // 2, :, : X // 2] = 3     # abdominal cavity (long enough for ABDOMEN at 5 mm)
    regions[Z // 2 - 4:, :, X // 2:] = 4   # thoracic cavity overlapping the abdomen range
    regions[Z // 2:, 0:6, 0:6] = 9         # mediastinum
    regions[Z // 2 + 2:Z - 8, 6:10, 0:6] = 7  # pericardium
    parts = rng.integers(0, 7, size=shape).astype(np.uint8)
    parts[:, Y // 4:3 * Y // 4, :] = 1
    return ct, total, regions, parts


def main():
    out = {}
    # ---------------------------------------------------------------- geometry + gaussian
    sw = S.load_by_path("ref_sw", S.EXT + "/nnunetv2/inference/sliding_window_prediction.py")
    cases = [((512, 512, 512), (128, 128, 128), 0.8), ((300, 512, 512), (128, 128, 128), 0.8),
             ((154, 512, 512), (128, 128, 128), 0.5), ((128, 128, 128), (128, 128, 128), 0.5),
             ((110, 64, 70), (64, 64, 64), 0.5), ((800, 1024, 1024), (192, 192, 192), 0.8),
             ((131, 129, 257), (128, 128, 128), 0.8), ((48, 64, 64), (32, 32, 32), 0.8), ((40, 48, 40), (32, 32, 32), 0.5)]
    geo = [{"image": c[0], "patch": c[1], "step": c[2],
            "steps": sw.compute_steps_for_sliding_window(c[0], c[1], c[2])} for c in cases]
    import torch
    gz = {}
    for tile in [(16, 16, 16), (8, 12, 10), (32, 32, 32)]:
        g = sw.compute_gaussian(tile, sigma_scale=1.0 / 8, value_scaling_factor=10, dtype=torch.float16,
                                device=torch.device("cpu"))
        gz["x".join(map(str, tile))] = g.numpy()
    g128 = sw.compute_gaussian((128, 128, 128), sigma_scale=1.0 / 8, value_scaling_factor=10, dtype=torch.float16,
                               device=torch.device("cpu")).numpy()
    geo_doc = {"steps": geo, "gaussian128": {"sha256": hashlib.sha256(g128.tobytes()).hexdigest(),
                                             "min": float(g128.min()), "max": float(g128.max()),
                                             "center": float(g128[64, 64, 64]), "corner": float(g128[0, 0, 0])}}
    json.dump(jsonable(geo_doc), open(os.path.join(HERE, "geometry.json"), "w"), indent=1)
    np.savez_compressed(os.path.join(HERE, "gaussian.npz"), **gz)

    # ---------------------------------------------------------------- CT normalisation
    norm = S.load_by_path("ref_norm", S.EXT + "/nnunetv2/preprocessing/normalization/default_normalization_schemes.py")
    props = {"mean": -370.00039267657144, "std": 436.5998675471528, "percentile_00_5": -1024.0, "percentile_99_5": 276.0}
    rng = np.random.default_rng(5)
    x = rng.integers(-1200, 3000, size=(6, 20, 24)).astype(np.int16)
    n = norm.CTNormalization(False, props, target_dtype=np.float32)
    y = n.run(x.astype(np.float32))
    np.savez_compressed(os.path.join(HERE, "ct_norm.npz"), x=x, y=y, props=json.dumps(props))

    # ---------------------------------------------------------------- tissue rules
    from body_composition_analysis.tissue.subclassification import subclassify_tissues
    import pathlib
    ct = rng.integers(-1100, 3100, size=(5, 32, 32)).astype(np.int16)
    # hit the inclusive bounds explicitly
    ct.flat[:12] = [-1000, -1001, 3000, 3001, -190, -191, -30, -29, 150, 151, -31, -189]
    reg = rng.integers(0, 12, size=ct.shape).astype(np.uint8)
    reg.flat[:12] = [5, 5, 5, 5, 1, 1, 3, 2, 2, 2, 2, 9]
    res = subclassify_tissues(S.FakeImage(ct), S.FakeImage(reg), pathlib.Path("/mem/tissues.nii.gz"))
    np.savez_compressed(os.path.join(HERE, "tissue.npz"), ct=ct, regions=reg, tissues=res.arr)

    # ---------------------------------------------------------------- total measurements
    meas = S.load_by_path("body_organ_analysis.compute.measurements", S.BOA + "/compute/measurements.py")
    shape = (44, 36, 40)
    ct, total, regions, parts = phantom(shape, 11)
    spacing = (1.5, 1.5, 1.5)
    S.FILES["/mem/ct.nii.gz"] = S.FakeImage(ct, spacing)
    folder = pathlib.Path("/mem/seg")

    S.FILES[str(folder / "total.nii.gz")] = S.FakeImage(total, spacing)
    orig_exists = pathlib.Path.exists
    pathlib.Path.exists = lambda self: str(self) in S.FILES or orig_exists(self)
    try:
        gold = meas.compute_measurements(pathlib.Path("/mem/ct.nii.gz"), folder, ["total"], cnr_adjustment=True)
    finally:
        pathlib.Path.exists = orig_exists
    pfav = S.FILES[str(folder / "ct_pfav.nii.gz")].arr
    json.dump(jsonable(gold), open(os.path.join(HERE, "measurements.json"), "w"), indent=0)
    np.savez_compressed(os.path.join(HERE, "phantom.npz"), ct=ct, total=total, regions=regions, parts=parts,
                        ct_pfav=pfav.astype(np.uint8), spacing=np.array(spacing))

    # ---------------------------------------------------------------- BCA report numerics
    from body_composition_analysis.report.builder import AggregatableBodyPart, Builder
    sp5 = (1.5, 1.5, 5.0)
    tissues = subclassify_tissues(S.FakeImage(ct, sp5), S.FakeImage(regions, sp5), pathlib.Path("/mem/t.nii.gz")).arr
    b = Builder(S.FakeImage(ct, sp5), S.FakeImage(parts, sp5), S.FakeImage(regions, sp5), S.FakeImage(tissues, sp5))
    part = AggregatableBodyPart.from_body_regions(S.FakeImage(regions, sp5))
    b.examined_body_part = part
    # create_vertebrae_info (commands.py:24-45) - commands.py itself needs nibabel; restate its loop with the
    # reference's class_map and AggregatableBodyPart
    from totalsegmentator.map_to_binary import class_map
    vmap = {v.removeprefix("vertebrae_"): k for k, v in class_map["total"].items() if v.startswith("vertebrae_")}
    vert = {}
    for vid, label in vmap.items():
        mask = np.where((total == label).any(axis=(1, 2)))[0]
        if len(mask) == 0:
            continue
        if (("C" in vid and AggregatableBodyPart.NECK not in part) or ("T" in vid and AggregatableBodyPart.THORAX not in part)
                or ("L" in vid and AggregatableBodyPart.ABDOMEN not in part)):
            continue
        vert[vid] = (int(mask.min()), int(mask.max() + 1))
    prepared = b.prepare(vert, total=S.FakeImage(total, sp5), total_measurements=None)
    js = b.create_json(**prepared)
    json.dump(jsonable({"json": js, "vertebrae": vert, "other_findings": prepared["other_findings"],
                        "body_part": int(part)}), open(os.path.join(HERE, "bca.json"), "w"), indent=0)
    np.savez_compressed(os.path.join(HERE, "phantom_bca.npz"), tissues=tissues, spacing=np.array(sp5))
    print("golden fixtures written:", sorted(f for f in os.listdir(HERE) if f.endswith((".json", ".npz"))))
    print("body part", part, "vertebrae", vert, "autochthon", gold["info"])


if __name__ == "__main__":
    main()
