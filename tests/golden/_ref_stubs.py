"""Stubs that let selected modules of the reference import in this image (no nibabel / SimpleITK / skimage / ...).

Used ONLY by tests/golden/make_golden.py (and tools/export_class_maps.py) in the development container, where
/root/reference exists; nothing here is imported at test or run time.  The fake SimpleITK wraps a numpy [z, y, x]
array + spacing - enough for the pure-numpy numerics of the reference to run unmodified.
"""
import sys
import types

import numpy as np
from scipy import ndimage

REF = "/root/reference"
BOA = REF + "/body_organ_analysis"
EXT = BOA + "/_external"


class FakeImage:
    def __init__(self, arr, spacing=(1.0, 1.0, 1.0)):
        self.arr = np.asarray(arr)
        self.spacing = tuple(float(s) for s in spacing)

    def GetSpacing(self):
        return self.spacing

    def GetDepth(self):
        return int(self.arr.shape[0])

    def GetSize(self):
        return tuple(int(s) for s in self.arr.shape[::-1])

    def CopyInformation(self, other):
        self.spacing = other.spacing


FILES = {}


def install():
    sitk = types.ModuleType("SimpleITK")
    sitk.Image = FakeImage
    sitk.GetArrayFromImage = lambda im: np.array(im.arr, copy=True)
    sitk.GetArrayViewFromImage = lambda im: im.arr
    sitk.GetImageFromArray = lambda a: FakeImage(a)
    sitk.ReadImage = lambda p, *a, **k: FILES[str(p)]
    sitk.WriteImage = lambda im, p, *a, **k: FILES.__setitem__(str(p), im)
    sys.modules["SimpleITK"] = sitk

    sk = types.ModuleType("skimage")
    morph = types.ModuleType("skimage.morphology")
    morph.pad_footprint = lambda fp, pad_end=True: np.pad(fp, [(0, 1)] * fp.ndim) if pad_end else np.pad(fp, [(1, 0)] * fp.ndim)
    # skimage 0.26 binary_erosion == scipy binary_erosion(structure=footprint, border_value=True)  
    morph.binary_erosion = lambda m, fp: ndimage.binary_erosion(m, structure=fp.astype(bool), border_value=True)
    measure = types.ModuleType("skimage.measure")
    measure.label = lambda m: ndimage.label(m, structure=np.ones((3, 3, 3)))[0]
    measure.regionprops = lambda lab: []
    sk.morphology, sk.measure = morph, measure
    sys.modules.update({"skimage": sk, "skimage.morphology": morph, "skimage.measure": measure})

    for name in ("weasyprint", "nibabel", "jinja2"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                m = types.ModuleType(name)
                if name == "jinja2":
                    m.Environment = lambda **k: None
                    m.FileSystemLoader = lambda *a, **k: None
                    m.select_autoescape = lambda *a, **k: None
                if name == "weasyprint":
                    m.HTML = object
                if name == "nibabel":
                    m.Nifti1Image = object
                    m.spatialimages = types.SimpleNamespace(SpatialImage=object)
                sys.modules[name] = m
    acvl = types.ModuleType("acvl_utils")
    cp = types.ModuleType("acvl_utils.cropping_and_padding")
    pad = types.ModuleType("acvl_utils.cropping_and_padding.padding")
    pad.pad_nd_image = None
    sys.modules.update({"acvl_utils": acvl, "acvl_utils.cropping_and_padding": cp,
                        "acvl_utils.cropping_and_padding.padding": pad})
    # plots: no-op stand-ins (visualisation is out of scope)
    base = "body_composition_analysis.report.plots"
    noop_img = lambda *a, **k: np.zeros((2, 2, 3), dtype=np.uint8)
    for mod, attrs in {
        "aggregation": {"create_aggregation_image": noop_img},
        "check": {"create_equidistant_overview": lambda *a, **k: []},
        "colors": {"BODY_REGION_COLOR_MAP": {}, "TISSUE_COLOR_MAP": {}, "TOTAL_COLOR_MAP": {}},
        "heatmaps": {"create_tissue_heatmaps": lambda *a, **k: []},
        "overview": {"create_tissue_summary": lambda *a, **k: types.SimpleNamespace(to_image=lambda **kw: b"")},
    }.items():
        m = types.ModuleType(f"{base}.{mod}")
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[f"{base}.{mod}"] = m
    plots = types.ModuleType(base)
    sys.modules[base] = plots
    ver = types.ModuleType("body_organ_analysis._version")
    ver.__githash__, ver.__version__ = "golden", "0"
    sys.modules["body_organ_analysis._version"] = ver
    for p in (EXT, REF):
        if p not in sys.path:
            sys.path.insert(0, p)


def load_by_path(name, path):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
