"""Drop-in for the reference's sliding-window predictor, running on libboa_b200.

Mirrors `nnUNetPredictor` (_external/nnunetv2/inference/predict_from_raw_data.py:36-680): same constructor keywords,
`initialize_from_trained_model_folder`, `predict_sliding_window_return_logits`, `predict_logits_from_preprocessed_data`
with the same argument meaning and error behaviour (RuntimeError on non-finite logits, :622-625), plus the native
entry `predict_labels` that never materialises / copies the C x V logits (the reference's D2H + numpy argmax,
:386 and export_prediction.py:38, is fused on the device).

PyTorch is used for device memory and streams only; every arithmetic step is a kernel of libboa_b200.  There is no
CPU path: device="cpu" raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib
from .geometry import compute_gaussian, pad_to_patch, sliding_window_origins
from .plans import ModelSpec, load_model_folder

DEFAULT_MAX_BATCH = int(os.environ.get("BOA_B200_BATCH", "4"))


def _device_index(device) -> int:
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("boa_b200 has no CPU implementation of the network; pass a CUDA device")
    return device.index if device.index is not None else torch.cuda.current_device()


class Network:
    """One set of PlainConvUNet weights resident on a GPU (a `boa_net*`)."""

    def __init__(self, arch: dict, state_dict: dict, device_index: int, max_batch: int, donor: "Network | None" = None):
        L = _lib.lib()
        a = _lib.BoaArch()
        n = len(arch["features"])
        if n > _lib.BOA_MAX_STAGES:
            raise NotImplementedError(f"{n} stages > {_lib.BOA_MAX_STAGES}")
        a.n_stages, a.in_channels, a.num_classes = n, arch["in_channels"], arch["num_classes"]
        for s in range(n):
            a.features[s] = arch["features"][s]
            a.n_conv_enc[s] = arch["n_conv_enc"][s]
            for k in range(3):
                a.strides[s][k] = arch["strides"][s][k]
                a.kernels[s][k] = arch["kernels"][s][k]
        for j in range(n - 1):
            a.n_conv_dec[j] = arch["n_conv_dec"][j]
        for k in range(3):
            a.patch[k] = arch["patch_size"][k]
        a.eps, a.leaky_slope = arch.get("eps", 1e-5), arch.get("leaky_slope", 0.01)
        self.arch, self.device_index, self.max_batch = arch, device_index, max_batch
        self.handle = C.c_void_p()
        _lib.check(L.boa_net_create(C.byref(a), device_index, max_batch, C.byref(self.handle)))
        try:
            if donor is not None:
                _lib.check(L.boa_net_share_workspace(self.handle, donor.handle))
                self._donor = donor  # keep alive
            for key, t in state_dict.items():
                arr = np.ascontiguousarray(t.detach().to("cpu", torch.float32).numpy())
                shape = (C.c_int64 * max(arr.ndim, 1))(*arr.shape)
                _lib.check(L.boa_net_set_tensor(self.handle, key.encode(), arr.ctypes.data_as(C.c_void_p), shape, arr.ndim))
            _lib.check(L.boa_net_finalize(self.handle))
            if os.environ.get("BOA_B200_MODE", "") == "simt":
                _lib.check(L.boa_net_set_mode(self.handle, 1))
            # plain launches by default: replaying the captured graphs measured 4 % slower than launching the ~100
            # kernels of a batch directly (each runs for ~100 us, so launch overhead is hidden either way)
            if os.environ.get("BOA_B200_GRAPH", "0") == "0":
                _lib.check(L.boa_net_set_graph(self.handle, 0))
        except Exception:
            L.boa_net_destroy(self.handle)
            self.handle = None
            raise

    @property
    def macs_per_patch(self) -> int:
        return int(_lib.lib().boa_net_macs_per_patch(self.handle))

    def set_mode(self, mode: int) -> None:
        _lib.check(_lib.lib().boa_net_set_mode(self.handle, mode))

    def set_graph(self, enable: bool) -> None:
        _lib.check(_lib.lib().boa_net_set_graph(self.handle, int(enable)))

    def describe(self):
        L = _lib.lib()
        cap, stride = 128, 96
        kinds = (C.c_int32 * cap)()
        macs = (C.c_double * cap)()
        names = C.create_string_buffer(cap * stride)
        n = L.boa_net_describe(self.handle, cap, kinds, macs, names, stride)
        return [(names.raw[i * stride:(i + 1) * stride].split(b"\0")[0].decode(), int(kinds[i]), float(macs[i]))
                for i in range(n)]

    def time_layers(self):
        L = _lib.lib()
        cap = 128
        ms = (C.c_float * cap)()
        _lib.check(L.boa_net_time_layers(self.handle, cap, ms, _lib.stream_ptr()))
        return [float(ms[i]) for i in range(len(self.describe()))]

    def enable_timing(self, enable: bool) -> None:
        """CUDA events around every conv kernel of the following forward_accumulate calls (single-lane schedule)."""
        _lib.check(_lib.lib().boa_net_enable_timing(self.handle, int(enable)))

    def read_timing_kinds(self, reset: bool = True):
        """[(ms, flop, launches)] per kernel kind (see include/boa_b200.h boa_net_read_timing_kinds)."""
        n = 7
        ms, fl, la = (C.c_double * n)(), (C.c_double * n)(), (C.c_int64 * n)()
        _lib.check(_lib.lib().boa_net_read_timing_kinds(self.handle, n, ms, fl, la, int(reset)))
        return [(float(ms[i]), float(fl[i]), int(la[i])) for i in range(n)]

    def forward_accumulate(self, vol: torch.Tensor, origins: np.ndarray, gaussian: torch.Tensor, acc) -> None:
        """vol fp32 [d0,d1,d2] (device), origins int32 [n,3] (host), gaussian fp32 [p0,p1,p2], acc fp32 [C,d0,d1,d2]:
        a tensor, or the device address (int) of such a buffer (dist.PeerExchange memory)."""
        assert vol.is_cuda and vol.dtype == torch.float32 and vol.is_contiguous()
        if isinstance(acc, torch.Tensor):
            assert acc.is_cuda and acc.dtype == torch.float32 and acc.is_contiguous()
            acc_ptr = _lib.ptr(acc)
        else:
            acc_ptr = C.c_void_p(int(acc))
        origins = np.ascontiguousarray(origins, dtype=np.int32)
        _lib.check(_lib.lib().boa_net_forward_accumulate(
            self.handle, _lib.ptr(vol), _lib.i32x3(vol.shape), origins.ctypes.data_as(C.POINTER(C.c_int32)),
            int(origins.shape[0]), _lib.ptr(gaussian), acc_ptr, _lib.stream_ptr()))

    def forward_logits(self, patches: torch.Tensor) -> torch.Tensor:
        """Parity entry: patches fp32 [n,Cin,p0,p1,p2] (device) -> raw logits fp32 [n,C,p0,p1,p2]."""
        assert patches.is_cuda and patches.dtype == torch.float32 and patches.is_contiguous()
        out = torch.empty((patches.shape[0], self.arch["num_classes"], *patches.shape[2:]), dtype=torch.float32,
                          device=patches.device)
        _lib.check(_lib.lib().boa_net_forward_logits(self.handle, _lib.ptr(patches), int(patches.shape[0]),
                                                     _lib.ptr(out), _lib.stream_ptr()))
        return out

    def close(self) -> None:
        if getattr(self, "handle", None):
            _lib.lib().boa_net_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_weight_sum_cache: dict = {}
_WEIGHT_SUM_CACHE_ENTRIES = 4


def weight_sum(shape, patch, origins: np.ndarray, gaussian: torch.Tensor, kind: str = "gaussian") -> torch.Tensor:
    """`n_predictions` (predict_from_raw_data.py:614): input independent, accumulated in slicer order; cached.
    `kind` names the content of `gaussian` ("gaussian": compute_gaussian(patch, 1/8, 10), "ones"): every predictor owns
    its own copy of the same map, so the key must not depend on which copy is passed."""
    key = (tuple(int(v) for v in shape), tuple(int(v) for v in patch), origins.tobytes(), kind, gaussian.device.index)
    hit = _weight_sum_cache.get(key)
    if hit is not None:
        _weight_sum_cache[key] = _weight_sum_cache.pop(key)  # most recently used last
        return hit
    while len(_weight_sum_cache) >= _WEIGHT_SUM_CACHE_ENTRIES:
        _weight_sum_cache.pop(next(iter(_weight_sum_cache)))
    w = torch.zeros(tuple(shape), dtype=torch.float32, device=gaussian.device)
    o = np.ascontiguousarray(origins, dtype=np.int32)
    _lib.check(_lib.lib().boa_accumulate_weights(o.ctypes.data_as(C.POINTER(C.c_int32)), int(o.shape[0]),
                                                 _lib.i32x3(patch), _lib.ptr(gaussian), _lib.ptr(w),
                                                 _lib.i32x3(shape), _lib.stream_ptr()))
    _weight_sum_cache[key] = w
    return w


INF_MESSAGE = ("Encountered inf in predicted array. Aborting... If this problem persists, reduce "
               "value_scaling_factor in compute_gaussian or increase the dtype of predicted_logits to fp32")


def raise_if_nonfinite(flags) -> None:
    """predict_from_raw_data.py:622-625 for flags collected with finalize_argmax(..., defer=flags): ONE device -> host
    read per task instead of one per network (each read drains the launch queue)."""
    if flags and int(torch.stack([f.reshape(()) for f in flags]).sum().item()) != 0:
        raise RuntimeError(INF_MESSAGE)


def finalize_argmax(acc: torch.Tensor, wsum: torch.Tensor, lut=None, label_inout: torch.Tensor | None = None,
                    overwrite_nonzero_only: bool = False, defer: list | None = None) -> torch.Tensor:
    """`logits /= n`, isinf check, argmax(0) (first max wins), part->global LUT, non-zero overwrite merge
    (predict_from_raw_data.py:620-625; label_handling.py:178; totalsegmentator/nnunet.py:553-556) in one pass.
    defer: a list that collects the non-finite flag instead of reading it here (see raise_if_nonfinite)."""
    Cn = acc.shape[0]
    V = wsum.numel()
    if label_inout is None:
        label_inout = torch.zeros(tuple(wsum.shape), dtype=torch.uint8, device=acc.device)
    lut_arr = (C.c_uint8 * Cn)(*(range(Cn) if lut is None else [int(v) for v in lut]))
    bad = torch.zeros(1, dtype=torch.int32, device=acc.device)
    _lib.check(_lib.lib().boa_finalize_argmax(_lib.ptr(acc), _lib.ptr(wsum), Cn, V, lut_arr,
                                              int(overwrite_nonzero_only), _lib.ptr(label_inout), _lib.ptr(bad),
                                              _lib.stream_ptr()))
    if defer is not None:
        defer.append(bad)
    elif int(bad.item()) != 0:
        raise RuntimeError(INF_MESSAGE)
    return label_inout


def finalize_argmax_resampled(acc: torch.Tensor, wsum: torch.Tensor, out_shape, separate_z: bool, lut=None,
                              defer: list | None = None) -> torch.Tensor:
    """finalize_argmax when the network ran on a grid nnU-Net resampled the volume to: logits / n are resampled with
    order 1 to out_shape before the argmax (export_prediction.py:25-38), fused into the argmax pass."""
    Cn = acc.shape[0]
    out = torch.zeros(tuple(int(v) for v in out_shape), dtype=torch.uint8, device=acc.device)
    lut_arr = (C.c_uint8 * Cn)(*(range(Cn) if lut is None else [int(v) for v in lut]))
    bad = torch.zeros(1, dtype=torch.int32, device=acc.device)
    _lib.check(_lib.lib().boa_finalize_argmax_resampled(_lib.ptr(acc), _lib.ptr(wsum), Cn, _lib.i32x3(wsum.shape),
                                                        _lib.i32x3(out_shape), int(separate_z), lut_arr, 0,
                                                        _lib.ptr(out), _lib.ptr(bad), _lib.stream_ptr()))
    if defer is not None:
        defer.append(bad)
    elif int(bad.item()) != 0:
        raise RuntimeError(INF_MESSAGE)
    return out


class nnUNetPredictor:
    def __init__(self, tile_step_size: float = 0.5, use_gaussian: bool = True, use_mirroring: bool = True,
                 perform_everything_on_device: bool = True, device=None, verbose: bool = False,
                 verbose_preprocessing: bool = False, allow_tqdm: bool = True, max_batch: int | None = None,
                 workspace_donor: "nnUNetPredictor | None" = None):
        self.tile_step_size = tile_step_size
        self.use_gaussian = use_gaussian
        self.use_mirroring = use_mirroring
        self.perform_everything_on_device = perform_everything_on_device
        self.verbose, self.verbose_preprocessing, self.allow_tqdm = verbose, verbose_preprocessing, allow_tqdm
        self.device_index = _device_index(device)
        self.device = torch.device("cuda", self.device_index)
        self.max_batch = max_batch or DEFAULT_MAX_BATCH
        self.networks: list[Network] = []
        self.spec: ModelSpec | None = None
        self._donor = workspace_donor
        self._gaussian = None

    # -- initialisation -------------------------------------------------------------------------------------------
    def initialize_from_trained_model_folder(self, model_training_output_dir: str, use_folds,
                                             checkpoint_name: str = "checkpoint_final.pth") -> None:
        spec = load_model_folder(model_training_output_dir, use_folds, checkpoint_name)
        self.manual_initialization(spec)

    def manual_initialization(self, spec: ModelSpec) -> None:
        if self.use_mirroring:
            # checkpoints of the *NoMirroring trainers disallow mirroring: the reference then runs without TTA
            # (predict_from_raw_data.py:542-557) and so do we.  Mirroring a checkpoint that allows it is not implemented.
            if getattr(spec, "allowed_mirroring_axes", None):
                raise NotImplementedError("test-time mirroring is not implemented: construct the predictor with "
                                          "use_mirroring=False (what the reference does on this path, tta=False)")
            self.use_mirroring = False
        self.spec = spec
        donor = self._donor.networks[0] if (self._donor is not None and self._donor.networks) else None
        if donor is not None and (donor.arch["features"] != spec.arch["features"] or
                                  donor.arch["patch_size"] != spec.arch["patch_size"] or
                                  donor.arch["strides"] != spec.arch["strides"] or
                                  donor.arch["n_conv_enc"] != spec.arch["n_conv_enc"] or
                                  donor.arch["n_conv_dec"] != spec.arch["n_conv_dec"] or
                                  donor.max_batch != self.max_batch or donor.device_index != self.device_index):
            donor = None
        self.networks = []
        for sd in spec.fold_weights:
            net = Network(spec.arch, sd, self.device_index, self.max_batch, donor)
            donor = donor or net
            self.networks.append(net)

    @property
    def patch_size(self):
        return list(self.spec.arch["patch_size"])

    @property
    def num_classes(self) -> int:
        return self.spec.arch["num_classes"]

    @property
    def gaussian_kind(self) -> str:
        return "gaussian" if self.use_gaussian else "ones"

    def gaussian(self) -> torch.Tensor:
        if self._gaussian is None:
            if self.use_gaussian:
                g = compute_gaussian(tuple(self.patch_size), 1.0 / 8, 10.0).astype(np.float32)
            else:
                g = np.ones(self.patch_size, dtype=np.float32)
            self._gaussian = torch.from_numpy(g).to(self.device)
        return self._gaussian

    # -- core -------------------------------------------------------------------------------------------------------
    def _prepare(self, input_image: torch.Tensor):
        if input_image.ndim != 4:
            raise ValueError("input_image must be a 4D tensor (c, x, y, z)")
        if input_image.shape[0] != 1:
            raise NotImplementedError("only single-channel input is implemented")
        vol = input_image[0].to(self.device, torch.float32)
        pads, unpad = pad_to_patch(vol.shape, self.patch_size)
        if any(b or a for b, a in pads):
            flat = [v for b, a in reversed(pads) for v in (b, a)]
            vol = torch.nn.functional.pad(vol, flat, mode="constant", value=0.0)
        vol = vol.contiguous()
        origins = sliding_window_origins(vol.shape, self.patch_size, self.tile_step_size)
        return vol, origins, unpad

    def accumulate(self, vol: torch.Tensor, origins: np.ndarray, acc=None):
        """Run every fold over the given patch origins, adding `logits * gaussian` into `acc` [C, *vol.shape] (a tensor
        or a device address, see Network.forward_accumulate)."""
        if acc is None:
            acc = torch.zeros((self.num_classes, *vol.shape), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device_index):
            for net in self.networks:
                net.forward_accumulate(vol, origins, self.gaussian(), acc)
        return acc

    @torch.inference_mode()
    def predict_sliding_window_return_logits(self, input_image: torch.Tensor) -> torch.Tensor:
        """[c,x,y,z] -> logits [C,x,y,z] fp32 on the device (mean over folds when several are loaded)."""
        with torch.cuda.device(self.device_index):
            vol, origins, unpad = self._prepare(input_image)
            acc = self.accumulate(vol, origins)
            w = weight_sum(vol.shape, self.patch_size, origins, self.gaussian(), self.gaussian_kind)
            bad = torch.zeros(1, dtype=torch.int32, device=self.device)
            _lib.check(_lib.lib().boa_normalize_logits(_lib.ptr(acc), _lib.ptr(w), self.num_classes, w.numel(),
                                                       float(len(self.networks)), _lib.ptr(bad), _lib.stream_ptr()))
            raise_if_nonfinite([bad])
            return acc[(slice(None), *unpad)]

    @torch.inference_mode()
    def predict_logits_from_preprocessed_data(self, data: torch.Tensor) -> torch.Tensor:
        return self.predict_sliding_window_return_logits(data).to("cpu")

    @torch.inference_mode()
    def predict_labels(self, input_image: torch.Tensor, lut=None, label_inout: torch.Tensor | None = None,
                       overwrite_nonzero_only: bool = False, defer: list | None = None, resample_to=None,
                       separate_z: bool = False) -> torch.Tensor:
        """[c,x,y,z] -> uint8 label map [x,y,z] on the device; logits never leave HBM.
        resample_to: shape of the volume BEFORE nnU-Net's preprocessing resampled it to the plan's spacing - the label
        map is produced on that grid (logits resampled with order 1 inside the argmax pass)."""
        with torch.cuda.device(self.device_index):
            vol, origins, unpad = self._prepare(input_image)
            acc = self.accumulate(vol, origins)
            w = weight_sum(vol.shape, self.patch_size, origins, self.gaussian(), self.gaussian_kind)
            padded = any(s.start != 0 or s.stop != d for s, d in zip(unpad, vol.shape))
            if resample_to is not None and tuple(resample_to) != tuple(input_image.shape[1:]):
                if padded:  # un-pad first (predict_from_raw_data.py:679), then resample
                    acc = acc[(slice(None), *unpad)].contiguous()
                    w = w[unpad].contiguous()
                lab = finalize_argmax_resampled(acc, w, resample_to, separate_z, lut, defer=defer)
                if label_inout is None:
                    return lab
                if overwrite_nonzero_only:
                    label_inout[lab != 0] = lab[lab != 0]
                else:
                    label_inout.copy_(lab)
                return label_inout
            if padded:
                lab = finalize_argmax(acc, w, lut, defer=defer)[unpad].contiguous()
                if label_inout is None:
                    return lab
                if overwrite_nonzero_only:
                    label_inout[lab != 0] = lab[lab != 0]
                else:
                    label_inout.copy_(lab)
                return label_inout
            return finalize_argmax(acc, w, lut, label_inout, overwrite_nonzero_only, defer=defer)
