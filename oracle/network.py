"""Oracle: PlainConvUNet forward, restated on torch.nn.functional (TEST INFRASTRUCTURE, see oracle/__init__.py).

The network class is NOT in the reference repo: it is `dynamic_network_architectures.architectures.unet.PlainConvUNet`
from the un-vendored PyPI package dynamic-network-architectures==0.4.3 (pyproject.toml:33), located by name at
_external/nnunetv2/utilities/get_network_from_plans.py:9-43 and called at
_external/nnunetv2/inference/predict_from_raw_data.py:543.  Its published structure, restated here:

  encoder stage s : n_conv[s] x [Conv3d(k, pad k//2, stride = strides[s] on the first conv of the stage, bias)
                                 -> InstanceNorm3d(eps, affine) -> LeakyReLU(0.01)]
  decoder level j : ConvTranspose3d(kernel = stride = strides[-(j+1)], bias) -> cat((up, skip), 1)
                    -> n_conv_dec[j] x [Conv3d(k, stride 1) -> InstanceNorm3d -> LeakyReLU]
  head            : Conv3d(features[0] -> num_classes, 1x1x1, bias) on the last decoder level (deep supervision is
                    switched off at inference, predict_from_raw_data.py:110)

The kwargs come from the "old plans" reconstruction at _external/nnunetv2/utilities/plans_handling/plans_handler.py:36-97
(features = min(base * 2**i, max), conv_bias True, InstanceNorm eps 1e-5 affine, LeakyReLU default slope 0.01).
State-dict key names follow the package (`encoder.stages.{s}.0.convs.{i}.conv.weight`, `...norm.weight`,
`decoder.transpconvs.{j}.weight`, `decoder.stages.{j}.convs.{i}...`, `decoder.seg_layers.{j}.weight`).

Parity: UNPINNED by the reference (no golden logits exist); anchored on the call sites above.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

LEAKY_SLOPE = 0.01  # torch.nn.LeakyReLU default; nonlin_kwargs only sets inplace (plans_handler.py:78-80)


def arch_from_plans(plans: dict, configuration: str, num_input_channels: int, num_classes: int) -> dict:
    """plans_handler.py:36-97 (old format) / :142-152 (new format `architecture.arch_kwargs`)."""
    cfg = plans["configurations"][configuration]
    if "architecture" in cfg:
        kw = cfg["architecture"]["arch_kwargs"]
        cls = cfg["architecture"]["network_class_name"].rsplit(".", 1)[-1]
        features = list(kw["features_per_stage"])
        kernels, strides = kw["kernel_sizes"], kw["strides"]
        n_enc, n_dec = kw["n_conv_per_stage"], kw["n_conv_per_stage_decoder"]
        eps = kw.get("norm_op_kwargs", {}).get("eps", 1e-5)
    else:
        cls = cfg["UNet_class_name"]
        n_enc, n_dec = cfg["n_conv_per_stage_encoder"], cfg["n_conv_per_stage_decoder"]
        features = [min(cfg["UNet_base_num_features"] * 2 ** i, cfg["unet_max_num_features"]) for i in range(len(n_enc))]
        kernels, strides = cfg["conv_kernel_sizes"], cfg["pool_op_kernel_sizes"]
        eps = 1e-5
    if cls != "PlainConvUNet":
        raise RuntimeError(f"oracle restates PlainConvUNet only, got {cls}")
    return {
        "in_channels": num_input_channels, "num_classes": num_classes, "features": features,
        "kernels": [list(k) for k in kernels], "strides": [list(s) for s in strides],
        "n_conv_enc": list(n_enc), "n_conv_dec": list(n_dec), "eps": eps,
        "patch_size": list(cfg["patch_size"]),
    }


def _h(t: torch.Tensor, emulate_fp16: bool) -> torch.Tensor:
    return t.half().float() if emulate_fp16 else t


def _conv_norm_act(x, sd, prefix, stride, eps, emulate_fp16):
    w = sd[prefix + ".conv.weight"].float()
    b = sd[prefix + ".conv.bias"].float()
    pad = [k // 2 for k in w.shape[2:]]
    y = F.conv3d(_h(x, emulate_fp16), _h(w, emulate_fp16), b, stride=stride, padding=pad)
    g = sd[prefix + ".norm.weight"].float()
    beta = sd[prefix + ".norm.bias"].float()
    if emulate_fp16:
        # product numerics: statistics from the fp32 accumulators, affine applied to the fp16-stored conv output
        mean = y.mean(dim=(2, 3, 4), keepdim=True, dtype=torch.float64)
        var = (y.double() ** 2).mean(dim=(2, 3, 4), keepdim=True) - mean ** 2
        rstd = 1.0 / torch.sqrt(var.clamp_min(0) + eps)
        a = (g.double().view(1, -1, 1, 1, 1) * rstd).float()
        s = (beta.double().view(1, -1, 1, 1, 1) - mean * g.double().view(1, -1, 1, 1, 1) * rstd).float()
        z = y.half().float() * a + s
        return F.leaky_relu(z, LEAKY_SLOPE).half().float()
    z = F.instance_norm(y, weight=g, bias=beta, eps=eps)
    return F.leaky_relu(z, LEAKY_SLOPE)


@torch.inference_mode()
def unet_forward(arch: dict, sd: dict, x: torch.Tensor, emulate_fp16: bool = False) -> torch.Tensor:
    """x [B, Cin, D, H, W] fp32 -> logits [B, num_classes, D, H, W] fp32."""
    n_stages = len(arch["features"])
    skips = []
    for s in range(n_stages):
        for i in range(arch["n_conv_enc"][s]):
            stride = arch["strides"][s] if i == 0 else [1, 1, 1]
            x = _conv_norm_act(x, sd, f"encoder.stages.{s}.0.convs.{i}", stride, arch["eps"], emulate_fp16)
        skips.append(x)
    x = skips[-1]
    for j in range(n_stages - 1):
        wt = sd[f"decoder.transpconvs.{j}.weight"].float()
        bt = sd[f"decoder.transpconvs.{j}.bias"].float()
        st = arch["strides"][-(j + 1)]
        up = F.conv_transpose3d(_h(x, emulate_fp16), _h(wt, emulate_fp16), bt, stride=st)
        up = _h(up, emulate_fp16)
        x = torch.cat((up, skips[-(j + 2)]), dim=1)
        for i in range(arch["n_conv_dec"][j]):
            x = _conv_norm_act(x, sd, f"decoder.stages.{j}.convs.{i}", [1, 1, 1], arch["eps"], emulate_fp16)
    wh = sd[f"decoder.seg_layers.{n_stages - 2}.weight"].float()
    bh = sd[f"decoder.seg_layers.{n_stages - 2}.bias"].float()
    return F.conv3d(_h(x, emulate_fp16), _h(wh, emulate_fp16), bh)


def count_macs(arch: dict, patch=None) -> int:
    """Algorithmic multiply-accumulates of one forward over one patch (SURVEY.md 8a layer table)."""
    shape = list(patch or arch["patch_size"])
    feats = arch["features"]
    total, cin, shapes = 0, arch["in_channels"], []
    for s, f in enumerate(feats):
        for i in range(arch["n_conv_enc"][s]):
            if i == 0:
                shape = [d // st for d, st in zip(shape, arch["strides"][s])]
            k = 1
            for kk in arch["kernels"][s]:
                k *= kk
            total += k * cin * f * shape[0] * shape[1] * shape[2]
            cin = f
        shapes.append(list(shape))
    for j in range(len(feats) - 1):
        below, skip = feats[-(j + 1)], feats[-(j + 2)]
        st = arch["strides"][-(j + 1)]
        vin = shapes[-(j + 1)]
        total += st[0] * st[1] * st[2] * below * skip * vin[0] * vin[1] * vin[2]
        vout = shapes[-(j + 2)]
        k = 1
        for kk in arch["kernels"][-(j + 2)]:
            k *= kk
        cin = 2 * skip
        for i in range(arch["n_conv_dec"][j]):
            total += k * cin * skip * vout[0] * vout[1] * vout[2]
            cin = skip
    v0 = shapes[0]
    total += feats[0] * arch["num_classes"] * v0[0] * v0[1] * v0[2]
    return total
