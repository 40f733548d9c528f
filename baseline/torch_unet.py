"""PlainConvUNet as torch.nn modules, laid out the way dynamic_network_architectures==0.4.3 lays it out
(encoder.stages.{s}.0.convs.{i}.{conv,norm,nonlin,all_modules}, decoder.{encoder,stages,transpconvs,seg_layers}), so
that a checkpoint's `network_weights` load with strict=True (alias keys included).  Kwargs as reconstructed by
_external/nnunetv2/utilities/plans_handling/plans_handler.py:36-97: conv bias, InstanceNorm3d(eps, affine),
LeakyReLU(inplace) with torch's default slope, no dropout.  MEASUREMENT / TEST INFRASTRUCTURE (see __init__.py)."""
from __future__ import annotations

import torch
from torch import nn


class ConvDropoutNormReLU(nn.Module):
    def __init__(self, cin, cout, ks, stride, eps):
        super().__init__()
        self.conv = nn.Conv3d(cin, cout, ks, stride, padding=[k // 2 for k in ks], bias=True)
        self.norm = nn.InstanceNorm3d(cout, eps=eps, affine=True)
        self.nonlin = nn.LeakyReLU(inplace=True)  # default negative_slope 0.01 (plans_handler.py:78-80)
        self.all_modules = nn.Sequential(self.conv, self.norm, self.nonlin)

    def forward(self, x):
        return self.all_modules(x)


class StackedConvBlocks(nn.Module):
    def __init__(self, n, cin, cout, ks, first_stride, eps):
        super().__init__()
        self.convs = nn.Sequential(*[ConvDropoutNormReLU(cin if i == 0 else cout, cout, ks,
                                                         first_stride if i == 0 else [1, 1, 1], eps) for i in range(n)])

    def forward(self, x):
        return self.convs(x)


class PlainConvEncoder(nn.Module):
    def __init__(self, arch):
        super().__init__()
        stages, cin = [], arch["in_channels"]
        for s, f in enumerate(arch["features"]):
            stages.append(nn.Sequential(StackedConvBlocks(arch["n_conv_enc"][s], cin, f, arch["kernels"][s],
                                                          arch["strides"][s], arch["eps"])))
            cin = f
        self.stages = nn.Sequential(*stages)

    def forward(self, x):
        skips = []
        for st in self.stages:
            x = st(x)
            skips.append(x)
        return skips


class UNetDecoder(nn.Module):
    def __init__(self, encoder, arch):
        super().__init__()
        self.encoder = encoder  # the package keeps a reference: the encoder's keys appear twice in a checkpoint
        feats, n = arch["features"], len(arch["features"])
        stages, ups, segs = [], [], []
        for j in range(n - 1):
            below, skip, st = feats[-(j + 1)], feats[-(j + 2)], arch["strides"][-(j + 1)]
            ups.append(nn.ConvTranspose3d(below, skip, st, st, bias=True))
            stages.append(StackedConvBlocks(arch["n_conv_dec"][j], 2 * skip, skip, arch["kernels"][-(j + 2)], [1, 1, 1],
                                            arch["eps"]))
            segs.append(nn.Conv3d(skip, arch["num_classes"], 1, 1, 0, bias=True))
        self.stages, self.transpconvs, self.seg_layers = nn.ModuleList(stages), nn.ModuleList(ups), nn.ModuleList(segs)

    def forward(self, skips):
        x = skips[-1]
        for j in range(len(self.stages)):
            x = self.transpconvs[j](x)
            x = torch.cat((x, skips[-(j + 2)]), 1)
            x = self.stages[j](x)
        return self.seg_layers[-1](x)  # deep supervision off: the last level only (predict_from_raw_data.py:110)


class PlainConvUNet(nn.Module):
    def __init__(self, arch):
        super().__init__()
        self.encoder = PlainConvEncoder(arch)
        self.decoder = UNetDecoder(self.encoder, arch)

    def forward(self, x):
        return self.decoder(self.encoder(x))


def build(arch: dict, state_dict: dict, device=None) -> PlainConvUNet:
    net = PlainConvUNet(arch).eval()
    net.load_state_dict(state_dict, strict=True)
    return net.to(device) if device is not None else net
