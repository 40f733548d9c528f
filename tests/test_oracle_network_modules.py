"""The network oracle (oracle/network.py, torch.nn.functional) against a second, structurally independent restatement
built from torch.nn MODULES laid out the way dynamic_network_architectures==0.4.3 lays out PlainConvUNet
(encoder.stages.{s}.0.convs.{i}.{conv,norm,nonlin,all_modules}, decoder.{encoder,stages,transpconvs,seg_layers}):
`load_state_dict(strict=True)` proves that the synthetic checkpoints carry exactly the key set such a module tree
registers (alias keys included), and the forwards must agree.  The package itself is not installed here (SURVEY.md 8c:
network parity is unpinned by the reference); this pins the two restatements against each other."""
import numpy as np
import torch
from torch import nn

from boa_b200 import zoo
from boa_b200.plans import arch_from_plans
from oracle.network import unet_forward


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


def test_module_restatement_loads_the_checkpoint_keys_and_agrees_with_the_oracle():
    for patch, stages, base, maxf, ncls in (((16, 16, 16), 3, 8, 32, 5), ((16, 24, 8), 2, 16, 16, 3)):
        arch = arch_from_plans(zoo.default_plans(patch, base, maxf, stages), "3d_fullres", 1, ncls)
        sd = zoo.random_state_dict(arch, 3)
        net = PlainConvUNet(arch).eval()
        missing, unexpected = net.load_state_dict(sd, strict=True)
        assert not missing and not unexpected
        x = torch.from_numpy(np.random.default_rng(0).standard_normal((2, 1, *patch)).astype(np.float32))
        with torch.inference_mode():
            ref = net(x)
        got = unet_forward(arch, sd, x, emulate_fp16=False)
        assert ref.shape == got.shape == (2, ncls, *patch)
        assert torch.allclose(got, ref, rtol=1e-5, atol=1e-5), float((got - ref).abs().max())
