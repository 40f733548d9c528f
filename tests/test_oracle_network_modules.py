"""The network oracle (oracle/network.py, torch.nn.functional) against a second, structurally independent restatement
built from torch.nn MODULES laid out the way dynamic_network_architectures==0.4.3 lays out PlainConvUNet
(encoder.stages.{s}.0.convs.{i}.{conv,norm,nonlin,all_modules}, decoder.{encoder,stages,transpconvs,seg_layers}):
`load_state_dict(strict=True)` proves that the synthetic checkpoints carry exactly the key set such a module tree
registers (alias keys included), and the forwards must agree.  The package itself is not installed here (SURVEY.md 8c:
network parity is unpinned by the reference); this pins the two restatements against each other."""
import numpy as np
import torch
from boa_b200 import zoo
from boa_b200.plans import arch_from_plans
from baseline.torch_unet import PlainConvUNet  # the module-tree restatement (also the bench's cuDNN baseline)
from oracle.network import unet_forward


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
