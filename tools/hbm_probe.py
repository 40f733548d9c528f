"""One launch of every HBM-bound pass on benchmark-sized inputs (512^3 volume, 25 classes, one 128^3 patch through the
head), for `ncu --set full` captures of their DRAM bytes / duration (profiles/rNN_hbm_passes.txt).  Development aid."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from boa_b200 import passes, zoo
from boa_b200.geometry import compute_gaussian
from boa_b200.plans import arch_from_plans
from boa_b200.predictor import Network, finalize_argmax

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
C, P = 25, 128
dev = torch.device("cuda", 0)
ct = torch.from_numpy(zoo.synthetic_ct((S, S, S), seed=3)).to(dev)
acc = torch.randn((C, S, S, S), device=dev)
w = torch.rand((S, S, S), device=dev) + 0.5
lab = torch.zeros((S, S, S), dtype=torch.uint8, device=dev)
# the head (+ Gaussian accumulate) as the network launches it: 8 patches of 128^3, C = 25
arch = arch_from_plans(zoo.default_plans((P, P, P), 32, 320, 6), "3d_fullres", 1, C)
net = Network(arch, zoo.random_state_dict(arch, 1), 0, 8)
net.set_graph(False)
g = torch.from_numpy(compute_gaussian((P, P, P)).astype(np.float32)).to(dev)
vol = torch.randn((S, S, S), device=dev)
origins = np.array([[0, 0, x] for x in (0, 96, 192, 288, 384)] + [[0, 96, x] for x in (0, 96, 192)], dtype=np.int32)
regions = (torch.arange(S * S * S, device=dev, dtype=torch.int32) % 12).to(torch.uint8).reshape(S, S, S)
mask = (regions > 3).to(torch.uint8)
# two identical rounds (the first warms up): per round 1 finalize, 1 tissue, 1 slice_stats, 1 label_hist, 3 erode_axis,
# 1 ct_normalize, 8 head launches = 16 launches of the kernels the ncu filter selects  (--launch-skip 16 --launch-count 16)
for _ in range(2):
    finalize_argmax(acc, w, list(range(C)), lab, True)
    tissues = passes.tissue_subclassify(ct, regions)
    passes.slice_label_stats(tissues, 8, ct=ct)
    passes.label_hu_hist(ct, regions, 118, -1024, 3072)
    passes.erode_box(mask)
    passes.ct_normalize(ct, -1024.0, 276.0, -370.0, 436.6)
    net.forward_accumulate(vol, origins, g, acc)
torch.cuda.synchronize()
print("hbm_probe done")
