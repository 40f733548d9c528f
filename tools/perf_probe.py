"""Per-layer timing of one forward at the TotalSegmentator geometry (development aid; run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from boa_b200 import zoo
from boa_b200.plans import arch_from_plans
from boa_b200.predictor import Network
from boa_b200.geometry import compute_gaussian

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
P = int(sys.argv[2]) if len(sys.argv) > 2 else 128
arch = arch_from_plans(zoo.default_plans((P, P, P), 32, 320, 6), "3d_fullres", 1, 25)
sd = zoo.random_state_dict(arch, 1)
net = Network(arch, sd, 0, B)
KIND = {0: "fold", 1: "taps", 2: "simt", 3: "tconv-taps", 4: "tconv-simt", 5: "first-simt"}
x = torch.randn(B, 1, P, P, P, device="cuda")
net.set_graph(False)
net.forward_logits(x[:1].contiguous()) if False else None
vol = torch.randn(P + 64, P + 64, P + 64, device="cuda")
g = torch.from_numpy(compute_gaussian((P, P, P)).astype(np.float32)).cuda()
acc = torch.zeros(25, *vol.shape, device="cuda")
origins = np.array([[0, 0, 0], [64, 64, 64], [0, 64, 0], [64, 0, 64]] * (B + 1), dtype=np.int32)[: 4 * B]
net.forward_accumulate(vol, origins, g, acc)
torch.cuda.synchronize()
desc = net.describe()
best = None
for rep in range(3):
    ms = net.time_layers()
    best = ms if best is None else [min(a, b) for a, b in zip(best, ms)]
tot_ms, tot_macs = 0, 0
for (name, kind, macs), t in zip(desc, best):
    tf = 2 * macs * B / (t * 1e-3) / 1e12
    print(f"{name:34s} {KIND[kind]:10s} {macs/1e9:8.2f} GMAC/patch  {t:8.3f} ms  {tf:8.1f} TFLOP/s")
    tot_ms += t; tot_macs += macs
print(f"conv kernels: {tot_ms:.3f} ms for B={B} -> {tot_ms/B:.3f} ms/patch, {2*tot_macs*B/(tot_ms*1e-3)/1e12:.1f} TFLOP/s")
for graph in (False, True):
    net.set_graph(graph)
    n = origins.shape[0]
    for _ in range(2):
        net.forward_accumulate(vol, origins, g, acc)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 3
    for _ in range(reps):
        net.forward_accumulate(vol, origins, g, acc)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * n)
    print(f"forward_accumulate graph={graph}: {ms:.3f} ms/patch  ({2*net.macs_per_patch/(ms*1e-3)/1e12:.1f} TFLOP/s whole forward)")
