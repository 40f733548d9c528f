"""Does an HBM-bound kernel overlap the persistent tcgen05 conv kernels? (development aid)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from boa_b200 import zoo
from boa_b200.plans import arch_from_plans
from boa_b200.predictor import Network

os.environ["BOA_B200_LANES"] = "1"
B, P = 8, 128
arch = arch_from_plans(zoo.default_plans((P, P, P), 32, 320, 6), "3d_fullres", 1, 25)
net = Network(arch, zoo.random_state_dict(arch, 1), 0, B)
net.set_graph(False)
x = torch.randn(B, 1, P, P, P, device="cuda")
a = torch.empty(512 * 1024 * 1024, dtype=torch.float16, device="cuda")
b = torch.empty_like(a)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def run_net(n):
    with torch.cuda.stream(s1):
        for _ in range(n):
            net.forward_logits(x)

def run_copy(n):
    with torch.cuda.stream(s2):
        for _ in range(n):
            b.copy_(a)

def timed(f):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    f()
    torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)

run_net(1); run_copy(1)
tn = timed(lambda: run_net(4))
tc = timed(lambda: run_copy(40))
tb = timed(lambda: (run_net(4), run_copy(40)))
print(f"net alone {tn:.1f} ms, copy alone {tc:.1f} ms ({40*2*a.numel()*2/tc/1e6:.0f} GB/s), both {tb:.1f} ms (sum {tn+tc:.1f}, max {max(tn,tc):.1f})")
