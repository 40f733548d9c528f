"""Do the HBM-bound passes of one forward co-run with the persistent tcgen05 conv kernels of another?  Two networks
with separate workspaces on two streams: one launches only its convolutions, the other only its passes
(BOA_B200_DEBUG_ONLY).  Prints each alone and both together (development aid; run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BOA_B200_LANES"] = "1"
import torch
from boa_b200 import zoo
from boa_b200.plans import arch_from_plans
from boa_b200.predictor import Network

B, P = 8, 128
arch = arch_from_plans(zoo.default_plans((P, P, P), 32, 320, 6), "3d_fullres", 1, 25)
sd = zoo.random_state_dict(arch, 1)
os.environ["BOA_B200_DEBUG_ONLY"] = "conv"
net_c = Network(arch, sd, 0, B)
os.environ["BOA_B200_DEBUG_ONLY"] = "thin"
net_t = Network(arch, sd, 0, B)
os.environ.pop("BOA_B200_DEBUG_ONLY")
for n in (net_c, net_t):
    n.set_graph(False)
x = torch.randn(B, 1, P, P, P, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
REPS = 4


def run(net, stream):
    with torch.cuda.stream(stream):
        for _ in range(REPS):
            net.forward_logits(x)


def timed(f):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    f()
    torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / REPS


run(net_c, s1); run(net_t, s2)
tc = timed(lambda: run(net_c, s1))
tt = timed(lambda: run(net_t, s2))
tb = timed(lambda: (run(net_c, s1), run(net_t, s2)))
tb2 = timed(lambda: (run(net_t, s2), run(net_c, s1)))
print(f"THIN={os.environ.get('BOA_B200_THIN', '1')}: convs alone {tc:.2f} ms/batch, passes alone {tt:.2f}, both {tb:.2f} / {tb2:.2f} "
      f"(sum {tc + tt:.2f}, max {max(tc, tt):.2f})")
