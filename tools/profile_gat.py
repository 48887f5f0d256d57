"""Per-kernel device time of one GATv2 forward+backward on the C4 graph (torch.profiler / CUPTI)."""
import sys, os, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import rmat_edge_index, C4
from keras_geometric_b200 import ops
from keras_geometric_b200.graph import GraphStructure
dev = torch.device("cuda", 0)
n, e = C4["nodes"], C4["edges"]
ei = rmat_edge_index(n, e, C4["rmat_scale"], 0, dev)
g = GraphStructure(ei, n, n, n)
g.csc
H, C = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 8)
h = torch.randn((n, H * C), device=dev, requires_grad=True)
att = (torch.randn((1, H, C), device=dev) * 0.3).requires_grad_(True)
R = torch.randn((n, H * C), device=dev)
def fb():
    o = ops.gatv2_aggregate(h, h, att, g, H, C, 0.2, None)
    torch.autograd.grad((o * R).sum(), [h, att])
for _ in range(2): fb()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2): fb()
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", ev.name)[:90]
        tot[name][0] += 1; tot[name][1] += ev.device_time
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{t/2/1e3:8.2f} ms  n={c//2:2d}  {k}")
