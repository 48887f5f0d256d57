"""Per-kernel device time of one CSR (by target) build of the C4 edge list (torch.profiler / CUPTI)."""
import sys, os, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from bench import rmat_edge_index, C4
from keras_geometric_b200.graph import GraphStructure
dev = torch.device("cuda", 0)
n, e = C4["nodes"], C4["edges"]
ei = rmat_edge_index(n, e, C4["rmat_scale"], 0, dev)
for _ in range(2):
    GraphStructure(ei, n, n, 0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        GraphStructure(ei, n, n, 0)
    torch.cuda.synchronize()
tot = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\(.*", "", ev.name)[:90]
        tot[name][0] += 1; tot[name][1] += ev.device_time
for k, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"{t/2/1e3:8.3f} ms  n={c//2:2d}  {k}")
