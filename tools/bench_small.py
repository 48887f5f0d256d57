"""Latency of the small (L2-resident, launch-bound) configs C1-C3 through the public layers."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import keras_geometric_b200 as kg
from keras_geometric_b200 import _lib
dev = torch.device("cuda", 0)

def sym_graph(n, e_dir, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    h = e_dir // 2
    s = torch.randint(0, n, (h,), device=dev, generator=g); d = torch.randint(0, n, (h,), device=dev, generator=g)
    return torch.stack([torch.cat([s, d]), torch.cat([d, s])]).to(torch.int32)

def timeit(step, reps=50):
    for _ in range(5): step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.load().kgb_launch_count()
    a.record()
    for _ in range(reps): step()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3, (_lib.load().kgb_launch_count() - l0) / reps

res = {}
# C1: 2-layer GCN on Cora-shaped graph
n, e = 2708, 10556
ei = sym_graph(n, e, 0); x = (torch.rand((n, 1433), device=dev) < 0.0127).float(); x = x / x.sum(1, keepdim=True).clamp(min=1)
l1, l2 = kg.GCNConv(16), kg.GCNConv(7); y = torch.randint(0, 7, (n,), device=dev)
def c1():
    h = l2([torch.relu(l1([x, ei])), ei]); loss = torch.nn.functional.cross_entropy(h, y); loss.backward()
c1(); us, nl = timeit(c1); res["C1_gcn_cora_fwd_bwd_us"] = us; res["C1_kernels_per_step"] = nl
# C2: 2-layer GATv2 8 heads on PubMed-shaped graph
n, e = 19717, 88648
ei = sym_graph(n, e, 0); x = torch.randn((n, 500), device=dev)
g1, g2 = kg.GATv2Conv(8, heads=8), kg.GATv2Conv(3, heads=1); y = torch.randint(0, 3, (n,), device=dev)
def c2():
    h = g2([torch.nn.functional.elu(g1([x, ei])), ei]); loss = torch.nn.functional.cross_entropy(h, y); loss.backward()
c2(); us, nl = timeit(c2); res["C2_gatv2_pubmed_fwd_bwd_us"] = us; res["C2_kernels_per_step"] = nl
# C3: 3-layer GIN on a batch of 4096 molecule-shaped graphs
rng = np.random.default_rng(0)
sizes = np.clip(np.round(rng.normal(25, 5, 4096)), 5, 60).astype(np.int64); offs = np.concatenate([[0], np.cumsum(sizes)])
src, dst = [], []
for gi, (o, s) in enumerate(zip(offs[:-1], sizes)):
    par = rng.integers(0, np.arange(1, s)); a = np.arange(1, s) + o; b = par + o
    extra = max(0, 27 - (s - 1)); ea = rng.integers(0, s, extra) + o; eb = rng.integers(0, s, extra) + o
    src += [a, b, ea, eb]; dst += [b, a, eb, ea]
ei = torch.from_numpy(np.stack([np.concatenate(src), np.concatenate(dst)]).astype(np.int32)).to(dev)
n = int(offs[-1]); x = torch.randn((n, 32), device=dev)
gin = [kg.GINConv(64, mlp_hidden=[64]) for _ in range(3)]
def c3():
    h = x
    for l in gin: h = l([h, ei])
    h.sum().backward()
c3(); us, nl = timeit(c3, 20); res["C3_gin_molecules_fwd_bwd_us"] = us; res["C3_kernels_per_step"] = nl; res["C3_nodes_edges"] = [n, int(ei.shape[1])]
print(json.dumps(res))
