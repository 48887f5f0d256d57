"""One launch of each of the north star's kernels on the C4 graph (after one warm-up pass) for an ncu capture:
   sum F=256 fwd/bwd, max F=256 fwd/bwd, GATv2 H=8 C=8 fwd/bwd.  Prints the launch labels in order."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import C4, rmat_edge_index
from keras_geometric_b200 import ops
from keras_geometric_b200.graph import GraphStructure
dev = torch.device("cuda", 0)
n, e = C4["nodes"], C4["edges"]
ei = rmat_edge_index(n, e, C4["rmat_scale"], 0, dev)
gen = torch.Generator(device=dev).manual_seed(11)
g0 = GraphStructure(ei, n, n, 0); g0.csc
g1 = GraphStructure(ei, n, n, n); g1.csc
x = torch.randn((n, 256), device=dev, generator=gen).requires_grad_(True)
R = torch.randn((n, 256), device=dev, generator=gen)
h = torch.randn((n, 64), device=dev, generator=gen).requires_grad_(True)
att = (torch.randn(64, device=dev, generator=gen) * 0.3).requires_grad_(True)
Rh = torch.randn((n, 64), device=dev, generator=gen)


def one_pass():
    o = ops.gather_reduce(x, g0, "sum"); torch.autograd.grad(o, x, R)
    o = ops.gather_reduce(x, g0, "max"); torch.autograd.grad(o, x, R)
    o = ops.gatv2_aggregate(h, h, att, g1, 8, 8); torch.autograd.grad(o, [h, att], Rh)
    torch.cuda.synchronize()


one_pass()
from keras_geometric_b200 import _lib
l0 = _lib.load().kgb_launch_count()
one_pass()
print("launches per pass:", _lib.load().kgb_launch_count() - l0)
