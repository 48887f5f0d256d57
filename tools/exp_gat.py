"""GATv2 / max-aggregation kernels on the C4 graph: time + algorithmic GB/s."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import rmat_edge_index, C4
from keras_geometric_b200 import ops, _lib
from keras_geometric_b200.graph import GraphStructure
dev = torch.device("cuda", 0)
div = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n, e = C4["nodes"] // div, C4["edges"] // div // 2 * 2
ei = rmat_edge_index(n, e, C4["rmat_scale"] - (div.bit_length() - 1), 0, dev)
g = GraphStructure(ei, n, n, n)   # with self loops like GATv2Conv
g.csc
nnz = g.nnz
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps): fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps
for H, C in [(8, 8), (4, 16), (1, 64), (8, 32)]:
    HC = H * C
    h = torch.randn((n, HC), device=dev, requires_grad=True)
    att = (torch.randn((1, H, C), device=dev) * 0.3).requires_grad_(True)
    R = torch.randn((n, HC), device=dev)
    out = ops.gatv2_aggregate(h, h, att, g, H, C, 0.2, None)
    tf = timeit(lambda: ops.gatv2_aggregate(h, h, att, g, H, C, 0.2, None))
    def fb():
        o = ops.gatv2_aggregate(h, h, att, g, H, C, 0.2, None)
        torch.autograd.grad((o * R).sum(), [h, att])
    tfb = timeit(fb)
    bf = nnz * (4 * HC + 4) + n * 4 * HC * 2 + n * H * 8 + (n + 1) * 8
    print(f"GATv2 H={H} C={C}: fwd {tf:.2f} ms ({bf/tf/1e6:.0f} GB/s)  fwd+bwd {tfb:.2f} ms  -> {nnz/tfb/1e6:.2f} GTEPS")
g2 = GraphStructure(ei, n, n, 0)
for F in (100, 256):
    x = torch.randn((n, F), device=dev, requires_grad=True)
    R = torch.randn((n, F), device=dev)
    for op in ("max", "sum", "mean"):
        tf = timeit(lambda: ops.gather_reduce(x, g2, op))
        def fb():
            o = ops.gather_reduce(x, g2, op)
            torch.autograd.grad((o * R).sum(), [x])
        tfb = timeit(fb)
        bf = e * (4 * F + 4) + n * 4 * F * (2 if op == "max" else 1) + (n + 1) * 8
        print(f"{op} F={F}: fwd {tf:.2f} ms ({bf/tf/1e6:.0f} GB/s)  fwd+bwd {tfb:.2f} ms -> {e/tfb/1e6:.2f} GTEPS")
