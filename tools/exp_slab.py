"""Experiment: feature-slab tiling of the gather-reduce (L2-resident slabs) on the C4 graph."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import rmat_edge_index, C4
from keras_geometric_b200 import ops, _lib
from keras_geometric_b200.graph import GraphStructure

dev = torch.device("cuda", 0)
div = int(sys.argv[1]) if len(sys.argv) > 1 else 1
noslab = "--noslab" in sys.argv
n, e = C4["nodes"] // div, C4["edges"] // div // 2 * 2
ei = rmat_edge_index(n, e, C4["rmat_scale"] - (div.bit_length() - 1), 0, dev)
g = GraphStructure(ei, n, n, 0)
print("graph", n, e, "hubs", g.csr.n_hubs, "chunks", g.csr.n_chunks, "maxdeg", int(g.csr.deg.max()))

def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(reps): fn()
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1) / reps

for F in (256, 128, 104, 100, 64, 48, 47):
    x = torch.randn((n, F), device=dev)
    out = torch.empty((n, F), device=dev)
    base = timeit(lambda: ops.gather_reduce_raw(x, g.csr, _lib.OP_MEAN, out=out))
    alg = (e * (4 * F + 4) + n * 4 * F + (n + 1) * 8) / 1e9
    print(f"F={F}: full-row {base:.3f} ms  {alg/base*1e3:.0f} GB/s")
    if F % 4 or noslab: continue
    for w in (4, 8, 16, 32, 64, 128):
        if w >= F or F % w: continue
        def run():
            for f0 in range(0, F, w):
                ops.gather_reduce_raw(x[:, f0:f0 + w], g.csr, _lib.OP_MEAN, out=out[:, f0:f0 + w])
        t = timeit(run)
        print(f"   slab w={w}: {t:.3f} ms  ({alg/t*1e3:.0f} GB/s algorithmic)")
