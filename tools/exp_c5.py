"""C5 at one GPU: 100 M nodes / 1 B edges RMAT (scale 27), F = 64: structure build + mean aggregation fwd/bwd."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import rmat_edge_index
from keras_geometric_b200 import ops, _lib
from keras_geometric_b200.graph import GraphStructure
dev = torch.device("cuda", 0)
n, e, F = 100_000_000, 1_000_000_000, 64
if len(sys.argv) > 1: n, e = n // int(sys.argv[1]), e // int(sys.argv[1])
t0 = time.time(); ei = rmat_edge_index(n, e, 27, 0, dev); torch.cuda.synchronize(); t_gen = time.time() - t0
t0 = time.time(); g = GraphStructure(ei, n, n, 0); torch.cuda.synchronize(); t_csr = time.time() - t0
t0 = time.time(); g.csc; torch.cuda.synchronize(); t_csc = time.time() - t0
print(f"gen {t_gen:.2f}s  csr {t_csr*1e3:.0f} ms  csc {t_csc*1e3:.0f} ms  hubs {g.csr.n_hubs} chunks {g.csr.n_chunks} maxdeg {int(g.csr.deg.max())} "
      f"mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
# integer check: in-degree via the kernel equals bincount
ones = torch.ones((n, 4), device=dev)
s, _ = ops.gather_reduce_raw(ones, g.csr, _lib.OP_SUM)
deg = g.csr.deg.to(torch.float32)
assert torch.equal(s[:, 0], deg) and int(g.csr.rowptr[-1]) == e, "degree check failed"
del ones, s
x = torch.randn((n, F), device=dev)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
out = torch.empty((n, F), device=dev)
tf = timeit(lambda: ops.gather_reduce_raw(x, g.csr, _lib.OP_MEAN, out=out))
tb = timeit(lambda: ops.gather_reduce_raw(x, g.csc, _lib.OP_SUM, src_scale=g.csr.inv_deg, out=out))
bf = e * (4 * F + 4) + n * 4 * F + (n + 1) * 8
print(json.dumps({"config": "C5 1 GPU", "nodes": n, "edges": e, "F": F, "fwd_ms": tf, "bwd_ms": tb,
                  "fwd_GBps": bf / tf / 1e6, "bwd_GBps": (bf + 4 * e) / tb / 1e6, "GTEPS_fwd_bwd": e / (tf + tb) / 1e6,
                  "csr_build_ms": t_csr * 1e3, "csc_build_ms": t_csc * 1e3}))
