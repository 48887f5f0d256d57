"""Gather locality experiment: does relabelling the nodes change the DRAM traffic / time of the
F = 256 mean gather on the C4 graph?  Variants: natural RMAT ids, degree-descending ids (hot rows contiguous),
random ids, and degree-descending + an L2 persisting window over the hottest rows.  Prints one JSON line."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import C4, rmat_edge_index
from keras_geometric_b200 import _lib, ops
from keras_geometric_b200.graph import GraphStructure
dev = torch.device("cuda", 0)
n, e, F = C4["nodes"], C4["edges"], 256
ei = rmat_edge_index(n, e, C4["rmat_scale"], 0, dev)
gen = torch.Generator(device=dev).manual_seed(3)
x = torch.randn((n, F), device=dev, generator=gen)


def timed(graph, xx, reps=5):
    for _ in range(2):
        ops.gather_reduce_raw(xx, graph.csr, _lib.OP_MEAN)
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.gather_reduce_raw(xx, graph.csr, _lib.OP_MEAN); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


res = {}
deg = torch.bincount(ei[1].long(), minlength=n) + torch.bincount(ei[0].long(), minlength=n)
res["natural"] = timed(GraphStructure(ei, n, n, 0), x)
for name, order in (("degree_descending", torch.argsort(deg, descending=True, stable=True)),
                    ("random", torch.randperm(n, device=dev, generator=gen))):
    new_id = torch.empty(n, dtype=torch.int32, device=dev)
    new_id[order] = torch.arange(n, dtype=torch.int32, device=dev)      # old id -> new id
    ei2 = new_id[ei.long()].contiguous()
    x2 = x[order].contiguous()                                            # row new_id holds the old node's features
    g2 = GraphStructure(ei2, n, n, 0)
    res[name] = timed(g2, x2)
    if name == "degree_descending":
        top = deg[order]
        csum = torch.cumsum(top.double(), 0) / float(top.sum())
        for rows in (40_000, 80_000, 120_000):
            res[f"edge_share_of_hottest_{rows}_rows"] = float(csum[rows - 1])
        # check: same result up to the permutation
        o1, _ = ops.gather_reduce_raw(x, GraphStructure(ei, n, n, 0).csr, _lib.OP_MEAN)
        o2, _ = ops.gather_reduce_raw(x2, g2.csr, _lib.OP_MEAN)
        res["max_abs_diff_vs_natural"] = float((o2 - o1[order]).abs().max())
print(json.dumps(res))

# ---- per-load L2 hints: hot rows (top-K by reference count) evict_last, the rest evict_first ----------------------
hint = {}
g = GraphStructure(ei, n, n, 0)
g.csc
for Fh in (256, 100, 48):
    xx = torch.randn((n, Fh), device=dev, generator=gen)
    for mb in ("0", "32", "64", "96"):
        os.environ["KGB200_HOT_MB"] = mb
        g.csr._hot.clear()
        hint[f"F{Fh}_hot{mb}MB"] = round(timed(g, xx), 3)
print(json.dumps({"l2_hints_mean_fwd_ms": hint}))
