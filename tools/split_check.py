"""Single-rank sanity of PartitionedGraph.split: local part + (empty) halo part reproduce the plain mean aggregation."""
import torch, numpy as np, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from keras_geometric_b200.dist import PartitionedGraph
from keras_geometric_b200 import ops
from keras_geometric_b200.graph import GraphStructure
rng = np.random.default_rng(0)
n, e = 3000, 40000
ei = torch.from_numpy(np.stack([rng.integers(0, n, e), rng.integers(0, n, e)]).astype(np.int32)).cuda()
pg = PartitionedGraph(ei[0], ei[1], n, 0, 1)
g_l, g_h, inv = pg.split
x = torch.randn((n, 32), device="cuda")
full = ops.gather_reduce(x, GraphStructure(ei, n, n, 0), "mean")
part = ops.gather_reduce(x, g_l, "sum", weight=(None, inv))
print("split ok", float((part - full).abs().max()), g_l.nnz, g_h.nnz, pg.n_halo)
