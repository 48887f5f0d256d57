"""F = 256 mean gather on the C4 graph without and with the L2 eviction hints, for an ncu capture of the DRAM bytes:
launches 0-2 plain (the third arms the hints), launches 3-4 with the tagged column array.
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct \
        -k regex:gather_reduce_kernel --csv python tools/profile_hints.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import C4, rmat_edge_index
from keras_geometric_b200 import _lib, ops
from keras_geometric_b200.graph import GraphStructure
dev = torch.device("cuda", 0)
n, e = C4["nodes"], C4["edges"]
ei = rmat_edge_index(n, e, C4["rmat_scale"], 0, dev)
x = torch.randn((n, 256), device=dev, generator=torch.Generator(device=dev).manual_seed(3))
g = GraphStructure(ei, n, n, 0)
for i in range(5):
    ops.gather_reduce_raw(x, g.csr, _lib.OP_MEAN)
    torch.cuda.synchronize()
    print("launch", i, "hints", any(isinstance(k, int) for k in g.csr._hot))
