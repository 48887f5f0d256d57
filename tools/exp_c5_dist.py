"""torchrun --nproc-per-node N tools/exp_c5_dist.py [div] : C5 (100 M nodes / 1 B edges RMAT, F = 64) STRONG-scaled over N
GPUs - two SAGEConv(64, mean) layers, forward + backward, on the 1-D node partition (keras_geometric_b200.dist).
Prints one JSON line: GTEPS = layers * E / t (max over ranks, CUDA events) next to the per-rank partition statistics.
SURVEY 8(d) target: >= 6x the 1-GPU figure of tools/exp_c5.py at 8 GPUs.  (Written at the end of round 1 after the GPU
budget was spent: not yet run.)"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench_dist
from bench import rmat_edge_index
from keras_geometric_b200 import SAGEConv
import keras_geometric_b200.dist as kd

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
div = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n, e, F = 100_000_000 // div, 1_000_000_000 // div // 2 * 2, 64
scale = 27 - (div.bit_length() - 1)
ei = rmat_edge_index(n, e, scale, 0, dev)            # every rank generates the same global list (8 GB at full size)
bounds = bench_dist.edge_balanced_bounds(ei[1], n, world)
lo, hi = bounds[rank], bounds[rank + 1]
mine = (ei[1] >= lo) & (ei[1] < hi)
src, dst = ei[0][mine].clone(), ei[1][mine].clone()
del ei, mine
torch.cuda.empty_cache()
kd.partition_bounds = lambda n_, w_, _b=bounds: _b
pg = kd.PartitionedGraph(src, dst, n, rank, world)
e_local = int(src.numel())
del src, dst
gen = torch.Generator(device=dev).manual_seed(1 + rank)
x = torch.randn((pg.n_local, F), device=dev, generator=gen).requires_grad_(True)
R = torch.randn((pg.n_local, F), device=dev, generator=gen)
torch.manual_seed(0)
layers = [SAGEConv(F, aggregator="mean"), SAGEConv(F, aggregator="mean")]


def step():
    h = x
    for lyr in layers:
        h = lyr([h, pg])
    params = [p for lyr in layers for p in lyr.trainable_weights]
    return torch.autograd.grad((h * R).sum(), [x] + params)


for _ in range(3):
    step()
torch.cuda.synchronize()
dist.barrier()
steps = 5
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(steps):
    step()
t1.record()
torch.cuda.synchronize()
dist.barrier()
ms = torch.tensor([t0.elapsed_time(t1) / steps], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
stats = torch.tensor([pg.n_local, pg.n_halo, e_local, pg.plan.n_send], device=dev, dtype=torch.float64)
allstats = [torch.zeros_like(stats) for _ in range(world)]
dist.all_gather(allstats, stats)
if rank == 0:
    print(json.dumps({"config": f"C5 strong-scaled over {world} GPUs", "nodes": n, "edges": e, "F": F, "layers": 2,
                      "ms_per_fwd_bwd": float(ms), "GTEPS": 2 * e / (float(ms) * 1e-3) / 1e9,
                      "per_rank": {"n_local": [int(s[0]) for s in allstats], "n_halo": [int(s[1]) for s in allstats],
                                   "edges": [int(s[2]) for s in allstats], "n_send": [int(s[3]) for s in allstats]},
                      "mem_GiB": torch.cuda.max_memory_allocated() / 2 ** 30}))
dist.destroy_process_group()
