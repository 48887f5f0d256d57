"""torchrun --nproc-per-node N tools/dist_check.py : partitioned SAGE / GCN / GIN / GATv2 layers (outputs, input and
weight gradients) vs the single-GPU layers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from keras_geometric_b200 import SAGEConv, GCNConv, GINConv, GATv2Conv
from keras_geometric_b200.dist import PartitionedGraph, partition_bounds

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
rng = np.random.default_rng(0)
n, e, fin = 5000, 80000, 40
dst = np.minimum((rng.pareto(1.3, e) * 40).astype(np.int64), n - 1)
ei = np.stack([rng.integers(0, n, e), dst]).astype(np.int32)
x = rng.standard_normal((n, fin)).astype(np.float32)
R = {12: rng.standard_normal((n, 12)).astype(np.float32), 64: rng.standard_normal((n, 64)).astype(np.float32)}
b = partition_bounds(n, world); lo, hi = b[rank], b[rank + 1]
mine = (ei[1] >= lo) & (ei[1] < hi)
eid = torch.from_numpy(ei).cuda()
ok = True
for name, mk, loops in [("sage_mean_reorder", lambda: SAGEConv(12, aggregator="mean"), False),
                        ("sage_mean_wide", lambda: SAGEConv(64, aggregator="mean"), False),
                        ("sage_sum_wide", lambda: SAGEConv(64, aggregator="sum", activation="relu"), False),
                        ("sage_sum_reorder", lambda: SAGEConv(12, aggregator="sum", activation="relu"), False),
                        ("sage_max", lambda: SAGEConv(64, aggregator="max"), False),
                        ("gcn", lambda: GCNConv(12), True),
                        ("gin_sum", lambda: GINConv(12, mlp_hidden=[16], aggregator="sum"), False),
                        ("gin_mean_eps", lambda: GINConv(12, mlp_hidden=[16], aggregator="mean", train_eps=True, eps_init=0.3), False),
                        ("gin_max", lambda: GINConv(64, aggregator="max"), False),
                        ("gatv2_concat", lambda: GATv2Conv(16, heads=4), True),
                        ("gatv2_head_mean", lambda: GATv2Conv(12, heads=2, concat=False), True)]:
    torch.manual_seed(7)
    layer = mk()
    xf = torch.from_numpy(x).cuda().requires_grad_(True)
    out_full = layer([xf, eid])
    Rf = torch.from_numpy(R[out_full.shape[1]]).cuda()
    gfull = torch.autograd.grad((out_full * Rf).sum(), [xf] + layer.trainable_weights)
    pg = PartitionedGraph(torch.from_numpy(ei[0][mine]).cuda(), torch.from_numpy(ei[1][mine]).cuda(), n, rank, world,
                          n_loops_local=loops)
    xl = torch.from_numpy(x[lo:hi]).cuda().requires_grad_(True)
    out = layer([xl, pg])
    g = torch.autograd.grad((out * Rf[lo:hi]).sum(), [xl] + layer.trainable_weights)
    gw = [t.clone() for t in g[1:]]
    for t in gw: dist.all_reduce(t)
    def chk(a, b_, what):
        global ok
        scale = float(b_.abs().max()) + 1e-30
        err = float((a - b_).abs().max()) / scale
        if err > 2e-5:
            ok = False; print(f"[rank {rank}] {name} {what}: rel err {err:.3e}")
    chk(out, out_full[lo:hi], "out"); chk(g[0], gfull[0][lo:hi], "grad_x")
    for i, (a, b_) in enumerate(zip(gw, gfull[1:])): chk(a, b_, f"grad_w{i}")
    if rank == 0: print(name, "checked; halo rows", pg.n_halo, "send", pg.plan.n_send)
flag = torch.tensor([1 if ok else 0], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0: print("DIST_CHECK", "PASS" if int(flag) else "FAIL")
dist.destroy_process_group()
sys.exit(0 if int(flag) else 1)
