"""Per-kernel timings of the GATv2 and max kernels on the C4 graph for the library named by KGB200_LIB (tuning
variants from tools/build_variant.sh): prints one JSON line  {label: ms}."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import C4, rmat_edge_index
from keras_geometric_b200 import ops
from keras_geometric_b200.graph import GraphStructure
dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "gat,max"
n, e = C4["nodes"], C4["edges"]
ei = rmat_edge_index(n, e, C4["rmat_scale"], 0, dev)
gen = torch.Generator(device=dev).manual_seed(11)
res = {"lib": os.environ.get("KGB200_LIB", "default")}


def collect(fn, reps=4):
    for _ in range(2):
        fn()
    ops.PROFILE = []
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    agg = {}
    for r in prof:
        agg.setdefault(r["label"], []).append(r["start"].elapsed_time(r["end"]))
    return {k: round(statistics.median(v), 3) for k, v in agg.items()}


if "gat" in which:
    g = GraphStructure(ei, n, n, n)
    g.csc
    for H, C in [(8, 8), (1, 64), (8, 32)]:
        h = torch.randn((n, H * C), device=dev, generator=gen).requires_grad_(True)
        att = (torch.randn(H * C, device=dev, generator=gen) * 0.3).requires_grad_(True)
        R = torch.randn((n, H * C), device=dev, generator=gen)

        def fb():
            o = ops.gatv2_aggregate(h, h, att, g, H, C)
            torch.autograd.grad(o, [h, att], R)
        res.update(collect(fb))
        del h, R
    del g
if "max" in which:
    g = GraphStructure(ei, n, n, 0)
    for F in (100, 256):
        x = torch.randn((n, F), device=dev, generator=gen).requires_grad_(True)
        R = torch.randn((n, F), device=dev, generator=gen)

        def fb():
            o = ops.gather_reduce(x, g, "max")
            torch.autograd.grad(o, x, R)
        res.update(collect(fb))
        del x, R
if "mean" in which:
    # mean gather forward + transposed backward at the C4 step's widths; run with KGB200_HOT_MB=0/auto to measure the
    # L2 eviction hints
    res["hot_mb"] = os.environ.get("KGB200_HOT_MB", "auto")
    g = GraphStructure(ei, n, n, 0)
    g.csc
    for F in (48, 100, 256):
        x = torch.randn((n, F), device=dev, generator=gen).requires_grad_(True)
        R = torch.randn((n, F), device=dev, generator=gen)

        def fb():
            o = ops.gather_reduce(x, g, "mean")
            torch.autograd.grad(o, x, R)
        fb()   # third use of the structure at this width arms the L2 hints (auto mode) before anything is timed
        res.update({k + ("_bwd" if k.endswith("_w") else "_fwd"): v for k, v in collect(fb, reps=5).items()})
        del x, R
print(json.dumps(res))
