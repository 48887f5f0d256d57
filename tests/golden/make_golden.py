"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference.

Run once in the build container (never on the GPU box - /root/reference is not there):

    python tests/golden/make_golden.py

The reference's sources are imported from /root/reference/src; its only missing dependency,
Keras, is satisfied by the stand-in package oracle/keras_shim (see oracle/__init__.py), so
the *layer logic* that produces these vectors is the reference's own code, executed with
torch on the CPU.  Each fixture stores inputs, weights, the forward output and the
gradients of ``sum(out * R)`` (R seeded) w.r.t. x and every weight.
"""

from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "keras_shim"))
sys.path.insert(0, "/root/reference/src")

import keras_geometric as kg  # noqa: E402  (the reference)
from keras_geometric.layers.aggregators import AggregatorFactory  # noqa: E402


def make_graph(rng, n, e, n_isolated=3, self_loops=2, dup=4):
    """Directed multigraph with duplicates, self-loops, and nodes without in-edges."""
    tgt_pool = np.arange(n - n_isolated)
    src = rng.integers(0, n, size=e)
    dst = rng.choice(tgt_pool, size=e)
    for k in range(self_loops):
        dst[k] = src[k] = tgt_pool[k]
    for k in range(dup):
        src[e - 1 - k] = src[k + self_loops]
        dst[e - 1 - k] = dst[k + self_loops]
    return np.stack([src, dst]).astype(np.int32)


def save(name, params, arrays):
    arrays = {k: np.asarray(v) for k, v in arrays.items() if v is not None}
    arrays["params"] = np.array(json.dumps(params))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
    print(f"wrote {name}.npz  ({', '.join(sorted(arrays))})")


def set_weights(named, rng, scale=0.5):
    out = {}
    for nm, p in named:
        w = (rng.standard_normal(tuple(p.shape)) * scale).astype(np.float32)
        with torch.no_grad():
            p.copy_(torch.from_numpy(w))
        out["w_" + nm] = w
    return out


def run_with_grads(layer, x_np, ei_np, named, rng, call=None):
    x = torch.from_numpy(x_np).clone().requires_grad_(True)
    ei = torch.from_numpy(ei_np)
    out = call(x, ei) if call is not None else layer([x, ei])
    r = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    loss = (out * torch.from_numpy(r)).sum()
    params = [p for _, p in named]
    grads = torch.autograd.grad(loss, [x] + params, allow_unused=True)
    res = {"out": out.detach().numpy(), "R": r,
           "grad_x": grads[0].numpy() if grads[0] is not None else np.zeros_like(x_np)}
    for (nm, p), g in zip(named, grads[1:]):
        res["grad_" + nm] = g.numpy() if g is not None else np.zeros(tuple(p.shape), np.float32)
    return res


def main():
    rng = np.random.default_rng(20261018)

    # ---- 1. aggregators called directly (layers/aggregators.py) -------------------------
    e, f, n = 160, 12, 23
    msgs = rng.standard_normal((e, f)).astype(np.float32)
    msgs = np.round(msgs, 1)  # force ties for max/min
    msgs[5, 3] = np.inf
    msgs[9, 4] = -np.inf
    tgt = rng.integers(0, n - 4, size=e).astype(np.int32)  # last 4 segments empty
    arrays = {"messages": msgs, "target_idx": tgt, "dim_size": n}
    for name in ["mean", "max", "sum", "min", "std"]:
        agg = AggregatorFactory.create(name)
        arrays["out_" + name] = agg.aggregate(torch.from_numpy(msgs), torch.from_numpy(tgt), n).numpy()
    # gradients on a finite copy (inf rows make autograd produce nan in the reference too)
    fin = np.where(np.isinf(msgs), 0.0, msgs).astype(np.float32)
    arrays["messages_finite"] = fin
    r = rng.standard_normal((n, f)).astype(np.float32)
    arrays["R"] = r
    for name in ["mean", "max", "sum", "min", "std"]:
        m = torch.from_numpy(fin).clone().requires_grad_(True)
        out = AggregatorFactory.create(name).aggregate(m, torch.from_numpy(tgt), n)
        (g,) = torch.autograd.grad((out * torch.from_numpy(r)).sum(), [m])
        arrays["outfin_" + name] = out.detach().numpy()
        arrays["grad_" + name] = g.numpy()
    save("aggregators", {"E": e, "F": f, "N": n}, arrays)

    # ---- 2. utils/main.py -----------------------------------------------------------------
    n, e = 41, 300
    ei = make_graph(rng, n, e)
    with_loops = kg.add_self_loops(torch.from_numpy(ei), n)
    w = kg.compute_gcn_normalization(with_loops, n)
    w_noloop = kg.compute_gcn_normalization(torch.from_numpy(ei), n)
    save("utils", {"N": n, "E": e}, {"edge_index": ei, "with_loops": with_loops.numpy(),
                                     "gcn_norm": w.numpy(), "gcn_norm_noloop": w_noloop.numpy()})

    # ---- 3. MessagePassing.propagate (layers/message_passing.py) ----------------------------
    n, e, f = 37, 180, 10
    ei = make_graph(rng, n, e)
    x = np.round(rng.standard_normal((n, f)), 1).astype(np.float32)
    arrays = {"x": x, "edge_index": ei}
    for name in ["mean", "max", "sum", "min", "std"]:
        layer = kg.MessagePassing(aggregator=name)
        res = run_with_grads(layer, x, ei, [], rng, call=lambda a, b, L=layer: L.propagate(x=a, edge_index=b))
        for k, v in res.items():
            arrays[f"{k}_{name}"] = v
    # bipartite: 9 targets, 37 sources
    nt = 9
    ei_b = np.stack([rng.integers(0, n, 60), rng.integers(0, nt, 60)]).astype(np.int32)
    xt = rng.standard_normal((nt, f)).astype(np.float32)
    out_b = kg.MessagePassing(aggregator="sum").propagate(
        x=(torch.from_numpy(xt), torch.from_numpy(x)), edge_index=torch.from_numpy(ei_b))
    arrays.update({"bip_edge_index": ei_b, "bip_x_target": xt, "bip_out_sum": out_b.numpy()})
    save("message_passing", {"N": n, "E": e, "F": f}, arrays)

    # ---- 4. GCNConv --------------------------------------------------------------------------
    n, e, fin_, fout = 33, 140, 9, 6
    ei = make_graph(rng, n, e)
    x = rng.standard_normal((n, fin_)).astype(np.float32)
    for tag, kw, layout in [("default", {}, "2E"), ("nonorm", {"normalize": False}, "2E"),
                            ("noloops_nobias", {"add_self_loops": False, "use_bias": False}, "2E"),
                            ("E2layout", {}, "E2")]:
        layer = kg.GCNConv(fout, **kw)
        ei_in = ei if layout == "2E" else np.ascontiguousarray(ei.T)
        layer([torch.from_numpy(x), torch.from_numpy(ei_in)])
        named = [("kernel", layer.kernel)] + ([("bias", layer.bias)] if layer.bias is not None else [])
        arrays = {"x": x, "edge_index": ei_in}
        arrays.update(set_weights(named, rng))
        arrays.update(run_with_grads(layer, x, ei_in, named, rng))
        save("gcn_" + tag, {"output_dim": fout, **kw}, arrays)

    # ---- 5. SAGEConv -------------------------------------------------------------------------
    n, e, fin_, fout = 35, 170, 8, 5
    ei = make_graph(rng, n, e)
    x = np.round(rng.standard_normal((n, fin_)), 1).astype(np.float32)
    for tag, kw in [("mean", {"aggregator": "mean"}), ("max", {"aggregator": "max"}),
                    ("sum", {"aggregator": "sum"}), ("min", {"aggregator": "min"}),
                    ("std", {"aggregator": "std"}),
                    ("pooling", {"aggregator": "pooling", "pool_hidden_dim": 7}),
                    ("mean_noroot_norm", {"aggregator": "mean", "root_weight": False,
                                          "normalize": True, "activation": None}),
                    ("mean_linear", {"aggregator": "mean", "activation": None})]:
        layer = kg.SAGEConv(fout, **kw)
        layer([torch.from_numpy(x), torch.from_numpy(ei)])
        named = [("lin_neigh", layer.lin_neigh.kernel)]
        if layer.lin_self is not None:
            named.append(("lin_self", layer.lin_self.kernel))
        if layer.pool_mlp is not None:
            named += [("pool_kernel", layer.pool_mlp.kernel), ("pool_bias", layer.pool_mlp.bias)]
        if layer.bias is not None:
            named.append(("bias", layer.bias))
        arrays = {"x": x, "edge_index": ei}
        arrays.update(set_weights(named, rng))
        arrays.update(run_with_grads(layer, x, ei, named, rng))
        save("sage_" + tag, {"output_dim": fout, **kw}, arrays)

    # ---- 6. GINConv --------------------------------------------------------------------------
    n, e, fin_, fout = 31, 120, 7, 6
    ei = make_graph(rng, n, e)
    x = np.round(rng.standard_normal((n, fin_)), 1).astype(np.float32)
    for tag, kw in [("sum", {"aggregator": "sum", "mlp_hidden": [11]}),
                    ("mean_eps", {"aggregator": "mean", "mlp_hidden": [11, 9], "train_eps": True,
                                  "eps_init": 0.25}),
                    ("max", {"aggregator": "max", "mlp_hidden": []})]:
        layer = kg.GINConv(fout, **kw)
        layer([torch.from_numpy(x), torch.from_numpy(ei)])
        named = []
        for i, d in enumerate([l for l in layer.mlp.layers if hasattr(l, "kernel")]):
            named += [(f"mlp{i}_kernel", d.kernel), (f"mlp{i}_bias", d.bias)]
        arrays = {"x": x, "edge_index": ei}
        arrays.update(set_weights(named, rng))
        if kw.get("train_eps"):
            named.append(("eps", layer.eps))
            arrays["w_eps"] = layer.eps.detach().numpy().copy()
        arrays.update(run_with_grads(layer, x, ei, named, rng))
        save("gin_" + tag, {"output_dim": fout, **kw}, arrays)

    # ---- 7. GATv2Conv ------------------------------------------------------------------------
    n, e, fin_ = 29, 130, 10
    ei = make_graph(rng, n, e)
    x = rng.standard_normal((n, fin_)).astype(np.float32)
    for tag, kw in [("h4c16", {"output_dim": 16, "heads": 4}), ("h8c8", {"output_dim": 8, "heads": 8}),
                    ("h1c3", {"output_dim": 3, "heads": 1}),
                    ("h2c5_mean", {"output_dim": 5, "heads": 2, "concat": False}),
                    ("h3c4_noloops", {"output_dim": 4, "heads": 3, "add_self_loops": False,
                                      "negative_slope": 0.1, "use_bias": False})]:
        layer = kg.GATv2Conv(**kw)
        layer([torch.from_numpy(x), torch.from_numpy(ei)])
        named = [("linear_transform", layer.linear_transform.kernel), ("att", layer.att)]
        if layer.bias is not None:
            named.append(("bias", layer.bias))
        arrays = {"x": x, "edge_index": ei}
        arrays.update(set_weights(named, rng))
        arrays.update(run_with_grads(layer, x, ei, named, rng))
        save("gatv2_" + tag, kw, arrays)


def extra():
    """next-1 / next-2 rows of SURVEY 8(f): pooling read-outs and batch_graphs."""
    from keras_geometric.layers.pooling import BatchGlobalPooling, GlobalPooling
    rng = np.random.default_rng(77)
    sizes = [5, 1, 9, 4, 12]
    n, f = sum(sizes), 6
    x = rng.standard_normal((n, f)).astype(np.float32)
    batch = np.repeat(np.arange(len(sizes)), sizes).astype(np.int32)
    arrays = {"x": x, "batch": batch}
    for pool in ["mean", "max", "sum"]:
        arrays["global_" + pool] = GlobalPooling(pooling=pool)(torch.from_numpy(x)).numpy()
        xt = torch.from_numpy(x).clone().requires_grad_(True)
        out = BatchGlobalPooling(pooling=pool)([xt, torch.from_numpy(batch)])
        r = rng.standard_normal(tuple(out.shape)).astype(np.float32)
        (g,) = torch.autograd.grad((out * torch.from_numpy(r)).sum(), [xt])
        arrays.update({"batch_" + pool: out.detach().numpy(), "R_" + pool: r, "grad_" + pool: g.numpy()})
    save("pooling", {"sizes": sizes}, arrays)

    graphs, arrays = [], {}
    for i, (nn, ee) in enumerate([(4, 6), (3, 0), (6, 10)]):
        gx = rng.standard_normal((nn, 3)).astype(np.float32)
        gei = np.stack([rng.integers(0, nn, ee), rng.integers(0, nn, ee)]).astype(np.int32)
        gy = rng.standard_normal((nn, 2)).astype(np.float32)
        ga = rng.standard_normal((ee, 2)).astype(np.float32)
        arrays.update({f"x{i}": gx, f"ei{i}": gei, f"y{i}": gy, f"ea{i}": ga})
        graphs.append(kg.GraphData(x=gx, edge_index=gei, edge_attr=ga, y=gy))
    b = kg.batch_graphs(graphs)
    arrays.update({"bx": b.x.numpy(), "bei": b.edge_index.numpy(), "by": b.y.numpy(), "bea": b.edge_attr.numpy(),
                   "bbatch": b.batch.numpy(), "bnum_nodes": b.num_nodes})
    save("batch_graphs", {"n_graphs": 3}, arrays)


if __name__ == "__main__":
    torch.manual_seed(0)
    main()
    extra()
