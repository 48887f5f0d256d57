"""Pieces of the benchmark that are not the headline step (imported by bench.py / bench_dist.py):

  kernel_suite   the north star's other kernels on the C4 graph - sum / max aggregation F = 100 / 256 forward +
                 backward, fused GATv2 (H = 8, C = 8) forward / backward, CSR + CSC build - each with algorithmic
                 bytes, CUDA-event time and fraction of the measured HBM peak;
  c5_strong      BASELINE configs[4]: 100 M nodes / 1 B edges RMAT, F = 64, SAGE-mean and GCN aggregation forward +
                 backward, strong-scaled over the ranks (halo exchange inside the timed region);
  parity_check   the partitioned layers on a small RMAT graph against the CPU oracle (checker only, rank 0's host).
"""
from __future__ import annotations

import json
import os
import statistics

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.abspath(__file__))
C5_N1_CACHE = "/tmp/kgb200_c5_n1.json"   # N = 1 figures of this box, read by the N > 1 runs that follow on it


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def _time(fn, reps=10, warm=3):
    """median CUDA-event time of fn() in ms on the current stream"""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


# ------------------------------------------------------------------------------------------ kernel suite
def kernel_suite(dev, ei, n, traffic=None):
    """Times the kernels the headline step does not exercise, on the same C4 graph.  Returns label -> record."""
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    peak, _ = hbm_peak()
    traffic = traffic or {}
    e = int(ei.shape[1])
    out = {}

    def rec(label, nbytes, ms, **extra):
        r = {"alg_bytes": int(nbytes), "ms": ms, "GBps": nbytes / ms / 1e6, "alg_frac": nbytes / ms / 1e6 / peak}
        tr = traffic.get(label)
        if tr:
            r["dram_bytes_ncu"] = tr
            r["dram_frac"] = tr / ms / 1e6 / peak
        r.update(extra)
        out[label] = r

    # K1: CSR (by target) and CSC (by source) build: 3 radix passes of ~20 B/edge + 12 B/edge of output
    passes = max(1, (max(n - 1, 1).bit_length() + 7) // 8)
    build_bytes = e * (20 * passes + 12)
    holder = {}

    def build_csr():
        holder["g"] = GraphStructure(ei, n, n, 0)

    ms = _time(build_csr, reps=10, warm=2)
    rec("csr_build", build_bytes, ms, note="kgb_csr_build + hub table + one host sync (status, hub counts)")
    graph = holder["g"]
    ms = _time(lambda: graph.__setattr__("_csc", None) or graph.csc, reps=10, warm=2)
    rec("csc_build", build_bytes, ms)
    gen = torch.Generator(device=dev).manual_seed(11)
    gtep = {}
    for F in (100, 256):
        x = torch.randn((n, F), device=dev, generator=gen).requires_grad_(True)
        R = torch.randn((n, F), device=dev, generator=gen)
        b_sum = e * (4 * F + 4) + n * 4 * F + (n + 1) * 8
        for op in ("sum", "max"):
            b_f = b_sum + (n * F * 4 if op == "max" else 0)
            b_b = n * F * 16 if op == "max" else b_sum
            t_f = _time(lambda: ops.gather_reduce(x, graph, op))
            o = ops.gather_reduce(x, graph, op)
            t_b = _time(lambda: torch.autograd.grad(o, x, R, retain_graph=True))
            rec(f"{op}_F{F}_fwd", b_f, t_f)
            rec(f"{op}_F{F}_bwd", b_b, t_b)
            rec(f"{op}_F{F}_fwd_bwd", b_f + b_b, t_f + t_b, GTEPS=e / (t_f + t_b) / 1e6)
            gtep[f"{op}_F{F}"] = e / (t_f + t_b) / 1e6
            del o
        del x, R
    # K6: fused GATv2, H = 8, C = 8 on the graph with self-loops (E' = E + N)
    H, C = 8, 8
    g_loops = GraphStructure(ei, n, n, n)
    g_loops.csc  # noqa: B018
    h = torch.randn((n, H * C), device=dev, generator=gen).requires_grad_(True)
    att = (torch.randn(H * C, device=dev, generator=gen) * 0.3).requires_grad_(True)
    R = torch.randn((n, H * C), device=dev, generator=gen)
    b_gat = ops.gat_bytes(g_loops.nnz, n, H, C)
    t_f = _time(lambda: ops.gatv2_aggregate(h, h, att, g_loops, H, C))
    o = ops.gatv2_aggregate(h, h, att, g_loops, H, C)
    ops.PROFILE = []
    t_b = _time(lambda: torch.autograd.grad(o, [h, att], R, retain_graph=True))
    prof, ops.PROFILE = ops.PROFILE, None
    rec("gatv2_H8_C8_fwd", b_gat, t_f)
    rec("gatv2_H8_C8_bwd", 2 * b_gat, t_b)
    for lab in ("gatv2_bwd_dst_H8_C8", "gatv2_bwd_src_H8_C8"):
        ts = [r["start"].elapsed_time(r["end"]) for r in prof if r["label"] == lab]
        if ts:
            rec(lab, b_gat, statistics.median(ts))
    rec("gatv2_H8_C8_fwd_bwd", 3 * b_gat, t_f + t_b, GTEPS=g_loops.nnz / (t_f + t_b) / 1e6)
    return out


# ------------------------------------------------------------------------------------------ C5 strong scaling
def c5_strong(world, rank, dev, div=1, reps=3, scramble=True):
    """BASELINE configs[4]: N = 100 M, E = 1 B directed RMAT edges (scale 27), x [N, 64] fp32.
    One aggregation layer forward + backward for SAGE-mean and for the GCN-normalised sum (self-loops appended),
    1-D node-partitioned over `world` ranks with the halo exchange inside the timed region.  GTEPS = E_agg / t."""
    from bench import rmat_edge_index
    from keras_geometric_b200 import ops
    from keras_geometric_b200.dist import PartitionedGraph, cost_balanced_bounds
    from keras_geometric_b200.graph import GraphStructure
    n, e, F = 100_000_000 // div, 1_000_000_000 // div // 2 * 2, 64
    scale = 27 - (div.bit_length() - 1)
    torch.cuda.empty_cache()
    ei = rmat_edge_index(n, e, scale, 0, dev)
    res = {"workload": f"C5: {n} nodes, {e} directed RMAT edges, F={F}, aggregation fwd+bwd incl. halo exchange",
           "n_gpus": world}
    gen = torch.Generator(device=dev).manual_seed(5 + rank)

    phases = {}

    def run(agg_fn, x, tag=""):
        def once():
            o = agg_fn(x)
            torch.autograd.grad(o, x, o.detach())   # seed = the output itself: no extra [N, F] buffer
        ops.PROFILE = []
        ms = torch.tensor([_time(once, reps=reps, warm=2)], device=dev)
        prof, ops.PROFILE = ops.PROFILE, None
        agg = {}
        for r in prof[-(len(prof) * reps // (reps + 2)):]:   # the timed repetitions only
            a = agg.setdefault(r["label"], [0.0, 0, 0])
            a[0] += r["start"].elapsed_time(r["end"]); a[1] += r["bytes"]; a[2] += 1
        phases[tag] = {k: {"ms_per_fwd_bwd": v[0] / reps, "GBps": (v[1] / v[0] / 1e6) if v[1] else None}
                       for k, v in agg.items()}
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    if world == 1:
        x = torch.randn((n, F), device=dev, generator=gen).requires_grad_(True)
        g = GraphStructure(ei, n, n, 0)
        g.csc  # noqa: B018
        t_sage = run(lambda t: ops.gather_reduce(t, g, "mean"), x, "sage_mean")
        del g
        torch.cuda.empty_cache()
        g = GraphStructure(ei, n, n, n)
        g.csc  # noqa: B018
        t_gcn = run(lambda t: ops.gather_reduce(t, g, "sum", weight="gcn"), x, "gcn")
        del g, x
        per_rank = None
    else:
        # aggregation only, measured at 2 GPUs: 0.095 ms per M edges, 0.48 ms per M owned nodes (output rows of both
        # passes, landing of the returned gradients) -> one node weighs 5 edges
        if scramble:   # hash partitioning (dist.scramble_ids): ranges balanced in nodes, edges and halo rows
            from keras_geometric_b200.dist import scramble_ids
            ei, _ = scramble_ids(ei, n)
            res["workload"] += ", node ids scrambled before the contiguous split (hash partitioning)"
        bounds = cost_balanced_bounds(ei[1], n, world, node_weight=5.0)
        lo, hi = bounds[rank], bounds[rank + 1]
        mine = (ei[1] >= lo) & (ei[1] < hi)
        src, dst = ei[0][mine].clone(), ei[1][mine].clone()
        del ei, mine
        torch.cuda.empty_cache()
        e_local = int(src.numel())
        pg = PartitionedGraph(src, dst, n, rank, world, bounds=bounds)
        x = torch.randn((pg.n_local, F), device=dev, generator=gen).requires_grad_(True)
        t_sage = run(lambda t: ops.aggregate_partitioned(t, pg, "mean"), x, "sage_mean")
        stats = torch.tensor([pg.n_local, pg.n_halo, e_local, pg.plan.n_send], device=dev, dtype=torch.float64)
        pg.close()
        del pg
        torch.cuda.empty_cache()
        pg = PartitionedGraph(src, dst, n, rank, world, bounds=bounds, n_loops_local=True)
        del src, dst
        t_gcn = run(lambda t: ops.aggregate_partitioned(t, pg, "gcn"), x, "gcn")
        pg.close()
        del pg, x
        allstats = [torch.zeros_like(stats) for _ in range(world)]
        dist.all_gather(allstats, stats)
        halo = [int(s[1]) for s in allstats]
        send = [int(s[3]) for s in allstats]
        link = max(max(h, s) for h, s in zip(halo, send)) * 4 * F * 2   # forward rows + backward gradient rows
        per_rank = {"n_local": [int(s[0]) for s in allstats], "n_halo": halo, "edges": [int(s[2]) for s in allstats],
                    "n_send": send, "halo_bytes_per_fwd_bwd_busiest_direction": link,
                    "nvlink_floor_ms_at_770GBps": link / 770e9 * 1e3}
    torch.cuda.empty_cache()
    res["sage_mean"] = {"ms_fwd_bwd": t_sage, "GTEPS": e / t_sage / 1e6}
    res["gcn"] = {"ms_fwd_bwd": t_gcn, "GTEPS": (e + n) / t_gcn / 1e6}
    res["per_rank"] = per_rank
    if world > 1:
        allph = [None] * world
        dist.all_gather_object(allph, phases)
        res["phases_per_rank"] = allph      # CUDA-event time of every launch group inside one fwd+bwd, per rank
    else:
        res["phases_per_rank"] = [phases]
    if rank == 0:
        if world == 1 and div == 1:
            try:
                json.dump({"sage_mean": res["sage_mean"]["GTEPS"], "gcn": res["gcn"]["GTEPS"]}, open(C5_N1_CACHE, "w"))
            except OSError:
                pass
        base = None
        if os.path.exists(C5_N1_CACHE):
            try:
                base = json.load(open(C5_N1_CACHE))
            except (OSError, ValueError):
                base = None
        if base and div == 1:
            res["speedup_vs_1gpu"] = {"sage_mean": res["sage_mean"]["GTEPS"] / base["sage_mean"],
                                      "gcn": res["gcn"]["GTEPS"] / base["gcn"],
                                      "n1_GTEPS": base, "n1_source": "the N=1 run on this box (" + C5_N1_CACHE + ")"}
        else:
            res["speedup_vs_1gpu"] = None
    return res


# ------------------------------------------------------------------------------------------ parity check
def _rel_err(got, want):
    """max |got - want| / (|want| + max|want|): <= 1e-5 is the north star's fp32 tolerance."""
    got, want = got.double(), want.double()
    scale = float(want.abs().max()) + 1e-300
    err = (got - want).abs() / (want.abs() + scale)
    return float(err.max()) if err.numel() else 0.0


def parity_check(world, rank, dev, n=50_000, e=600_000, fin=32, tol=1e-5):
    """Every rank runs the partitioned layers (forward + all gradients) on its slice of a small RMAT graph; rank 0
    gathers the slices and compares them with oracle/reference_path.py computed on its host."""
    import numpy as np

    from bench import rmat_edge_index
    from keras_geometric_b200 import GATv2Conv, GCNConv, GINConv, SAGEConv
    from keras_geometric_b200.dist import PartitionedGraph, cost_balanced_bounds
    cpu = torch.device("cpu")
    ei = rmat_edge_index(n, e, 16, 123, cpu)
    gen = torch.Generator().manual_seed(321)
    x = torch.randn((n, fin), generator=gen)
    bounds = cost_balanced_bounds(ei[1], n, world, node_weight=4.0)
    lo, hi = bounds[rank], bounds[rank + 1]
    mine = (ei[1] >= lo) & (ei[1] < hi)
    src, dst = ei[0][mine].to(dev), ei[1][mine].to(dev)
    pgs = {False: PartitionedGraph(src, dst, n, rank, world, bounds=bounds),
           True: PartitionedGraph(src, dst, n, rank, world, bounds=bounds, n_loops_local=True)}
    cases = [("sage_mean_wide", lambda: SAGEConv(64, aggregator="mean", activation=None), False),
             ("sage_mean_narrow", lambda: SAGEConv(12, aggregator="mean", activation=None), False),
             ("sage_mean_square_relu", lambda: SAGEConv(fin, aggregator="mean", activation="relu"), False),
             ("sage_sum_relu", lambda: SAGEConv(64, aggregator="sum", activation="relu"), False),
             ("sage_max", lambda: SAGEConv(64, aggregator="max", activation=None), False),
             ("gcn", lambda: GCNConv(16), True),
             ("gin_sum", lambda: GINConv(16, mlp_hidden=[], aggregator="sum"), False),
             ("gatv2", lambda: GATv2Conv(8, heads=4), True)]
    report, worst = {}, 0.0
    for name, make, loops in cases:
        torch.manual_seed(7)
        layer = make()
        xl = x[lo:hi].to(dev).requires_grad_(True)
        out = layer([xl, pgs[loops]])
        width = int(out.shape[1])
        R = torch.randn((n, width), generator=torch.Generator().manual_seed(99))
        if name.startswith("sage"):
            weights = [layer.lin_neigh.kernel, layer.lin_self.kernel, layer.bias]
        elif name == "gcn":
            weights = [layer.kernel, layer.bias]
        elif name == "gin_sum":
            weights = [layer.mlp.layers[0].kernel, layer.mlp.layers[0].bias]
        else:
            weights = [layer.linear_transform.kernel, layer.att, layer.bias]
        grads = torch.autograd.grad((out * R[lo:hi].to(dev)).sum(), [xl] + weights)
        gws = [g.clone() for g in grads[1:]]
        for g in gws:
            dist.all_reduce(g)
        h_gpu = None
        if name == "gatv2":   # the GPU's h = x W, to pin the oracle's LeakyReLU sign pattern (oracle/kink.py)
            from keras_geometric_b200 import ops
            with torch.no_grad():
                h_gpu = ops.linear(xl, layer.linear_transform.kernel).cpu().numpy()
        parts = [None] * world
        dist.all_gather_object(parts, (out.detach().cpu().numpy(), grads[0].cpu().numpy(), h_gpu))
        if rank != 0:
            continue
        from oracle import reference_path as ref
        out_full = torch.from_numpy(np.concatenate([p[0] for p in parts]))
        gx_full = torch.from_numpy(np.concatenate([p[1] for p in parts]))
        xo = x.clone().requires_grad_(True)
        wc = [w.detach().cpu().clone().requires_grad_(True) for w in weights]
        if name.startswith("sage"):
            wn, ws, b = wc
            agg = name.split("_")[1]
            if name.endswith("relu"):
                # ReLU kinks: the derivative mask is taken from the GPU's own output, so both sides differentiate the
                # same piecewise-linear function; the forward comparison covers the mask itself
                pre = ref.sage_conv(xo, ei, wn, ws, b, agg, None)
                fwd_err = _rel_err(out_full, torch.relu(pre).detach())
                want = pre * (out_full > 0).to(pre.dtype)
            else:
                want = ref.sage_conv(xo, ei, wn, ws, b, agg, None)
                fwd_err = _rel_err(out_full, want.detach())
        elif name == "gcn":
            want = ref.gcn_conv(xo, ei, wc[0], wc[1])
            fwd_err = _rel_err(out_full, want.detach())
        elif name == "gin_sum":
            want = ref.gin_conv(xo, ei, lambda t: t @ wc[0] + wc[1], 0.0, "sum")
            fwd_err = _rel_err(out_full, want.detach())
        else:
            from oracle.kink import pinned_matmul
            fwd_err = _rel_err(out_full, ref.gatv2_conv(xo, ei, wc[0], wc[1], wc[2], heads=4).detach())
            with pinned_matmul([torch.from_numpy(np.concatenate([p[2] for p in parts]))]):
                want = ref.gatv2_conv(xo, ei, wc[0], wc[1], wc[2], heads=4)
        gwant = torch.autograd.grad((want * R).sum(), [xo] + wc)
        errs = {"out": fwd_err, "grad_x": _rel_err(gx_full, gwant[0])}
        for i, (a, b_) in enumerate(zip(gws, gwant[1:])):
            errs[f"grad_w{i}"] = _rel_err(a.cpu().reshape(b_.shape), b_)
        report[name] = max(errs.values())
        worst = max(worst, report[name])
    for pg in pgs.values():
        pg.close()
    ok = torch.tensor([1 if worst <= tol else 0], device=dev)
    dist.broadcast(ok, src=0)
    return {"vs": "oracle", "ok": bool(int(ok)), "max_rel_err": worst, "tol": tol,
            "metric": "max |got - want| / (|want| + max|want|) over outputs, input gradients and weight gradients",
            "graph": {"nodes": n, "edges": e, "feats": fin, "rmat_scale": 16},
            "cases": report,
            "kinks": "derivative discontinuities are evaluated on identical sign patterns: the ReLU mask is taken from the "
                     "GPU output, GATv2's h = xW is pinned to the GPU's values in the oracle (oracle/kink.py); no "
                     "tolerance is loosened"}
