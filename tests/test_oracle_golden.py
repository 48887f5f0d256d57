"""Pin the CPU oracle (oracle/reference_path.py) before anything is compared with it.

(1) the reference's own hand-computed KATs, re-expressed without Keras
    (/root/reference/tests/test_message_passing.py:54-155, :168-179, :342-360;
     /root/reference/tests/test_graphsage_conv.py:431-537);
(2) golden vectors produced by the UNMODIFIED reference layer code (tests/golden/make_golden.py).
Tolerances: bit-exact for integer/index outputs, rtol 1e-5 / atol 1e-6 for float32.
"""
import numpy as np
import pytest
import torch

from oracle import keras_ops as kops
from oracle import reference_path as ref

from conftest import load_golden

RTOL, ATOL = 1e-5, 1e-6


def t(a):
    return torch.from_numpy(np.asarray(a))


# ------------------------------------------------------------------ reference KATs
def test_kat_mean():  # tests/test_message_passing.py:54-80
    out = ref.aggregate("mean", t(np.array([[1, 2], [3, 4], [5, 6]], np.float32)), t(np.array([0, 0, 1], np.int32)), 3)
    np.testing.assert_allclose(out.numpy(), [[2, 3], [5, 6], [0, 0]], rtol=1e-5)


def test_kat_max():  # :82-101
    out = ref.aggregate("max", t(np.array([[1, 5], [3, 2], [2, 4]], np.float32)), t(np.array([0, 0, 1], np.int32)), 3)
    np.testing.assert_allclose(out.numpy()[:2], [[3, 5], [2, 4]], rtol=1e-5)
    np.testing.assert_array_equal(out.numpy()[2], [0, 0])


def test_kat_sum_min_std():  # :103-155
    m = t(np.array([[1, 2], [3, 4], [5, 6]], np.float32))
    idx = t(np.array([0, 0, 1], np.int32))
    np.testing.assert_allclose(ref.aggregate("sum", m, idx, 3).numpy()[0], [4, 6], rtol=1e-5)
    m2 = t(np.array([[1, 5], [3, 2], [2, 4]], np.float32))
    np.testing.assert_allclose(ref.aggregate("min", m2, idx, 3).numpy()[0], [1, 2], rtol=1e-5)
    m3 = t(np.array([[1, 2], [3, 4], [5, 6], [7, 8]], np.float32))
    out = ref.aggregate("std", m3, t(np.array([0, 0, 1, 1], np.int32)), 2)
    np.testing.assert_allclose(out.numpy(), [[1, 1], [1, 1]], rtol=1e-5)


def test_kat_empty_and_no_edges():  # :157-179
    assert ref.propagate(t(np.zeros((0, 8), np.float32)), t(np.zeros((2, 0), np.int32)), "mean").shape == (0, 8)
    x = np.random.default_rng(0).standard_normal((5, 8)).astype(np.float32)
    out = ref.propagate(t(x), t(np.zeros((2, 0), np.int32)), "mean")
    np.testing.assert_array_equal(out.numpy(), np.zeros((5, 8), np.float32))


def test_kat_extreme_values():  # :342-360
    for v in (1e10, 1e-10):
        out = ref.aggregate("mean", t(np.full((100, 10), v, np.float32)), t(np.zeros(100, np.int32)), 1).numpy()
        assert np.isfinite(out).all()


def test_kat_sage_mean_numpy():  # tests/test_graphsage_conv.py:431-537 (graph :114-120, seed 45)
    np.random.seed(45)
    x = np.random.randn(7, 16).astype(np.float32)
    ei = np.array([[0, 1, 1, 2, 3, 4, 4, 5, 0, 3, 6, 5, 1, 6], [1, 0, 2, 1, 4, 3, 5, 4, 2, 5, 5, 6, 6, 0]], np.int64)
    rng = np.random.default_rng(1)
    wn, ws = rng.standard_normal((16, 8)).astype(np.float32), rng.standard_normal((16, 8)).astype(np.float32)
    b = rng.standard_normal(8).astype(np.float32)
    out = ref.sage_conv(t(x), t(ei), t(wn), t(ws), t(b), "mean", None, False).numpy()
    agg = np.zeros((7, 16), np.float32)
    for i in range(7):
        nb = ei[0][ei[1] == i]
        if len(nb):
            agg[i] = x[nb].mean(0)
    np.testing.assert_allclose(out, x @ ws + agg @ wn + b, rtol=1e-5, atol=1e-5)


def test_invalid_aggregator_message():  # aggregators.py:312-317
    with pytest.raises(ValueError, match="Invalid aggregator"):
        ref.aggregate("median", t(np.ones((1, 1), np.float32)), t(np.zeros(1, np.int32)), 1)


# ------------------------------------------------------------------ golden vectors
def test_golden_aggregators():
    g = load_golden("aggregators")
    n = int(g["dim_size"])
    for name in ["mean", "max", "sum", "min", "std"]:
        out = ref.aggregate(name, t(g["messages"]), t(g["target_idx"]), n).numpy()
        np.testing.assert_allclose(out, g["out_" + name], rtol=RTOL, atol=ATOL, equal_nan=True)
        m = t(g["messages_finite"]).clone().requires_grad_(True)
        o = ref.aggregate(name, m, t(g["target_idx"]), n)
        (gr,) = torch.autograd.grad((o * t(g["R"])).sum(), [m])
        np.testing.assert_allclose(gr.numpy(), g["grad_" + name], rtol=RTOL, atol=ATOL)
    np.testing.assert_array_equal(ref.aggregate("max", t(g["messages"]), t(g["target_idx"]), n).numpy(), g["out_max"])


def test_golden_utils():
    g = load_golden("utils")
    n = g["params"]["N"]
    wl = ref.add_self_loops(t(g["edge_index"]), n)
    np.testing.assert_array_equal(wl.numpy(), g["with_loops"])
    np.testing.assert_array_equal(ref.compute_gcn_normalization(wl, n).numpy(), g["gcn_norm"])
    np.testing.assert_array_equal(ref.compute_gcn_normalization(t(g["edge_index"]), n).numpy(), g["gcn_norm_noloop"])


def test_golden_message_passing():
    g = load_golden("message_passing")
    for name in ["mean", "max", "sum", "min", "std"]:
        x = t(g["x"]).clone().requires_grad_(True)
        out = ref.propagate(x, t(g["edge_index"]), name)
        np.testing.assert_allclose(out.detach().numpy(), g["out_" + name], rtol=RTOL, atol=ATOL)
        (gx,) = torch.autograd.grad((out * t(g["R_" + name])).sum(), [x])
        np.testing.assert_allclose(gx.numpy(), g["grad_x_" + name], rtol=RTOL, atol=ATOL)
    out = ref.propagate((t(g["bip_x_target"]), t(g["x"])), t(g["bip_edge_index"]), "sum")
    np.testing.assert_allclose(out.numpy(), g["bip_out_sum"], rtol=RTOL, atol=ATOL)


def _check(out, params, g, names):
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=RTOL, atol=ATOL)
    grads = torch.autograd.grad((out * t(g["R"])).sum(), params, allow_unused=True)
    for nm, gr in zip(names, grads):
        want = g["grad_" + nm]
        got = gr.numpy() if gr is not None else np.zeros_like(want)
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-5, err_msg=nm)


@pytest.mark.parametrize("tag", ["default", "nonorm", "noloops_nobias", "E2layout"])
def test_golden_gcn(tag):
    g = load_golden("gcn_" + tag)
    p = g["params"]
    x = t(g["x"]).clone().requires_grad_(True)
    k = t(g["w_kernel"]).clone().requires_grad_(True)
    b = t(g["w_bias"]).clone().requires_grad_(True) if "w_bias" in g else None
    out = ref.gcn_conv(x, t(g["edge_index"]), k, b, p.get("add_self_loops", True), p.get("normalize", True))
    _check(out, [x, k] + ([b] if b is not None else []), g, ["x", "kernel"] + (["bias"] if b is not None else []))


ACT = {"relu": torch.relu, None: None}


@pytest.mark.parametrize("tag", ["mean", "max", "sum", "min", "std", "pooling", "mean_noroot_norm", "mean_linear"])
def test_golden_sage(tag):
    g = load_golden("sage_" + tag)
    p = g["params"]
    names = ["x", "lin_neigh"]
    x = t(g["x"]).clone().requires_grad_(True)
    wn = t(g["w_lin_neigh"]).clone().requires_grad_(True)
    params = [x, wn]
    ws = b = pw = pb = None
    if "w_lin_self" in g:
        ws = t(g["w_lin_self"]).clone().requires_grad_(True); params.append(ws); names.append("lin_self")
    if "w_pool_kernel" in g:
        pw = t(g["w_pool_kernel"]).clone().requires_grad_(True); pb = t(g["w_pool_bias"]).clone().requires_grad_(True)
        params += [pw, pb]; names += ["pool_kernel", "pool_bias"]
    if "w_bias" in g:
        b = t(g["w_bias"]).clone().requires_grad_(True); params.append(b); names.append("bias")
    out = ref.sage_conv(x, t(g["edge_index"]), wn, ws, b, p["aggregator"], ACT[p.get("activation", "relu")],
                        p.get("normalize", False), pw, pb, torch.relu)
    _check(out, params, g, names)


@pytest.mark.parametrize("tag", ["sum", "mean_eps", "max"])
def test_golden_gin(tag):
    g = load_golden("gin_" + tag)
    p = g["params"]
    n_dense = len(p["mlp_hidden"]) + 1
    ws = [(t(g[f"w_mlp{i}_kernel"]).clone().requires_grad_(True), t(g[f"w_mlp{i}_bias"]).clone().requires_grad_(True))
          for i in range(n_dense)]
    eps = t(g["w_eps"]).clone().requires_grad_(True) if "w_eps" in g else p.get("eps_init", 0.0)

    def mlp(h):
        for i, (k, b) in enumerate(ws):
            h = h @ k + b
            if i < n_dense - 1:
                h = torch.relu(h)
        return h

    x = t(g["x"]).clone().requires_grad_(True)
    out = ref.gin_conv(x, t(g["edge_index"]), mlp, eps, p["aggregator"])
    params, names = [x], ["x"]
    for i, (k, b) in enumerate(ws):
        params += [k, b]; names += [f"mlp{i}_kernel", f"mlp{i}_bias"]
    if "w_eps" in g:
        params.append(eps); names.append("eps")
    _check(out, params, g, names)


@pytest.mark.parametrize("tag", ["h4c16", "h8c8", "h1c3", "h2c5_mean", "h3c4_noloops"])
def test_golden_gatv2(tag):
    g = load_golden("gatv2_" + tag)
    p = g["params"]
    x = t(g["x"]).clone().requires_grad_(True)
    w = t(g["w_linear_transform"]).clone().requires_grad_(True)
    a = t(g["w_att"]).clone().requires_grad_(True)
    b = t(g["w_bias"]).clone().requires_grad_(True) if "w_bias" in g else None
    out = ref.gatv2_conv(x, t(g["edge_index"]), w, a, b, p.get("heads", 1), p.get("concat", True),
                         p.get("negative_slope", 0.2), p.get("add_self_loops", True))
    _check(out, [x, w, a] + ([b] if b is not None else []), g,
           ["x", "linear_transform", "att"] + (["bias"] if b is not None else []))


def test_stable_csr_matches_scatter_order():
    rng = np.random.default_rng(3)
    ei = np.stack([rng.integers(0, 50, 400), rng.integers(0, 50, 400)]).astype(np.int32)
    rowptr, col, perm, deg = ref.stable_csr(ei, 50)
    assert rowptr[-1] == 400 and (np.diff(rowptr) == deg).all()
    for i in range(50):
        seg = perm[rowptr[i]:rowptr[i + 1]]
        assert (np.diff(seg) > 0).all() and (ei[1][seg] == i).all()
        assert (col[rowptr[i]:rowptr[i + 1]] == ei[0][seg]).all()


def test_keras_segment_semantics():
    # out-of-range ids are silently dropped, empty segments are 0 / -inf
    d = t(np.array([[1.0], [2.0], [3.0]], np.float32))
    ids = t(np.array([0, 5, -1], np.int32))
    np.testing.assert_array_equal(kops.segment_sum(d, ids, 2).numpy(), [[1.0], [0.0]])
    np.testing.assert_array_equal(kops.segment_max(d, ids, 2).numpy(), [[1.0], [-np.inf]])


def test_golden_pooling_and_batch_graphs():
    g = load_golden("pooling")
    for pool in ["mean", "max", "sum"]:
        np.testing.assert_allclose(ref.global_pooling(t(g["x"]), pool).numpy(), g["global_" + pool], rtol=RTOL, atol=ATOL)
        x = t(g["x"]).clone().requires_grad_(True)
        out = ref.batch_global_pooling(x, t(g["batch"]), pool)
        np.testing.assert_allclose(out.detach().numpy(), g["batch_" + pool], rtol=RTOL, atol=ATOL)
        (gr,) = torch.autograd.grad((out * t(g["R_" + pool])).sum(), [x])
        np.testing.assert_allclose(gr.numpy(), g["grad_" + pool], rtol=RTOL, atol=ATOL)
    g = load_golden("batch_graphs")
    bx, bei, bb = ref.batch_graphs([g[f"x{i}"] for i in range(3)], [g[f"ei{i}"] for i in range(3)])
    np.testing.assert_array_equal(bx, g["bx"])
    np.testing.assert_array_equal(bei, g["bei"])
    np.testing.assert_array_equal(bb, g["bbatch"])


# ---------------------------------------------------- independent restatement by explicit edge loops (float64)
def _loop_aggregate(name, msgs, dst, n):
    """Per-target lists built edge by edge, reduced with the formulas of layers/aggregators.py:56-228."""
    buckets = [[] for _ in range(n)]
    for e, i in enumerate(dst):
        if 0 <= i < n:
            buckets[i].append(msgs[e].astype(np.float64))
    out = np.zeros((n, msgs.shape[1]))
    for i, b in enumerate(buckets):
        if not b:
            continue  # empty segment: 0 for every aggregator (max/min: -inf/+inf rewritten to 0)
        m = np.stack(b)
        if name == "sum":
            out[i] = m.sum(0)
        elif name == "mean":
            out[i] = m.sum(0) / max(float(len(b)), 1e-8)
        elif name == "max":
            out[i] = np.where(np.isinf(m.max(0)), 0.0, m.max(0))
        elif name == "min":
            out[i] = np.where(np.isinf(m.min(0)), 0.0, m.min(0))
        elif name == "std":
            out[i] = 0.0 if len(b) <= 1 else np.sqrt(np.maximum(((m - m.mean(0)) ** 2).mean(0), 0.0))
    return out


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_matches_edge_loops(seed):
    """The oracle's aggregators, GCN, SAGE, GIN-style sum and GATv2 against a restatement that shares no code with
    it: Python loops over edges in float64, written from the formulas in SURVEY.md Appendix A."""
    rng = np.random.default_rng(100 + seed)
    n, e, fin, fout = 23, 140, 6, 5
    ei = np.stack([rng.integers(0, n, e), rng.integers(0, n - 3, e)]).astype(np.int32)  # last rows stay empty
    x = rng.standard_normal((n, fin)).astype(np.float32)
    msgs = rng.standard_normal((e, 4)).astype(np.float32)
    for name in ("sum", "mean", "max", "min", "std"):
        got = ref.aggregate(name, t(msgs), t(ei[1]), n).numpy()
        np.testing.assert_allclose(got, _loop_aggregate(name, msgs, ei[1], n), rtol=1e-5, atol=2e-6, err_msg=name)

    # GCN: out_i = sum_e dis[i] dis[j] (x_j W) + b over edges + appended self-loops, dis = (deg + 1e-12)^-1/2
    w = (rng.standard_normal((fin, fout)) * 0.4).astype(np.float32)
    b = rng.standard_normal(fout).astype(np.float32)
    src = np.concatenate([ei[0], np.arange(n)])
    dst = np.concatenate([ei[1], np.arange(n)])
    deg = np.zeros(n)
    for i in dst:
        deg[i] += 1
    dis = (deg + 1e-12) ** -0.5
    xw = x.astype(np.float64) @ w.astype(np.float64)
    want = np.tile(b.astype(np.float64), (n, 1))
    for j, i in zip(src, dst):
        want[i] += dis[i] * dis[j] * xw[j]
    np.testing.assert_allclose(ref.gcn_conv(t(x), t(ei), t(w), t(b)).numpy(), want, rtol=1e-5, atol=1e-5)

    # SAGE (mean, relu): act(x W_self + mean_j(x_j) W_neigh + b)
    ws = (rng.standard_normal((fin, fout)) * 0.4).astype(np.float32)
    agg = _loop_aggregate("mean", x[ei[0]], ei[1], n)
    want = np.maximum(x.astype(np.float64) @ ws + agg @ w.astype(np.float64) + b, 0.0)
    got = ref.sage_conv(t(x), t(ei), t(w), t(ws), t(b), "mean", torch.relu).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)

    # GATv2: s = att . LeakyReLU(h_i + h_j); alpha = exp(s - max_i) / (sum_i exp + 1e-10); out_i = sum alpha h_j
    H, C = 2, 3
    wg = (rng.standard_normal((fin, H * C)) * 0.5).astype(np.float32)
    att = (rng.standard_normal((1, H, C)) * 0.5).astype(np.float32)
    bg = rng.standard_normal(H * C).astype(np.float32)
    h = (x.astype(np.float64) @ wg.astype(np.float64)).reshape(n, H, C)
    s_e = np.zeros((len(src), H))
    for k, (j, i) in enumerate(zip(src, dst)):
        z = h[i] + h[j]
        z = np.where(z > 0, z, 0.2 * z)
        s_e[k] = (z * att[0]).sum(-1)
    mx = np.full((n, H), -np.inf)
    for k, i in enumerate(dst):
        mx[i] = np.maximum(mx[i], s_e[k])
    den = np.zeros((n, H))
    for k, i in enumerate(dst):
        den[i] += np.exp(s_e[k] - mx[i])
    want = np.zeros((n, H, C))
    for k, (j, i) in enumerate(zip(src, dst)):
        alpha = np.exp(s_e[k] - mx[i]) / (den[i] + 1e-10)
        want[i] += alpha[:, None] * h[j]
    got = ref.gatv2_conv(t(x), t(ei), t(wg), t(att), t(bg), heads=H, concat=True).numpy()
    np.testing.assert_allclose(got, want.reshape(n, H * C) + bg, rtol=1e-5, atol=1e-5)
    got = ref.gatv2_conv(t(x), t(ei), t(wg), t(att), t(bg[:C]), heads=H, concat=False).numpy()
    np.testing.assert_allclose(got, want.mean(1) + bg[:C], rtol=1e-5, atol=1e-5)
