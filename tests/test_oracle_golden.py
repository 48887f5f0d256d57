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
