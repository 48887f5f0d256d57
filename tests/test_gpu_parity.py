"""GPU parity tests: the sm_100a kernels (through the C ABI, via ctypes) against the CPU oracle
and the committed golden vectors.  Bit-exact for integer/index work and for max/min values;
float32 results (forward outputs AND every gradient, also through the tensor-core GEMMs) within the north star's
1e-5 relative tolerance of the oracle: |got - want| <= 1e-5 * (|want| + max|want|), i.e. numpy's
rtol = 1e-5 plus atol = 1e-5 x the output scale.  No test uses a looser bound.
"""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import reference_path as ref

pytestmark = pytest.mark.gpu

RTOL = 1e-5


def close(got, want, rtol=RTOL, atol_scale=1e-5, msg=""):
    got = got.detach().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    want = want.detach().cpu().numpy() if isinstance(want, torch.Tensor) else np.asarray(want)
    assert got.shape == want.shape, f"{msg}: shape {got.shape} vs {want.shape}"
    finite = np.abs(want[np.isfinite(want)])
    scale = float(finite.max()) if finite.size else 1.0
    np.testing.assert_allclose(got, want, rtol=rtol, atol=atol_scale * max(scale, 1e-30), err_msg=msg, equal_nan=True)


def cuda(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return t.to(dtype) if dtype is not None else t


def rand_graph(rng, n_dst, n_src, e, hub=None):
    src = rng.integers(0, n_src, e)
    dst = rng.integers(0, max(n_dst - 2, 1), e)  # last rows stay empty
    if hub:
        dst[:hub] = 1
    return np.stack([src, dst]).astype(np.int32)


# ------------------------------------------------------------------------------------- K1 / K2
@pytest.mark.parametrize("n,e,loops", [(1, 0, 0), (1, 5, 1), (7, 0, 7), (50, 400, 0), (50, 400, 50),
                                       (300, 5000, 300), (70000, 300000, 70000), (5, 100000, 0)])
def test_csr_build_bit_exact(n, e, loops):
    from keras_geometric_b200.graph import GraphStructure
    rng = np.random.default_rng(n * 7 + e)
    ei = np.stack([rng.integers(0, n, e), rng.integers(0, n, e)]).astype(np.int32)
    g = GraphStructure(cuda(ei), n, n, loops)
    full = ei if not loops else np.concatenate([ei, np.stack([np.arange(loops), np.arange(loops)]).astype(np.int32)], 1)
    for by_source, s in ((False, g.csr), (True, g.csc)):
        rowptr, col, perm, deg = ref.stable_csr(full, n, by_source)
        np.testing.assert_array_equal(s.rowptr.cpu().numpy(), rowptr)
        np.testing.assert_array_equal(s.deg.cpu().numpy(), deg)
        np.testing.assert_array_equal(s.perm.cpu().numpy(), perm)
        np.testing.assert_array_equal(s.col.cpu().numpy(), col)


def test_csr_build_power_law_and_hubs():
    from keras_geometric_b200.graph import GraphStructure, HUB_THRESHOLD
    rng = np.random.default_rng(5)
    n, e = 20000, 600000
    dst = np.minimum((rng.pareto(1.2, e) * 3).astype(np.int64), n - 1)
    ei = np.stack([rng.integers(0, n, e), dst]).astype(np.int32)
    g = GraphStructure(cuda(ei), n, n, 0)
    rowptr, col, perm, deg = ref.stable_csr(ei, n)
    np.testing.assert_array_equal(g.csr.perm.cpu().numpy(), perm)
    np.testing.assert_array_equal(g.csr.col.cpu().numpy(), col)
    n_hubs = int((deg > HUB_THRESHOLD).sum())
    assert g.csr.n_hubs == n_hubs and n_hubs > 0
    hub_rows = sorted(g.csr.hub_row[:n_hubs].cpu().tolist())
    assert hub_rows == sorted(np.nonzero(deg > HUB_THRESHOLD)[0].tolist())


def test_csr_out_of_range_raises():
    from keras_geometric_b200.graph import GraphStructure
    ei = np.array([[0, 1, 9], [1, 2, 0]], np.int32)
    with pytest.raises(IndexError):
        GraphStructure(cuda(ei), 3, 3, 0)
    ei = np.array([[0, 1, 2], [1, -1, 0]], np.int32)
    with pytest.raises(IndexError):
        GraphStructure(cuda(ei), 3, 3, 0)


def test_utils_golden_bit_exact():
    import keras_geometric_b200 as kg
    g = load_golden("utils")
    n = g["params"]["N"]
    wl = kg.add_self_loops(g["edge_index"], n)
    np.testing.assert_array_equal(wl.cpu().numpy(), g["with_loops"])
    np.testing.assert_array_equal(kg.compute_gcn_normalization(wl, n).cpu().numpy(), g["gcn_norm"])
    np.testing.assert_array_equal(kg.compute_gcn_normalization(g["edge_index"], n).cpu().numpy(), g["gcn_norm_noloop"])


def test_gcn_norm_large_bit_exact():
    import keras_geometric_b200 as kg
    rng = np.random.default_rng(11)
    n, e = 5000, 200000
    ei = np.stack([rng.integers(0, n, e), np.minimum((rng.pareto(1.1, e) * 2).astype(np.int64), n - 1)]).astype(np.int32)
    want = ref.compute_gcn_normalization(torch.from_numpy(ei), n).numpy()
    np.testing.assert_array_equal(kg.compute_gcn_normalization(ei, n).cpu().numpy(), want)


# ------------------------------------------------------------------------------ K3 / K4 / K5
@pytest.mark.parametrize("F", [1, 3, 7, 16, 32, 64, 100, 128, 256, 516, 1030])
@pytest.mark.parametrize("op", ["sum", "mean", "max", "min"])
def test_gather_reduce_vs_oracle(F, op):
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    rng = np.random.default_rng(F * 13 + len(op))
    n_dst, n_src, e = 211, 157, 6000
    ei = rand_graph(rng, n_dst, n_src, e, hub=2500)  # row 1 is a hub -> chunked path
    x = np.round(rng.standard_normal((n_src, F)), 1).astype(np.float32)  # rounding forces max ties
    R = rng.standard_normal((n_dst, F)).astype(np.float32)
    graph = GraphStructure(cuda(ei), n_dst, n_src, 0)
    assert graph.csr.n_hubs >= 1
    xg = cuda(x).requires_grad_(True)
    out = ops.gather_reduce(xg, graph, op)
    xo = torch.from_numpy(x).requires_grad_(True)
    want = ref.propagate((torch.zeros(n_dst, F), xo), torch.from_numpy(ei), op)
    if op in ("max", "min"):
        np.testing.assert_array_equal(out.detach().cpu().numpy(), want.detach().numpy())
    else:
        close(out, want, msg=f"{op} F={F}")
    (gx,) = torch.autograd.grad((out * cuda(R)).sum(), [xg])
    (gw,) = torch.autograd.grad((want * torch.from_numpy(R)).sum(), [xo])
    close(gx, gw, msg=f"grad {op} F={F}")


@pytest.mark.parametrize("op", ["max", "min"])
@pytest.mark.parametrize("F", [8, 100, 256])
def test_max_backward_run_to_run_identical(op, F):
    """The max/min backward scatters into source entries that MANY targets select (few sources, many targets, values
    rounded so that ties between different sources are common, one hub row).  The fixed-point accumulation makes the
    result independent of the order in which the atomics land: repeated runs must agree bit for bit, and match the
    oracle (torch amax backward: even split between tied maxima) within the fp32 tolerance."""
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    rng = np.random.default_rng(F + len(op))
    n_dst, n_src, e = 20000, 64, 200000          # ~3000 targets pick each (source, feature) entry
    ei = rand_graph(rng, n_dst, n_src, e, hub=3000)
    x = np.round(rng.standard_normal((n_src, F)), 1).astype(np.float32)
    R = (rng.standard_normal((n_dst, F)) * np.exp(rng.uniform(-8, 8, (n_dst, 1)))).astype(np.float32)  # wide range
    graph = GraphStructure(cuda(ei), n_dst, n_src, 0)
    xg = cuda(x).requires_grad_(True)
    Rg = cuda(R)
    runs = []
    for _ in range(4):
        out = ops.gather_reduce(xg, graph, op)
        (gx,) = torch.autograd.grad((out * Rg).sum(), [xg])
        runs.append(gx.cpu().numpy())
    for r in runs[1:]:
        np.testing.assert_array_equal(runs[0], r)
    xo = torch.from_numpy(x).requires_grad_(True)
    want = ref.propagate((torch.zeros(n_dst, F), xo), torch.from_numpy(ei), op)
    (gw,) = torch.autograd.grad((want * torch.from_numpy(R)).sum(), [xo])
    close(runs[0], gw, msg=f"{op} backward F={F}")
    # non-finite gradients propagate like in the reference (NaN / inf in -> NaN / inf out), finite ones are untouched
    R2 = R.copy()
    R2[5, 0], R2[6, 1] = np.inf, np.nan
    out = ops.gather_reduce(xg, graph, op)
    (g2,) = torch.autograd.grad((out * cuda(R2)).sum(), [xg])
    (w2,) = torch.autograd.grad((ref.propagate((torch.zeros(n_dst, F), xo), torch.from_numpy(ei), op)
                                 * torch.from_numpy(R2)).sum(), [xo])
    g2, w2 = g2.cpu().numpy(), w2.numpy()
    np.testing.assert_array_equal(np.isnan(g2), np.isnan(w2))
    np.testing.assert_array_equal(np.isinf(g2), np.isinf(w2))


@pytest.mark.parametrize("F", [1, 7, 32, 100, 256, 516])
def test_fused_std_vs_oracle(F):
    """KGB_OP_SQDEV (second pass of the std aggregator) fused with the gather and as the generic segment aggregator:
    forward and gradient vs the oracle, including a hub row (chunked path), empty rows, single-message rows (output
    forced to 0) and the reference's NaN gradients for rows whose variance is exactly 0."""
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    rng = np.random.default_rng(F)
    n_dst, n_src, e = 211, 157, 6000
    ei = rand_graph(rng, n_dst, n_src, e, hub=2500)
    ei[1, -1] = n_dst - 1          # (rand_graph leaves the last two rows empty) exactly one message: std 0,
    ei[1, -3:-1] = n_dst - 2       # NaN gradient like the reference; two equal messages: variance exactly 0
    ei[0, -3:-1] = 5
    x = rng.standard_normal((n_src, F)).astype(np.float32)
    R = rng.standard_normal((n_dst, F)).astype(np.float32)
    graph = GraphStructure(cuda(ei), n_dst, n_src, 0)
    assert graph.csr.n_hubs >= 1
    xo = torch.from_numpy(x).requires_grad_(True)
    eio = torch.from_numpy(ei)
    want = ref.aggregate("std", xo[eio[0].long()], eio[1], n_dst)
    (gw,) = torch.autograd.grad((want * torch.from_numpy(R)).sum(), [xo])
    xg = cuda(x).requires_grad_(True)
    out = ops.gather_std(xg, graph)
    close(out, want, msg=f"std F={F}")
    (gx,) = torch.autograd.grad((out * cuda(R)).sum(), [xg])
    assert np.isnan(gw.numpy()).any()          # the reference's 0 * inf on zero-variance rows
    close(gx, gw, msg=f"std grad F={F}")       # equal_nan: NaN positions must coincide
    # generic aggregator over materialised messages
    m = rng.standard_normal((e, F)).astype(np.float32)
    mo = torch.from_numpy(m).requires_grad_(True)
    want2 = ref.aggregate("std", mo, eio[1], n_dst)
    (gw2,) = torch.autograd.grad((want2 * torch.from_numpy(R)).sum(), [mo])
    mg = cuda(m).requires_grad_(True)
    out2 = ops.segment_std(mg, graph)
    close(out2, want2, msg=f"segment std F={F}")
    (gm,) = torch.autograd.grad((out2 * cuda(R)).sum(), [mg])
    close(gm, gw2, msg=f"segment std grad F={F}")


def test_unweighted_sum_matches_host_order_bitwise():
    """Non-hub rows accumulate in CSR (= original edge) order, like the sequential host scatter."""
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    rng = np.random.default_rng(2)
    n, e, F = 500, 6000, 64
    ei = np.stack([rng.integers(0, n, e), rng.integers(0, n, e)]).astype(np.int32)
    x = rng.standard_normal((n, F)).astype(np.float32)
    graph = GraphStructure(cuda(ei), n, n, 0)
    assert graph.csr.n_hubs == 0
    out = ops.gather_reduce(cuda(x), graph, "sum").cpu().numpy()
    want = ref.aggregate("sum", torch.from_numpy(x[ei[0]]), torch.from_numpy(ei[1]), n).numpy()
    np.testing.assert_array_equal(out, want)


def test_gather_reduce_special_values_and_determinism():
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    rng = np.random.default_rng(9)
    n, e, F = 40, 300, 8
    ei = rand_graph(rng, n, n, e)
    x = rng.standard_normal((n, F)).astype(np.float32)
    x[3, 2] = np.inf
    x[5, 1] = -np.inf
    x[7, 4] = np.nan
    graph = GraphStructure(cuda(ei), n, n, 0)
    for op in ["sum", "mean", "max", "min"]:
        a = ops.gather_reduce(cuda(x), graph, op).cpu().numpy()
        b = ops.gather_reduce(cuda(x), graph, op).cpu().numpy()
        np.testing.assert_array_equal(a, b)  # run-to-run identical
        want = ref.propagate(torch.from_numpy(x), torch.from_numpy(ei), op).numpy()
        np.testing.assert_allclose(a, want, rtol=1e-5, atol=1e-6, equal_nan=True)


def test_fused_epilogue_and_gcn_weights():
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    rng = np.random.default_rng(4)
    n, e, F = 300, 4000, 48
    ei = rand_graph(rng, n, n, e, hub=700)
    x = rng.standard_normal((n, F)).astype(np.float32)
    ad = rng.standard_normal((n, F)).astype(np.float32)
    b = rng.standard_normal(F).astype(np.float32)
    R = rng.standard_normal((n, F)).astype(np.float32)
    graph = GraphStructure(cuda(ei), n, n, n)  # with self-loops
    xg, ag, bg = cuda(x).requires_grad_(True), cuda(ad).requires_grad_(True), cuda(b).requires_grad_(True)
    out = ops.gather_reduce(xg, graph, "sum", weight="gcn", addend=ag, addend_scale=1.5, bias=bg, act="relu")
    xo, ao, bo = (torch.from_numpy(t).requires_grad_(True) for t in (x, ad, b))
    full = ref.add_self_loops(torch.from_numpy(ei), n)
    w = ref.compute_gcn_normalization(full, n)
    agg = ref.aggregate("sum", xo[full[0].long()] * w[:, None], full[1], n)
    want = torch.relu(agg + 1.5 * ao + bo)
    close(out, want, msg="fused fwd")
    got = torch.autograd.grad((out * cuda(R)).sum(), [xg, ag, bg])
    exp = torch.autograd.grad((want * torch.from_numpy(R)).sum(), [xo, ao, bo])
    for a_, b_, nm in zip(got, exp, ["x", "addend", "bias"]):
        close(a_, b_, msg="fused grad " + nm)
    # explicit COO edge weights take the edge_w path
    wv = rng.random(e + n).astype(np.float32)
    out2 = ops.gather_reduce(cuda(x), graph, "sum", weight=cuda(wv))
    want2 = ref.aggregate("sum", torch.from_numpy(x)[full[0].long()] * torch.from_numpy(wv)[:, None], full[1], n)
    close(out2, want2, msg="edge_w fwd")


def test_aggregator_kats_on_device():
    """The reference's hand-computed KATs (tests/test_message_passing.py:54-155) through the product."""
    from keras_geometric_b200 import MessagePassing
    idx = np.array([0, 0, 1], np.int32)
    r = MessagePassing(aggregator="mean").aggregate(np.array([[1, 2], [3, 4], [5, 6]], np.float32), idx, num_nodes=3)
    np.testing.assert_allclose(r.cpu().numpy(), [[2, 3], [5, 6], [0, 0]], rtol=1e-5)
    r = MessagePassing(aggregator="max").aggregate(np.array([[1, 5], [3, 2], [2, 4]], np.float32), idx, num_nodes=3)
    np.testing.assert_array_equal(r.cpu().numpy(), [[3, 5], [2, 4], [0, 0]])
    r = MessagePassing(aggregator="sum").aggregate(np.array([[1, 2], [3, 4], [5, 6]], np.float32), idx, num_nodes=3)
    np.testing.assert_allclose(r.cpu().numpy()[0], [4, 6], rtol=1e-5)
    r = MessagePassing(aggregator="min").aggregate(np.array([[1, 5], [3, 2], [2, 4]], np.float32), idx, num_nodes=3)
    np.testing.assert_array_equal(r.cpu().numpy()[0], [1, 2])
    r = MessagePassing(aggregator="std").aggregate(np.array([[1, 2], [3, 4], [5, 6], [7, 8]], np.float32),
                                                   np.array([0, 0, 1, 1], np.int32), num_nodes=2)
    np.testing.assert_allclose(r.cpu().numpy(), [[1, 1], [1, 1]], rtol=1e-5)
    lyr = MessagePassing(aggregator="mean")
    assert tuple(lyr.propagate(x=np.zeros((0, 8), np.float32), edge_index=np.zeros((2, 0), np.int32)).shape) == (0, 8)
    out = lyr.propagate(x=np.random.randn(5, 8).astype(np.float32), edge_index=np.zeros((2, 0), np.int32))
    np.testing.assert_array_equal(out.cpu().numpy(), np.zeros((5, 8), np.float32))
    for v in (1e10, 1e-10):
        o = lyr.aggregate(np.full((100, 10), v, np.float32), np.zeros(100, np.int32), num_nodes=1).cpu().numpy()
        assert np.isfinite(o).all()


def test_golden_aggregators_and_message_passing():
    from keras_geometric_b200 import MessagePassing
    from keras_geometric_b200.layers import AggregatorFactory
    g = load_golden("aggregators")
    n = int(g["dim_size"])
    for name in ["mean", "max", "sum", "min", "std"]:
        out = AggregatorFactory.create(name).aggregate(g["messages"], g["target_idx"], n)
        np.testing.assert_allclose(out.cpu().numpy(), g["out_" + name], rtol=1e-5, atol=1e-6, equal_nan=True)
        m = cuda(g["messages_finite"]).requires_grad_(True)
        o = AggregatorFactory.create(name).aggregate(m, cuda(g["target_idx"]), n)
        (gr,) = torch.autograd.grad((o * cuda(g["R"])).sum(), [m])
        close(gr, g["grad_" + name], msg="agg grad " + name)
    np.testing.assert_array_equal(
        AggregatorFactory.create("max").aggregate(g["messages"], g["target_idx"], n).cpu().numpy(), g["out_max"])
    g = load_golden("message_passing")
    for name in ["mean", "max", "sum", "min", "std"]:
        x = cuda(g["x"]).requires_grad_(True)
        out = MessagePassing(aggregator=name).propagate(x=x, edge_index=cuda(g["edge_index"]))
        close(out, g["out_" + name], msg="propagate " + name)
        (gx,) = torch.autograd.grad((out * cuda(g["R_" + name])).sum(), [x])
        close(gx, g["grad_x_" + name], msg="propagate grad " + name)
    out = MessagePassing(aggregator="sum").propagate(x=(g["bip_x_target"], g["x"]), edge_index=g["bip_edge_index"])
    close(out, g["bip_out_sum"], msg="bipartite")


def test_user_subclass_generic_path():
    """Overridden hooks force the materialised path (reference tests/test_message_passing.py:256-312)."""
    from keras_geometric_b200 import MessagePassing

    class Custom(MessagePassing):
        def message(self, x_i, x_j, **kwargs):
            return x_j * 2.0 + x_i

        def update(self, aggregated, x=None):
            return aggregated + x

    rng = np.random.default_rng(8)
    n, e, F = 30, 200, 6
    ei = rand_graph(rng, n, n, e)
    x = rng.standard_normal((n, F)).astype(np.float32)
    xg = cuda(x).requires_grad_(True)
    out = Custom(aggregator="mean")([xg, ei])
    xo = torch.from_numpy(x).requires_grad_(True)
    want = ref.propagate(xo, torch.from_numpy(ei), "mean", lambda xi, xj: xj * 2.0 + xi, lambda a, x0: a + x0)
    close(out, want, msg="custom fwd")
    (gx,) = torch.autograd.grad(out.sum(), [xg])
    (gw,) = torch.autograd.grad(want.sum(), [xo])
    close(gx, gw, msg="custom grad")


# --------------------------------------------------------------------------------- conv layers
def _grads(out, R, params):
    return torch.autograd.grad((out * cuda(R)).sum(), params, allow_unused=True)


def _check_grads(grads, g, names, rtol=RTOL):
    for nm, gr in zip(names, grads):
        want = g["grad_" + nm]
        got = gr if gr is not None else torch.zeros(want.shape)
        close(got.reshape(want.shape), want, rtol=rtol, msg="grad " + nm)


def _set(p, arr):
    with torch.no_grad():
        p.copy_(cuda(arr))


@pytest.mark.parametrize("tag", ["default", "nonorm", "noloops_nobias", "E2layout"])
def test_golden_gcn(tag):
    from keras_geometric_b200 import GCNConv
    g = load_golden("gcn_" + tag)
    layer = GCNConv(**g["params"])
    x = cuda(g["x"]).requires_grad_(True)
    layer.build([tuple(g["x"].shape), tuple(g["edge_index"].shape)])
    layer.built = True
    _set(layer.kernel, g["w_kernel"])
    params, names = [x, layer.kernel], ["x", "kernel"]
    if "w_bias" in g:
        _set(layer.bias, g["w_bias"])
        params.append(layer.bias); names.append("bias")
    out = layer([x, g["edge_index"]])
    close(out, g["out"], msg="gcn fwd " + tag)
    _check_grads(_grads(out, g["R"], params), g, names)


@pytest.mark.parametrize("tag", ["mean", "max", "sum", "min", "std", "pooling", "mean_noroot_norm", "mean_linear"])
def test_golden_sage(tag):
    from keras_geometric_b200 import SAGEConv
    g = load_golden("sage_" + tag)
    layer = SAGEConv(**g["params"])
    x = cuda(g["x"]).requires_grad_(True)
    layer.build([tuple(g["x"].shape), tuple(g["edge_index"].shape)])
    layer.built = True
    _set(layer.lin_neigh.kernel, g["w_lin_neigh"])
    params, names = [x, layer.lin_neigh.kernel], ["x", "lin_neigh"]
    if "w_lin_self" in g:
        _set(layer.lin_self.kernel, g["w_lin_self"]); params.append(layer.lin_self.kernel); names.append("lin_self")
    if "w_pool_kernel" in g:
        _set(layer.pool_mlp.kernel, g["w_pool_kernel"]); _set(layer.pool_mlp.bias, g["w_pool_bias"])
        params += [layer.pool_mlp.kernel, layer.pool_mlp.bias]; names += ["pool_kernel", "pool_bias"]
    if "w_bias" in g:
        _set(layer.bias, g["w_bias"]); params.append(layer.bias); names.append("bias")
    out = layer([x, g["edge_index"]])
    close(out, g["out"], msg="sage fwd " + tag)
    _check_grads(_grads(out, g["R"], params), g, names)


def test_sage_reordered_fused_path():
    """output_dim < input_dim with a linear aggregator: aggregate after lin_neigh, epilogue fused."""
    from keras_geometric_b200 import SAGEConv
    rng = np.random.default_rng(21)
    n, e, fin, fout = 400, 5000, 40, 12
    ei = rand_graph(rng, n, n, e, hub=900)
    x = rng.standard_normal((n, fin)).astype(np.float32)
    wn, ws = (rng.standard_normal((fin, fout)).astype(np.float32) * 0.3 for _ in range(2))
    b = rng.standard_normal(fout).astype(np.float32)
    R = rng.standard_normal((n, fout)).astype(np.float32)
    for aggr in ("mean", "sum"):
        layer = SAGEConv(fout, aggregator=aggr)
        layer.build([(n, fin), (2, e)]); layer.built = True
        _set(layer.lin_neigh.kernel, wn); _set(layer.lin_self.kernel, ws); _set(layer.bias, b)
        xg = cuda(x).requires_grad_(True)
        out = layer([xg, ei])
        ts = [torch.from_numpy(t).requires_grad_(True) for t in (x, wn, ws, b)]
        want = ref.sage_conv(ts[0], torch.from_numpy(ei), ts[1], ts[2], ts[3], aggr, torch.relu)
        close(out, want, msg="sage reorder " + aggr)
        got = torch.autograd.grad((out * cuda(R)).sum(), [xg, layer.lin_neigh.kernel, layer.lin_self.kernel, layer.bias])
        exp = torch.autograd.grad((want * torch.from_numpy(R)).sum(), ts)
        for a_, b_ in zip(got, exp):
            close(a_, b_, msg="sage reorder grad")


@pytest.mark.parametrize("fin,fout,act", [(40, 64, "relu"), (64, 64, None), (100, 256, "relu"), (256, 48, "relu")])
def test_sage_one_node_paths_vs_oracle(fin, fout, act):
    """Sizes that take the single-autograd-node paths (ops.sage_layer / ops.linear_pair on the tcgen05 kernels):
    output and all four gradients against the oracle's SAGEConv."""
    from keras_geometric_b200 import SAGEConv
    rng = np.random.default_rng(fin * 7 + fout)
    n, e = 700, 9000
    ei = rand_graph(rng, n, n, e, hub=1200)
    x = rng.standard_normal((n, fin)).astype(np.float32)
    wn, ws = (rng.standard_normal((fin, fout)).astype(np.float32) * 0.2 for _ in range(2))
    b = rng.standard_normal(fout).astype(np.float32)
    R = rng.standard_normal((n, fout)).astype(np.float32)
    for aggr in ("mean", "sum"):
        layer = SAGEConv(fout, aggregator=aggr, activation=act)
        layer.build([(n, fin), (2, e)]); layer.built = True
        _set(layer.lin_neigh.kernel, wn); _set(layer.lin_self.kernel, ws); _set(layer.bias, b)
        xg = cuda(x).requires_grad_(True)
        out = layer([xg, ei])
        ts = [torch.from_numpy(t).requires_grad_(True) for t in (x, wn, ws, b)]
        want = ref.sage_conv(ts[0], torch.from_numpy(ei), ts[1], ts[2], ts[3], aggr, torch.relu if act else None)
        close(out, want, msg=f"sage one-node {aggr}")
        got = torch.autograd.grad((out * cuda(R)).sum(), [xg, layer.lin_neigh.kernel, layer.lin_self.kernel, layer.bias])
        exp = torch.autograd.grad((want * torch.from_numpy(R)).sum(), ts)
        for a_, b_, nm in zip(got, exp, ("dx", "dWn", "dWs", "db")):
            close(a_, b_, msg=f"sage one-node grad {nm} {aggr}")
        # x without gradient (first layer of a model): weight gradients only
        out2 = layer([cuda(x), ei])
        got2 = torch.autograd.grad((out2 * cuda(R)).sum(), [layer.lin_neigh.kernel, layer.lin_self.kernel, layer.bias])
        for a_, b_ in zip(got2, exp[1:]):
            close(a_, b_, msg="sage one-node grad (x const)")


@pytest.mark.parametrize("rows,F", [(1, 4), (63, 48), (1000, 100), (70001, 256), (5000, 1024)])
def test_relu_bwd_colsum(rows, F):
    from keras_geometric_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(rows + F)
    g = torch.randn((rows, F), device="cuda", generator=gen)
    y = torch.randn((rows, F), device="cuda", generator=gen)
    gp, col = ops.relu_bwd_colsum(g, y)
    want = torch.where(y > 0, g, torch.zeros_like(g))
    assert torch.equal(gp, want)
    close(col, want.double().sum(0).float(), rtol=1e-5, atol_scale=1e-5, msg="colsum")
    gp2, col2 = ops.relu_bwd_colsum(g, y)
    assert torch.equal(col, col2)          # fixed reduction order
    same, col3 = ops.relu_bwd_colsum(g, None)
    assert same is g
    close(col3, g.double().sum(0).float(), rtol=1e-5, atol_scale=1e-5, msg="plain colsum")
    # padded rows (a [:, :F] view of a wider buffer)
    wide = torch.randn((rows, F + 4), device="cuda", generator=gen)
    _, col4 = ops.relu_bwd_colsum(wide[:, :F], None)
    close(col4, wide[:, :F].double().sum(0).float(), rtol=1e-5, atol_scale=1e-5, msg="strided colsum")


@pytest.mark.parametrize("rows,F", [(1, 1), (7, 3), (300, 47), (1000, 256), (33, 1000), (5, 1500)])
def test_l2_normalize_rows(rows, F):
    """ops.l2_normalize == keras.ops.normalize(axis=-1, order=2) (oracle/keras_ops.py:276) forward and backward,
    including all-zero rows (norm below eps: y = 0, gradient g / eps) and a tiny row."""
    from keras_geometric_b200 import ops
    from oracle import keras_ops as kops
    rng = np.random.default_rng(rows * 7 + F)
    x = rng.standard_normal((rows, F)).astype(np.float32)
    x[0] = 0.0
    if rows > 2:
        x[2] *= 1e-20
    R = rng.standard_normal((rows, F)).astype(np.float32)
    xg = cuda(x).requires_grad_(True)
    out = ops.l2_normalize(xg)
    xo = torch.from_numpy(x).requires_grad_(True)
    want = kops.normalize(xo, axis=-1, order=2)
    close(out, want, msg="l2 normalize fwd")
    (gx,) = torch.autograd.grad((out * cuda(R)).sum(), [xg])
    (gw,) = torch.autograd.grad((want * torch.from_numpy(R)).sum(), [xo])
    # rows whose norm is far below eps have gradients of 1e12 * g; compare row by row at the row's own scale
    got, exp = gx.cpu().numpy(), gw.numpy()
    for r in range(rows):
        close(got[r], exp[r], msg=f"l2 normalize bwd row {r}")


@pytest.mark.parametrize("rows,C", [(1, 2), (100, 47), (5000, 47), (3000, 7), (2000, 130), (700, 1000)])
def test_softmax_cross_entropy(rows, C):
    from keras_geometric_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(rows * 3 + C)
    pad = (-C) % 4
    base = torch.randn((rows, C + pad), device="cuda", generator=gen) * 3
    logits = base[:, :C].detach().requires_grad_(True)   # strided rows like a padded layer output
    y = torch.randint(0, C, (rows,), device="cuda", generator=gen)
    loss = ops.softmax_cross_entropy(logits, y)
    (gl,) = torch.autograd.grad(loss * 1.7, [logits])
    ref_logits = base[:, :C].double().detach().requires_grad_(True)
    want = torch.nn.functional.cross_entropy(ref_logits, y)
    (gw,) = torch.autograd.grad(want * 1.7, [ref_logits])
    assert abs(float(loss) - float(want)) <= 1e-5 * max(1.0, abs(float(want)))
    close(gl, gw.float(), msg="xent grad")


@pytest.mark.parametrize("tag", ["sum", "mean_eps", "max"])
def test_golden_gin(tag):
    from keras_geometric_b200 import GINConv
    g = load_golden("gin_" + tag)
    layer = GINConv(**g["params"])
    x = cuda(g["x"]).requires_grad_(True)
    layer.build([tuple(g["x"].shape), tuple(g["edge_index"].shape)])
    layer.built = True
    dense = [l for l in layer.mlp.layers if hasattr(l, "kernel")]
    params, names = [x], ["x"]
    for i, d in enumerate(dense):
        _set(d.kernel, g[f"w_mlp{i}_kernel"]); _set(d.bias, g[f"w_mlp{i}_bias"])
        params += [d.kernel, d.bias]; names += [f"mlp{i}_kernel", f"mlp{i}_bias"]
    if "w_eps" in g:
        _set(layer.eps, g["w_eps"]); params.append(layer.eps); names.append("eps")
    out = layer([x, g["edge_index"]])
    close(out, g["out"], msg="gin fwd " + tag)
    _check_grads(_grads(out, g["R"], params), g, names)


@pytest.mark.parametrize("tag", ["h4c16", "h8c8", "h1c3", "h2c5_mean", "h3c4_noloops"])
def test_golden_gatv2(tag):
    from keras_geometric_b200 import GATv2Conv
    g = load_golden("gatv2_" + tag)
    layer = GATv2Conv(**g["params"])
    x = cuda(g["x"]).requires_grad_(True)
    layer.build([tuple(g["x"].shape), tuple(g["edge_index"].shape)])
    layer.built = True
    _set(layer.linear_transform.kernel, g["w_linear_transform"]); _set(layer.att, g["w_att"])
    params, names = [x, layer.linear_transform.kernel, layer.att], ["x", "linear_transform", "att"]
    if "w_bias" in g:
        _set(layer.bias, g["w_bias"]); params.append(layer.bias); names.append("bias")
    out = layer([x, g["edge_index"]])
    close(out, g["out"], msg="gat fwd " + tag)
    _check_grads(_grads(out, g["R"], params), g, names)


@pytest.mark.parametrize("H,C", [(1, 1), (1, 64), (2, 7), (4, 32), (8, 8), (3, 20), (1, 256), (16, 4), (2, 160)])
def test_gatv2_kernel_vs_oracle(H, C):
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    rng = np.random.default_rng(H * 100 + C)
    n, e = 150, 2500
    ei = rand_graph(rng, n, n, e, hub=600)
    h = (rng.standard_normal((n, H * C)) * 0.7).astype(np.float32)
    att = (rng.standard_normal((1, H, C)) * 0.5).astype(np.float32)
    b = rng.standard_normal(H * C).astype(np.float32)
    R = rng.standard_normal((n, H * C)).astype(np.float32)
    graph = GraphStructure(cuda(ei), n, n, n)
    hg, ag, bg = cuda(h).requires_grad_(True), cuda(att).requires_grad_(True), cuda(b).requires_grad_(True)
    out = ops.gatv2_aggregate(hg, hg, ag, graph, H, C, 0.2, bg)
    ho, ao, bo = (torch.from_numpy(t).requires_grad_(True) for t in (h, att, b))
    want = ref.gatv2_conv(ho, torch.from_numpy(ei), torch.eye(H * C), ao, bo, H, True, 0.2, True)
    close(out, want, msg=f"gat H={H} C={C}")
    got = torch.autograd.grad((out * cuda(R)).sum(), [hg, ag, bg])
    exp = torch.autograd.grad((want * torch.from_numpy(R)).sum(), [ho, ao, bo])
    for a_, b_, nm in zip(got, exp, ["h", "att", "bias"]):
        close(a_, b_, msg=f"gat grad {nm} H={H} C={C}")


@pytest.mark.parametrize("H,C", [(8, 8), (1, 64), (4, 32), (2, 7)])
def test_gatv2_record_path(H, C, monkeypatch):
    """The opt-in per-edge record backward (KGB200_GAT_REC=1) gives the gradients of the recomputing default."""
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    rng = np.random.default_rng(H * 31 + C)
    n, e = 300, 6000
    ei = rand_graph(rng, n, n, e, hub=700)
    graph = GraphStructure(cuda(ei), n, n, n)
    h = cuda((rng.standard_normal((n, H * C)) * 0.7).astype(np.float32))
    att = cuda((rng.standard_normal((1, H, C)) * 0.5).astype(np.float32))
    R = cuda(rng.standard_normal((n, H * C)).astype(np.float32))
    grads = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("KGB200_GAT_REC", flag)
        hg, ag = h.clone().requires_grad_(True), att.clone().requires_grad_(True)
        out = ops.gatv2_aggregate(hg, hg, ag, graph, H, C, 0.2, None)
        grads[flag] = torch.autograd.grad((out * R).sum(), [hg, ag])
    for a_, b_, nm in zip(grads["1"], grads["0"], ["h", "att"]):
        close(a_, b_.cpu(), msg=f"gat record path grad {nm} H={H} C={C}")


def test_gatv2_bipartite_and_dropout_path():
    from keras_geometric_b200 import GATv2Conv
    rng = np.random.default_rng(3)
    xt, xs = rng.standard_normal((9, 6)).astype(np.float32), rng.standard_normal((20, 6)).astype(np.float32)
    ei = np.stack([rng.integers(0, 20, 50), rng.integers(0, 9, 50)]).astype(np.int32)
    layer = GATv2Conv(5, heads=2)
    out = layer.propagate(x=(xt, xs), edge_index=ei)
    assert tuple(out.shape) == (9, 10)
    want = None
    w, att, b = layer.linear_transform.kernel.detach().cpu(), layer.att.detach().cpu(), layer.bias.detach().cpu()
    hi, hj = torch.from_numpy(xt) @ w, torch.from_numpy(xs) @ w
    src, dst = torch.from_numpy(ei[0]).long(), torch.from_numpy(ei[1]).long()
    z = torch.nn.functional.leaky_relu(hi[dst].reshape(-1, 2, 5) + hj[src].reshape(-1, 2, 5), 0.2)
    s = (z * att).sum(-1)
    from oracle import keras_ops as kops
    m = kops.segment_max(s, dst, 9)
    p = torch.exp(s - m[dst])
    alpha = p / (kops.segment_sum(p, dst, 9)[dst] + 1e-10)
    want = kops.segment_sum((alpha.unsqueeze(-1) * hj[src].reshape(-1, 2, 5)).reshape(-1, 10), dst, 9) + b
    close(out, want, msg="gat bipartite")
    # dropout path: same expectation when the rate is tiny and training, shape/finite only otherwise
    lyr = GATv2Conv(4, heads=2, dropout=0.5)
    o = lyr([rng.standard_normal((20, 6)).astype(np.float32), np.stack([ei[0], ei[0]])], training=True)
    assert tuple(o.shape) == (20, 8) and torch.isfinite(o).all()


def test_layer_reuse_cache_and_inplace_update():
    """One layer, many graphs (reference tests/unit/test_error_handling.py:335-356) + in-place edits."""
    from keras_geometric_b200 import GCNConv
    rng = np.random.default_rng(17)
    layer = GCNConv(5)
    for n, e in [(10, 30), (25, 100), (7, 12)]:
        x = rng.standard_normal((n, 4)).astype(np.float32)
        ei = np.stack([rng.integers(0, n, e), rng.integers(0, n, e)]).astype(np.int32)
        out = layer([x, ei])
        want = ref.gcn_conv(torch.from_numpy(x), torch.from_numpy(ei), layer.kernel.detach().cpu(), layer.bias.detach().cpu())
        close(out, want, msg="reuse")
    n = 12
    x = cuda(rng.standard_normal((n, 4)).astype(np.float32))
    ei = cuda(np.stack([rng.integers(0, n, 40), rng.integers(0, n, 40)]).astype(np.int32))
    a = layer([x, ei]).clone()
    ei[1, :20] = 0  # in-place mutation must invalidate the cached structure
    b = layer([x, ei])
    want = ref.gcn_conv(x.cpu(), ei.cpu(), layer.kernel.detach().cpu(), layer.bias.detach().cpu())
    close(b, want, msg="after in-place edit")
    assert not torch.allclose(a, b)


def test_full_size_properties_products_slice():
    """Size-independent checks at a scale the oracle cannot reach quickly: linearity of sum,
    mean(ones) == [deg>0], sum(ones) == in-degree (exact integers), permutation invariance of max."""
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    gen = torch.Generator(device="cuda").manual_seed(0)
    n, e, F = 200_000, 5_000_000, 100
    dst = (torch.rand(e, device="cuda", generator=gen) ** 3 * n).long().clamp_(max=n - 1)
    src = torch.randint(0, n, (e,), device="cuda", generator=gen)
    ei = torch.stack([src, dst]).to(torch.int32)
    graph = GraphStructure(ei, n, n, 0)
    assert graph.csr.n_hubs > 0
    ones = torch.ones((n, F), device="cuda")
    deg = torch.bincount(dst, minlength=n).to(torch.float32)
    s = ops.gather_reduce(ones, graph, "sum")
    assert torch.equal(s, deg[:, None].expand(n, F))
    m = ops.gather_reduce(ones, graph, "mean")
    assert torch.equal(m, (deg > 0).float()[:, None].expand(n, F))
    x = torch.randn((n, F), device="cuda", generator=gen)
    y = torch.randn((n, F), device="cuda", generator=gen)
    lhs = ops.gather_reduce(x + y, graph, "sum")
    rhs = ops.gather_reduce(x, graph, "sum") + ops.gather_reduce(y, graph, "sum")
    assert torch.allclose(lhs, rhs, rtol=1e-4, atol=1e-3)
    perm = torch.randperm(e, device="cuda", generator=gen)
    graph2 = GraphStructure(ei[:, perm].contiguous(), n, n, 0)
    assert torch.equal(ops.gather_reduce(x, graph, "max"), ops.gather_reduce(x, graph2, "max"))
    # transposed structure: <A x, y> == <x, A^T y>
    xr = x.clone().requires_grad_(True)
    (gx,) = torch.autograd.grad((ops.gather_reduce(xr, graph, "sum") * y).sum(), [xr])
    ref_dot = (ops.gather_reduce(x, graph, "sum") * y).sum()
    assert torch.allclose((gx * x).sum(), ref_dot, rtol=1e-4)


def test_csc_prefetch_side_stream_and_l2_hints_change_nothing(monkeypatch):
    """The source-major orientation built on the side stream beside the forward pass (GraphStructure.prefetch_csc) is
    bit-identical to the one built in line, and neither it nor the L2 eviction hints (armed at the third use of a
    structure) change a single bit of outputs or gradients."""
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    gen = torch.Generator(device="cuda").manual_seed(5)
    n, e, F = 300_000, 3_000_000, 256
    dst = (torch.rand(e, device="cuda", generator=gen) ** 3 * n).long().clamp_(max=n - 1)
    src = (torch.rand(e, device="cuda", generator=gen) ** 2 * n).long().clamp_(max=n - 1)
    ei = torch.stack([src, dst]).to(torch.int32)
    x = torch.randn((n, F), device="cuda", generator=gen)
    R = torch.randn((n, F), device="cuda", generator=gen)

    def run(graph, reps=1):
        outs = []
        for _ in range(reps):
            xr = x.clone().requires_grad_(True)
            out = ops.gather_reduce(xr, graph, "mean")
            (gx,) = torch.autograd.grad(out, [xr], R)
            outs.append((out, gx))
        return outs

    monkeypatch.setenv("KGB200_PREFETCH_CSC", "0")
    monkeypatch.setenv("KGB200_HOT_MB", "0")
    g0 = GraphStructure(ei, n, n, 0)
    (o0, gx0), = run(g0)
    assert g0._csc_pending is None
    monkeypatch.setenv("KGB200_PREFETCH_CSC", "1")
    monkeypatch.setenv("KGB200_HOT_MB", "auto")
    g1 = GraphStructure(ei, n, n, 0)
    xr = x.clone().requires_grad_(True)
    out = ops.gather_reduce(xr, g1, "mean")
    assert g1._csc_pending is not None and g1._csc is None      # on its way, the host has not waited
    (gx,) = torch.autograd.grad(out, [xr], R)
    assert g1._csc_pending is None and g1._csc is not None
    for name in ("rowptr", "col", "perm", "deg"):
        assert torch.equal(getattr(g1.csc, name), getattr(g0.csc, name)), name
    assert (g1.csc.n_hubs, g1.csc.n_chunks) == (g0.csc.n_hubs, g0.csc.n_chunks)
    assert torch.equal(out, o0) and torch.equal(gx, gx0)
    for o, g in run(g1, reps=3):                                 # third and later uses run with the tagged columns
        assert torch.equal(o, o0) and torch.equal(g, gx0)
    assert any(isinstance(k, int) for k in g1.csr._hot), "hints were never armed"


# ------------------------------------------------------------------------------ fused dropout (training paths)
def test_dropout_mask_statistics_and_reproducibility():
    """The in-kernel Philox mask: keep rate 1 - p within 5 sigma for every feature column and head, a pure function of
    (edge id, element, seed), different for different seeds and for the per-head stream."""
    from keras_geometric_b200 import ops
    E, F, p = 200_000, 20, 0.3
    m1 = ops.dropout_mask(None, E, F, p, 1234).cpu().numpy()
    m2 = ops.dropout_mask(None, E, F, p, 1234).cpu().numpy()
    m3 = ops.dropout_mask(None, E, F, p, 1235).cpu().numpy()
    mh = ops.dropout_mask(None, E, F, p, 1234, per_head=True).cpu().numpy()
    np.testing.assert_array_equal(m1, m2)
    assert set(np.unique(m1)) == {0.0, np.float32(1.0 / (1.0 - p))}
    sigma = np.sqrt(p * (1 - p) / E)
    for m in (m1, m3, mh):
        keep = (m > 0).mean(axis=0)
        assert np.all(np.abs(keep - (1 - p)) < 5 * sigma), keep
    assert 0.4 < ((m1 > 0) != (m3 > 0)).mean() < 0.44         # independent masks disagree with prob. 2 p (1 - p)
    assert 0.4 < ((m1 > 0) != (mh > 0)).mean() < 0.44
    # edge ids select the rows of the same mask
    ids = torch.randperm(E, device="cuda")[:1000].to(torch.int32)
    sub = ops.dropout_mask(ids, 1000, F, p, 1234).cpu().numpy()
    np.testing.assert_array_equal(sub, m1[ids.cpu().numpy()])


@pytest.mark.parametrize("F,op,weight", [(7, "sum", None), (64, "mean", None), (100, "sum", "gcn"), (260, "mean", None)])
def test_fused_gather_dropout_matches_explicit_mask(F, op, weight):
    """Fused element-wise dropout of the gathered rows (GCN / SAGE training path): identical, forward and backward, to
    the reference's formulation with the SAME mask written out explicitly ([E, F] messages x mask, then weighted and
    reduced), including hub rows and (GCN) appended self-loops."""
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    rng = np.random.default_rng(F)
    n, e, p, seed = 300, 9000, 0.4, 99 + F
    ei = rand_graph(rng, n, n, e, hub=3000)
    loops = n if weight == "gcn" else 0
    x = rng.standard_normal((n, F)).astype(np.float32)
    R = rng.standard_normal((n, F)).astype(np.float32)
    graph = GraphStructure(cuda(ei), n, n, loops)
    assert graph.csr.n_hubs >= 1
    xg = cuda(x).requires_grad_(True)
    out = ops.gather_reduce(xg, graph, op, weight=weight, dropout=p, dropout_seed=seed)
    (gx,) = torch.autograd.grad((out * cuda(R)).sum(), [xg])
    mask = ops.dropout_mask(None, e + loops, F, p, seed).cpu()       # edge ids: real edges, then the self-loops
    eio = torch.from_numpy(ei)
    if loops:
        eio = ref.add_self_loops(eio, n)
    xo = torch.from_numpy(x).requires_grad_(True)
    msg = xo[eio[0].long()] * mask
    if weight == "gcn":
        msg = msg * ref.compute_gcn_normalization(eio, n).unsqueeze(1)
    want = ref.aggregate(op, msg, eio[1], n)
    (gw,) = torch.autograd.grad((want * torch.from_numpy(R)).sum(), [xo])
    close(out, want, msg=f"dropout fwd {op} F={F}")
    close(gx, gw, msg=f"dropout grad {op} F={F}")
    # without a seed every call draws a new mask; the expectation is the plain aggregation
    a = ops.gather_reduce(xg, graph, op, weight=weight, dropout=p)
    b = ops.gather_reduce(xg, graph, op, weight=weight, dropout=p)
    assert not torch.equal(a, b)


@pytest.mark.parametrize("H,C", [(1, 8), (4, 16), (8, 8), (3, 5)])
def test_fused_gatv2_attention_dropout_matches_explicit_mask(H, C):
    """GATv2 attention dropout inside the fused kernels vs the reference's per-edge formulation with the same
    (edge, head) mask written out: alpha = softmax over ALL edges, then alpha * mask (gatv2_conv.py:241-335)."""
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    from oracle import keras_ops as kops
    rng = np.random.default_rng(H * 100 + C)
    n, e, p, seed = 200, 5000, 0.35, 7 + H
    ei = rand_graph(rng, n, n, e, hub=1500)
    h = rng.standard_normal((n, H * C)).astype(np.float32)
    att = (rng.standard_normal((1, H, C)) * 0.5).astype(np.float32)
    R = rng.standard_normal((n, H * C)).astype(np.float32)
    graph = GraphStructure(cuda(ei), n, n, n)
    hg, ag = cuda(h).requires_grad_(True), cuda(att).requires_grad_(True)
    out = ops.gatv2_aggregate(hg, hg, ag, graph, H, C, 0.2, None, dropout=p, dropout_seed=seed)
    gh, ga = torch.autograd.grad((out * cuda(R)).sum(), [hg, ag])
    mask = ops.dropout_mask(None, e + n, H, p, seed, per_head=True).cpu()
    eio = ref.add_self_loops(torch.from_numpy(ei), n)
    src, dst = eio[0].long(), eio[1].int()
    ho, ao = torch.from_numpy(h).requires_grad_(True), torch.from_numpy(att).requires_grad_(True)
    hh = ho.reshape(n, H, C)
    z = torch.nn.functional.leaky_relu(hh[dst.long()] + hh[src], 0.2)
    s = (z * ao).sum(-1)
    m = kops.segment_max(s, dst, num_segments=n)
    pe = torch.exp(s - m[dst.long()])
    d = kops.segment_sum(pe, dst, num_segments=n)
    alpha = pe / (d[dst.long()] + 1e-10) * mask
    want = kops.segment_sum((alpha.unsqueeze(-1) * hh[src]).reshape(-1, H * C), dst, num_segments=n)
    gwh, gwa = torch.autograd.grad((want * torch.from_numpy(R)).sum(), [ho, ao])
    close(out, want, msg=f"gat dropout fwd H={H} C={C}")
    close(gh, gwh, msg=f"gat dropout grad h H={H} C={C}")
    close(ga.reshape(gwa.shape), gwa, msg=f"gat dropout grad att H={H} C={C}")


def test_layers_training_dropout_is_unbiased_and_inference_unaffected():
    """GCNConv / SAGEConv / GATv2Conv with dropout: training=False equals the dropout-free layer exactly, training=True
    differs from call to call and averages to it (inverted dropout is unbiased in the linear layers)."""
    import keras_geometric_b200 as kg
    rng = np.random.default_rng(3)
    n, e = 400, 4000
    ei = rand_graph(rng, n, n, e)
    x = cuda(rng.standard_normal((n, 16)).astype(np.float32))
    for make in (lambda: kg.GCNConv(8, dropout_rate=0.5), lambda: kg.SAGEConv(8, aggregator="mean", activation=None,
                                                                               root_weight=False, dropout_rate=0.5)):
        torch.manual_seed(0)
        layer = make()
        base = layer([x, ei], training=False)
        a, b = layer([x, ei], training=True), layer([x, ei], training=True)
        assert not torch.equal(a, b) and not torch.equal(a, base)
        avg = torch.stack([layer([x, ei], training=True) for _ in range(400)]).mean(0)
        err = (avg - base).abs().max() / base.abs().max()
        assert float(err) < 0.25, float(err)
    gat = kg.GATv2Conv(8, heads=2, dropout=0.5)
    base = gat([x, ei], training=False)
    a = gat([x, ei], training=True)
    assert tuple(a.shape) == tuple(base.shape) and not torch.equal(a, base)
    wide = kg.GATv2Conv(516, heads=1)           # per-head width beyond the fused layout: per-edge path, still correct
    o = wide([x, ei])
    w_, a_, b_ = (t.detach().cpu() for t in (wide.linear_transform.kernel, wide.att, wide.bias))
    close(o, ref.gatv2_conv(x.cpu(), torch.from_numpy(ei), w_, a_, b_, heads=1), msg="gat wide head")


# ---------------------------------------------------------------------------------------- K8
@pytest.mark.parametrize("M,K,N", [(1, 4, 4), (300, 100, 256), (5000, 256, 48), (20001, 48, 256), (70000, 64, 64),
                                   (513, 12, 7), (1000, 1433, 16), (128, 32, 64), (40000, 260, 132), (9999, 512, 300),
                                   (5, 3, 2), (17, 1433, 7), (2708, 16, 7), (19717, 500, 64), (19717, 64, 3),
                                   (100, 600, 520), (90001, 256, 256), (80000, 100, 200)])   # last two: CTA-pair kernel
def test_linear_tensor_core_gemm(M, K, N):
    """X @ W (+ addend) and its two gradient GEMMs vs float64 for every shape class: ragged widths (zero-padded to
    multiples of 4), widths above 256 (column slabs), fewer rows than one tile (TMA zero fill).  Every launch must be
    one of the library's own kernels - there is no cuBLAS / CUTLASS path."""
    from keras_geometric_b200 import _lib, ops
    l0 = _lib.load().kgb_launch_count()
    gen = torch.Generator(device="cuda").manual_seed(M + K + N)
    x = torch.randn((M, K), device="cuda", generator=gen, requires_grad=True)
    w = (torch.randn((K, N), device="cuda", generator=gen) * 0.2).requires_grad_(True)
    ad = torch.randn((M, N), device="cuda", generator=gen, requires_grad=True)
    R = torch.randn((M, N), device="cuda", generator=gen)
    out = ops.linear(x, w, addend=ad)
    gx, gw, gad = torch.autograd.grad((out * R).sum(), [x, w, ad])
    assert _lib.load().kgb_launch_count() - l0 >= 3 * max(1, -(-N // 256))   # fwd + dX + dW slabs, all kgb:: kernels
    x64, w64 = x.detach().double(), w.detach().double()
    # stated tolerance: 1e-5 relative to the output scale (3xTF32 emulation of fp32)
    close(out, (x64 @ w64 + ad.detach().double()).float(), rtol=1e-5, atol_scale=1e-5, msg="gemm fwd")
    close(gx, (R.double() @ w64.t()).float(), rtol=1e-5, atol_scale=1e-5, msg="gemm dX")
    close(gw, (x64.t() @ R.double()).float(), rtol=1e-5, atol_scale=1e-5, msg="gemm dW")
    close(gad, R, msg="gemm d addend")


@pytest.mark.parametrize("M,K,Na,Nb", [(5000, 256, 48, 48), (777, 100, 12, 64), (40000, 64, 128, 96), (300, 256, 200, 100)])
def test_two_weight_gradients_one_pass(M, K, Na, Nb):
    """ops._dw_tc2: (X^T Ga, X^T Gb) with X loaded and split once (kgb_linear_tc_dw2) against fp64; shapes beyond one
    launch (ceil32(Na) + Nb > 256) fall back to two kgb_linear_tc_dw calls."""
    from keras_geometric_b200 import ops
    rng = np.random.default_rng(M + K + Na)
    x = cuda(rng.standard_normal((M, K)).astype(np.float32))
    ga = cuda(rng.standard_normal((M, Na)).astype(np.float32))
    gb = cuda(rng.standard_normal((M, Nb)).astype(np.float32))
    wa, wb = ops._dw_tc2(x, ga, gb)
    close(wa, (x.double().t() @ ga.double()).float(), msg="dW a")
    close(wb, (x.double().t() @ gb.double()).float(), msg="dW b")
    wa2, wb2 = ops._dw_tc2(x, ga, gb)
    assert torch.equal(wa, wa2) and torch.equal(wb, wb2)          # run-to-run identical


@pytest.mark.parametrize("M,Ka,Kb,N", [(5000, 100, 100, 256), (777, 12, 64, 48), (40000, 128, 128, 64), (300, 200, 100, 32)])
def test_two_feature_operands_one_weight_gradient_pass(M, Ka, Kb, N):
    """ops._dw_tc_x2: (Xa^T G, Xb^T G) with G loaded and split once (kgb_linear_tc_dw_x2) against fp64; shapes beyond
    one launch (ceil32(Ka) + Kb > 256) fall back to two kgb_linear_tc_dw calls."""
    from keras_geometric_b200 import ops
    rng = np.random.default_rng(M + Ka + N)
    xa = cuda(rng.standard_normal((M, Ka)).astype(np.float32))
    xb = cuda(rng.standard_normal((M, Kb)).astype(np.float32))
    g = cuda(rng.standard_normal((M, N)).astype(np.float32))
    wa, wb = ops._dw_tc_x2(xa, xb, g)
    close(wa, (xa.double().t() @ g.double()).float(), msg="dW a")
    close(wb, (xb.double().t() @ g.double()).float(), msg="dW b")
    wa2, wb2 = ops._dw_tc_x2(xa, xb, g)
    assert torch.equal(wa, wa2) and torch.equal(wb, wb2)          # run-to-run identical


@pytest.mark.parametrize("M,K1,K2,N", [(128, 4, 4, 4), (1000, 100, 100, 256), (4097, 256, 256, 256), (3000, 48, 48, 256),
                                       (2000, 32, 64, 64), (777, 36, 8, 132), (77001, 100, 100, 256)])
def test_linear_two_operands_one_pass(M, K1, K2, N):
    """[A1 | A2] @ [W1 ; W2] + bias (ReLU) through kgb_linear_tc2 (K-concatenated operands) vs float64."""
    from keras_geometric_b200 import ops
    gen = torch.Generator(device="cuda").manual_seed(M + K1 + N)
    a1 = torch.randn((M, K1), device="cuda", generator=gen)
    a2 = torch.randn((M, K2), device="cuda", generator=gen)
    w1 = torch.randn((K1, N), device="cuda", generator=gen) * 0.2
    w2 = torch.randn((K2, N), device="cuda", generator=gen) * 0.2
    b = torch.randn(N, device="cuda", generator=gen)
    want = a1.double() @ w1.double() + a2.double() @ w2.double() + b.double()
    hi, lo = ops._split_weight_pair(w1, w2, transpose=True)
    close(ops.linear_tc2(a1, a2, hi, lo, N, bias=b), want.float(), rtol=1e-5, atol_scale=1e-5, msg="tc2")
    close(ops.linear_tc2(a1, a2, hi, lo, N, bias=b, relu=True), torch.relu(want).float(), rtol=1e-5, atol_scale=1e-5,
          msg="tc2 relu")
    # the untransposed pairing used by the backward: [G1 | G2] @ [W1^T ; W2^T]
    g1 = torch.randn((M, N), device="cuda", generator=gen)
    g2 = torch.randn((M, N), device="cuda", generator=gen)
    if K1 == K2:
        hi, lo = ops._split_weight_pair(w1, w2, transpose=False)
        want = g1.double() @ w1.double().t() + g2.double() @ w2.double().t()
        close(ops.linear_tc2(g1, g2, hi, lo, K1), want.float(), rtol=1e-5, atol_scale=1e-5, msg="tc2 dX")


# ------------------------------------------------------------------ SURVEY 8(f) next-1 / next-2
def test_golden_pooling():
    from keras_geometric_b200.layers import BatchGlobalPooling, GlobalPooling
    g = load_golden("pooling")
    for pool in ["mean", "max", "sum"]:
        close(GlobalPooling(pooling=pool)(g["x"]), g["global_" + pool], msg="global " + pool)
        x = cuda(g["x"]).requires_grad_(True)
        out = BatchGlobalPooling(pooling=pool)([x, g["batch"]])
        close(out, g["batch_" + pool], msg="batch " + pool)
        (gx,) = torch.autograd.grad((out * cuda(g["R_" + pool])).sum(), [x])
        close(gx, g["grad_" + pool], msg="batch grad " + pool)
    with pytest.raises(ValueError, match="pooling must be one of"):
        GlobalPooling(pooling="median")
    assert BatchGlobalPooling().compute_output_shape([(10, 4), (10,)]) == (None, 4)


def test_golden_batch_graphs():
    import keras_geometric_b200 as kg
    g = load_golden("batch_graphs")
    graphs = [kg.GraphData(x=g[f"x{i}"], edge_index=g[f"ei{i}"], edge_attr=g[f"ea{i}"], y=g[f"y{i}"]) for i in range(3)]
    b = kg.batch_graphs(graphs)
    np.testing.assert_array_equal(b.x.cpu().numpy(), g["bx"])
    np.testing.assert_array_equal(b.edge_index.cpu().numpy(), g["bei"])
    np.testing.assert_array_equal(b.batch.cpu().numpy(), g["bbatch"])
    np.testing.assert_array_equal(b.y.cpu().numpy(), g["by"])
    np.testing.assert_array_equal(b.edge_attr.cpu().numpy(), g["bea"])
    assert b.num_nodes == int(g["bnum_nodes"]) and b.num_edges == g["bei"].shape[1]
    assert b.to_inputs()[1].dtype == torch.int32


# ------------------------------------------------------------------------- edge cases through the layers
def test_layer_edge_cases_match_reference_conventions():
    """N == 0, E == 0, [E,2] layout, float64 / int64 inputs, isolated nodes."""
    import keras_geometric_b200 as kg
    rng = np.random.default_rng(5)
    x = rng.standard_normal((6, 4)).astype(np.float32)
    e0 = np.zeros((2, 0), np.int32)
    # N == 0
    for layer in (kg.GCNConv(3), kg.GINConv(3), kg.GATv2Conv(3, heads=2)):
        out = layer([np.zeros((0, 4), np.float32), e0])
        assert tuple(out.shape) == (0, 3 if not isinstance(layer, kg.GATv2Conv) else 6)
    # E == 0: GCN without self-loops = xW + b; with self-loops every node sees itself
    g = kg.GCNConv(3, add_self_loops=False)
    out = g([x, e0])
    close(out, x @ g.kernel.detach().cpu().numpy() + g.bias.detach().cpu().numpy(), msg="gcn no edges")
    g2 = kg.GCNConv(3)
    close(g2([x, e0]), x @ g2.kernel.detach().cpu().numpy() + g2.bias.detach().cpu().numpy(), msg="gcn loops only")
    # SAGE with no edges: aggregated = 0 -> act(lin_self(x) + b)
    s = kg.SAGEConv(3, activation=None)
    close(s([x, e0]), x @ s.lin_self.kernel.detach().cpu().numpy() + s.bias.detach().cpu().numpy(), msg="sage no edges")
    # GIN with no edges: mlp((1 + eps) x)
    gin = kg.GINConv(3, mlp_hidden=[5], eps_init=0.5)
    out_gin = gin([x, e0])  # builds the MLP
    want = ref.gin_conv(torch.from_numpy(x), torch.from_numpy(e0.astype(np.int64)), lambda h: torch.relu(
        h @ gin.mlp.layers[0].kernel.detach().cpu() + gin.mlp.layers[0].bias.detach().cpu()) @ gin.mlp.layers[1].kernel.detach().cpu()
        + gin.mlp.layers[1].bias.detach().cpu(), 0.5)
    close(out_gin, want, msg="gin no edges")
    # GATv2 without self-loops and without edges: zeros, no bias (gatv2_conv.py:204-210)
    gat = kg.GATv2Conv(3, heads=2, add_self_loops=False)
    assert float(gat([x, e0]).abs().max()) == 0.0
    # dtype / layout tolerance: float64 features, int64 edges, [E,2] layout (GCN, SAGE)
    ei = np.array([[0, 1], [1, 2], [2, 0], [5, 0]], np.int64)  # [E,2]
    a = g2([x.astype(np.float64), ei])
    b = g2([x, np.ascontiguousarray(ei.T).astype(np.int32)])
    assert a.dtype == torch.float32 and torch.equal(a, b)
    with pytest.raises(ValueError, match="edge_index must have shape"):
        g2([x, np.zeros((3, 5), np.int32)])
    # NaN / inf propagate (tests/unit/test_error_handling.py:233-258)
    xn = x.copy(); xn[2, 1] = np.nan
    assert torch.isnan(g2([xn, np.ascontiguousarray(ei.T).astype(np.int32)])).any()


def test_property_random_graphs_vs_oracle():
    """Randomised structural cases (duplicates, self-loops, isolated nodes, hubs) - CSR bit-exact, aggregations in tolerance."""
    from keras_geometric_b200 import ops
    from keras_geometric_b200.graph import GraphStructure
    rng = np.random.default_rng(123)
    for trial in range(12):
        n = int(rng.integers(1, 400))
        e = int(rng.integers(0, 3000))
        F = int(rng.choice([1, 2, 5, 8, 20, 33, 64, 130]))
        skew = rng.random() < 0.5
        dst = (rng.pareto(1.0, e) * 2).astype(np.int64) % n if skew else rng.integers(0, n, e)
        src = rng.integers(0, n, e)
        if e > 10:
            src[:5] = dst[:5]            # self-loops
            src[5:10], dst[5:10] = src[0], dst[0]  # duplicates
        ei = np.stack([src, dst]).astype(np.int32)
        loops = n if rng.random() < 0.3 else 0
        g = GraphStructure(cuda(ei), n, n, loops)
        full = ei if not loops else np.concatenate([ei, np.stack([np.arange(n), np.arange(n)]).astype(np.int32)], 1)
        rowptr, col, perm, deg = ref.stable_csr(full, n)
        np.testing.assert_array_equal(g.csr.perm.cpu().numpy(), perm)
        np.testing.assert_array_equal(g.csr.rowptr.cpu().numpy(), rowptr)
        x = np.round(rng.standard_normal((n, F)), 1).astype(np.float32)
        if full.shape[1] == 0:
            continue
        for op in ("sum", "mean", "max", "min"):
            xg = cuda(x).requires_grad_(True)
            out = ops.gather_reduce(xg, g, op)
            xo = torch.from_numpy(x).requires_grad_(True)
            want = ref.propagate(xo, torch.from_numpy(full), op)
            close(out, want, msg=f"trial {trial} {op} n={n} e={e} F={F}")
            (gx,) = torch.autograd.grad(out.sum(), [xg])
            (gw,) = torch.autograd.grad(want.sum(), [xo])
            close(gx, gw, msg=f"trial {trial} grad {op}")


def test_processed_npz_to_device(tmp_path):
    import keras_geometric_b200 as kg
    from keras_geometric_b200.data_utils import load_processed_npz, save_processed_npz
    rng = np.random.default_rng(1)
    gs = [kg.GraphData(x=rng.standard_normal((n, 3)).astype(np.float32),
                       edge_index=rng.integers(0, n, (2, e)).astype(np.int32), y=rng.integers(0, 2, n))
          for n, e in [(7, 12), (4, 5)]]
    path = str(tmp_path / "ds.npz")
    save_processed_npz(path, gs, num_classes=2)
    back, nc = load_processed_npz(path)
    assert nc == 2 and len(back) == 2
    for a, b in zip(gs, back):
        assert torch.equal(a.x, b.x) and torch.equal(a.edge_index, b.edge_index) and b.x.is_cuda
    out = kg.GCNConv(2)([back[0].x, back[0].edge_index])
    assert tuple(out.shape) == (7, 2)
