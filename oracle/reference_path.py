"""Functional CPU restatement of the reference's message-passing hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain torch on the host, float32, same
operation order as the reference so that torch autograd over these functions *is* the
reference backward.  Weights are passed explicitly (no layer objects), every function
cites the reference file:line (relative to /root/reference/src/keras_geometric/).
"""

from __future__ import annotations

import torch

from . import keras_ops as ops

# --------------------------------------------------------------------------- utils/main.py


def add_self_loops(edge_index, num_nodes: int):
    """utils/main.py:8-16 - append [0..N-1]->[0..N-1] after the real edges, unconditionally."""
    edge_index = ops.convert_to_tensor(edge_index)
    if ops.shape(edge_index)[0] != 2:
        edge_index = ops.stack([edge_index[0], edge_index[1]], axis=0)
    loop = ops.arange(0, num_nodes, dtype=edge_index.dtype)
    return ops.concatenate([edge_index, ops.stack([loop, loop], axis=0)], axis=1)


def compute_gcn_normalization(edge_index, num_nodes: int):
    """utils/main.py:20-33 - deg over *targets*, (deg+1e-12)^-0.5, w = dis[dst]*dis[src]."""
    source, target = edge_index[0], edge_index[1]
    ones = ops.ones_like(source, dtype=ops.floatx())
    degrees = ops.segment_sum(data=ones, segment_ids=target, num_segments=num_nodes)
    dis = ops.power(ops.add(degrees, 1e-12), -0.5)
    dis = ops.where(ops.isinf(dis), ops.zeros_like(dis), dis)
    return ops.multiply(ops.take(dis, target, axis=0), ops.take(dis, source, axis=0))


# ------------------------------------------------------------------ layers/aggregators.py


def aggregate(name: str, messages, target_idx, dim_size: int):
    """layers/aggregators.py: Mean :56-85, Max :99-112, Sum :126-137, Min :151-167,
    Std :182-228.  ``messages`` [E,F], ``target_idx`` [E] unsorted, result [dim_size,F]."""
    messages = ops.convert_to_tensor(messages)
    if ops.shape(messages)[0] == 0:
        return ops.zeros((dim_size, ops.shape(messages)[1]), dtype=messages.dtype)
    target_idx = ops.cast(target_idx, "int32")
    if name == "sum":
        return ops.segment_sum(messages, target_idx, dim_size)
    if name == "mean":
        one = ops.ones((ops.shape(messages)[0], 1), dtype=messages.dtype)
        degree = ops.segment_sum(one, target_idx, dim_size)
        total = ops.segment_sum(messages, target_idx, dim_size)
        degree = ops.maximum(degree, ops.convert_to_tensor(1e-8, dtype=degree.dtype))
        return total / degree
    if name == "max":
        aggr = ops.segment_max(messages, target_idx, dim_size)
        return ops.where(ops.isinf(aggr), ops.zeros_like(aggr), aggr)
    if name == "min":
        aggr = ops.negative(ops.segment_max(ops.negative(messages), target_idx, dim_size))
        return ops.where(ops.isinf(aggr), ops.zeros_like(aggr), aggr)
    if name == "std":
        one = ops.ones((ops.shape(messages)[0], 1), dtype=messages.dtype)
        count = ops.segment_sum(one, target_idx, dim_size)
        total = ops.segment_sum(messages, target_idx, dim_size)
        safe = ops.maximum(count, ops.convert_to_tensor(1e-8, dtype=count.dtype))
        mean = total / safe
        sq = ops.square(messages - ops.take(mean, target_idx, axis=0))
        var = ops.segment_sum(sq, target_idx, dim_size) / safe
        std = ops.sqrt(ops.maximum(var, ops.zeros_like(var)))
        return ops.where(count <= 1, ops.zeros_like(std), std)
    raise ValueError(f"Invalid aggregator: {name}. Available aggregators: "
                     f"['mean', 'max', 'sum', 'min', 'std']")


# -------------------------------------------------------------- layers/message_passing.py


def propagate(x, edge_index, aggregator: str = "mean", message_fn=None, update_fn=None):
    """layers/message_passing.py:147-220 with the default hooks (:47-145).

    ``x`` is a tensor or an ``(x_i, x_j)`` tuple (bipartite).  Both gathers are
    materialised like the reference does (:195-196)."""
    if isinstance(x, (list, tuple)):
        x_i, x_j = ops.convert_to_tensor(x[0]), ops.convert_to_tensor(x[1])
    else:
        x_i = x_j = ops.convert_to_tensor(x)
    edge_index = ops.convert_to_tensor(edge_index)
    n = ops.shape(x_i)[0]
    if n == 0:
        f = ops.shape(x_i)[1] if len(ops.shape(x_i)) > 1 else 1
        return ops.zeros((0, f), dtype=x_i.dtype)
    if ops.shape(edge_index)[1] == 0:
        return ops.zeros((n, ops.shape(x_i)[1]), dtype=x_i.dtype)
    src, dst = edge_index[0], edge_index[1]
    xj = ops.take(x_j, src, axis=0)
    xi = ops.take(x_i, dst, axis=0)
    msg = message_fn(xi, xj) if message_fn is not None else xj
    out = aggregate(aggregator, msg, dst, n)
    return update_fn(out, x_i) if update_fn is not None else out


def _canon_edge_index(edge_index, allow_transpose: bool):
    edge_index = ops.cast(edge_index, "int32")
    shp = ops.shape(edge_index)
    if allow_transpose and shp[0] != 2:
        if shp[1] == 2:
            edge_index = ops.transpose(edge_index)
        else:
            raise ValueError(f"edge_index must have shape [2, E] or [E, 2], but got {shp}")
    return edge_index


# --------------------------------------------------------------------- layers/gcn_conv.py


def gcn_conv(x, edge_index, kernel, bias=None, add_loops: bool = True, normalize: bool = True):
    """layers/gcn_conv.py:275-364 (call), :208-250 (message: per-edge x_j @ W, then * w_e),
    :252-272 (update: + bias).  Inference path (dropout inactive)."""
    x = ops.cast(x, "float32")
    edge_index = _canon_edge_index(edge_index, True)
    n = ops.shape(x)[0]
    out_dim = ops.shape(kernel)[1]
    if n == 0:
        return ops.zeros((0, out_dim), dtype=x.dtype)
    if add_loops:
        edge_index = add_self_loops(edge_index, n)
    e = ops.shape(edge_index)[1]
    if e == 0:
        out = ops.matmul(x, kernel)
        return ops.add(out, bias) if bias is not None else out
    w = compute_gcn_normalization(edge_index, n) if normalize else ops.ones((e,), "float32")

    def message(x_i, x_j):
        return ops.matmul(x_j, kernel) * ops.expand_dims(w, axis=1)

    def update(agg, _x):
        return ops.add(agg, bias) if bias is not None else agg

    return propagate(x, edge_index, "sum", message, update)


# -------------------------------------------------------------------- layers/sage_conv.py


def sage_conv(x, edge_index, w_neigh, w_self=None, bias=None, aggregator: str = "mean",
              activation=None, normalize: bool = False, pool_w=None, pool_b=None,
              pool_activation=None):
    """layers/sage_conv.py:351-439 (call), :300-348 (aggregate_neighbors), :259-298 (message).
    ``aggregator == 'pooling'``: Dense(pool)+act on the gathered x_j, then segment max with
    -inf -> 0 (aggregators.py:254-274)."""
    x = ops.cast(x, "float32")
    edge_index = _canon_edge_index(edge_index, True)
    n = ops.shape(x)[0]
    if ops.shape(edge_index)[1] == 0:
        f = ops.shape(pool_w)[1] if aggregator == "pooling" else ops.shape(x)[1]
        agg = ops.zeros((n, f), dtype=x.dtype)
    else:
        src, dst = edge_index[0], edge_index[1]
        xj = ops.take(x, src, axis=0)
        _xi = ops.take(x, dst, axis=0)  # materialised-but-unused, as in the reference (:332)
        if aggregator == "pooling":
            t = ops.matmul(xj, pool_w)
            if pool_b is not None:
                t = ops.add(t, pool_b)
            if pool_activation is not None:
                t = pool_activation(t)
            agg = ops.segment_max(t, ops.cast(dst, "int32"), n)
            agg = ops.where(ops.isinf(agg), ops.zeros_like(agg), agg)
        else:
            agg = aggregate(aggregator, xj, dst, n)
    out = ops.matmul(agg, w_neigh)
    if w_self is not None:
        out = ops.add(ops.matmul(x, w_self), out)
    if bias is not None:
        out = ops.add(out, bias)
    if activation is not None:
        out = activation(out)
    if normalize:
        out = ops.normalize(out, axis=-1, order=2)
    return out


# --------------------------------------------------------------------- layers/gin_conv.py


def gin_conv(x, edge_index, mlp, eps=0.0, aggregator: str = "sum"):
    """layers/gin_conv.py:228-300 (call), :195-225 (update): mlp((1+eps)*x + AGG_j x_j).
    ``mlp`` is a callable (the Dense stack)."""
    x = ops.convert_to_tensor(x)
    edge_index = ops.convert_to_tensor(edge_index)
    n = ops.shape(x)[0]
    if n == 0:
        raise ValueError("caller handles N == 0 (needs output_dim)")
    if ops.shape(edge_index)[1] == 0:
        return mlp((1 + eps) * x)
    edge_index = ops.cast(edge_index, "int32")
    return propagate(x, edge_index, aggregator, None, lambda agg, x0: mlp((1 + eps) * x0 + agg))


# ------------------------------------------------------------------- layers/gatv2_conv.py


def gatv2_conv(x, edge_index, w, att, bias=None, heads: int = 1, concat: bool = True,
               negative_slope: float = 0.2, add_loops: bool = True):
    """layers/gatv2_conv.py:129-174 (call), :176-266 (_gatv2_propagate), :268-289
    (_compute_attention), :291-311 (_softmax_by_target), :313-335 (_aggregate_messages),
    :337-352 (_final_update).  ``w`` [F, H*C] (shared), ``att`` [1, H, C]."""
    x = ops.convert_to_tensor(x)
    edge_index = ops.cast(ops.convert_to_tensor(edge_index), "int32")
    n = ops.shape(x)[0]
    c = ops.shape(att)[2]
    if add_loops:
        edge_index = add_self_loops(edge_index, n)
    e = ops.shape(edge_index)[1]
    width = heads * c if concat else c
    if n == 0:
        return ops.zeros((0, width), dtype=x.dtype)
    if e == 0:
        return ops.zeros((n, width), dtype=x.dtype)
    h = ops.reshape(ops.matmul(x, w), [n, heads, c])
    src = ops.cast(edge_index[0], "int32")
    dst = ops.cast(edge_index[1], "int32")
    h_j = ops.take(h, src, axis=0)
    h_i = ops.take(h, dst, axis=0)
    z = ops.leaky_relu(ops.add(h_i, h_j), negative_slope=negative_slope)
    s = ops.sum(ops.multiply(z, att), axis=-1)
    m = ops.segment_max(s, dst, num_segments=n)
    p = ops.exp(ops.subtract(s, ops.take(m, dst, axis=0)))
    d = ops.segment_sum(p, dst, num_segments=n)
    alpha = ops.divide(p, ops.add(ops.take(d, dst, axis=0), 1e-10))
    msg = ops.expand_dims(alpha, -1) * h_j
    agg = ops.segment_sum(ops.reshape(msg, [e, heads * c]), dst, num_segments=n)
    agg = ops.reshape(agg, [n, heads, c])
    out = ops.reshape(agg, [n, heads * c]) if concat else ops.mean(agg, axis=1)
    if bias is not None:
        out = out + bias
    return out


# ----------------------------------------------------------- graph-structure integer work


def stable_csr(edge_index, num_segments: int, by_source: bool = False):
    """Integer oracle for the COO -> CSR grouping implied by ``scatter_reduce`` on the CPU
    (edges of a segment are visited in their original order): returns
    (rowptr int64 [n+1], col int32 [E], perm int32 [E], deg int32 [n])."""
    import numpy as np

    ei = ops.convert_to_numpy(edge_index).astype(np.int64)
    key, val = (ei[0], ei[1]) if by_source else (ei[1], ei[0])
    perm = np.argsort(key, kind="stable")
    deg = np.bincount(key, minlength=num_segments).astype(np.int64)
    rowptr = np.zeros(num_segments + 1, dtype=np.int64)
    np.cumsum(deg, out=rowptr[1:])
    return rowptr, val[perm].astype(np.int32), perm.astype(np.int32), deg.astype(np.int32)


# ------------------------------------------------------- layers/pooling/global_pooling.py, utils/data_utils.py


def global_pooling(x, pooling: str = "mean"):
    """layers/pooling/global_pooling.py:57-80 - whole-tensor mean/max/sum with keepdims -> [1, F]."""
    fn = {"mean": ops.mean, "max": ops.max, "sum": ops.sum}[pooling]
    return fn(ops.convert_to_tensor(x), axis=0, keepdims=True)


def batch_global_pooling(x, batch, pooling: str = "mean"):
    """layers/pooling/global_pooling.py:200-251 - segment mean/max/sum over the batch vector;
    num_graphs = max(batch) + 1, mean divides by max(count, 1), max is the raw segment_max."""
    x, batch = ops.convert_to_tensor(x), ops.convert_to_tensor(batch)
    g = int(ops.max(batch)) + 1
    if pooling == "sum":
        return ops.segment_sum(x, batch, num_segments=g)
    if pooling == "max":
        return ops.segment_max(x, batch, num_segments=g)
    total = ops.segment_sum(x, batch, num_segments=g)
    cnt = ops.maximum(ops.segment_sum(ops.ones_like(batch, dtype=x.dtype), batch, num_segments=g), 1.0)
    return total / ops.expand_dims(cnt, axis=1)


def batch_graphs(xs, edge_indices):
    """utils/data_utils.py:139-272 (node features, shifted edge_index, batch vector of the disjoint union)."""
    import numpy as np

    off, bx, bei, bb = 0, [], [], []
    for i, (x, ei) in enumerate(zip(xs, edge_indices)):
        bx.append(np.asarray(x))
        bei.append(np.asarray(ei) + off)
        bb.append(np.full(len(x), i, np.int32))
        off += len(x)
    return np.concatenate(bx), np.concatenate(bei, axis=1), np.concatenate(bb)
