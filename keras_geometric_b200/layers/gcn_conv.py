"""GCNConv - same constructor, call contract and config keys as the reference
(/root/reference/src/keras_geometric/layers/gcn_conv.py).

The reference transforms every *edge* (``matmul(x_j, W)`` on [E, F_in], gcn_conv.py:233); here
``X @ W`` is computed once per node and the normalised aggregation
``out_i = sum_e dis_i * dis_src(e) * (XW)[src(e)] + b`` runs in one fused gather-reduce launch
(self-loops are appended by the CSR builder, normalisation and bias live in the kernel).
"""
from __future__ import annotations

from typing import Any

import torch

from .. import ops
from .._compat import Dropout, constraints, initializers, regularizers, to_device_tensor, value_of
from ..graph import get_graph
from .message_passing import MessagePassing


def canonical_edge_index(edge_index, allow_transpose: bool = True) -> torch.Tensor:
    """int32 [2, E]; an [E, 2] input is transposed (gcn_conv.py:307-318, sage_conv.py:382-393)."""
    edge_index = to_device_tensor(edge_index, what="edge_index")
    if edge_index.dtype != torch.int32:
        edge_index = edge_index.to(torch.int32)
    if edge_index.dim() != 2:
        raise ValueError(f"edge_index must have shape [2, E] or [E, 2], but got {tuple(edge_index.shape)}")
    if edge_index.shape[0] != 2:
        if allow_transpose and edge_index.shape[1] == 2:
            edge_index = edge_index.t().contiguous()
        else:
            raise ValueError(f"edge_index must have shape [2, E] or [E, 2], but got {tuple(edge_index.shape)}")
    return edge_index


def input_dim_from_shape(input_shape, what="node features") -> int:
    if not isinstance(input_shape, (list, tuple)) or len(input_shape) < 2:
        raise ValueError("Expected input_shape to be a list/tuple with at least 2 elements "
                         f"[(node_features_shape), (edge_index_shape)], but got {input_shape}")
    node_shape = input_shape[0]
    if hasattr(node_shape, "as_list"):
        node_shape = node_shape.as_list()
    if not isinstance(node_shape, (list, tuple)) or len(node_shape) < 2:
        raise ValueError(f"Expected {what} shape to be (N, F), but got {node_shape}")
    input_dim = node_shape[-1]
    if input_dim is None or int(input_dim) <= 0:
        raise ValueError(f"Input dimension must be a positive integer, but got {input_dim}")
    return int(input_dim)


class GCNConv(MessagePassing):
    """H' = D^-1/2 (A + I) D^-1/2 X W + b   (gcn_conv.py:11-61)."""

    def __init__(self, output_dim: int, use_bias: bool = True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None,
                 kernel_constraint=None, bias_constraint=None, add_self_loops: bool = True,
                 normalize: bool = True, dropout_rate: float = 0.0, **kwargs: Any) -> None:
        kwargs["aggregator"] = "sum"  # gcn_conv.py:78-80
        super().__init__(**kwargs)
        self.output_dim = output_dim
        self.use_bias = use_bias
        self.add_self_loops = add_self_loops
        self.normalize = normalize
        self.dropout_rate = dropout_rate
        self.kernel_initializer = initializers.get(kernel_initializer)
        self.bias_initializer = initializers.get(bias_initializer)
        self.kernel_regularizer = regularizers.get(kernel_regularizer)
        self.bias_regularizer = regularizers.get(bias_regularizer)
        self.kernel_constraint = constraints.get(kernel_constraint)
        self.bias_constraint = constraints.get(bias_constraint)
        self.kernel = None
        self.bias = None
        self._current_edge_weights = None
        self._current_training = None

    def build(self, input_shape) -> None:  # gcn_conv.py:105-185
        if input_shape is None:
            return
        input_dim = input_dim_from_shape(input_shape)
        self.kernel = self.add_weight(shape=(input_dim, self.output_dim), initializer=self.kernel_initializer,
                                      regularizer=self.kernel_regularizer, constraint=self.kernel_constraint,
                                      name="kernel", trainable=True, dtype=self.dtype)
        if self.use_bias:
            self.bias = self.add_weight(shape=(self.output_dim,), initializer=self.bias_initializer,
                                        regularizer=self.bias_regularizer, constraint=self.bias_constraint,
                                        name="bias", trainable=True, dtype=self.dtype)
        else:
            self.bias = None
        super().build(input_shape)

    def compute_output_shape(self, input_shape) -> tuple:  # gcn_conv.py:187-206
        if isinstance(input_shape, (list, tuple)) and len(input_shape) >= 1:
            node_shape = input_shape[0]
            if hasattr(node_shape, "as_list"):
                node_shape = node_shape.as_list()
            batch = node_shape[0] if isinstance(node_shape, (list, tuple)) else None
            return (batch, self.output_dim)
        return (None, self.output_dim)

    # generic-path hooks, kept for API compatibility (gcn_conv.py:208-272)
    def message(self, x_i, x_j, edge_attr=None, edge_index=None, size=None, **kwargs):
        x_j_t = ops.linear(x_j, value_of(self.kernel))
        if self.dropout_rate > 0 and self._current_training:
            x_j_t = Dropout(self.dropout_rate)(x_j_t, training=self._current_training)
        if edge_attr is not None:
            return x_j_t * edge_attr.unsqueeze(1)
        return x_j_t

    def update(self, aggregated, x=None):
        if self.use_bias and self.bias is not None:
            return aggregated + value_of(self.bias)
        return aggregated

    def call(self, inputs, training=None, mask=None):  # gcn_conv.py:275-364
        if not isinstance(inputs, (list, tuple)) or len(inputs) < 2:
            raise ValueError("GCNConv expects inputs to be a list/tuple of [node_features, edge_index]")
        x = to_device_tensor(inputs[0], torch.float32, "node features")
        src_obj = inputs[1]
        from ..dist import PartitionedGraph
        if isinstance(src_obj, PartitionedGraph):
            return self._call_partitioned(x, src_obj, training)
        edge_index = src_obj if isinstance(src_obj, torch.Tensor) and src_obj.is_cuda and src_obj.dtype == torch.int32 \
            and src_obj.dim() == 2 and src_obj.shape[0] == 2 else self._canonical_cached(src_obj)
        num_nodes = int(x.shape[0])
        if num_nodes == 0:
            return torch.zeros((0, self.output_dim), dtype=x.dtype, device=x.device)
        kernel, bias = value_of(self.kernel), value_of(self.bias) if self.use_bias else None
        n_loops = num_nodes if self.add_self_loops else 0
        num_edges = int(edge_index.shape[1]) + n_loops
        dropping = self.dropout_rate > 0 and bool(training)
        if num_edges == 0:  # gcn_conv.py:332-347
            out = ops.linear(x, kernel)
            if dropping:
                out = Dropout(self.dropout_rate)(out, training=training)
            return out + bias if bias is not None else out
        graph = get_graph(edge_index, num_nodes, num_nodes, n_loops)
        self._current_training = training
        fast = type(self).message is GCNConv.message and type(self).update is GCNConv.update
        if fast:
            h = ops.linear(x, kernel)  # dense transform once per node on the tensor cores (K8)
            # training: the reference drops the per-edge transformed messages element-wise (gcn_conv.py:238-242);
            # the same dropout is generated inside the gather (Philox on the edge id), no [E, F] tensor exists
            return ops.gather_reduce(h, graph, "sum", weight="gcn" if self.normalize else None, bias=bias,
                                     dropout=float(self.dropout_rate) if dropping else 0.0)
        # per-edge path: dropout on the transformed messages (gcn_conv.py:238-242) or user overrides
        w = graph.gcn_norm()[1] if self.normalize else torch.ones(graph.nnz, dtype=x.dtype, device=x.device)
        self._current_edge_weights = w
        x_j = ops.take_rows(x, graph, "src")
        x_i = ops.take_rows(x, graph, "dst")
        messages = self.pre_aggregate(self.message(x_i, x_j, edge_attr=w, edge_index=graph.full_edge_index(),
                                                   size=(num_nodes, num_nodes), training=training))
        aggregated = ops.segment_reduce(messages, graph, "sum")
        return self.post_update(x, self.update(aggregated, x=x))

    def _call_partitioned(self, x, pg, training=None):
        """1-D node partition: transform locally, exchange the (narrow) transformed halo rows, aggregate.
        ``pg`` must have been built with ``n_loops_local=self.add_self_loops``."""
        if bool(pg._n_loops) != bool(self.add_self_loops):
            raise ValueError("PartitionedGraph(n_loops_local=...) must match GCNConv(add_self_loops=...)")
        kernel, bias = value_of(self.kernel), value_of(self.bias) if self.use_bias else None
        h = ops.linear(x, kernel)
        dropping = self.dropout_rate > 0 and bool(training)     # the same on every rank: the path choice stays uniform
        if pg.world > 1 and pg.any_halo and not dropping:
            # local-source edges (and the self-loops) are reduced while the transformed halo rows are in flight
            return ops.aggregate_partitioned(h, pg, "gcn" if self.normalize else "sum", bias=bias)
        # training with message dropout (gcn_conv.py:238-242): the [local | halo] source space, where the fused gather
        # draws the element-wise Philox mask from the rank-local edge id (regenerated by the transposed pass)
        h_ext = pg.exchange(h)
        weight = None
        if self.normalize:
            dis_local, dis_ext = pg.gcn_dis_ext()
            weight = (dis_ext, dis_local)
        return ops.gather_reduce(h_ext, pg.graph, "sum", weight=weight, bias=bias,
                                 dropout=float(self.dropout_rate) if dropping else 0.0)

    def _canonical_cached(self, edge_index):
        key = (id(edge_index), getattr(edge_index, "_version", None))
        if self._cached_edge_idx is None or self._cached_edge_idx_hash != key:
            self._cached_edge_idx = canonical_edge_index(edge_index, True)
            self._cached_edge_idx_hash = key
            self._cached_edge_src = edge_index
        return self._cached_edge_idx

    def get_config(self) -> dict[str, Any]:  # gcn_conv.py:366-389
        config = super().get_config()
        config.update({
            "output_dim": self.output_dim,
            "use_bias": self.use_bias,
            "kernel_initializer": initializers.serialize(self.kernel_initializer),
            "bias_initializer": initializers.serialize(self.bias_initializer),
            "kernel_regularizer": regularizers.serialize(self.kernel_regularizer),
            "bias_regularizer": regularizers.serialize(self.bias_regularizer),
            "kernel_constraint": constraints.serialize(self.kernel_constraint),
            "bias_constraint": constraints.serialize(self.bias_constraint),
            "add_self_loops": self.add_self_loops,
            "normalize": self.normalize,
            "dropout_rate": self.dropout_rate,
        })
        return config

    @classmethod
    def from_config(cls, config: dict[str, Any]) -> "GCNConv":  # gcn_conv.py:391-426
        config = config.copy()
        config["kernel_initializer"] = initializers.deserialize(config.get("kernel_initializer", "glorot_uniform"))
        config["bias_initializer"] = initializers.deserialize(config.get("bias_initializer", "zeros"))
        config["kernel_regularizer"] = regularizers.deserialize(config.get("kernel_regularizer"))
        config["bias_regularizer"] = regularizers.deserialize(config.get("bias_regularizer"))
        config["kernel_constraint"] = constraints.deserialize(config.get("kernel_constraint"))
        config["bias_constraint"] = constraints.deserialize(config.get("bias_constraint"))
        config.pop("aggregator", None)
        return cls(**config)
