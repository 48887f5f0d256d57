"""SAGEConv - same constructor, call contract and config keys as the reference
(/root/reference/src/keras_geometric/layers/sage_conv.py).

out = act(lin_self(x) + lin_neigh(AGG_j x_j) + b), optional L2 normalisation.
The neighbour aggregation never materialises the [E, F] gather (sage_conv.py:331-332): it is
one fused gather-reduce launch.  For the linear aggregators (mean / sum) the layer aggregates on
whichever side of ``lin_neigh`` is narrower; when that is the output side the root term, bias and
ReLU are applied in the kernel epilogue.
"""
from __future__ import annotations

from typing import Any

import torch

from .. import ops
from .._compat import (Dense, Dropout, activations, apply_dense, constraints, initializers, regularizers,
                       to_device_tensor, value_of)
from ..graph import get_graph
from .aggregators import AggregatorFactory
from .gcn_conv import canonical_edge_index, input_dim_from_shape
from .message_passing import MessagePassing


class SAGEConv(MessagePassing):
    """sage_conv.py:10-77."""

    def __init__(self, output_dim: int, aggregator: str = "mean", normalize: bool = False,
                 root_weight: bool = True, use_bias: bool = True, activation="relu",
                 pool_activation="relu", pool_hidden_dim=None, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", kernel_regularizer=None, bias_regularizer=None,
                 kernel_constraint=None, bias_constraint=None, dropout_rate: float = 0.0, **kwargs: Any) -> None:
        valid = ["mean", "max", "sum", "min", "std", "pooling"]
        if aggregator not in valid:  # sage_conv.py:99-103
            raise ValueError(f"Invalid aggregator '{aggregator}'. Must be one of {valid}")
        super().__init__(aggregator="mean" if aggregator == "pooling" else aggregator, **kwargs)
        self.actual_aggregator = aggregator
        self.output_dim = output_dim
        self.normalize = normalize
        self.root_weight = root_weight
        self.use_bias = use_bias
        self.pool_hidden_dim = pool_hidden_dim
        self.dropout_rate = dropout_rate
        self.activation = activations.get(activation)
        self._activation_id = activation
        self.pool_activation = activations.get(pool_activation)
        self.kernel_initializer = initializers.get(kernel_initializer)
        self.bias_initializer = initializers.get(bias_initializer)
        self.kernel_regularizer = regularizers.get(kernel_regularizer)
        self.bias_regularizer = regularizers.get(bias_regularizer)
        self.kernel_constraint = constraints.get(kernel_constraint)
        self.bias_constraint = constraints.get(bias_constraint)
        self.lin_neigh = None
        self.lin_self = None
        self.pool_mlp = None
        self.bias = None

    def build(self, input_shape) -> None:  # sage_conv.py:142-236
        if input_shape is None:
            return
        input_dim = input_dim_from_shape(input_shape, "node")
        dense_kw = dict(kernel_initializer=self.kernel_initializer, kernel_regularizer=self.kernel_regularizer,
                        kernel_constraint=self.kernel_constraint, dtype=self.dtype)
        if self.actual_aggregator == "pooling":
            pool_dim = self.pool_hidden_dim or input_dim
            self.pool_mlp = Dense(units=pool_dim, activation=self.pool_activation, use_bias=self.use_bias,
                                  bias_initializer=self.bias_initializer, bias_regularizer=self.bias_regularizer,
                                  bias_constraint=self.bias_constraint, name="pool_mlp", **dense_kw)
            self.pool_mlp.build((None, input_dim))
            self.pool_mlp.built = True
        neigh_in = (self.pool_hidden_dim or input_dim) if self.actual_aggregator == "pooling" else input_dim
        self.lin_neigh = Dense(units=self.output_dim, use_bias=False, name="linear_neigh", **dense_kw)
        self.lin_neigh.build((None, neigh_in))
        self.lin_neigh.built = True
        if self.root_weight:
            self.lin_self = Dense(units=self.output_dim, use_bias=False, name="linear_self", **dense_kw)
            self.lin_self.build((None, input_dim))
            self.lin_self.built = True
        if self.use_bias:
            self.bias = self.add_weight(shape=(self.output_dim,), initializer=self.bias_initializer,
                                        regularizer=self.bias_regularizer, constraint=self.bias_constraint,
                                        name="bias", trainable=True, dtype=self.dtype)
        super().build(input_shape)

    def compute_output_shape(self, input_shape) -> tuple:  # sage_conv.py:238-257
        if isinstance(input_shape, (list, tuple)) and len(input_shape) >= 1:
            node_shape = input_shape[0]
            if hasattr(node_shape, "as_list"):
                node_shape = node_shape.as_list()
            batch = node_shape[0] if isinstance(node_shape, (list, tuple)) else None
            return (batch, self.output_dim)
        return (None, self.output_dim)

    def message(self, x_i, x_j, edge_attr=None, edge_index=None, size=None, **kwargs):  # sage_conv.py:259-298
        training = kwargs.get("training", None)
        if self.dropout_rate > 0 and training:
            x_j = Dropout(self.dropout_rate)(x_j, training=training)
        return x_j

    def aggregate_neighbors(self, x, edge_index, num_nodes: int, training=None):  # sage_conv.py:300-348
        x = to_device_tensor(x, torch.float32, "node features")
        agg_name = self.actual_aggregator
        if int(edge_index.shape[1]) == 0:
            width = int(self.pool_mlp.units) if (agg_name == "pooling" and self.pool_mlp is not None) else int(x.shape[1])
            return torch.zeros((num_nodes, width), dtype=x.dtype, device=x.device)
        graph = get_graph(edge_index, num_nodes, num_nodes, 0)
        dropping = self.dropout_rate > 0 and bool(training)
        default_hooks = type(self).message is SAGEConv.message and self._uses_default("aggregate", "update")
        if default_hooks and dropping and agg_name in ("sum", "mean"):
            # per-edge dropout of x_j (sage_conv.py:295-297) generated inside the gather: no [E, F] messages
            return ops.gather_reduce(x, graph, agg_name, dropout=float(self.dropout_rate))
        if default_hooks and not dropping:
            if agg_name in ("sum", "mean", "max", "min"):
                return ops.gather_reduce(x, graph, agg_name)
            if agg_name == "std":
                return ops.gather_std(x, graph)          # fused with the gather: no [E, F] messages
            if agg_name == "pooling":
                # Dense+act commute with the row gather: transform once per node, then fused max
                return ops.gather_reduce(apply_dense(self.pool_mlp, x), graph, "max")
        # generic path: per-edge dropout, std, user-overridden hooks
        x_j = ops.take_rows(x, graph, "src")
        x_i = ops.take_rows(x, graph, "dst")
        messages = self.message(x_i, x_j, training=training)
        from .aggregators import graph_hint
        target_idx = graph.full_edge_index()[1]
        with graph_hint(target_idx, graph):
            if agg_name == "pooling":
                aggregated = AggregatorFactory.create_pooling(self.pool_mlp).aggregate(messages, target_idx, num_nodes)
            else:
                aggregated = MessagePassing.aggregate(self, messages, target_idx, num_nodes=num_nodes)
        return self.update(aggregated)

    def call(self, inputs, training=None, mask=None):  # sage_conv.py:351-439
        if not isinstance(inputs, (list, tuple)) or len(inputs) < 2:
            raise ValueError("SAGEConv expects inputs to be a list/tuple of [node_features, edge_index]")
        x = to_device_tensor(inputs[0], torch.float32, "node features")
        src_obj = inputs[1]
        from ..dist import PartitionedGraph
        if isinstance(src_obj, PartitionedGraph):
            return self._call_partitioned(x, src_obj, training)
        if isinstance(src_obj, torch.Tensor) and src_obj.is_cuda and src_obj.dtype == torch.int32 \
                and src_obj.dim() == 2 and src_obj.shape[0] == 2:
            edge_index = src_obj
        else:
            key = (id(src_obj), getattr(src_obj, "_version", None))
            if self._cached_edge_idx is None or self._cached_edge_idx_hash != key:
                self._cached_edge_idx = canonical_edge_index(src_obj, True)
                self._cached_edge_idx_hash = key
                self._cached_edge_src = src_obj
            edge_index = self._cached_edge_idx
        if not self.built:
            self.build([tuple(x.shape), tuple(edge_index.shape)])
            self.built = True
        num_nodes = int(x.shape[0])
        w_neigh = value_of(self.lin_neigh.kernel)
        w_self = value_of(self.lin_self.kernel) if (self.root_weight and self.lin_self is not None) else None
        bias = value_of(self.bias) if (self.use_bias and self.bias is not None) else None
        dropping = self.dropout_rate > 0 and bool(training)
        act_is_relu = self._activation_id == "relu"
        act_is_none = self._activation_id in (None, "linear")

        # linear aggregator and a narrower output side: aggregate after lin_neigh, fuse the rest
        reorder = (self.actual_aggregator in ("mean", "sum") and not dropping and num_nodes > 0
                   and int(edge_index.shape[1]) > 0 and self.output_dim < int(x.shape[1])
                   and (act_is_relu or act_is_none)
                   and type(self).message is SAGEConv.message and self._uses_default("aggregate", "update"))
        one_node = (not reorder and self.actual_aggregator in ("mean", "sum") and not dropping and num_nodes > 0
                    and int(edge_index.shape[1]) > 0 and w_self is not None and (act_is_relu or act_is_none)
                    and type(self).message is SAGEConv.message and self._uses_default("aggregate", "update")
                    and ops.sage_layer_ok(x, w_neigh, w_self))
        if reorder:
            graph = get_graph(edge_index, num_nodes, num_nodes, 0)
            out = self._aggregate_after_transform(x, graph, w_neigh, w_self, bias, act_is_relu)
        elif one_node:
            # aggregate -> two accumulating GEMMs (bias + ReLU in the epilogue) recorded as one autograd node
            out = ops.sage_layer(x, w_neigh, w_self, bias, get_graph(edge_index, num_nodes, num_nodes, 0),
                                 self.actual_aggregator, act_is_relu)
        else:
            aggregated = self.aggregate_neighbors(x, edge_index, num_nodes, training=training)
            out = self._dense_update(aggregated, x, w_neigh, w_self, bias, dropping)
        if self.normalize:  # ops.normalize(axis=-1, order=2): x / max(||x||, 1e-12), one fused row pass
            out = ops.l2_normalize(out, 1e-12)
        return out

    def _aggregate_after_transform(self, x, graph, w_neigh, w_self, bias, act_is_relu, exchange=None):
        """mean/sum commute with lin_neigh: transform first (narrower rows to gather), then ONE kernel does
        aggregate + root term + bias + ReLU.  The width is padded to a multiple of 4 floats so the gather
        uses 128-bit loads; the pad columns are exact zeros and are sliced off."""
        fout = int(w_neigh.shape[1])
        pad = (-fout) % 4
        if pad:
            w_neigh = torch.nn.functional.pad(w_neigh, (0, pad))
            w_self = torch.nn.functional.pad(w_self, (0, pad)) if w_self is not None else None
            bias = torch.nn.functional.pad(bias, (0, pad)) if bias is not None else None
        if exchange is None and w_self is not None:
            z, root = ops.linear_pair(x, w_neigh, w_self)
        else:
            z = ops.linear(x, w_neigh)
            if exchange is not None:
                z = exchange[0](z)          # halo all-to-all left in flight ...
            root = ops.linear(x, w_self) if w_self is not None else None
            if exchange is not None:
                exchange[1]()               # ... while the root transform runs
        out = ops.gather_reduce(z, graph, self.actual_aggregator, addend=root, bias=bias,
                                act="relu" if act_is_relu else None)
        return out[:, :fout] if pad else out

    def _dense_update(self, aggregated, x, w_neigh, w_self, bias, dropping=False):
        """act(lin_self(x) + lin_neigh(agg) + b) (sage_conv.py:411-433) as two accumulating GEMMs; the second one
        carries addend, bias and ReLU in its epilogue."""
        fuse_relu = self._activation_id == "relu"
        if w_self is None:
            out = ops.linear(aggregated, w_neigh, bias=bias, act="relu" if fuse_relu else None)
        else:
            out = ops.linear(aggregated, w_neigh)
            x_self = Dropout(self.dropout_rate)(x, training=True) if dropping else x
            out = ops.linear(x_self, w_self, addend=out, bias=bias, act="relu" if fuse_relu else None)
        if self.activation is not None and not fuse_relu:
            out = self.activation(out)
        return out

    def _call_partitioned(self, x, pg, training=None):
        """Same layer on a 1-D node partition: ``x`` is this rank's [n_local, F] slice, halo rows are
        exchanged (at the narrower of the two widths for linear aggregators) before the aggregation."""
        if not self.built:
            self.build([tuple(x.shape), (2, 0)])
            self.built = True
        dropping = self.dropout_rate > 0 and bool(training)     # the same on every rank: the path choice stays uniform
        w_neigh = value_of(self.lin_neigh.kernel)
        w_self = value_of(self.lin_self.kernel) if (self.root_weight and self.lin_self is not None) else None
        bias = value_of(self.bias) if (self.use_bias and self.bias is not None) else None
        act_is_relu = self._activation_id == "relu"
        act_is_none = self._activation_id in (None, "linear")
        linear_agg = self.actual_aggregator in ("mean", "sum")
        fuse_act = act_is_relu or act_is_none
        if dropping:
            # training with dropout (sage_conv.py:295-297, 411-433): the [local | halo] source space.  mean / sum drop
            # the gathered rows inside the fused gather (Philox on the rank-local edge id, regenerated by the
            # transposed pass); the other aggregators drop explicit per-edge messages like the reference
            src = x
            x_ext = pg.exchange(src)
            agg_name = self.actual_aggregator
            if agg_name in ("mean", "sum"):
                aggregated = ops.gather_reduce(x_ext, pg.graph, agg_name, dropout=float(self.dropout_rate))
            else:
                x_j = Dropout(self.dropout_rate)(ops.take_rows(x_ext, pg.graph, "src"), training=True)
                if agg_name == "pooling":
                    x_j = apply_dense(self.pool_mlp, x_j)
                if agg_name == "std":
                    aggregated = ops.segment_std(x_j, pg.graph)
                else:
                    aggregated = ops.segment_reduce(x_j, pg.graph, "max" if agg_name == "pooling" else agg_name)
            out = self._dense_update(aggregated, x, w_neigh, w_self, bias, dropping=True)
            if self.normalize:
                out = ops.l2_normalize(out, 1e-12)
            return out
        if linear_agg and pg.world > 1 and pg.any_halo:   # rank-uniform choice of the path (collectives stay matched)
            # local-source edges are reduced while the halo rows are in flight; the halo part is added on top
            g_local, g_halo, inv_deg = pg.split
            scale = (None, inv_deg) if self.actual_aggregator == "mean" else None
            act = "relu" if act_is_relu else None
            reorder = self.output_dim < int(x.shape[1]) and fuse_act
            fout = int(w_neigh.shape[1])
            pad = ((-fout) % 4) if reorder else 0
            if pad:
                w_neigh = torch.nn.functional.pad(w_neigh, (0, pad))
                w_self = torch.nn.functional.pad(w_self, (0, pad)) if w_self is not None else None
                bias = torch.nn.functional.pad(bias, (0, pad)) if bias is not None else None
            if fuse_act and w_self is not None and ops.sage_layer_ok(x, w_neigh, w_self):
                # one autograd node that schedules the exchange / compute overlap itself (both directions)
                out = ops.sage_partitioned(x, w_neigh, w_self, bias, pg, self.actual_aggregator, act_is_relu, reorder)
                out = out[:, :fout] if pad else out
            elif reorder:   # aggregate after lin_neigh (narrower rows travel)
                z = ops.linear(x, w_neigh)
                halo = pg.halo_start(z)
                root = ops.linear(x, w_self) if w_self is not None else None
                part = ops.gather_reduce(z, g_local, "sum", weight=scale, addend=root)
                pg.halo_finish()
                out = ops.gather_reduce(halo, g_halo, "sum", weight=scale, addend=part, bias=bias, act=act)
                out = out[:, :fout] if pad else out
            else:
                halo = pg.halo_start(x)
                root = ops.linear(x, w_self) if w_self is not None else None
                part = ops.gather_reduce(x, g_local, "sum", weight=scale)
                pg.halo_finish()
                aggregated = ops.gather_reduce(halo, g_halo, "sum", weight=scale, addend=part)
                out = ops.linear(aggregated, w_neigh, addend=root, bias=bias, act=act if fuse_act else None)
                if self.activation is not None and not fuse_act:
                    out = self.activation(out)
        elif linear_agg and self.output_dim < int(x.shape[1]) and fuse_act:
            out = self._aggregate_after_transform(x, pg.graph, w_neigh, w_self, bias, act_is_relu,
                                                   exchange=(pg.exchange_start, pg.exchange_finish))
        else:
            # max / min / std / pooling: the [local | halo] source space (their backward needs the gathered rows)
            src = apply_dense(self.pool_mlp, x) if self.actual_aggregator == "pooling" else x
            x_ext = pg.exchange_start(src)                    # exchange in flight ...
            root = ops.linear(x, w_self) if w_self is not None else None   # ... behind the root transform
            pg.exchange_finish()
            if self.actual_aggregator == "std":
                aggregated = ops.gather_std(x_ext, pg.graph)
            else:
                aggregated = ops.gather_reduce(x_ext, pg.graph,
                                               "max" if self.actual_aggregator == "pooling" else self.actual_aggregator)
            out = ops.linear(aggregated, w_neigh, addend=root, bias=bias, act="relu" if act_is_relu else None)
            if self.activation is not None and not act_is_relu:
                out = self.activation(out)
        if self.normalize:  # ops.normalize(axis=-1, order=2): x / max(||x||, 1e-12), one fused row pass
            out = ops.l2_normalize(out, 1e-12)
        return out

    def get_config(self) -> dict[str, Any]:  # sage_conv.py:441-473
        config = super().get_config()
        config.update({
            "output_dim": self.output_dim,
            "normalize": self.normalize,
            "root_weight": self.root_weight,
            "use_bias": self.use_bias,
            "activation": activations.serialize(self.activation),
            "pool_activation": activations.serialize(self.pool_activation),
            "pool_hidden_dim": self.pool_hidden_dim,
            "kernel_initializer": initializers.serialize(self.kernel_initializer),
            "bias_initializer": initializers.serialize(self.bias_initializer),
            "kernel_regularizer": regularizers.serialize(self.kernel_regularizer),
            "bias_regularizer": regularizers.serialize(self.bias_regularizer),
            "kernel_constraint": constraints.serialize(self.kernel_constraint),
            "bias_constraint": constraints.serialize(self.bias_constraint),
            "dropout_rate": self.dropout_rate,
        })
        config["aggregator"] = self.actual_aggregator
        return config

    @classmethod
    def from_config(cls, config: dict[str, Any]) -> "SAGEConv":  # sage_conv.py:475-509
        config = config.copy()
        config["activation"] = activations.deserialize(config.get("activation"))
        config["pool_activation"] = activations.deserialize(config.get("pool_activation"))
        config["kernel_initializer"] = initializers.deserialize(config.get("kernel_initializer", "glorot_uniform"))
        config["bias_initializer"] = initializers.deserialize(config.get("bias_initializer", "zeros"))
        config["kernel_regularizer"] = regularizers.deserialize(config.get("kernel_regularizer"))
        config["bias_regularizer"] = regularizers.deserialize(config.get("bias_regularizer"))
        config["kernel_constraint"] = constraints.deserialize(config.get("kernel_constraint"))
        config["bias_constraint"] = constraints.deserialize(config.get("bias_constraint"))
        return cls(**config)
