"""GINConv - same constructor, call contract and config keys as the reference
(/root/reference/src/keras_geometric/layers/gin_conv.py):  h' = MLP((1 + eps) * x + AGG_j x_j).
The neighbour sum and the (1 + eps) * x term are one fused gather-reduce launch."""
from __future__ import annotations

from typing import Any

import torch

from .. import ops
from .._compat import Dense, Dropout, Sequential, apply_mlp, initializers, to_device_tensor, value_of
from ..graph import get_graph
from .message_passing import MessagePassing


class GINConv(MessagePassing):
    """gin_conv.py:10-84."""

    def __init__(self, output_dim: int, mlp_hidden=None, aggregator: str = "sum", eps_init: float = 0.0,
                 train_eps: bool = False, use_bias: bool = True, dropout: float = 0.0,
                 kernel_initializer="glorot_uniform", bias_initializer="zeros", activation="relu",
                 **kwargs: Any) -> None:
        super().__init__(aggregator=aggregator, **kwargs)
        self.output_dim = output_dim
        self.mlp_hidden = mlp_hidden if mlp_hidden is not None else []
        self.eps_init = eps_init
        self.train_eps = train_eps
        self.use_bias = use_bias
        self.dropout_rate = dropout
        self.kernel_initializer = kernel_initializer
        self.bias_initializer = bias_initializer
        self.activation = activation
        self.mlp = None
        self.eps = None
        if self.aggregator not in ["mean", "max", "sum"]:  # gin_conv.py:80-84
            raise ValueError(f"Invalid aggregator: {self.aggregator}. Must be one of ['mean', 'max', 'sum']")

    def build(self, input_shape: Any) -> None:  # gin_conv.py:86-164
        node_shape = input_shape[0] if isinstance(input_shape, (list, tuple)) and len(input_shape) >= 1 \
            and isinstance(input_shape[0], (list, tuple)) else input_shape
        if hasattr(node_shape, "as_list"):
            node_shape = node_shape.as_list()
        if not hasattr(node_shape, "__len__") or len(node_shape) < 2:
            raise ValueError(f"Expected node features shape (N, F), got {node_shape}")
        input_dim = node_shape[1]
        if input_dim is None:
            raise ValueError("Input feature dimension cannot be None")
        input_dim = int(input_dim)
        if self.train_eps:
            self.eps = self.add_weight(name="eps", shape=(1,), initializer=initializers.Constant(self.eps_init),
                                       trainable=True)
        else:
            self.eps = self.eps_init
        mlp_layers = []
        for i, hidden_dim in enumerate(self.mlp_hidden):
            mlp_layers.append(Dense(units=hidden_dim, activation=self.activation,
                                    kernel_initializer=self.kernel_initializer,
                                    bias_initializer=self.bias_initializer, use_bias=self.use_bias,
                                    name=f"mlp_hidden_{i}"))
            if self.dropout_rate > 0:
                mlp_layers.append(Dropout(self.dropout_rate))
        mlp_layers.append(Dense(units=self.output_dim, activation=None, kernel_initializer=self.kernel_initializer,
                                bias_initializer=self.bias_initializer, use_bias=self.use_bias, name="mlp_output"))
        self.mlp = Sequential(mlp_layers, name="gin_mlp")
        self.mlp.build((None, input_dim))
        super().build(input_shape)

    def message(self, x_i, x_j, edge_attr=None, edge_index=None, size=None, **kwargs):  # gin_conv.py:166-193
        return x_j

    def update(self, aggregated, x=None):  # gin_conv.py:195-225
        if x is None:
            raise ValueError("Original node features x are required for GIN update")
        if self.mlp is None:
            raise RuntimeError("MLP not initialized. Call build() first.")
        eps = value_of(self.eps) if self.train_eps else self.eps_init
        return apply_mlp(self.mlp, (1 + eps) * x + aggregated, training=getattr(self, "_current_training", None))

    def call(self, inputs, edge_attr=None, training=None):  # gin_conv.py:228-300
        if not isinstance(inputs, (list, tuple)):
            raise ValueError("Inputs must be a list or tuple containing [x, edge_index]")
        if len(inputs) < 2:
            raise ValueError("Inputs must contain at least [x, edge_index]")
        x = to_device_tensor(inputs[0], what="x")
        if x.is_floating_point() and x.dtype != torch.float32:
            x = x.to(torch.float32)
        num_nodes = int(x.shape[0])
        self._current_training = training
        from ..dist import PartitionedGraph
        if isinstance(inputs[1], PartitionedGraph):
            return self._call_partitioned(x, inputs[1], training)
        if num_nodes == 0:
            return torch.zeros((0, self.output_dim), dtype=x.dtype, device=x.device)
        edge_index = inputs[1]
        if not (isinstance(edge_index, torch.Tensor) and edge_index.is_cuda and edge_index.dtype == torch.int32):
            edge_index = self._cast_edge_index(edge_index)
        if self.mlp is None:
            raise RuntimeError("MLP not initialized. This indicates a build issue.")
        if int(edge_index.shape[1]) == 0:  # gin_conv.py:269-280
            eps = value_of(self.eps) if self.train_eps else self.eps_init
            return apply_mlp(self.mlp, (1 + eps) * x, training=training)
        fused = (not self.train_eps and self.aggregator in ("sum", "mean")
                 and type(self).message is GINConv.message and type(self).update is GINConv.update
                 and self._uses_default("pre_aggregate", "aggregate", "post_update"))
        if fused:
            graph = get_graph(edge_index, num_nodes, num_nodes, 0)
            h = ops.gather_reduce(x, graph, self.aggregator, addend=x, addend_scale=1.0 + float(self.eps_init))
            return apply_mlp(self.mlp, h, training=training)
        return self.propagate(x=x, edge_index=edge_index, edge_attr=edge_attr, training=training)

    def _call_partitioned(self, x, pg, training=None):
        """Same layer on a 1-D node partition: ``x`` is this rank's [n_local, F] slice.  sum / mean reduce the
        local-source edges while the halo rows are in flight and add the halo part on top; max uses the
        concatenated [local | halo] source space."""
        if not self.built:
            self.build([tuple(x.shape), (2, 0)])
            self.built = True
        if self.mlp is None:
            raise RuntimeError("MLP not initialized. This indicates a build issue.")
        if not (type(self).message is GINConv.message and type(self).update is GINConv.update
                and self._uses_default("pre_aggregate", "aggregate", "post_update")):
            raise NotImplementedError("partitioned GINConv supports the default message/aggregate/update hooks")
        eps = value_of(self.eps) if self.train_eps else float(self.eps_init)
        if self.aggregator in ("sum", "mean") and pg.world > 1 and pg.any_halo:   # rank-uniform choice of the path
            g_local, g_halo, inv_deg = pg.split
            scale = (None, inv_deg) if self.aggregator == "mean" else None
            halo = pg.halo_start(x)
            if self.train_eps:
                part = ops.gather_reduce(x, g_local, "sum", weight=scale) + (1 + eps) * x
            else:
                part = ops.gather_reduce(x, g_local, "sum", weight=scale, addend=x, addend_scale=1.0 + eps)
            pg.halo_finish()
            h = ops.gather_reduce(halo, g_halo, "sum", weight=scale, addend=part)
        else:
            x_ext = pg.exchange(x)
            h = (1 + eps) * x + ops.gather_reduce(x_ext, pg.graph, self.aggregator)
        return apply_mlp(self.mlp, h, training=training)

    def compute_output_shape(self, input_shape):  # gin_conv.py:303-322
        x_shape = input_shape[0] if isinstance(input_shape, list) else (
            input_shape[0] if len(input_shape) > 0 else input_shape)
        return (x_shape[0], self.output_dim)

    def get_config(self) -> dict[str, Any]:  # gin_conv.py:324-346
        config = super().get_config()
        config.update({
            "output_dim": self.output_dim,
            "mlp_hidden": self.mlp_hidden,
            "eps_init": float(self.eps_init),
            "train_eps": self.train_eps,
            "use_bias": self.use_bias,
            "dropout": self.dropout_rate,
            "kernel_initializer": self.kernel_initializer,
            "bias_initializer": self.bias_initializer,
            "activation": self.activation,
        })
        return config

    @classmethod
    def from_config(cls, config: dict[str, Any]) -> "GINConv":
        return cls(**config)
