"""MessagePassing base layer - same contract as the reference
(/root/reference/src/keras_geometric/layers/message_passing.py), B200-native underneath.

When ``message`` / ``pre_aggregate`` / ``aggregate`` are not overridden (and there are no edge
attributes) ``propagate`` never materialises an [E, F] tensor: it runs the fused
gather-reduce kernel over the cached CSR.  Otherwise the per-edge tensors are produced by
``kgb_gather_rows`` and reduced by the generic segment kernel, so user subclasses keep working.
"""
from __future__ import annotations

from typing import Any

import torch

from .. import ops
from .._compat import Layer, to_device_tensor
from ..graph import get_graph
from .aggregators import Aggregator, AggregatorFactory, graph_hint

_FUSED_OPS = ("sum", "mean", "max", "min")


class MessagePassing(Layer):
    """message_passing.py:9-45.  ``aggregator`` in ['mean', 'max', 'sum', 'min', 'std']."""

    def __init__(self, aggregator: str = "mean", **kwargs) -> None:
        super().__init__(**kwargs)
        self.aggregator_name: str = aggregator
        self._aggregator: Aggregator = AggregatorFactory.create(aggregator)
        self.aggregator: str = aggregator
        self.supported_aggregators = AggregatorFactory.get_available_aggregators()
        self._cached_edge_idx = None
        self._cached_edge_idx_hash = None
        self.message_kwargs: dict[str, Any] = {}

    # ---- overridable hooks (message_passing.py:47-145) ---------------------------------------
    def message(self, x_i, x_j, edge_attr=None, edge_index=None, size=None, **kwargs):
        if edge_attr is not None:
            return torch.cat([x_j, to_device_tensor(edge_attr, torch.float32, "edge_attr")], dim=-1)
        return x_j

    def aggregate(self, messages, target_idx, num_nodes: int, dim_size: int | None = None):
        if dim_size is None:
            dim_size = num_nodes
        return self._aggregator.aggregate(messages, target_idx, dim_size)

    def update(self, aggregated, x=None):
        return aggregated

    def pre_aggregate(self, messages):
        return messages

    def post_update(self, x, x_updated):
        return x_updated

    # ---- helpers -----------------------------------------------------------------------------
    def _cast_edge_index(self, edge_index):
        """int32 cast with the reference's identity-keyed cache (message_passing.py:256-266); the
        cache additionally checks the tensor version so in-place edits are not missed."""
        key = (id(edge_index), getattr(edge_index, "_version", None))
        if self._cached_edge_idx is None or self._cached_edge_idx_hash != key:
            self._cached_edge_idx = to_device_tensor(edge_index, torch.int32, "edge_index")
            self._cached_edge_idx_hash = key
            self._cached_edge_src = edge_index  # keeps id() from being recycled while cached
        return self._cached_edge_idx

    def _uses_default(self, *names) -> bool:
        return all(getattr(type(self), n) is getattr(MessagePassing, n) for n in names)

    # ---- propagate (message_passing.py:147-220) ------------------------------------------------
    def propagate(self, x, edge_index, edge_attr=None, size=None, **kwargs):
        if isinstance(x, (list, tuple)):
            x_i = to_device_tensor(x[0], what="x_i")
            x_j = to_device_tensor(x[1], what="x_j")
        else:
            x_i = x_j = to_device_tensor(x, what="x")
        if x_i.is_floating_point() and x_i.dtype != torch.float32:
            x_i, x_j = x_i.to(torch.float32), x_j.to(torch.float32)
        size = (int(x_i.shape[0]), int(x_j.shape[0]))
        num_nodes = size[0]
        if num_nodes == 0:
            feature_dim = int(x_i.shape[1]) if x_i.dim() > 1 else 1
            return torch.zeros((0, feature_dim), dtype=x_i.dtype, device=x_i.device)
        edge_index = to_device_tensor(edge_index, what="edge_index")
        if edge_index.dtype != torch.int32:
            edge_index = edge_index.to(torch.int32)
        if int(edge_index.shape[1]) == 0:
            return torch.zeros((num_nodes, int(x_i.shape[1])), dtype=x_i.dtype, device=x_i.device)
        graph = get_graph(edge_index, size[0], size[1], 0)

        fused = (edge_attr is None and self.aggregator in _FUSED_OPS
                 and self._uses_default("message", "pre_aggregate", "aggregate"))
        if fused:
            aggregated = ops.gather_reduce(x_j, graph, self.aggregator)
        else:
            x_j_g = ops.take_rows(x_j, graph, "src")
            x_i_g = ops.take_rows(x_i, graph, "dst")
            messages = self.message(x_i_g, x_j_g, edge_attr=edge_attr, edge_index=edge_index, size=size, **kwargs)
            messages = self.pre_aggregate(messages)
            target_idx = edge_index[1]
            with graph_hint(target_idx, graph):
                aggregated = self.aggregate(messages, target_idx, num_nodes, dim_size=size[0])
        updated = self.update(aggregated, x=x_i)
        return self.post_update(x_i, updated)

    # ---- call (message_passing.py:223-275) -----------------------------------------------------
    def call(self, inputs, edge_attr=None, training=None):
        if not isinstance(inputs, (list, tuple)):
            raise ValueError("Inputs must be a list or tuple containing [x, edge_index]")
        if len(inputs) < 2:
            raise ValueError("Inputs must contain at least [x, edge_index]")
        x, edge_index = inputs[0], inputs[1]
        if len(inputs) >= 3 and inputs[2] is not None:
            edge_attr = inputs[2]
        edge_index = self._cast_edge_index(edge_index)
        self.message_kwargs = {}
        return self.propagate(x=x, edge_index=edge_index, edge_attr=edge_attr, training=training)

    def compute_output_shape(self, input_shape):
        """message_passing.py:277-296 (returns the node-feature shape)."""
        if isinstance(input_shape, list):
            return input_shape[0]
        return input_shape[0] if len(input_shape) > 0 else input_shape

    def get_config(self) -> dict[str, Any]:
        config = super().get_config()
        config.update({"aggregator": self.aggregator})
        return config

    @classmethod
    def from_config(cls, config: dict[str, Any]):
        return cls(**config)
