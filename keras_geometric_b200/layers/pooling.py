"""Graph read-outs with the reference's interface
(/root/reference/src/keras_geometric/layers/pooling/global_pooling.py).  SURVEY 8(f) "next-1": the step right
after the conv stack; same segment kernels as the aggregators (the ``batch`` vector is the segment id)."""
from __future__ import annotations

from typing import Any

import torch

from .. import ops
from .._compat import Layer, to_device_tensor
from .aggregators import _structure_for

_POOLS = ["mean", "max", "sum"]


class GlobalPooling(Layer):
    """Whole-graph read-out [N, F] -> [1, F] (global_pooling.py:9-137)."""

    def __init__(self, pooling: str = "mean", **kwargs) -> None:
        super().__init__(**kwargs)
        if pooling not in _POOLS:
            raise ValueError(f"pooling must be one of ['mean', 'max', 'sum'], got {pooling}")
        self.pooling = pooling

    def call(self, inputs: Any, **kwargs: Any):
        x = to_device_tensor(inputs, what="node features")
        if x.is_floating_point() and x.dtype != torch.float32:
            x = x.to(torch.float32)
        n = int(x.shape[0])
        seg = torch.zeros(n, dtype=torch.int32, device=x.device)
        graph, _ = _structure_for(seg, 1)
        return ops.segment_reduce(x, graph, "max_raw" if self.pooling == "max" else self.pooling)

    def compute_output_shape(self, input_shape):
        if len(input_shape) != 2:
            raise ValueError("Expected input shape to be 2D (num_nodes, num_features), "
                             f"got {len(input_shape)}D")
        return (1, input_shape[1])

    def get_config(self) -> dict[str, Any]:
        config = super().get_config()
        config.update({"pooling": self.pooling})
        return config

    @classmethod
    def from_config(cls, config: dict[str, Any]):
        return cls(**config)


class BatchGlobalPooling(Layer):
    """Per-graph read-out over a batch vector: [N, F], [N] -> [G, F] (global_pooling.py:140-316).
    ``G = max(batch) + 1`` like the reference (one host read); mean divides by max(count, 1)."""

    def __init__(self, pooling: str = "mean", **kwargs) -> None:
        super().__init__(**kwargs)
        if pooling not in _POOLS:
            raise ValueError(f"pooling must be one of ['mean', 'max', 'sum'], got {pooling}")
        self.pooling = pooling

    def call(self, inputs, **kwargs):
        if not isinstance(inputs, (list, tuple)) or len(inputs) != 2:
            raise ValueError("inputs must be a list/tuple of [node_features, batch], "
                             f"got {type(inputs)} with length {len(inputs) if hasattr(inputs, '__len__') else 'unknown'}")
        x = to_device_tensor(inputs[0], what="node features")
        if x.is_floating_point() and x.dtype != torch.float32:
            x = x.to(torch.float32)
        batch = to_device_tensor(inputs[1], what="batch")
        num_graphs = int(batch.max().item()) + 1
        graph, keep = _structure_for(batch, num_graphs)
        if keep is not None:
            x = x[keep]
        return ops.segment_reduce(x, graph, "max_raw" if self.pooling == "max" else self.pooling)

    def compute_output_shape(self, input_shape):
        if not isinstance(input_shape, (list, tuple)) or len(input_shape) != 2:
            raise ValueError("input_shape must be a list/tuple of 2 shapes for [node_features, batch]")
        node_features_shape, batch_shape = input_shape
        if isinstance(node_features_shape, int):
            raise ValueError("input_shape must be a list/tuple of 2 shapes for [node_features, batch], "
                             f"got single shape {input_shape}")
        if len(node_features_shape) != 2:
            raise ValueError("Expected node_features shape to be 2D (total_nodes, num_features), "
                             f"got {len(node_features_shape)}D")
        if len(batch_shape) != 1:
            raise ValueError(f"Expected batch shape to be 1D (total_nodes,), got {len(batch_shape)}D")
        return (None, node_features_shape[1])

    def get_config(self) -> dict[str, Any]:
        config = super().get_config()
        config.update({"pooling": self.pooling})
        return config

    @classmethod
    def from_config(cls, config: dict[str, Any]):
        return cls(**config)
