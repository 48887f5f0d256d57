"""Aggregation strategies with the reference's interface
(/root/reference/src/keras_geometric/layers/aggregators.py), executed by the sm_100a
segment kernels (kgb_gather_reduce with col = perm, include/kgb200.h)."""
from __future__ import annotations

import threading
from abc import ABC, abstractmethod

import torch

from .. import ops
from .._compat import apply_dense, to_device_tensor
from ..graph import GraphStructure

_hint = threading.local()


class graph_hint:
    """Lets ``propagate`` hand its prebuilt structure to ``Aggregator.aggregate`` when the
    ``target_idx`` it passes on is the very tensor the structure was built from."""

    def __init__(self, target_idx, graph):
        self.pair = (target_idx, graph)

    def __enter__(self):
        self.prev = getattr(_hint, "pair", None)
        _hint.pair = self.pair

    def __exit__(self, *exc):
        _hint.pair = self.prev


def _structure_for(target_idx: torch.Tensor, dim_size: int):
    """(graph, keep_mask or None).  Segment ids outside [0, dim_size) are dropped silently, which
    is what keras.ops.segment_* does on the torch backend."""
    pair = getattr(_hint, "pair", None)
    if pair is not None and pair[0] is target_idx and pair[1].n_dst == dim_size:
        return pair[1], None
    t = target_idx.to(torch.int32).reshape(-1)
    ei = torch.stack([torch.zeros_like(t), t]).contiguous()
    try:
        return GraphStructure(ei, dim_size, 1, 0), None
    except IndexError:
        keep = (t >= 0) & (t < dim_size)
        t = t[keep]
        ei = torch.stack([torch.zeros_like(t), t]).contiguous()
        return GraphStructure(ei, dim_size, 1, 0), keep


class Aggregator(ABC):
    """aggregators.py:16-45."""

    @abstractmethod
    def aggregate(self, messages, target_idx, dim_size: int):
        ...

    @property
    @abstractmethod
    def name(self) -> str:
        ...


class _KernelAggregator(Aggregator):
    _op = "sum"

    def aggregate(self, messages, target_idx, dim_size: int):
        messages = to_device_tensor(messages, what="messages")
        if messages.shape[0] == 0:  # aggregators.py:59-61
            return torch.zeros((dim_size, messages.shape[1]), dtype=messages.dtype, device=messages.device)
        target_idx = to_device_tensor(target_idx, what="target_idx")
        graph, keep = _structure_for(target_idx, int(dim_size))
        if keep is not None:
            messages = messages[keep]
        out = ops.segment_reduce(messages, graph, self._op)
        return out if out.dtype == messages.dtype else out.to(messages.dtype)

    @property
    def name(self) -> str:
        return self._op


class MeanAggregator(_KernelAggregator):
    """aggregators.py:48-89: sum / max(count, 1e-8)."""
    _op = "mean"


class MaxAggregator(_KernelAggregator):
    """aggregators.py:92-116: segment max, -inf (and real +-inf) -> 0."""
    _op = "max"


class SumAggregator(_KernelAggregator):
    """aggregators.py:119-141."""
    _op = "sum"


class MinAggregator(_KernelAggregator):
    """aggregators.py:144-171: -segment_max(-m), inf -> 0."""
    _op = "min"


class StdAggregator(Aggregator):
    """aggregators.py:174-232: two-pass population std; count <= 1 -> 0.  Two launches of the segment kernel: the
    mean, then KGB_OP_SQDEV (squared deviations from the row's mean, sqrt and the count <= 1 rule in its epilogue)."""

    def aggregate(self, messages, target_idx, dim_size: int):
        messages = to_device_tensor(messages, what="messages")
        if messages.shape[0] == 0:
            return torch.zeros((dim_size, messages.shape[1]), dtype=messages.dtype, device=messages.device)
        target_idx = to_device_tensor(target_idx, what="target_idx")
        graph, keep = _structure_for(target_idx, int(dim_size))
        if keep is not None:
            messages = messages[keep]
        return ops.segment_std(messages, graph)   # two fused passes (mean, squared deviations), no [E, F] temporaries

    @property
    def name(self) -> str:
        return "std"


class PoolingAggregator(Aggregator):
    """aggregators.py:235-278: max over Dense-transformed messages."""

    def __init__(self, pool_mlp) -> None:
        self.pool_mlp = pool_mlp

    def aggregate(self, messages, target_idx, dim_size: int):
        messages = to_device_tensor(messages, what="messages")
        if messages.shape[0] == 0:
            width = self._transform(torch.zeros((1, messages.shape[1]), dtype=messages.dtype,
                                                device=messages.device)).shape[1]
            return torch.zeros((dim_size, width), dtype=messages.dtype, device=messages.device)
        return MaxAggregator().aggregate(self._transform(messages), target_idx, dim_size)

    def _transform(self, messages):
        if getattr(self.pool_mlp, "kernel", None) is not None and hasattr(self.pool_mlp, "units"):
            return apply_dense(self.pool_mlp, messages)   # Dense on the tensor-core kernel (K8)
        return self.pool_mlp(messages)

    @property
    def name(self) -> str:
        return "pooling"


class AggregatorFactory:
    """aggregators.py:281-343."""

    _AGGREGATORS = {"mean": MeanAggregator, "max": MaxAggregator, "sum": SumAggregator,
                    "min": MinAggregator, "std": StdAggregator}

    @classmethod
    def create(cls, aggregator_name: str, **kwargs) -> Aggregator:
        if aggregator_name not in cls._AGGREGATORS:
            raise ValueError(f"Invalid aggregator: {aggregator_name}. "
                             f"Available aggregators: {list(cls._AGGREGATORS.keys())}")
        return cls._AGGREGATORS[aggregator_name](**kwargs)

    @classmethod
    def create_pooling(cls, pool_mlp) -> PoolingAggregator:
        return PoolingAggregator(pool_mlp)

    @classmethod
    def get_available_aggregators(cls) -> list:
        return list(cls._AGGREGATORS.keys())
