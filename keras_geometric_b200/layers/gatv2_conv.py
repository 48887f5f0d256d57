"""GATv2Conv - same constructor, call contract and config keys as the reference
(/root/reference/src/keras_geometric/layers/gatv2_conv.py).

h = X W (one shared W, no bias) is computed per node; logits, the per-target softmax and the
alpha-weighted aggregation (gatv2_conv.py:241-335, ~9 materialised [E,H,C] tensors in the
reference) are ONE fused sm_100a kernel (kgb_gatv2_fwd).
"""
from __future__ import annotations

from typing import Any

import torch

from .. import ops
from .._compat import Dense, Dropout, initializers, to_device_tensor, value_of
from ..graph import get_graph
from .message_passing import MessagePassing


class GATv2Conv(MessagePassing):
    """gatv2_conv.py:11-75."""

    def __init__(self, output_dim: int, heads: int = 1, concat: bool = True, negative_slope: float = 0.2,
                 dropout: float = 0.0, use_bias: bool = True, kernel_initializer="glorot_uniform",
                 bias_initializer="zeros", att_initializer="glorot_uniform", add_self_loops: bool = True,
                 **kwargs) -> None:
        super().__init__(aggregator="sum", **kwargs)
        self.output_dim = output_dim
        self.heads = heads
        self.concat = concat
        self.negative_slope = negative_slope
        self.dropout_rate = dropout
        self.dropout_layer = Dropout(dropout) if dropout > 0 else None
        self.use_bias = use_bias
        self.kernel_initializer = kernel_initializer
        self.bias_initializer = bias_initializer
        self.att_initializer = att_initializer
        self.add_self_loops_flag = add_self_loops
        self.features_per_head = output_dim
        self.linear_transform = None
        self.att = None
        self.bias = None

    def build(self, input_shape) -> None:  # gatv2_conv.py:77-127
        shape = input_shape[0] if isinstance(input_shape, list) and len(input_shape) >= 1 else input_shape
        if hasattr(shape, "as_list"):
            shape = shape.as_list()
        elif hasattr(shape, "__len__"):
            shape = tuple(shape)
        if not isinstance(shape, (list, tuple)) or len(shape) != 2:
            raise ValueError(f"Expected features input shape like (N, F), but got {shape}")
        node_feature_dim = shape[1]
        if node_feature_dim is None:
            raise ValueError("Input feature dimension cannot be None.")
        self.linear_transform = Dense(self.heads * self.features_per_head,
                                      kernel_initializer=self.kernel_initializer, use_bias=False,
                                      name="linear_transform")
        self.linear_transform.build((None, int(node_feature_dim)))
        self.linear_transform.built = True
        self.att = self.add_weight(shape=(1, self.heads, self.features_per_head),
                                   initializer=initializers.get(self.att_initializer), name="att", trainable=True)
        if self.use_bias:
            bias_shape = (self.heads * self.features_per_head,) if self.concat else (self.features_per_head,)
            self.bias = self.add_weight(shape=bias_shape, initializer=initializers.get(self.bias_initializer),
                                        name="final_bias", trainable=True)
        else:
            self.bias = None
        super().build(input_shape)

    def call(self, inputs, edge_attr=None, training=None):  # gatv2_conv.py:129-174
        if isinstance(inputs, (list, tuple)) and len(inputs) >= 2:
            x, edge_index = inputs[0], inputs[1]
        else:
            raise ValueError(f"Expected inputs to be [x, edge_index], got {inputs}")
        from ..dist import PartitionedGraph
        if isinstance(edge_index, PartitionedGraph):
            return self._call_partitioned(x, edge_index, training)
        if not (isinstance(edge_index, torch.Tensor) and edge_index.is_cuda and edge_index.dtype == torch.int32):
            edge_index = self._cast_edge_index(edge_index)
        return self._gatv2_propagate(x=x, edge_index=edge_index, training=training,
                                     n_loops_from_flag=self.add_self_loops_flag)

    def _call_partitioned(self, x, pg, training=None):
        """Same layer on a 1-D node partition: ``h = x W`` is computed by the owner, the halo rows of ``h`` are
        exchanged (width H*C) and the fused edge kernel runs on the rank's rows over the [local | halo] sources."""
        x = to_device_tensor(x, what="x")
        if x.dtype != torch.float32:
            x = x.to(torch.float32)
        if bool(pg._n_loops) != bool(self.add_self_loops_flag):
            raise ValueError("PartitionedGraph(n_loops_local=...) must match GATv2Conv(add_self_loops=...)")
        if self.linear_transform is None or self.att is None:
            self.build(tuple(x.shape))
            self.built = True
        n, H, C = int(x.shape[0]), self.heads, self.features_per_head
        h = ops.linear(x, value_of(self.linear_transform.kernel))
        h_ext = pg.exchange(h)
        att = value_of(self.att)
        bias = value_of(self.bias) if (self.use_bias and self.bias is not None) else None
        fuse_bias = bias is not None and (self.concat or H == 1)
        # attention dropout (training): the fused kernels draw one Philox decision per (local edge id, head); forward and
        # both backward passes regenerate it, exactly as on one GPU (the mask itself differs from a 1-GPU run's - the
        # edge ids are rank-local - which no reference fixes either)
        drop = float(self.dropout_rate) if (self.dropout_layer is not None and training) else 0.0
        out = ops.gatv2_aggregate(h_ext, h, att, pg.graph, H, C, self.negative_slope, bias if fuse_bias else None,
                                  dropout=drop)
        if not self.concat and H > 1:
            out = out.reshape(n, H, C).mean(dim=1)
            if bias is not None:
                out = out + bias
        return out

    def propagate(self, x, edge_index, edge_attr=None, size=None, **kwargs):  # gatv2_conv.py:162-174
        return self._gatv2_propagate(x=x, edge_index=edge_index, training=kwargs.get("training", None))

    def _gatv2_propagate(self, x, edge_index, training=None, n_loops_from_flag: bool = False):
        """gatv2_conv.py:176-266."""
        if isinstance(x, (list, tuple)):
            x_i, x_j = to_device_tensor(x[0], what="x_i"), to_device_tensor(x[1], what="x_j")
        else:
            x_i = x_j = to_device_tensor(x, what="x")
        if x_i.dtype != torch.float32:
            x_i = x_i.to(torch.float32)
            x_j = x_i if isinstance(x, torch.Tensor) or not isinstance(x, (list, tuple)) else x_j.to(torch.float32)
        n, n_src = int(x_i.shape[0]), int(x_j.shape[0])
        H, C = self.heads, self.features_per_head
        width = H * C if self.concat else C
        edge_index = to_device_tensor(edge_index, what="edge_index")
        if edge_index.dtype != torch.int32:
            edge_index = edge_index.to(torch.int32)
        n_loops = n if n_loops_from_flag else 0
        e = int(edge_index.shape[1]) + n_loops
        if n == 0 or e == 0:  # gatv2_conv.py:195-210 (no bias on these paths)
            return torch.zeros((n, width), dtype=x_i.dtype, device=x_i.device)
        if self.linear_transform is None or self.att is None:
            self.build([tuple(x_i.shape), tuple(x_j.shape)] if x_i is not x_j else tuple(x_i.shape))
            self.built = True
        if self.linear_transform is None:
            raise RuntimeError("Linear transform layer not built")
        w = value_of(self.linear_transform.kernel)
        h_i = ops.linear(x_i, w)
        h_j = h_i if x_i is x_j else ops.linear(x_j, w)
        graph = get_graph(edge_index, n, n_src, n_loops)
        att = value_of(self.att)
        bias = value_of(self.bias) if (self.use_bias and self.bias is not None) else None
        if C > 512 or (C > 128 and C % 4 != 0):
            # wider than the fused kernels' lane layout (per-head width / vector <= 128): the reference's per-edge
            # formulation on top of the segment kernels (any output_dim is accepted, like the reference)
            return self._propagate_with_dropout(h_i, h_j, graph, att, bias, training)
        # attention dropout (gatv2_conv.py:252-253, training only) is generated inside the fused kernels: alpha is
        # never materialised
        drop = float(self.dropout_rate) if (self.dropout_layer is not None and training) else 0.0
        fuse_bias = bias is not None and (self.concat or H == 1)
        out = ops.gatv2_aggregate(h_j, h_i, att, graph, H, C, self.negative_slope, bias if fuse_bias else None,
                                  dropout=drop)
        if not self.concat and H > 1:
            out = out.reshape(n, H, C).mean(dim=1)
            if bias is not None:
                out = out + bias
        return out

    def _propagate_with_dropout(self, h_i, h_j, graph, att, bias, training):
        """The reference's per-edge formulation (gatv2_conv.py:241-335) on top of the segment kernels, alpha
        materialised: used for per-head widths beyond the fused kernels' layout."""
        n, H, C = int(h_i.shape[0]), self.heads, self.features_per_head
        hj_e = ops.take_rows(h_j, graph, "src").reshape(-1, H, C)
        hi_e = ops.take_rows(h_i, graph, "dst").reshape(-1, H, C)
        z = torch.nn.functional.leaky_relu(hi_e + hj_e, self.negative_slope)
        s = (z * att).sum(dim=-1)
        dst = graph.full_edge_index()[1].long()
        m = ops.segment_reduce(s, graph, "max").detach()
        # rows whose max is +-inf were rewritten to 0 by the max aggregator; harmless for finite logits
        p = torch.exp(s - m.index_select(0, dst))
        d = ops.segment_reduce(p, graph, "sum")
        alpha = p / (d.index_select(0, dst) + 1e-10)
        if self.dropout_layer is not None and training:
            alpha = self.dropout_layer(alpha, training=training)
        msg = (alpha.unsqueeze(-1) * hj_e).reshape(-1, H * C)
        agg = ops.segment_reduce(msg, graph, "sum").reshape(n, H, C)
        return self._final_update(agg, bias)

    def _final_update(self, aggregated, bias=None):  # gatv2_conv.py:337-352
        n = int(aggregated.shape[0])
        out = aggregated.reshape(n, self.heads * self.features_per_head) if self.concat else aggregated.mean(dim=1)
        if bias is None and self.use_bias and self.bias is not None:
            bias = value_of(self.bias)
        return out + bias if bias is not None else out

    def message(self, x_i, x_j, edge_attr=None, edge_index=None, size=None, **kwargs):  # gatv2_conv.py:354-367
        return x_j

    def get_config(self) -> dict[str, Any]:  # gatv2_conv.py:369-387
        config = super().get_config()
        config.update({
            "output_dim": self.output_dim,
            "heads": self.heads,
            "concat": self.concat,
            "negative_slope": self.negative_slope,
            "dropout": self.dropout_rate,
            "use_bias": self.use_bias,
            "kernel_initializer": self.kernel_initializer,
            "bias_initializer": self.bias_initializer,
            "att_initializer": self.att_initializer,
            "add_self_loops": self.add_self_loops_flag,
        })
        return config

    @classmethod
    def from_config(cls, config: dict[str, Any]) -> "GATv2Conv":  # gatv2_conv.py:389-399
        config = dict(config)
        config.pop("aggregator", None)
        return cls(**config)
