"""Layer exports mirror /root/reference/src/keras_geometric/layers/__init__.py (hot-path subset)."""
from .aggregators import (Aggregator, AggregatorFactory, MaxAggregator, MeanAggregator, MinAggregator,
                          PoolingAggregator, StdAggregator, SumAggregator)
from .gatv2_conv import GATv2Conv
from .gcn_conv import GCNConv
from .gin_conv import GINConv
from .message_passing import MessagePassing
from .pooling import BatchGlobalPooling, GlobalPooling
from .sage_conv import SAGEConv

__all__ = ["MessagePassing", "GCNConv", "GINConv", "GATv2Conv", "SAGEConv", "Aggregator", "AggregatorFactory",
           "MeanAggregator", "MaxAggregator", "SumAggregator", "MinAggregator", "StdAggregator",
           "PoolingAggregator", "GlobalPooling", "BatchGlobalPooling"]
