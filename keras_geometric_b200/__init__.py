"""keras_geometric_b200 - B200-native (sm_100a) message-passing hot path behind
keras-geometric's layer API (MessagePassing, GCNConv, GINConv, SAGEConv, GATv2Conv)."""
from .layers import (GATv2Conv, GCNConv, GINConv, MessagePassing, SAGEConv)  # noqa: F401
from .data_utils import GraphData, batch_graphs  # noqa: F401
from .utils import add_self_loops, compute_gcn_normalization  # noqa: F401

__version__ = "0.1.0"
__all__ = ["__version__", "GCNConv", "GINConv", "GATv2Conv", "SAGEConv", "MessagePassing", "add_self_loops",
           "compute_gcn_normalization", "GraphData", "batch_graphs"]
