"""``add_self_loops`` / ``compute_gcn_normalization`` with the reference's signatures
(/root/reference/src/keras_geometric/utils/main.py), computed on the device."""
from __future__ import annotations

import torch

from ._compat import to_device_tensor
from .graph import GraphStructure


def add_self_loops(edge_index, num_nodes: int):
    """utils/main.py:8-16: append [0..N-1]->[0..N-1] after the existing edges (never dedups)."""
    edge_index = to_device_tensor(edge_index, what="edge_index")
    if edge_index.shape[0] != 2:
        edge_index = torch.stack([edge_index[0], edge_index[1]], dim=0)
    loop = torch.arange(0, int(num_nodes), dtype=edge_index.dtype, device=edge_index.device)
    return torch.cat([edge_index, torch.stack([loop, loop], dim=0)], dim=1)


def compute_gcn_normalization(edge_index, num_nodes: int):
    """utils/main.py:20-33: w_e = dis[dst_e] * dis[src_e], dis = (in_degree + 1e-12)^-0.5, via
    kgb_csr_build (degrees) + kgb_gcn_norm.  ``edge_index`` already contains any self-loops."""
    edge_index = to_device_tensor(edge_index, torch.int32, "edge_index")
    if edge_index.shape[1] == 0:
        return torch.zeros((0,), dtype=torch.float32, device=edge_index.device)
    g = GraphStructure(edge_index.contiguous(), int(num_nodes), int(num_nodes), 0)
    return g.gcn_norm()[1]
