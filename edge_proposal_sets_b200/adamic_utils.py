"""Drop-in for the two functions of the reference's ``adamic_utils.py`` that sit on the scoring path
(/root/reference/adamic_utils.py:8-25) and for ``train_and_eval.resource_allocation``
(/root/reference/train_and_eval.py:195-216), with the same call signatures:

    A = get_A(adj_t, num_nodes)
    pred, edge_index = AA(A, edge_index, batch_size=2000)
    pred = resource_allocation(A, link_list, batch_size)

``A`` is the device-resident ``SparseAdj`` itself (the reference converts to scipy CSR on the host
and scores single-threaded); ``batch_size`` is accepted and ignored — the K3 kernel takes the whole
pair list in one launch.  Scores come back as a float32 tensor on the graph's device.
"""
from __future__ import annotations

import torch

from . import ops
from .graph import SparseAdj


def get_A(adj: SparseAdj, num_nodes: int) -> SparseAdj:
    assert adj.n == num_nodes
    return adj


def _grouped(edge_index: torch.Tensor) -> bool:
    """True when equal destination ids come in runs (the column-major candidate order)."""
    v = edge_index[1]
    if v.numel() < 4096:
        return False
    runs = int((v[1:] != v[:-1]).sum().item()) + 1
    return v.numel() >= 16 * runs


def AA(A: SparseAdj, edge_index: torch.Tensor, batch_size: int = 2000, grouped_by_v=None):
    """sum_k A[u,k] * (A[v,k] / log(colsum_k)), 1/log = inf -> 0; no sigmoid (adamic_utils.py:13-25)."""
    e = edge_index.to(A.device)
    g = _grouped(e) if grouped_by_v is None else grouped_by_v
    pred = ops.cn_aa(A, e, A.aa_ogb_weights(), use_values=True, grouped_by_v=g)
    return pred, edge_index


def resource_allocation(A: SparseAdj, link_list: torch.Tensor, batch_size: int = 32768, grouped_by_v=None):
    """sum_k A[u,k] * (A[v,k] / colsum_k) for ``link_list [m,2]`` (train_and_eval.py:195-216)."""
    e = link_list.t().to(A.device)
    g = _grouped(e) if grouped_by_v is None else grouped_by_v
    return ops.cn_aa(A, e, A.ra_weights(), use_values=True, grouped_by_v=g)
