"""The consumer side of the proposal set: prefix selection, graph augmentation and the rank-side
evaluation, on the same kernels.

  sweep schedule / prefix        /root/reference/rank.py:260-272, 294
  ``--valid_proposal`` surgery   /root/reference/rank.py:222-251
  augmented graphs               /root/reference/rank.py:299-314 (adj_t, full_adj_t)
  evaluation                     /root/reference/train_and_eval.py:98-156 (test), 158-193
                                 (test_adamic), 218-270 (test_resource_allocation)
  Hits@K                         ogb Evaluator: 1.0 if len(neg) < K else
                                 mean(pos > K-th largest neg), strict '>'

Training of a parameterised rank model: ``train_step.train`` (the epoch loop lives in rank.py).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

from . import ops
from .filter_step import score_edges
from .graph import SparseAdj, add_edges

HITS = {"collab": [10, 50, 100], "reddit": [10, 50, 100], "ppa": [10, 100, 200], "ddi": [10, 20, 30],
        "email": [10, 20, 30], "twitch": [10, 50, 100], "fb": [10, 20, 30]}   # train_and_eval.py:20-29


def sweep_index_ends(sweep_num, sweep_min, sweep_max, num_sorted_edge) -> List[int]:
    """rank.py:260-272."""
    if sweep_num:
        sweep_min = 0 if sweep_min is None else sweep_min
        sweep_max = (sweep_num - 1) * 1000 if sweep_max is None else sweep_max
        return [sweep_min + int(i * (sweep_max - sweep_min) / sweep_num) for i in range(sweep_num + 1)]
    if num_sorted_edge:
        return [num_sorted_edge]
    return [0]


def prefix_edges(sorted_edges: torch.Tensor, index_end: int) -> torch.Tensor:
    """rank.py:294-297."""
    extra = sorted_edges[: int(index_end), :2].t().long()
    assert extra.size(0) == 2 and extra.size(1) == index_end
    return extra


def valid_proposal(sorted_edges: torch.Tensor, valid_pos: torch.Tensor) -> torch.Tensor:
    """rank.py:222-251: both directions of every validation edge go on top with score 100000.0 and
    are removed from the body (vectorised; the reference loops over python sets)."""
    vp = valid_pos.reshape(-1, 2).long().cpu()
    both = torch.cat([vp, vp.flip(1)], 0)
    n = int(max(both.max().item(), sorted_edges[:, :2].max().item())) + 1
    bkey = torch.unique(both[:, 0] * n + both[:, 1])
    top = torch.stack([torch.div(bkey, n, rounding_mode="floor").double(), (bkey % n).double(),
                       torch.full((bkey.numel(),), 100000.0, dtype=torch.float64)], 1)
    se = sorted_edges.cpu().double()
    key = se[:, 0].long() * n + se[:, 1].long()
    keep = ~torch.isin(key, bkey)
    out = torch.cat([top, se[keep]], 0)
    assert bkey.numel() == top.shape[0]
    return out


def augmented_graphs(dataset: str, edge_index, edge_weight, extra_edges, split_edge, num_nodes, device,
                     eval_extra=None):
    """(adj_t, full_adj_t) of rank.py:299-314: full adds both directions of the validation edges
    for collab / email / reddit.  ``eval_extra``: the proposal prefix that still enters
    ``full_adj_t`` when ``--only_supervision`` keeps it out of ``adj_t`` (rank.py:299-312)."""
    ei, ew = edge_index.to(device), edge_weight.to(device)
    adj = add_edges(dataset, ei, ew, extra_edges.to(device), num_nodes)
    if dataset.split("-shape")[0] in ("collab", "email", "reddit"):
        v = split_edge["valid"]["edge"].t().to(device)
        vboth = torch.unique(torch.cat([v, v.flip(0)], 1), dim=1)          # to_undirected
        fx = (extra_edges if eval_extra is None else eval_extra).to(device)
        full = add_edges(dataset, ei, ew, torch.cat([fx, vboth], 1), num_nodes)
    else:
        full = adj
    return adj, full


def hits_at_k(pos: torch.Tensor, neg: torch.Tensor, K: int) -> float:
    if neg.numel() < K:
        return 1.0
    kth = ops.topk(neg.float(), K)[1][-1]
    return float((pos.float() > kth).sum().item()) / float(pos.numel())


@torch.no_grad()
def evaluate(model_name: str, model, x, adj: SparseAdj, full_adj: SparseAdj, split_edge, dataset: str
             ) -> Dict[str, Tuple[float, float, float]]:
    """{'Hits@K': (train, valid, test)} exactly as the reference's test / test_adamic /
    test_resource_allocation compute it: valid edges on adj_t, test edges on full_adj_t; the
    heuristic variants score 'train' with constant ones (train_and_eval.py:170,250)."""
    dev = adj.device
    e = lambda name, key: split_edge[name][key].t().to(dev)
    sc = lambda edges, a: score_edges(model_name, model, x, a, edges, grouped_by_v=False).reshape(-1)
    pos_valid, neg_valid = sc(e("valid", "edge"), adj), sc(e("valid", "edge_neg"), adj)
    pos_test, neg_test = sc(e("test", "edge"), full_adj), sc(e("test", "edge_neg"), full_adj)
    if model_name in ("adamic_ogb", "resource_allocation"):
        pos_train = torch.ones(split_edge["train"]["edge"].size(0), device=dev)
    else:
        pos_train = sc(e("eval_train", "edge"), adj)
    base = dataset.split("-shape")[0]
    out = {}
    for K in HITS[base]:
        out[f"Hits@{K}"] = (hits_at_k(pos_train, neg_valid, K), hits_at_k(pos_valid, neg_valid, K),
                            hits_at_k(pos_test, neg_test, K))
    return out


class RunLog:
    """Per-run history of (train, valid, test) Hits for one K, with the summaries the reference
    prints (/root/reference/logger.py:4-47): best validation epoch decides the reported test value."""

    def __init__(self, runs: int):
        self.results = [[] for _ in range(runs)]

    def add_result(self, run: int, result):
        assert len(result) == 3 and 0 <= run < len(self.results)
        self.results[run].append(tuple(float(v) for v in result))

    def _best(self, run: int):
        r = 100 * torch.tensor(self.results[run], dtype=torch.float32)
        am = int(r[:, 1].argmax())
        return float(r[:, 0].max()), float(r[:, 1].max()), float(r[am, 0]), float(r[am, 2])

    def curve_point(self, run: int, index_end: int):
        """[index_end, valid, test] at the best-validation epoch (rank.py:377-380)."""
        r = 100 * torch.tensor(self.results[run], dtype=torch.float32)
        am = int(r[:, 1].argmax())
        return [index_end, r[am, 1], r[am, 2]]

    def print_statistics(self, run=None):
        if run is not None:
            hi_tr, hi_va, fin_tr, fin_te = self._best(run)
            print(f"Run {run + 1:02d}:")
            print(f"Highest Train: {hi_tr:.2f}")
            print(f"Highest Valid: {hi_va:.2f}")
            print(f"  Final Train: {fin_tr:.2f}")
            print(f"   Final Test: {fin_te:.2f}")
            return
        best = torch.tensor([self._best(i) for i in range(len(self.results)) if self.results[i]])
        print("All runs:")
        for j, label in enumerate(["Highest Train", "Highest Valid", "  Final Train", "   Final Test"]):
            c = best[:, j]
            std = c.std() if c.numel() > 1 else torch.tensor(float("nan"))
            print(f"{label}: {c.mean():.2f} ± {std:.2f}")
