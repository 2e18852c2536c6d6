"""The filter step: score every 2-hop non-edge with the filter model, keep the k best.

B200-native restatement of the body of /root/reference/filter.py:92-166:

  reference                                         here
  ---------                                         ----
  A2 = adj@adj on one CPU thread, scipy masking     K6 one-pass bitmap enumeration per owner slab
  for batch in DataLoader(range(N), B):             one K2 / K3 launch per slab
      model(x, edges, adj)   # GNN re-run per batch   embeddings computed once (LinkGNN.embed)
      cat([edges.t(), score]).cpu()                   scores stay on the device
  all_scores[:,2].sort(descending=True)  # CPU      K4 radix-select + stable sort of k rows
  torch.save([N,3])                                 [k,3] float32 (k = N keeps the full list)

Slabs: owners are cut into contiguous ranges whose candidate BOUND is at most ``slab_pairs``
(``iter_slabs``); every slab is folded into the running proposal set by one K4 selection over
(running list ++ slab) (``RunningTopK``) — positions in that concatenation preserve the global
candidate order, so the final stable sort equals a single global stable sort.  Multi-GPU: each rank takes a contiguous owner range (parallel.partition_by_work)
and the per-rank lists are merged with one all-gather (parallel.merge_topk).
"""
from __future__ import annotations

from typing import Iterator, Optional, Tuple

import torch

from . import adamic_utils, candidates, ops, parallel
from .graph import SparseAdj, add_edges
from .models import GNN_MODELS, LinkGNN


def score_edges(model_name: str, model, x, adj: SparseAdj, edges: torch.Tensor,
                grouped_by_v: bool = True, ra_adj: Optional[SparseAdj] = None) -> torch.Tensor:
    """Scores of ``edges [2,M]`` under the filter model (/root/reference/filter.py:113-142)."""
    if model_name in GNN_MODELS:
        assert isinstance(model, LinkGNN)
        h = model.embed(x, adj)
        return model.linkpred.score_pairs(h, edges)
    if model_name in ("simple", "adamic"):
        model.grouped_by_v = grouped_by_v
        return model(x, edges, adj)
    if model_name == "adamic_ogb":
        return adamic_utils.AA(adamic_utils.get_A(adj, adj.n), edges, grouped_by_v=grouped_by_v)[0]
    if model_name == "resource_allocation":
        A = ra_adj if ra_adj is not None else adj
        return ops.cn_aa(A, edges, A.ra_weights(), use_values=True, grouped_by_v=grouped_by_v)
    raise ValueError(f"model {model_name!r} is not a filter model on this path")


def same_structure(a: SparseAdj, b: SparseAdj) -> bool:
    """True when two unweighted adjacencies hold the same entries (e.g. the RA graph rebuilt from
    the raw train split, filter.py:130-139, when no proposal edges were added)."""
    return (a is b) or (a.val is None and b.val is None and a.n == b.n and a.nnz == b.nnz
                        and bool(torch.equal(a.rowptr, b.rowptr)) and bool(torch.equal(a.col, b.col)))


def heuristic_table(model_name: str, adj: SparseAdj, ra_adj: Optional[SparseAdj] = None):
    """(graph, weight table, sigmoid, use_values) of a heuristic filter model when it can take the fused
    enumerate+score kernel (K6+K3, candidates.two_hop_scored): the scoring graph IS the graph whose 2-hop
    neighbourhood defines the candidates and, if weighted (collab), its values are bitwise symmetric.
    ``None`` otherwise (RA on a rebuilt multigraph) -> two_hop + ops.cn_aa."""
    if not candidates.values_symmetric(adj):
        return None
    if model_name == "simple":
        return adj, None, False, True                        # models.py:536-542 (weighted on collab)
    if model_name == "adamic":
        return adj, adj.adamic_weights(), True, False        # models.py:544-554 (indices only)
    if model_name == "adamic_ogb":
        return adj, adj.aa_ogb_weights(), False, True        # adamic_utils.py:13-25 (values of A)
    if model_name == "resource_allocation" and adj.val is None and (ra_adj is None or same_structure(ra_adj, adj)):
        return adj, adj.ra_weights(), False, True            # train_and_eval.py:195-216
    return None


def ra_graph_from_train_edges(train_edges: torch.Tensor, num_nodes: int) -> SparseAdj:
    """filter.py:130-139: A rebuilt from the raw train split, both directions, duplicate pairs
    SUMMED into integer weights (scipy csr_matrix constructor semantics)."""
    e = train_edges.reshape(-1, 2).long()
    both = torch.cat([e, e.flip(1)], 0).t()
    n = int(num_nodes)
    key = both[0] * n + both[1]
    ukey, cnt = torch.unique(key, sorted=True, return_counts=True)
    row = torch.div(ukey, n, rounding_mode="floor")
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=e.device)
    rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    val = cnt.float()
    return SparseAdj(rowptr.int(), (ukey - row * n).int(), None if bool((cnt == 1).all()) else val, n)


def iter_slabs(adj: SparseAdj, v_lo: int, v_hi: int, slab_pairs: int) -> Iterator[Tuple[int, int, int]]:
    """Yield (lo, hi, cap) owner ranges whose candidate count is at most ``cap <= slab_pairs``
    (a single owner whose bound exceeds that gets a slab of its own).  ``cap`` is the sum of the
    per-owner upper bounds (candidates.owner_bounds) — it sizes the one-pass kernel's outputs, so no
    count pass over the graph is needed to plan the slabs."""
    if v_hi <= v_lo:
        return
    cs = torch.cumsum(candidates.owner_bounds(adj)[v_lo:v_hi], 0).cpu()
    lo = 0
    n_own = v_hi - v_lo
    base = 0
    while lo < n_own:
        hi = int(torch.searchsorted(cs, torch.tensor(base + slab_pairs), right=True).item())
        hi = max(hi, lo + 1)
        hi = min(hi, n_own)
        end = int(cs[hi - 1])
        yield v_lo + lo, v_lo + hi, end - base
        base = end
        lo = hi


class RunningTopK:
    """The proposal set while owner slabs stream by.  State: the current best ``<= k`` candidates as
    SoA (u, v int32, score fp32) in CANDIDATE ORDER (= the reference's column-major order, which is the
    tie rule of the final sort).  ``update`` selects the k best of (state ++ slab) with K4's radix
    select + ordered compaction over the two segments in place — no concatenation, no sort;
    ``result`` sorts once (stable, score descending) and packs the float32 ``[k,3]`` list.
    Equal to one global stable sort of all candidates, bit for bit (tests/test_gpu_topk.py)."""

    def __init__(self, k: Optional[int]):
        self.k = k                      # None: keep every candidate, like the reference
        self.u = self.v = self.score = None
        self.kth_key = None             # order key of the current k-th score (device), once the list is full
        self.seen = 0

    def update(self, edges: torch.Tensor, score: torch.Tensor) -> None:
        M = score.numel()
        if M == 0:
            return
        self.seen += M
        pu, pv = ops._pairs(edges)
        score = score.contiguous().float()
        have = 0 if self.score is None else self.score.numel()
        kk = have + M if self.k is None else min(self.k, have + M)
        if kk == have + M:                                   # everything survives: plain append
            if self.score is None:
                self.u, self.v, self.score = pu.clone(), pv.clone(), score.clone()
            else:
                self.u, self.v = torch.cat([self.u, pu]), torch.cat([self.v, pv])
                self.score = torch.cat([self.score, score])
            return
        # a full list lets the select skip every slab element that is already worse than its k-th score
        prune = self.kth_key if (have == kk and self.kth_key is not None) else None
        idx, sc, self.kth_key = ops.topk_select2(self.score, score, kk, prune_key=prune, want_kth_key=True)
        self.u, self.v = ops.gather_pairs2(None if self.score is None else (self.u, self.v), (pu, pv), idx)
        self.score = sc

    def result(self, device=None) -> torch.Tensor:
        if self.score is None or self.score.numel() == 0:
            return torch.empty((0, 3), dtype=torch.float32, device=device)
        return ops.topk_edges(torch.stack([self.u, self.v]), self.score, self.score.numel())


@torch.no_grad()
def filter_topk(model_name: str, model, x, adj: SparseAdj, k: Optional[int] = None,
                slab_pairs: int = 1 << 27, distributed: bool = False, ra_adj: Optional[SparseAdj] = None,
                stats: Optional[dict] = None) -> torch.Tensor:
    """Sorted proposal list, float32 ``[k,3]`` rows (u, v, score) on the device: score descending,
    ties by the reference's candidate order.  ``k=None`` keeps every candidate like the reference."""
    rank, world = parallel.world_info() if distributed else (0, 1)
    v_lo, v_hi = 0, adj.n
    if world > 1:
        bounds = parallel.partition_by_work(candidates.two_path_work(adj), world)
        v_lo, v_hi = bounds[rank], bounds[rank + 1]
    running = RunningTopK(k)
    fused = heuristic_table(model_name, adj, ra_adj) if model_name not in GNN_MODELS else None
    for lo, hi, cap in iter_slabs(adj, v_lo, v_hi, slab_pairs):
        if fused is not None:
            # CN / AA / RA are the values of A@A: one walk over the owners' 2-paths yields the
            # candidates and their scores together (bit-identical to scoring the pairs with K3)
            edges, score = candidates.two_hop_scored(fused[0], fused[1], lo, hi, sigmoid=fused[2], cap=cap,
                                                     use_values=fused[3])
        else:
            edges = candidates.two_hop(adj, lo, hi, cap=cap)
            if edges.shape[1]:
                score = score_edges(model_name, model, x, adj, edges, True, ra_adj)
        if edges.shape[1]:
            running.update(edges, score)
        del edges
    n_scored = running.seen
    running = running.result(adj.device)
    if stats is not None:
        stats["candidates_scored"] = n_scored
    if world > 1:
        assert k is not None, "distributed filter needs a proposal size k"
        running = parallel.merge_topk(running, k)
    return running


def load_extra_edges(path: str, num_sorted_edge: int) -> torch.Tensor:
    """filter.py:82 / rank.py:294: the first ``num_sorted_edge`` rows of a saved sorted list."""
    t = torch.load(path)
    extra = t[:num_sorted_edge, :2].t().long()
    assert extra.size(0) == 2 and extra.size(1) == num_sorted_edge
    return extra
