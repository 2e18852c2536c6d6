"""2-hop candidate enumeration (K6) — the GPU replacement for /root/reference/filter.py:96-109.

``two_hop(adj)`` returns the reference's ``all_edges`` as int32 ``[2, N]`` on the device, in the
reference's column-major order (sorted by (all_edges[1], all_edges[0])).  ``owner_counts`` /
``two_hop(adj, v_lo, v_hi)`` work on a contiguous owner range so callers can stream slabs of a
graph whose full candidate list would not fit (ogbl-ppa scale) or shard owners across GPUs.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import EPS_CN_SIGMOID, EpsError, check
from .graph import SparseAdj
from .ops import _need_cuda, _ptr, _stream, _ws


def owner_counts(adj: SparseAdj, v_lo: int = 0, v_hi: int | None = None) -> torch.Tensor:
    """#candidates of every owner v in [v_lo, v_hi) as int64."""
    _need_cuda(adj.col)
    lib = _lib.load()
    v_hi = adj.n if v_hi is None else v_hi
    counts = torch.zeros(max(v_hi - v_lo, 0), dtype=torch.int32, device=adj.device)
    if v_hi <= v_lo:
        return counts.long()
    ws = _ws(lib.eps_twohop_workspace_bytes(), adj.device)
    check(lib.eps_twohop_candidates(_ptr(adj.rowptr), _ptr(adj.col), adj.n, v_lo, v_hi, None, _ptr(counts),
                                    None, None, _ptr(ws), ws.numel(), _stream()), "eps_twohop_candidates")
    from . import ops as _ops
    _ops.LAUNCHES["n"] += 1 if v_hi > v_lo else 0
    return counts.long() & 0xFFFFFFFF


def two_hop(adj: SparseAdj, v_lo: int = 0, v_hi: int | None = None, counts: torch.Tensor | None = None):
    """(u, v) candidate list of owners [v_lo, v_hi): int32 [2, N], column-major reference order."""
    _need_cuda(adj.col)
    lib = _lib.load()
    v_hi = adj.n if v_hi is None else v_hi
    if counts is None:
        counts = owner_counts(adj, v_lo, v_hi)
    offsets = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=adj.device)
    torch.cumsum(counts, 0, out=offsets[1:])
    N = int(offsets[-1].item())
    if N >= 2**31:
        raise EpsError(f"{N} candidates in one slab; split the owner range (see filter_step.iter_slabs)")
    edges = torch.empty((2, N), dtype=torch.int32, device=adj.device)
    if N == 0:
        return edges
    ws = _ws(lib.eps_twohop_workspace_bytes(), adj.device)
    check(lib.eps_twohop_candidates(_ptr(adj.rowptr), _ptr(adj.col), adj.n, v_lo, v_hi, _ptr(offsets), None,
                                    _ptr(edges[0]), _ptr(edges[1]), _ptr(ws), ws.numel(), _stream()),
          "eps_twohop_candidates")
    from . import ops as _ops
    _ops.LAUNCHES["n"] += 1
    return edges


def two_hop_scored(adj: SparseAdj, wtable: torch.Tensor | None = None, v_lo: int = 0, v_hi: int | None = None,
                   counts: torch.Tensor | None = None, sigmoid: bool = False, want_count: bool = False):
    """K6+K3 fused: the candidates of owners [v_lo, v_hi) AND their heuristic scores from one walk
    over the owners' 2-paths (``wtable=None`` -> CN count, else sum of ``wtable[k]`` over the common
    neighbours: AA with 1/log deg, RA with 1/deg).  Returns ``(edges int32 [2,N], score fp32 [N])``
    (+ ``count int32 [N]`` with ``want_count``); scores are bit-identical to ``ops.cn_aa`` on the
    same pairs.  Unweighted adjacency only — weighted graphs go through two_hop + ops.cn_aa."""
    _need_cuda(adj.col, wtable)
    if adj.val is not None:
        raise EpsError("two_hop_scored: weighted adjacency (collab) is scored by two_hop + ops.cn_aa")
    lib = _lib.load()
    v_hi = adj.n if v_hi is None else v_hi
    if counts is None:
        counts = owner_counts(adj, v_lo, v_hi)
    offsets = torch.zeros(counts.numel() + 1, dtype=torch.int64, device=adj.device)
    torch.cumsum(counts, 0, out=offsets[1:])
    N = int(offsets[-1].item())
    if N >= 2**31:
        raise EpsError(f"{N} candidates in one slab; split the owner range (see filter_step.iter_slabs)")
    edges = torch.empty((2, N), dtype=torch.int32, device=adj.device)
    score = torch.empty(N, dtype=torch.float32, device=adj.device)
    count = torch.empty(N, dtype=torch.int32, device=adj.device) if want_count else None
    if N:
        wt = None if wtable is None else wtable.contiguous().float()
        ws = _ws(lib.eps_twohop_scored_workspace_bytes(N), adj.device)
        check(lib.eps_twohop_scored(_ptr(adj.rowptr), _ptr(adj.col), _ptr(wt), adj.n, v_lo, v_hi, _ptr(offsets), N,
                                    EPS_CN_SIGMOID if sigmoid else 0, _ptr(edges[0]), _ptr(edges[1]), _ptr(score),
                                    _ptr(count), _ptr(ws), ws.numel(), _stream()), "eps_twohop_scored")
        from . import ops as _ops
        _ops.LAUNCHES["n"] += 2
    return (edges, score, count) if want_count else (edges, score)


def two_path_work(adj: SparseAdj) -> torch.Tensor:
    """Per-owner work estimate sum_{k in N(v)} deg(k) (the 2-path count), used to cut owner ranges
    into equal-work slabs / GPU shards.  Torch ops only; device-agnostic."""
    deg = adj.degree().long()
    contrib = deg[adj.col.long()]
    out = torch.zeros(adj.n, dtype=torch.int64, device=adj.device)
    return out.index_add_(0, adj.row(), contrib)
