"""Oracle: graph construction and 2-hop candidate enumeration (numpy / scipy).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

Restates
  * ``rank.add_edges``            /root/reference/rank.py:28-36
  * candidate enumeration         /root/reference/filter.py:96-109
and the torch_sparse semantics they rely on.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import scipy.sparse as ssp


@dataclass
class CSR:
    """Symmetric adjacency in CSR form; columns ascending inside every row."""

    rowptr: np.ndarray  # int64 [n+1]
    col: np.ndarray     # int64 [nnz]
    val: np.ndarray     # float32 [nnz]
    n: int

    @property
    def nnz(self) -> int:
        return int(self.col.shape[0])

    def degree(self) -> np.ndarray:
        return np.diff(self.rowptr)

    def to_scipy(self) -> ssp.csr_matrix:
        m = ssp.csr_matrix((self.val, self.col, self.rowptr), shape=(self.n, self.n))
        m.has_sorted_indices = True
        return m


def add_edges(dataset: str, edge_index: np.ndarray, edge_weight: np.ndarray,
              extra_edges: np.ndarray, num_nodes: int) -> CSR:
    """rank.py:28-36.

    ``cat(edge_index, extra_edges)`` with weights ``cat(edge_weight, ones)``;
    ``SparseTensor.from_edge_index`` keeps duplicates; ``to_symmetric()`` forms
    the multiset {(r,c,w)} U {(c,r,w)} and sums the weights of equal (r,c)
    (SURVEY A.2); unless ``dataset == "collab"`` every stored value is then
    reset to 1.0 (rank.py:34-35).
    """
    ei = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    ex = np.asarray(extra_edges, dtype=np.int64).reshape(2, -1)
    w = np.concatenate([np.asarray(edge_weight, dtype=np.float32).reshape(-1),
                        np.ones(ex.shape[1], dtype=np.float32)])
    full = np.concatenate([ei, ex], axis=1)
    assert w.shape[0] == full.shape[1]
    n = int(num_nodes)
    r = np.concatenate([full[0], full[1]])
    c = np.concatenate([full[1], full[0]])
    ww = np.concatenate([w, w])
    key = r * n + c
    order = np.argsort(key, kind="stable")
    key = key[order]
    ww = ww[order]
    if key.size:
        first = np.concatenate([[True], key[1:] != key[:-1]])
    else:
        first = np.zeros(0, dtype=bool)
    starts = np.flatnonzero(first)
    ukey = key[starts]
    if key.size:
        # fp32 segment sum in sorted order (integer-valued weights in every
        # reference dataset => exact regardless of order)
        val = np.add.reduceat(ww, starts).astype(np.float32)
    else:
        val = np.zeros(0, dtype=np.float32)
    if dataset != "collab":
        val = np.ones_like(val)
    row = ukey // n
    col = ukey % n
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, row + 1, 1)
    rowptr = np.cumsum(rowptr)
    return CSR(rowptr=rowptr, col=col.astype(np.int64), val=val, n=n)


def two_hop_candidates(g: CSR, return_values: bool = False):
    """filter.py:96-109: every (u, v), u != v, with (A@A)[u,v] != 0 and A[u,v] == 0.

    Returned as int64 ``[2, N]`` with row 0 = ``all_edges[:,0]`` (u) and row 1 =
    ``all_edges[:,1]`` (v), in the reference's order: the scipy CSC->COO walk,
    i.e. sorted by (v, u) ascending.  Both (u,v) and (v,u) appear.
    ``return_values`` additionally returns the A@A value (the CN count / the
    weighted product sum on collab) that the reference computes and discards.
    """
    A = g.to_scipy()
    A2 = (A @ A).tocsc()
    A2.sort_indices()
    indptr, rows, vals = A2.indptr, A2.indices.astype(np.int64), A2.data
    cols = np.repeat(np.arange(g.n, dtype=np.int64), np.diff(indptr))
    keep = (rows != cols) & (vals != 0)          # remove_diag, values.nonzero()
    # A2[adj > 0] = 0  -> drop pairs that are stored (positive) entries of A
    Akeys = np.repeat(np.arange(g.n, dtype=np.int64), np.diff(g.rowptr)) * g.n + g.col
    Akeys = Akeys[g.val > 0]                      # CSR keys are already ascending
    k = rows * g.n + cols                         # (u=row, v=col) ; A symmetric
    pos = np.searchsorted(Akeys, k)
    pos[pos >= Akeys.size] = max(Akeys.size - 1, 0)
    is_edge = (Akeys[pos] == k) if Akeys.size else np.zeros_like(keep)
    keep &= ~is_edge
    out = np.stack([rows[keep], cols[keep]])
    if return_values:
        return out, vals[keep]
    return out


def degree_sorted_is_valid(g: CSR) -> bool:
    """Structural invariants every kernel assumes: ascending columns, no duplicates."""
    d = np.diff(g.col)
    row_start = np.zeros(g.nnz, dtype=bool)
    rs = g.rowptr[:-1][np.diff(g.rowptr) > 0]
    row_start[rs] = True
    return bool(np.all((d > 0) | row_start[1:]))
