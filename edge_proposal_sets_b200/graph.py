"""Graph container for the scoring path: the symmetric adjacency every kernel reads.

Mirrors what the reference gets from torch_sparse (``SparseTensor.from_edge_index``,
``to_symmetric``, ``fill_value``) through ``rank.add_edges`` (/root/reference/rank.py:28-36),
plus the derived tables the kernels consume:

* ``SparseAdj``          CSR with int32 ``rowptr`` / ``col`` (ascending inside a row, coalesced)
                         and optional fp32 ``val`` (``None`` = all ones, every dataset but collab);
* ``SparseAdj.gcn_norm`` the self-looped, symmetrically normalised matrix GCNConv multiplies by
                          — built ONCE per graph, not once per batch as the
                         reference's ``cached=False`` GCNConv does (/root/reference/models.py:169-173);
* weight tables for Adamic-Adar / Resource-Allocation (/root/reference/adamic_utils.py:15-16,
  models.py:546-550, train_and_eval.py:203-204).

Only torch tensor ops are used here (device-agnostic plumbing: sort / unique / cumsum), so the
same code builds the graph on ``cuda:k`` for the kernels and on the CPU for host-logic tests.
"""
from __future__ import annotations

import itertools
from typing import Optional

import torch

_UID = itertools.count(1)


class SparseAdj:
    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor], n: int):
        assert rowptr.dtype == torch.int32 and col.dtype == torch.int32
        assert val is None or val.dtype == torch.float32
        self.rowptr, self.col, self.val, self.n = rowptr.contiguous(), col.contiguous(), val, int(n)
        if self.val is not None:
            self.val = self.val.contiguous()
        self._cache = {}
        self.uid = next(_UID)           # identity of this graph object for caches (never reused, unlike id())

    # -- torch_sparse.SparseTensor look-alikes used on the path -------------------------------
    @property
    def device(self):
        return self.col.device

    @property
    def nnz(self) -> int:
        return int(self.col.numel())

    def sparse_sizes(self):
        return (self.n, self.n)

    def to(self, device) -> "SparseAdj":
        device = torch.device(device)
        if device == self.device:
            return self
        return SparseAdj(self.rowptr.to(device), self.col.to(device),
                         None if self.val is None else self.val.to(device), self.n)

    def cpu(self) -> "SparseAdj":
        return self.to("cpu")

    def row(self) -> torch.Tensor:
        deg = (self.rowptr[1:] - self.rowptr[:-1]).long()
        return torch.repeat_interleave(torch.arange(self.n, device=self.device), deg)

    def values(self) -> torch.Tensor:
        if self.val is None:
            return torch.ones(self.nnz, dtype=torch.float32, device=self.device)
        return self.val

    def coo(self):
        """(row, col, value) as int64/int64/fp32 — what ``adamic_utils.get_A`` unpacks
        (/root/reference/adamic_utils.py:9)."""
        return self.row(), self.col.long(), self.values()

    def degree(self) -> torch.Tensor:
        return (self.rowptr[1:] - self.rowptr[:-1])

    def sum(self, dim: int = -1) -> torch.Tensor:
        """Row sums (``adj.sum(-1)``, /root/reference/models.py:546); the matrix is symmetric so
        column sums (``A.sum(0)``, adamic_utils.py:15) are the same numbers."""
        if self.val is None:
            return self.degree().float()
        out = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        return out.index_add_(0, self.row(), self.val)

    # -- derived tables ------------------------------------------------------------------------
    def gcn_norm(self):
        """(rowptr, col, val) of D^-1/2 (A with diag := 1) D^-1/2, int32/int32/fp32."""
        if "gcn" in self._cache:
            return self._cache["gcn"]
        if self.col.is_cuda:
            self._cache["gcn"] = self._gcn_norm_cuda()
            return self._cache["gcn"]
        # host tensors (CPU tests of the host logic): the same construction in torch ops
        n, dev = self.n, self.device
        row, col, w = self.row(), self.col.long(), self.values()
        off = row != col
        ar = torch.arange(n, device=dev)
        r = torch.cat([row[off], ar])
        c = torch.cat([col[off], ar])
        w = torch.cat([w[off], torch.ones(n, dtype=torch.float32, device=dev)])
        order = torch.argsort(r * n + c)
        r, c, w = r[order], c[order], w[order]
        deg = torch.zeros(n, dtype=torch.float32, device=dev).index_add_(0, r, w)
        dinv = deg.pow(-0.5)
        dinv[torch.isinf(dinv)] = 0
        val = (w * dinv[r]) * dinv[c]
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        rowptr[1:] = torch.cumsum(torch.bincount(r, minlength=n), 0)
        out = (rowptr.int(), c.int().contiguous(), val.contiguous())
        self._cache["gcn"] = out
        return out

    def _gcn_norm_cuda(self):
        """K1b kernels (csrc/gcn_norm.cu): count -> prefix sum -> fill."""
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        n, dev = self.n, self.device
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        dinv = torch.empty(n, dtype=torch.float32, device=dev)
        newlen = torch.empty(n, dtype=torch.int32, device=dev)
        _lib.check(lib.eps_gcn_norm_count(p(self.rowptr), p(self.col), p(self.val), n, p(dinv), p(newlen), st),
                   "eps_gcn_norm_count")
        rowptr2 = torch.zeros(n + 1, dtype=torch.int32, device=dev)
        torch.cumsum(newlen, 0, out=rowptr2[1:])
        nnz2 = self.nnz + n      # upper bound (every row gains at most the diagonal): no host sync
        col2 = torch.empty(nnz2, dtype=torch.int32, device=dev)
        val2 = torch.empty(nnz2, dtype=torch.float32, device=dev)
        _lib.check(lib.eps_gcn_norm_fill(p(self.rowptr), p(self.col), p(self.val), n, p(dinv), p(rowptr2),
                                         p(col2), p(val2), st), "eps_gcn_norm_fill")
        return rowptr2, col2, val2

    def gcn_norm_transposed_values(self):
        """Values of (GCN-normalised matrix)^T in its own CSR order, for the SpMM backward
        (autograd.py).  Unweighted: dinv_i*dinv_j is commutative -> ``None`` (same array)."""
        if self.val is None:
            return None
        if "gcn_t" not in self._cache:
            from .autograd import transposed_values
            rowptr, col, val = self.gcn_norm()
            self._cache["gcn_t"] = transposed_values(rowptr, col, val, self.n)
        return self._cache["gcn_t"]

    def inv_degree(self) -> torch.Tensor:
        """1/row-length (0 for isolated nodes): the scale of SAGEConv's neighbour mean."""
        if "inv_deg" not in self._cache:
            d = self.degree().float()
            self._cache["inv_deg"] = torch.where(d > 0, 1.0 / d, torch.zeros_like(d))
        return self._cache["inv_deg"]

    def aa_ogb_weights(self) -> torch.Tensor:
        """1/log(A.sum(0)), inf -> 0 (/root/reference/adamic_utils.py:15-16)."""
        if "aa_ogb" not in self._cache:
            w = 1.0 / torch.log(self.sum(0))
            w[torch.isinf(w)] = 0
            self._cache["aa_ogb"] = w.contiguous()
        return self._cache["aa_ogb"]

    def adamic_weights(self) -> torch.Tensor:
        """1/log(adj.sum(-1) + 1e-6) (/root/reference/models.py:546,550)."""
        if "adamic" not in self._cache:
            self._cache["adamic"] = (1.0 / torch.log(self.sum(-1) + 1e-6)).contiguous()
        return self._cache["adamic"]

    def ra_weights(self) -> torch.Tensor:
        """1/A.sum(0), inf -> 0 (/root/reference/train_and_eval.py:203-204; fp64 there, the
        table is rounded to fp32 once here)."""
        if "ra" not in self._cache:
            w = 1.0 / self.sum(0).double()
            w[torch.isinf(w)] = 0
            self._cache["ra"] = w.float().contiguous()
        return self._cache["ra"]


def add_edges(dataset: str, edge_index: torch.Tensor, edge_weight: torch.Tensor,
              extra_edges: torch.Tensor, num_nodes: int) -> SparseAdj:
    """rank.add_edges (/root/reference/rank.py:28-36) without torch_sparse.

    ``cat(edge_index, extra_edges)`` / ``cat(edge_weight, ones)``; symmetrise as the multiset
    {(r,c,w)} U {(c,r,w)} with equal (r,c) SUMMED (``to_symmetric``); reset all values to 1
    unless ``dataset == "collab"``.  The result lives on ``edge_index.device``.
    """
    dev = edge_index.device
    n = int(num_nodes)
    ei = edge_index.reshape(2, -1).long()
    ex = extra_edges.reshape(2, -1).long().to(dev)
    full = torch.cat([ei, ex], dim=1)
    keep_w = dataset == "collab"
    r = torch.cat([full[0], full[1]])
    c = torch.cat([full[1], full[0]])
    if full.numel():
        assert int(full.min()) >= 0 and int(full.max()) < n, "node id out of range"
    key = r * n + c
    if keep_w:
        w = torch.cat([edge_weight.reshape(-1).float().to(dev),
                       torch.ones(ex.shape[1], dtype=torch.float32, device=dev)])
        assert w.numel() == full.shape[1]
        ukey, inv = torch.unique(key, sorted=True, return_inverse=True)
        val = torch.zeros(ukey.numel(), dtype=torch.float32, device=dev).index_add_(0, inv, torch.cat([w, w]))
    else:
        ukey = torch.unique(key, sorted=True)
        val = None
    row = torch.div(ukey, n, rounding_mode="floor")
    col = (ukey - row * n).int()
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    if row.numel():
        rowptr[1:] = torch.cumsum(torch.bincount(row, minlength=n), 0)
    assert ukey.numel() < 2**31, "nnz must fit int32"
    return SparseAdj(rowptr.int(), col.contiguous(), val, n)


def from_scipy(A, keep_values: bool = False) -> SparseAdj:
    A = A.tocsr()
    A.sort_indices()
    val = torch.from_numpy(A.data.astype("float32")) if keep_values else None
    return SparseAdj(torch.from_numpy(A.indptr.astype("int32")), torch.from_numpy(A.indices.astype("int32")),
                     val, A.shape[0])
