"""Oracle: Common-Neighbour / Adamic-Adar / Resource-Allocation pair scores.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

Restates
  * ``CommonNeighborsPredictor.forward`` type 'simple'  /root/reference/models.py:536-542
  * ``CommonNeighborsPredictor.forward`` type 'adamic'  /root/reference/models.py:544-554
  * ``adamic_utils.get_A`` / ``AA`` ('adamic_ogb')      /root/reference/adamic_utils.py:8-25
  * ``train_and_eval.resource_allocation``              /root/reference/train_and_eval.py:195-216

Two independent formulations are kept:
  ``*_pairs``      a vectorised sorted-key membership algorithm written for the
                   oracle (no scipy fancy indexing), used by the parity tests;
  ``aa_scipy``     the reference's own scipy formulation (row-index two CSR
                   matrices, element-wise multiply, row sum), used to time the
                   CPU baseline because that is what the reference executes.
Both are pinned against the real ``adamic_utils.AA`` in tests/golden.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as ssp

from .graph import CSR


# ----------------------------------------------------------------------------
# weight tables
# ----------------------------------------------------------------------------

def aa_ogb_weights(g: CSR) -> np.ndarray:
    """adamic_utils.py:15-16: ``1/np.log(A.sum(0))`` in fp32, inf -> 0."""
    colsum = np.asarray(g.to_scipy().sum(0)).reshape(-1).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = (np.float32(1.0) / np.log(colsum)).astype(np.float32)
    w[np.isinf(w)] = 0
    return w


def adamic_gpu_weights(g: CSR) -> np.ndarray:
    """models.py:546,550: ``1/log(adj.sum(-1) + 1e-6)`` in fp32 (row sums)."""
    rowsum = np.asarray(g.to_scipy().sum(1)).reshape(-1).astype(np.float32)
    deg = (rowsum + np.float32(1e-6)).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = (np.float32(1.0) / np.log(deg)).astype(np.float32)
    return w


def ra_weights(g: CSR) -> np.ndarray:
    """train_and_eval.py:203-204: ``1/A.sum(axis=0)``, inf -> 0 (float64 in the reference)."""
    colsum = np.asarray(g.to_scipy().sum(0)).reshape(-1).astype(np.float64)
    with np.errstate(divide="ignore"):
        w = 1.0 / colsum
    w[np.isinf(w)] = 0
    return w


# ----------------------------------------------------------------------------
# membership-based pair scorer (oracle's own formulation)
# ----------------------------------------------------------------------------

def _common_neighbour_terms(g: CSR, edges: np.ndarray):
    """For pairs ``edges[2,B]`` list every (pair, k, A[u,k], A[v,k]) with k in N(u)&N(v).

    Walks N(u) for each pair and looks the key (v, k) up in the sorted CSR key
    array.  Terms come out ordered by (pair, k ascending).
    """
    u = np.asarray(edges[0], dtype=np.int64)
    v = np.asarray(edges[1], dtype=np.int64)
    B = u.shape[0]
    n = g.n
    keys = np.repeat(np.arange(n, dtype=np.int64), np.diff(g.rowptr)) * n + g.col
    start = g.rowptr[u]
    ln = g.rowptr[u + 1] - start
    tot = int(ln.sum())
    pair = np.repeat(np.arange(B, dtype=np.int64), ln)
    off = np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(ln) - ln, ln)
    pos_u = np.repeat(start, ln) + off
    k = g.col[pos_u]
    q = v[pair] * n + k
    pos_v = np.searchsorted(keys, q)
    pos_v[pos_v >= keys.size] = max(keys.size - 1, 0)
    hit = keys[pos_v] == q if keys.size else np.zeros(tot, dtype=bool)
    return pair[hit], k[hit], g.val[pos_u[hit]], g.val[pos_v[hit]], B


def _seq_fp32_segment_sum(pair: np.ndarray, terms: np.ndarray, B: int, order: str = "numpy") -> np.ndarray:
    """Sum fp32 ``terms`` per pair (terms arrive grouped by pair, k ascending).

    order="numpy"       ``np.add.reduceat`` — the routine scipy's row sum reaches
                        (``_minor_reduce``), i.e. bit-for-bit what the reference computes: numpy
                        evaluates t0 + (t1 + t2 + ...) and switches to an 8-way unrolled pairwise
                        scheme inside the bracket from 8 terms on.
    order="sequential"  strict left fold ((t0 + t1) + t2) + ... in fp32; identical bits to
                        "numpy" up to 2 terms, within a few ulp beyond.
    order="exact"       what the CUDA kernels compute (include/eps.h, K3 and K6+K3): every fp32 term
                        becomes the integer RN(t * 2^38), the integers are added exactly, and the sum
                        is rounded ONCE to fp32 — order-independent; for terms >= 2^-15 it is the
                        correctly rounded exact sum (measured <= 2.4e-7 relative from "numpy" on twitch).
    """
    out = np.zeros(B, dtype=np.float32)
    if pair.size == 0:
        return out
    terms = terms.astype(np.float32)
    first = np.concatenate([[True], pair[1:] != pair[:-1]])
    starts = np.flatnonzero(first)
    if order == "numpy":
        out[pair[starts]] = np.add.reduceat(terms, starts)
        return out
    if order == "exact":
        fx = np.rint(terms.astype(np.float64) * float(1 << FX_FRAC_BITS)).astype(np.int64)   # exact: 24-bit x 2^38
        sums = np.add.reduceat(fx, starts)
        assert np.all(np.abs(sums) < (1 << 53)), "oracle: fixed-point sum beyond exact float64 range"
        # int64 -> float64 exact below 2^53, float64 -> float32 is ONE round-to-nearest-even
        out[pair[starts]] = (sums.astype(np.float64) * 2.0 ** -FX_FRAC_BITS).astype(np.float32)
        return out
    lens = np.diff(np.concatenate([starts, [pair.size]]))
    acc = terms[starts].copy()
    for j in range(1, int(lens.max())):
        live = lens > j
        acc[live] = (acc[live] + terms[starts[live] + j]).astype(np.float32)
    out[pair[starts]] = acc
    return out


FX_FRAC_BITS = 38   # edge_proposal_sets_b200/csrc/eps_common.cuh EPS_FX_FRAC_BITS


def _batched(fn, edges: np.ndarray, batch: int) -> np.ndarray:
    edges = np.asarray(edges)
    B = edges.shape[1]
    outs = [fn(edges[:, i:i + batch]) for i in range(0, B, batch)]
    if not outs:
        return np.zeros(0, dtype=np.float32)
    return np.concatenate(outs)


def cn_count_pairs(g: CSR, edges: np.ndarray, batch: int = 1 << 16) -> np.ndarray:
    """|N(u) & N(v)| as int32 (structure only, weights ignored)."""
    def one(e):
        pair, _, _, _, B = _common_neighbour_terms(g, e)
        return np.bincount(pair, minlength=B).astype(np.int32)
    edges = np.asarray(edges)
    if edges.shape[1] == 0:
        return np.zeros(0, dtype=np.int32)
    return np.concatenate([one(edges[:, i:i + batch]) for i in range(0, edges.shape[1], batch)])


def cn_scores_pairs(g: CSR, edges: np.ndarray, batch: int = 1 << 16, order: str = "numpy") -> np.ndarray:
    """models.py:536-542 ('simple'): sum_k A[u,k]*A[v,k] in fp32, no sigmoid."""
    def one(e):
        pair, _, au, av, B = _common_neighbour_terms(g, e)
        return _seq_fp32_segment_sum(pair, (au * av).astype(np.float32), B, order)
    return _batched(one, edges, batch)


def adamic_sigmoid_pairs(g: CSR, edges: np.ndarray, batch: int = 1 << 16,
                         return_presigmoid: bool = False, order: str = "numpy") -> np.ndarray:
    """models.py:544-554 ('adamic'): sigmoid(sum_{k in CN} 1/log(deg_k + 1e-6)).

    Only the *indices* of the common neighbours are used (models.py:544), so
    edge weights enter through ``deg`` alone.
    """
    w = adamic_gpu_weights(g)

    def one(e):
        pair, k, _, _, B = _common_neighbour_terms(g, e)
        return _seq_fp32_segment_sum(pair, w[k], B, order)
    s = _batched(one, edges, batch)
    if return_presigmoid:
        return s
    return (np.float32(1.0) / (np.float32(1.0) + np.exp(-s, dtype=np.float32))).astype(np.float32)


def aa_ogb_pairs(g: CSR, edges: np.ndarray, batch: int = 1 << 16, order: str = "numpy",
                 weights: np.ndarray | None = None) -> np.ndarray:
    """adamic_utils.py:13-25 ('adamic_ogb'): sum_k A[u,k] * (A[v,k] * w_k), fp32, no sigmoid."""
    w = aa_ogb_weights(g) if weights is None else weights

    def one(e):
        pair, k, au, av, B = _common_neighbour_terms(g, e)
        terms = (au * (av * w[k]).astype(np.float32)).astype(np.float32)
        return _seq_fp32_segment_sum(pair, terms, B, order)
    return _batched(one, edges, batch)


def ra_pairs(g: CSR, edges: np.ndarray, batch: int = 1 << 16) -> np.ndarray:
    """train_and_eval.py:195-216: sum_k A[u,k]*A[v,k]/colsum_k in float64, cast to fp32 at the end."""
    w = ra_weights(g)

    def one(e):
        pair, k, au, av, B = _common_neighbour_terms(g, e)
        out = np.zeros(B, dtype=np.float64)
        np.add.at(out, pair, au.astype(np.float64) * (av.astype(np.float64) * w[k]))
        return out.astype(np.float32)
    return _batched(one, edges, batch)


# ----------------------------------------------------------------------------
# the reference's own scipy formulation (CPU baseline)
# ----------------------------------------------------------------------------

def aa_scipy(g: CSR, edges: np.ndarray, batch_size: int = 2000, weights: str = "aa") -> np.ndarray:
    """What ``adamic_utils.AA`` (batch 2000) / ``resource_allocation`` execute:
    scale the columns of A by the weight table once, then per batch row-select
    A[src] and A_[dst], multiply element-wise and row-sum.  Single-threaded
    scipy sparsetools, as in the reference (adamic_utils.py:17-23).
    """
    A = g.to_scipy()
    if weights == "aa":
        A_ = A.multiply(aa_ogb_weights(g)).tocsr()
    elif weights == "ra":
        A = A.astype(np.float64)
        A_ = A.multiply(ra_weights(g)).tocsr()
    elif weights == "cn":
        A_ = A
    else:
        raise ValueError(weights)
    src_all = np.asarray(edges[0], dtype=np.int64)
    dst_all = np.asarray(edges[1], dtype=np.int64)
    out = []
    for i in range(0, src_all.shape[0], batch_size):
        src, dst = src_all[i:i + batch_size], dst_all[i:i + batch_size]
        out.append(np.asarray(A[src].multiply(A_[dst]).sum(1)).reshape(-1))
    if not out:
        return np.zeros(0, dtype=np.float32)
    return np.concatenate(out).astype(np.float32)


def cn_loops(g: CSR, u: int, v: int):
    """Pure-python two-pointer merge for tiny hand-checked cases: (count, [k...])."""
    a = g.col[g.rowptr[u]:g.rowptr[u + 1]]
    b = g.col[g.rowptr[v]:g.rowptr[v + 1]]
    i = j = 0
    ks = []
    while i < len(a) and j < len(b):
        if a[i] == b[j]:
            ks.append(int(a[i])); i += 1; j += 1
        elif a[i] < b[j]:
            i += 1
        else:
            j += 1
    return len(ks), ks
