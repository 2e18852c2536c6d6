"""Oracle: GCN / SAGE / LinkPredictor forward on the CPU (torch, fp32 or fp64).

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

Restates
  * ``GCN.forward``            /root/reference/models.py:181-187
  * ``SAGE.forward``           /root/reference/models.py:434-440
  * ``LinkPredictor.forward``  /root/reference/models.py:478-485
  * ``LinkGNN.forward``        /root/reference/models.py:500-506
and the torch_geometric 1.7.0 ``GCNConv`` / ``SAGEConv`` arithmetic described in
SURVEY.md Appendix A.3 / A.4 (third-party, not under /root/reference:
**parity unpinned** for that part; LinkPredictor is plain torch).

State-dict layout (filter.py:69, rank.py:359-361):
  emb.weight [n,H]; gnn.convs.{i}.weight [in,out] + .bias (GCN, PyG 1.7);
  gnn.convs.{i}.lin_l.weight [out,in], .lin_l.bias, .lin_r.weight [out,in] (SAGE);
  linkpred.lins.{i}.weight [out,in], .bias.
"""
from __future__ import annotations

import numpy as np
import torch

from .graph import CSR


def _t(x, dtype):
    return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(dtype)


def gcn_norm(g: CSR):
    """A.3: diag *set* to 1 (inserted where missing), deg = rowsum, dinv = deg^-1/2 (inf->0),
    value = (w * dinv[row]) * dinv[col] with two fp32 roundings in that order.
    Returns (rowptr, col, val_norm fp32) of the self-looped matrix, columns ascending.
    """
    n = g.n
    row = np.repeat(np.arange(n, dtype=np.int64), np.diff(g.rowptr))
    offdiag = row != g.col
    r = np.concatenate([row[offdiag], np.arange(n, dtype=np.int64)])
    c = np.concatenate([g.col[offdiag], np.arange(n, dtype=np.int64)])
    w = np.concatenate([g.val[offdiag], np.ones(n, dtype=np.float32)]).astype(np.float32)
    order = np.argsort(r * n + c, kind="stable")
    r, c, w = r[order], c[order], w[order]
    deg = np.zeros(n, dtype=np.float32)
    # fp32 row sums in column order (torch_sparse sum -> segment reduce)
    starts = np.flatnonzero(np.concatenate([[True], r[1:] != r[:-1]]))
    deg[r[starts]] = np.add.reduceat(w, starts)
    with np.errstate(divide="ignore"):
        dinv = np.power(deg, np.float32(-0.5)).astype(np.float32)
    dinv[np.isinf(dinv)] = 0
    val = ((w * dinv[r]).astype(np.float32) * dinv[c]).astype(np.float32)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(rowptr, r + 1, 1)
    return np.cumsum(rowptr), c, val


def spmm(rowptr, col, val, x: torch.Tensor, reduce: str = "sum") -> torch.Tensor:
    """Y[i] = sum_j val_ij x[col_j] (or mean over stored entries, values ignored)."""
    n = rowptr.shape[0] - 1
    row = torch.as_tensor(np.repeat(np.arange(n, dtype=np.int64), np.diff(rowptr)))
    colt = torch.as_tensor(np.asarray(col, dtype=np.int64))
    out = torch.zeros((n, x.shape[1]), dtype=x.dtype)
    if reduce == "sum":
        v = torch.as_tensor(np.asarray(val)).to(x.dtype)
        out.index_add_(0, row, x[colt] * v[:, None])
    else:
        out.index_add_(0, row, x[colt])
        cnt = torch.as_tensor(np.diff(rowptr)).to(x.dtype).clamp(min=1)
        out = out / cnt[:, None]
    return out


def spmm_sequential(rowptr, col, val, x: np.ndarray, reduce: str = "sum") -> np.ndarray:
    """Row-by-row fp32 left fold in ascending-column order with fused multiply-add emulated in
    float64-then-round (exact for one FMA step).  Small inputs only; used to check that the CUDA
    kernel keeps torch_sparse's accumulation order (A.3 order-sensitive note)."""
    n = rowptr.shape[0] - 1
    F = x.shape[1]
    out = np.zeros((n, F), dtype=np.float32)
    for i in range(n):
        acc = np.zeros(F, dtype=np.float32)
        for p in range(rowptr[i], rowptr[i + 1]):
            xv = x[col[p]].astype(np.float64)
            if reduce == "sum":
                acc = (np.float64(val[p]) * xv + acc.astype(np.float64)).astype(np.float32)
            else:
                acc = (xv + acc.astype(np.float64)).astype(np.float32)
        if reduce == "mean":
            acc = (acc / np.float32(max(rowptr[i + 1] - rowptr[i], 1))).astype(np.float32)
        out[i] = acc
    return out


def gcn_forward(g: CSR, x, sd: dict, num_layers: int, dtype=torch.float32, prefix="gnn.") -> torch.Tensor:
    rp, c, v = gcn_norm(g)
    h = _t(x, dtype)
    for i in range(num_layers):
        W = _t(sd[f"{prefix}convs.{i}.weight"], dtype)   # [in,out]
        b = _t(sd[f"{prefix}convs.{i}.bias"], dtype)
        h = spmm(rp, c, v, h @ W, "sum") + b
        if i != num_layers - 1:
            h = torch.relu(h)
    return h


def sage_forward(g: CSR, x, sd: dict, num_layers: int, dtype=torch.float32, prefix="gnn.") -> torch.Tensor:
    h = _t(x, dtype)
    for i in range(num_layers):
        Wl = _t(sd[f"{prefix}convs.{i}.lin_l.weight"], dtype)  # [out,in]
        bl = _t(sd[f"{prefix}convs.{i}.lin_l.bias"], dtype)
        Wr = _t(sd[f"{prefix}convs.{i}.lin_r.weight"], dtype)
        agg = spmm(g.rowptr, g.col, None, h, "mean")
        h = agg @ Wl.t() + bl + h @ Wr.t()
        if i != num_layers - 1:
            h = torch.relu(h)
    return h


def linkpred_forward(h: torch.Tensor, edges, sd: dict, num_layers: int, dtype=torch.float32,
                     prefix="linkpred.", return_logit: bool = False) -> torch.Tensor:
    e = torch.as_tensor(np.asarray(edges, dtype=np.int64))
    h = h.to(dtype)
    z = h[e[0]] * h[e[1]]
    for i in range(num_layers - 1):
        z = torch.relu(z @ _t(sd[f"{prefix}lins.{i}.weight"], dtype).t() + _t(sd[f"{prefix}lins.{i}.bias"], dtype))
    i = num_layers - 1
    z = z @ _t(sd[f"{prefix}lins.{i}.weight"], dtype).t() + _t(sd[f"{prefix}lins.{i}.bias"], dtype)
    z = z.reshape(-1)
    return z if return_logit else torch.sigmoid(z)


def link_gnn_input(sd: dict, x):
    """models.py:501-504: emb.weight | cat([emb.weight, x], 1) | x."""
    emb = sd.get("emb.weight")
    if x is None:
        return emb
    if emb is None:
        return x
    return torch.cat([torch.as_tensor(emb), torch.as_tensor(x)], dim=1)


def random_state_dict(model: str, n: int, f_in_feat: int, hidden: int, num_layers: int,
                      use_emb: bool = True, seed: int = 1234) -> dict:
    """Seeded weights with the reference's default inits: glorot GCN weight / zero
    bias, Kaiming-uniform nn.Linear, N(0,1) embedding."""
    gen = torch.Generator().manual_seed(seed)
    sd = {}
    f_in = f_in_feat + (hidden if use_emb else 0)
    if use_emb:
        sd["emb.weight"] = torch.randn(n, hidden, generator=gen)

    def lin(out_c, in_c, bias=True, key=""):
        bound = 1.0 / np.sqrt(in_c)
        sd[key + ".weight"] = (torch.rand(out_c, in_c, generator=gen) * 2 - 1) * bound
        if bias:
            sd[key + ".bias"] = (torch.rand(out_c, generator=gen) * 2 - 1) * bound

    for i in range(num_layers):
        ic = f_in if i == 0 else hidden
        if model == "gcn":
            a = np.sqrt(6.0 / (ic + hidden))
            sd[f"gnn.convs.{i}.weight"] = (torch.rand(ic, hidden, generator=gen) * 2 - 1) * a
            # a trained checkpoint has non-zero bias; keep it small but non-zero so the test sees it
            sd[f"gnn.convs.{i}.bias"] = (torch.rand(hidden, generator=gen) * 2 - 1) * 0.05
        else:
            lin(hidden, ic, True, f"gnn.convs.{i}.lin_l")
            lin(hidden, ic, False, f"gnn.convs.{i}.lin_r")
    for i in range(num_layers):
        oc = 1 if i == num_layers - 1 else hidden
        lin(oc, hidden, True, f"linkpred.lins.{i}")
    return sd
