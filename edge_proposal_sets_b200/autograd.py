"""Autograd wrappers over the K1 SpMM kernel, so the reference's ``train()`` loop
(/root/reference/train_and_eval.py:31-96) runs on the same aggregation kernel as the scoring path
(SURVEY §8f row 4).

  forward   y = A · x            K1 (``eps_spmm_csr_f32``), exactly the inference arithmetic
  backward  dx = Aᵀ · dy         K1 again, on the transposed matrix

Every adjacency on this path is structurally symmetric (``add_edges`` symmetrises, the GCN
normalisation adds the diagonal), so Aᵀ shares ``rowptr`` / ``col`` with A and only the VALUES have
to be transposed:
  * unweighted GCN norm  val_ij = dinv_i·dinv_j  — commutative, the same array;
  * weighted GCN norm    val_ij = (w·dinv_i)·dinv_j — permuted once per graph (``transposed_values``);
  * SAGE mean            A = D⁻¹·S with S the 0/1 structure  =>  Aᵀ·dy = S·(D⁻¹ dy): the rows of dy
                         are scaled by 1/deg and summed with the structure-only kernel.
The LinkPredictor's input z0 = h[u] * h[v] and its backward (scatter-add into dh) are K7
(``eps_pair_hadamard_f32`` / ``_bwd_f32``): one fused pass each instead of two index_select + mul and
their three autograd backward launches.  The dense layers of a training step (x·W, the
LinkPredictor's Linear layers, Adam) stay in torch / cuBLAS: plain library GEMMs.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


def transposed_values(rowptr: torch.Tensor, col: torch.Tensor, val: torch.Tensor, n: int) -> torch.Tensor:
    """Values of Aᵀ in A's own CSR order, for a structurally symmetric A: entry p = (i, j) of the
    CSR receives the value stored at (j, i)."""
    nnz = int(rowptr[-1].item())
    deg = (rowptr[1:] - rowptr[:-1]).long()
    row = torch.repeat_interleave(torch.arange(n, device=col.device), deg)
    c = col[:nnz].long()
    perm = torch.argsort(c * n + row)          # (j, i) sorted ascending == CSR order of the transpose
    out = val.clone()
    out[:nnz] = val[:nnz][perm]
    return out


class _SpMM(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rowptr, col, val, val_t, reduce, inv_deg):
        ctx.rowptr, ctx.col, ctx.val_t, ctx.reduce, ctx.inv_deg = rowptr, col, val_t, reduce, inv_deg
        return ops.spmm_csr(rowptr, col, val, x, reduce)

    @staticmethod
    def backward(ctx, gy):
        gy = gy.contiguous()
        if ctx.reduce == "mean":
            gy = gy * ctx.inv_deg[:, None]     # D^-1 dy, then the structure-only sum
            gx = ops.spmm_csr(ctx.rowptr, ctx.col, None, gy, "sum")
        else:
            gx = ops.spmm_csr(ctx.rowptr, ctx.col, ctx.val_t, gy, "sum")
        return gx, None, None, None, None, None, None


def spmm(x: torch.Tensor, rowptr: torch.Tensor, col: torch.Tensor, val: Optional[torch.Tensor],
         reduce: str = "sum", val_t: Optional[torch.Tensor] = None,
         inv_deg: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Differentiable (w.r.t. ``x``) K1 SpMM.  ``val_t`` = values of the transpose in CSR order
    (``None``: the matrix is numerically symmetric); ``inv_deg`` = 1/row-length for ``reduce="mean"``
    (0 for empty rows)."""
    return _SpMM.apply(x, rowptr, col, val, val if val_t is None else val_t, reduce, inv_deg)


class _PairHadamard(torch.autograd.Function):
    @staticmethod
    def forward(ctx, h, edges):
        h = h.contiguous()
        ctx.save_for_backward(h)
        ctx.edges = edges
        return ops.pair_hadamard(h, edges)

    @staticmethod
    def backward(ctx, gz):
        (h,) = ctx.saved_tensors
        return ops.pair_hadamard_bwd(h, ctx.edges, gz), None


def pair_hadamard(h: torch.Tensor, edges: torch.Tensor) -> torch.Tensor:
    """Differentiable (w.r.t. ``h``) K7: h[edges[0]] * h[edges[1]] -> [B,H]."""
    return _PairHadamard.apply(h, edges)
