"""Multi-GPU sharding of the filter step: one process per GPU, candidates owned by
contiguous ranges of v = all_edges[:,1], graph and embeddings replicated, and exactly one
exchange step — an all-gather of the per-GPU top-k rows followed by the same K4 select on every
rank.  No data-path collective is needed while scoring.

``torch.distributed`` is the plumbing (NCCL over NVLink on the GPUs, gloo in the CPU tests); the
raw-NCCL C-ABI variant of the same merge is ``eps_topk_merge_allgather`` (csrc/comm.cu).
"""
from __future__ import annotations

import math
from typing import List, Optional

import torch
import torch.distributed as dist


def partition_by_work(work: torch.Tensor, parts: int) -> List[int]:
    """Cut [0, n) into ``parts`` contiguous ranges of near-equal total ``work``; returns parts+1
    boundaries.  ``work`` is the per-owner 2-path count (candidates.two_path_work)."""
    n = work.numel()
    if parts <= 1 or n == 0:
        return [0, n]
    cs = torch.cumsum(work.double().cpu(), 0)
    total = float(cs[-1])
    bounds = [0]
    for p in range(1, parts):
        target = total * p / parts
        b = int(torch.searchsorted(cs, torch.tensor(target, dtype=cs.dtype)).item()) + 1
        bounds.append(min(max(b, bounds[-1]), n))
    bounds.append(n)
    return bounds


def world_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def pad_rows(rows: torch.Tensor, k_local: int) -> torch.Tensor:
    """Pad a sorted [m,3] (u, v, score) list to exactly k_local rows with score = -inf."""
    m = rows.shape[0]
    if m == k_local:
        return rows.contiguous()
    assert m < k_local
    pad = torch.zeros((k_local - m, 3), dtype=rows.dtype, device=rows.device)
    pad[:, 2] = -math.inf
    return torch.cat([rows, pad], 0).contiguous()


def allgather_rows(local: torch.Tensor, group=None) -> torch.Tensor:
    """[k_local,3] on every rank -> [world*k_local,3] in rank order (NCCL or gloo)."""
    rank, world = world_info()
    if world == 1:
        return local
    out = torch.empty((world * local.shape[0], 3), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def merge_topk(local_sorted: torch.Tensor, k: int, k_local: Optional[int] = None, group=None,
               select=None) -> torch.Tensor:
    """Global top-k of the per-rank sorted (u, v, score) lists; identical on every rank.

    ``select(score, k) -> idx`` is the K4 select (ops.topk on the GPU); because ranks own
    ascending owner ranges, position in the gathered array breaks ties exactly like the global
    candidate index does."""
    k_local = k if k_local is None else k_local
    gathered = allgather_rows(pad_rows(local_sorted[:k_local], k_local), group)
    if select is None:
        from . import ops
        select = lambda s, kk: ops.topk(s, kk)[0]
    valid = int((gathered[:, 2] > -math.inf).sum().item())
    idx = select(gathered[:, 2].contiguous(), min(k, valid))
    return gathered[idx]
