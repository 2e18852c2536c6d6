"""Multi-GPU sharding of the filter step: one process per GPU.

* candidates are owned by contiguous ranges of v = all_edges[:,1] of equal 2-path work
  (``partition_by_work``); the graph is replicated and no collective runs while scoring;
* the GNN embeddings are computed ROW-SHARDED (``sharded_gnn_embed``): every rank runs the dense
  x·W blocks and the K1 SpMM for its block-aligned row range only and the ``n/G x H`` slabs are
  all-gathered per layer, so that every rank again holds the full ``h`` — bit-identical to the
  single-GPU embeddings (same row blocks, same per-row accumulation order);
* the per-GPU proposal lists are merged after an exchange of the GLOBAL k-th score
  (``global_kth_key``: two all-reduced 65,536-bin histograms): only rows that can be in the global
  top-k are gathered, then one K4 select runs on every rank (``merge_topk``).

``torch.distributed`` is the plumbing (NCCL over NVLink on the GPUs, gloo in the CPU tests); the
raw-NCCL C-ABI variant of the plain all-gather merge is ``eps_topk_merge_allgather`` (csrc/comm.cu).
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


# ------------------------------------------------------------------------------------------------
# row-sharded GNN embeddings
# ------------------------------------------------------------------------------------------------

def row_block(n: int) -> int:
    """Row-block size of the dense x·W products: a function of n ONLY (never of the world size), so a
    row's block — hence the cuBLAS problem shape that produces it, hence its bits — is the same on 1
    and on G GPUs.  8,192 rows for large graphs, n/64 rounded down to a power of two (>= 256) below."""
    if n >= 8192 * 64:
        return 8192
    r = 256
    while r * 2 * 64 <= n:
        r *= 2
    return r


def row_partition(rowptr: torch.Tensor, n: int, parts: int) -> List[int]:
    """Block-aligned row ranges of near-equal nnz: parts+1 boundaries (multiples of row_block(n))."""
    R = row_block(n)
    nb = (n + R - 1) // R
    if parts <= 1:
        return [0, n]
    ends = torch.arange(1, nb + 1, device=rowptr.device).mul_(R).clamp_(max=n)
    cs = rowptr[ends.long()].double().cpu()                 # nnz before the end of every block
    total = float(cs[-1]) if nb else 0.0
    bounds = [0]
    for p in range(1, parts):
        b = int(torch.searchsorted(cs, torch.tensor(total * p / parts, dtype=cs.dtype)).item()) + 1
        bounds.append(min(max(b, bounds[-1]), nb))
    bounds.append(nb)
    return [min(b * R, n) for b in bounds]


def block_matmul(x: torch.Tensor, w: torch.Tensor, lo: int, hi: int, n: int,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """rows [lo, hi) of ``x_full @ w`` where ``x`` holds exactly those rows, computed block by block
    (global blocks of row_block(n) rows; lo is block-aligned).  Library GEMM (cuBLAS fp32)."""
    R = row_block(n)
    assert lo % R == 0 and x.shape[0] == hi - lo
    y = out if out is not None else torch.empty((hi - lo, w.shape[1]), dtype=x.dtype, device=x.device)
    for b0 in range(lo, hi, R):
        b1 = min(b0 + R, hi)
        torch.mm(x[b0 - lo:b1 - lo], w, out=y[b0 - lo:b1 - lo])
    return y


def allgather_row_slabs(local: torch.Tensor, bounds: List[int], group=None) -> torch.Tensor:
    """rows [bounds[r], bounds[r+1]) from every rank r -> the full [n, F] matrix on every rank.
    One ``all_gather_into_tensor`` of slabs padded to the largest range (NCCL wants equal sizes).  The number
    of ranks is what ``bounds`` says (a non-distributed call inside an initialised process group has
    ``bounds == [0, n]`` and gathers nothing)."""
    world = len(bounds) - 1
    n, F = bounds[-1], local.shape[1]
    if world == 1:
        return local
    sizes = [bounds[r + 1] - bounds[r] for r in range(world)]
    mx = max(sizes)
    if all(s == mx for s in sizes):
        out = torch.empty((n, F), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous(), group=group)
        return out
    send = local
    if local.shape[0] != mx:
        send = torch.zeros((mx, F), dtype=local.dtype, device=local.device)
        send[: local.shape[0]] = local
    buf = torch.empty((world * mx, F), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, send.contiguous(), group=group)
    out = torch.empty((n, F), dtype=local.dtype, device=local.device)
    for r in range(world):
        out[bounds[r]:bounds[r + 1]] = buf[r * mx: r * mx + sizes[r]]
    return out


# ------------------------------------------------------------------------------------------------
# proposal-list merge
# ------------------------------------------------------------------------------------------------

def order_keys(score: torch.Tensor) -> torch.Tensor:
    """K4's order key of every score as int64 in [0, 2^32): ascending key == descending score,
    -0.0 folded into +0.0 (csrc/topk.cu score_key).  Torch ops only (device-agnostic)."""
    b = (score.float() + 0.0).contiguous().view(torch.int32).long() & 0xFFFFFFFF
    asc = torch.where(b >= 2**31, b ^ 0xFFFFFFFF, b ^ 0x80000000)
    return 0xFFFFFFFF - asc


def global_kth_key(score_local: torch.Tensor, k: int, group=None) -> torch.Tensor:
    """Order key (int32 [1] on the scores' device, uint32 payload) of the k-th best score over ALL ranks'
    ``score_local``: a two-level radix select whose two 65,536-bin histograms are summed with
    ``all_reduce``.  0xFFFFFFFF (keep everything) when fewer than k scores exist in total."""
    keys = order_keys(score_local)
    dev = score_local.device
    hi = keys >> 16
    h1 = torch.bincount(hi, minlength=65536)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(h1, group=group)
    cs = torch.cumsum(h1, 0)
    if int(cs[-1]) < k:
        return torch.tensor([-1], dtype=torch.int32, device=dev)
    b = int(torch.searchsorted(cs, torch.tensor(k, device=dev, dtype=cs.dtype)).item())
    k_rem = k - (int(cs[b - 1]) if b else 0)
    lo = keys[hi == b] & 0xFFFF
    h2 = torch.bincount(lo, minlength=65536)
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(h2, group=group)
    cs2 = torch.cumsum(h2, 0)
    b2 = int(torch.searchsorted(cs2, torch.tensor(k_rem, device=dev, dtype=cs2.dtype)).item())
    key = (b << 16) | b2
    return torch.tensor([key - 2**32 if key >= 2**31 else key], dtype=torch.int32, device=dev)


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
               select=None, exchange_threshold: bool = True) -> torch.Tensor:
    """Global top-k of the per-rank sorted (u, v, score) lists; identical on every rank.

    With ``exchange_threshold`` the ranks first agree on the global k-th score (``global_kth_key``) and
    gather only their rows that reach it — a PREFIX of each sorted list, ties at the k-th score
    included — so ~k rows travel instead of world*k.  ``select(score, k) -> idx`` is the K4 select
    (ops.topk on the GPU); because ranks own ascending owner ranges and every list is sorted with ties in
    candidate order, position in the gathered array breaks ties exactly like the global candidate index."""
    rank, world = world_info()
    k_local = k if k_local is None else k_local
    rows = local_sorted[:k_local]
    if world > 1 and exchange_threshold:
        gkey = int(global_kth_key(rows[:, 2].contiguous(), k, group).item()) & 0xFFFFFFFF
        keep = int((order_keys(rows[:, 2]) <= gkey).sum().item())          # a prefix: the list is sorted
        cnt = torch.tensor([keep], dtype=torch.int64, device=rows.device)
        dist.all_reduce(cnt, op=dist.ReduceOp.MAX, group=group)
        k_local = max(int(cnt.item()), 1)
        rows = rows[:keep]
    gathered = allgather_rows(pad_rows(rows, k_local), group)
    if select is None:
        from . import ops
        select = lambda s, kk: ops.topk(s, kk)[0]
    valid = int((gathered[:, 2] > -math.inf).sum().item())
    idx = select(gathered[:, 2].contiguous(), min(k, valid))
    return gathered[idx]
