"""One training epoch of a parameterised filter / rank model on the B200 kernels — the
counterpart of ``train_and_eval.train`` (/root/reference/train_and_eval.py:31-96), so that
``rank.py --save_models`` can produce the checkpoint ``filter.py`` consumes without torch_geometric
(SURVEY §8f row 4).

  batches            ``DataLoader(range(E), batch_size, shuffle=True)``      -> ``torch.randperm`` chunks
  positives          ``to_undirected(pos_train_edge[perm].t())``  [PyG]      -> ``to_undirected`` below
  negatives          gcn/sage: PyG ``negative_sampling(method='dense')`` (uniform distinct NON-edges of
                     the current ``adj_t``, self pairs allowed, possibly fewer than asked);
                     collab: ``torch.randint`` pairs (train_and_eval.py:50-52)
  loss               ``-log(pos+1e-8).mean() - log(1-neg+1e-8).mean()``, grad-norm clip 1.0, optimizer step
The forward/backward of the neighbour aggregation is K1 (``autograd.spmm``); everything here is
device-side torch plumbing.  Random streams differ from the reference (python ``random`` / PyG there,
``torch.Generator`` here), the distributions are the same.
"""
from __future__ import annotations

from typing import Optional

import torch

from .graph import SparseAdj


def to_undirected(edge_index: torch.Tensor, num_nodes: int) -> torch.Tensor:
    """PyG ``to_undirected``: both directions, duplicates merged, sorted by (row, col)."""
    e = edge_index.reshape(2, -1).long()
    key = torch.cat([e[0] * num_nodes + e[1], e[1] * num_nodes + e[0]])
    key = torch.unique(key, sorted=True)
    return torch.stack([torch.div(key, num_nodes, rounding_mode="floor"), key % num_nodes])


def adjacency_keys(adj: SparseAdj) -> torch.Tensor:
    """Sorted int64 keys row*n+col of the adjacency (CSR order is already (row, col) ascending)."""
    if "keys" not in adj._cache:
        adj._cache["keys"] = adj.row() * adj.n + adj.col.long()
    return adj._cache["keys"]


def negative_sampling(adj: SparseAdj, num_neg_samples: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Uniform distinct non-edges ``[2, <=num_neg_samples]`` — PyG 1.7 ``negative_sampling(..., method=
    'dense')`` semantics: oversample by alpha = |1/(1-1.1*density)|, drop existing edges, truncate."""
    n = adj.n
    size = n * n
    num = min(int(num_neg_samples), size - adj.nnz)
    if num <= 0:
        return torch.zeros((2, 0), dtype=torch.long, device=adj.device)
    alpha = abs(1.0 / (1.0 - 1.1 * (adj.nnz / size)))
    keys = adjacency_keys(adj)
    draw = min(int(alpha * num) + 1, size)
    cand = torch.randint(0, size, (draw,), device=adj.device, dtype=torch.long, generator=generator)
    cand = torch.unique(cand)                                   # ``random.sample``: distinct indices
    pos = torch.searchsorted(keys, cand).clamp_(max=max(keys.numel() - 1, 0))
    cand = cand[keys[pos] != cand] if keys.numel() else cand
    cand = cand[torch.randperm(cand.numel(), device=adj.device, generator=generator)][:num]
    return torch.stack([torch.div(cand, n, rounding_mode="floor"), cand % n])


def link_loss(pos_out: torch.Tensor, neg_out: torch.Tensor) -> torch.Tensor:
    """train_and_eval.py:62-68."""
    return -torch.log(pos_out + 1e-8).mean() - torch.log(1 - neg_out + 1e-8).mean()


def train(model, data, dataset_name, split_edge, optimizer, batch_size, use_params, model_str, device,
          generator: Optional[torch.Generator] = None, max_batches: Optional[int] = None) -> float:
    """One epoch; returns the example-weighted mean loss like the reference."""
    model.train()
    pos_train_edge = split_edge["train"]["edge"].to(device)
    adj = data.adj_t
    n = data.num_nodes
    total_loss = total_examples = 0.0
    order = torch.randperm(pos_train_edge.size(0), device=device, generator=generator)
    for b, perm in enumerate(order.split(int(batch_size))):
        if max_batches is not None and b >= max_batches:
            break
        if use_params:
            optimizer.zero_grad()
        pos_edge = to_undirected(pos_train_edge[perm].t(), n)
        if model_str in ("gcn", "sage"):
            if dataset_name in ("collab",):
                neg_edge = torch.randint(0, n, pos_edge.size(), dtype=torch.long, device=device, generator=generator)
            else:
                neg_edge = negative_sampling(adj, pos_edge.size(1), generator)
        else:
            neg_dst = torch.randint(0, n, (pos_edge.size(1),), dtype=torch.long, device=device, generator=generator)
            neg_edge = torch.stack([pos_edge[0], neg_dst])
        out = model(data.x, torch.cat([pos_edge, neg_edge], 1), adj).reshape(-1)
        pos_out, neg_out = out[: pos_edge.size(1)], out[pos_edge.size(1):]
        loss = link_loss(pos_out, neg_out)
        if use_params:
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
            optimizer.step()
        num_examples = pos_out.size(0)
        total_loss += float(loss.item()) * num_examples
        total_examples += num_examples
    return total_loss / max(total_examples, 1)
