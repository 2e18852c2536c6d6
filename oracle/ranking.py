"""Oracle: global ordering, top-k prefix, sweep schedule, Hits@K.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.

Restates
  * ordering + packing        /root/reference/filter.py:113-121,160-165
  * sweep schedule / prefix   /root/reference/rank.py:260-272,294
  * ``--valid_proposal``      /root/reference/rank.py:222-251
  * OGB Hits@K                /root/reference/train_and_eval.py:138-154 (+ SURVEY A.7)
"""
from __future__ import annotations

import numpy as np


def stable_order_desc(score: np.ndarray) -> np.ndarray:
    """Permutation of ``sort(descending=True, stable=True)``: score descending, ties by
    candidate index ascending (SURVEY §8 a11 — the reference's unstable sort is defined only up
    to permutation inside ties; this is the contract every top-k result is held to)."""
    s = np.asarray(score, dtype=np.float32)
    assert not np.isnan(s).any(), "scores on this path are never NaN"
    # -0.0 and +0.0 compare equal; lexsort: last key is primary
    idx = np.arange(s.shape[0], dtype=np.int64)
    return np.lexsort((idx, -(s + np.float32(0.0)).astype(np.float64)))


def topk_desc(score: np.ndarray, k: int):
    order = stable_order_desc(score)[:k]
    return order, np.asarray(score, dtype=np.float32)[order]


def sorted_edges(all_edges: np.ndarray, score: np.ndarray, k: int | None = None) -> np.ndarray:
    """filter.py:119,160-161: float32 ``[N,3]`` rows (u, v, score) sorted by score descending.
    Node ids are stored as float32 exactly as the reference does (exact below 2**24)."""
    order = stable_order_desc(score)
    if k is not None:
        order = order[:k]
    e = np.asarray(all_edges)
    out = np.empty((order.shape[0], 3), dtype=np.float32)
    out[:, 0] = e[0, order]
    out[:, 1] = e[1, order]
    out[:, 2] = np.asarray(score, dtype=np.float32)[order]
    return out


def sweep_index_ends(sweep_num, sweep_min, sweep_max, num_sorted_edge):
    """rank.py:260-272."""
    ends = []
    if sweep_num:
        if sweep_min is None:
            sweep_min = 0
        if sweep_max is None:
            sweep_max = (sweep_num - 1) * 1000
        for i in range(sweep_num + 1):
            ends.append(sweep_min + int(i * (sweep_max - sweep_min) / sweep_num))
    elif num_sorted_edge:
        ends.append(num_sorted_edge)
    else:
        ends.append(0)
    return ends


def prefix_edges(sorted_e: np.ndarray, index_end: int) -> np.ndarray:
    """rank.py:294: ``sorted[:index_end, :2].t().long()``."""
    return np.asarray(sorted_e)[: int(index_end), :2].T.astype(np.int64)


def hits_at_k(pos: np.ndarray, neg: np.ndarray, K: int) -> float:
    """A.7: 1.0 if len(neg) < K else mean(pos > K-th largest neg) (strict >)."""
    pos = np.asarray(pos, dtype=np.float32)
    neg = np.asarray(neg, dtype=np.float32)
    if neg.shape[0] < K:
        return 1.0
    kth = np.partition(neg, neg.shape[0] - K)[neg.shape[0] - K]
    return float((pos > kth).sum()) / float(pos.shape[0])


def valid_proposal_surgery(sorted_e: np.ndarray, valid_pos: np.ndarray) -> np.ndarray:
    """rank.py:222-251: put both directions of every validation edge on top with score
    100000.0 and drop them from the body (the set iteration order on top is arbitrary in the
    reference; rows are emitted here sorted by (u, v) — only the *set* of the first
    ``len(valid_pos_set)`` rows is asserted by the reference, rank.py:246-250)."""
    vp = np.asarray(valid_pos, dtype=np.int64).reshape(-1, 2)
    both = set()
    for a, b in vp:
        both.add((int(a), int(b)))
        both.add((int(b), int(a)))
    und = {tuple(sorted(t)) for t in both}
    top = np.array([[u, v, 100000.0] for (u, v) in sorted(both)], dtype=np.float64).reshape(-1, 3)
    body = [t for t in np.asarray(sorted_e)
            if (int(t[0]), int(t[1])) not in und and (int(t[1]), int(t[0])) not in und]
    body = np.asarray(body, dtype=np.float64).reshape(-1, 3)
    return np.concatenate([top, body], axis=0)
