"""Dataset plumbing for the CLI drop-ins: what ``rank.get_dataset`` / ``rank.get_data`` return
(/root/reference/rank.py:38-56, 83-126), without torch_geometric / ogb at import time.

  twitch / fb     the bundled MUSAE files with the reference's seeded split
                  (/root/reference/twitch/data.py:39-64, fb/data.py: same code): read from the CWD
                  layout the reference uses (``twitch/musae_DE_edges.csv`` ...) or, failing that,
                  from this repo's committed fixtures (tests/golden/{twitch,fb}.npz: the split edges;
                  {twitch,fb}_features.npz: the binary feature matrix as CSR index lists);
  ddi/collab/ppa  through ``ogb`` when it is installed (not in this image, no network);
  <name>-shape    seeded synthetic graph of that dataset's shape (synth.py), 80/10/10 split.

Negative validation/test edges: OGB ships them; for the local datasets the reference draws them
with ``random.sample`` (twitch/data.py:156-166) — here they are drawn with a seeded numpy
generator (same count, same exclusion rule).  They feed only the rank-side Hits@K evaluation.
"""
from __future__ import annotations

import json
import os
import random
from types import SimpleNamespace

import numpy as np
import torch

from . import synth
from .graph import add_edges

_LOCAL = {
    "twitch": dict(edges="twitch/musae_DE_edges.csv", feats="twitch/musae_DE_features.json", n=9498, width=3170),
    "fb": dict(edges="fb/musae_facebook_edges.csv", feats="fb/musae_facebook_features.json", n=22470, width=4714),
}
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _to_undirected(train: np.ndarray, n: int) -> np.ndarray:
    r = np.concatenate([train[:, 0], train[:, 1]])
    c = np.concatenate([train[:, 1], train[:, 0]])
    key = np.unique(r * n + c)
    return np.stack([key // n, key % n])


def _negatives(n: int, count: int, exist: set, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    have = np.fromiter((a * n + b for a, b in exist), dtype=np.int64, count=len(exist))
    have.sort()
    out = np.zeros(0, dtype=np.int64)
    while out.size < count:
        a = rng.integers(0, n, size=int((count - out.size) * 1.2) + 16)
        b = rng.integers(0, n, size=a.size)
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        key = (lo * n + hi)[lo != hi]
        pos = np.minimum(np.searchsorted(have, key), max(have.size - 1, 0))
        key = key[have[pos] != key] if have.size else key
        _, first = np.unique(key, return_index=True)
        key = key[np.sort(first)]                              # keep draw order, drop repeats
        key = key[~np.isin(key, out)]
        out = np.concatenate([out, key])[:count]
    exist.update((int(k // n), int(k % n)) for k in out)
    return np.stack([out // n, out % n], axis=1)


class LinkDataset:
    def __init__(self, name, n, train, valid, test, x=None, edge_weight=None, valid_neg=None, test_neg=None):
        self.name, self.num_nodes = name, n
        self.train, self.valid, self.test = train, valid, test
        ei = _to_undirected(train, n) if edge_weight is None else np.concatenate([train.T, train[:, ::-1].T], 1)
        w = None if edge_weight is None else np.concatenate([edge_weight, edge_weight])
        self.data = SimpleNamespace(num_nodes=n, edge_index=torch.from_numpy(ei),
                                    x=None if x is None else torch.from_numpy(x).float(),
                                    edge_weight=None if w is None else torch.from_numpy(w).float())
        exist = {(int(a), int(b)) for a, b in np.concatenate([train, valid, test])}
        self.valid_neg = valid_neg if valid_neg is not None else _negatives(n, len(valid), exist, 42)
        # the reference draws val_edges.size(1) negatives for BOTH splits (twitch/data.py make_edge_split)
        self.test_neg = test_neg if test_neg is not None else _negatives(n, len(valid), exist, 43)

    def __getitem__(self, i):
        return self.data

    def get_edge_split(self):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).long()
        return {"train": {"edge": t(self.train)},
                "valid": {"edge": t(self.valid), "edge_neg": t(self.valid_neg)},
                "test": {"edge": t(self.test), "edge_neg": t(self.test_neg)}}


def _load_local(name: str) -> LinkDataset:
    spec = _LOCAL[name]
    n = spec["n"]
    x = None
    if os.path.exists(spec["edges"]):
        import pandas as pd
        random.seed(42)
        edges = pd.read_csv(spec["edges"]).values.tolist()
        edges = [sorted((int(a), int(b))) for a, b in edges]
        edges = [e for e in edges if e[0] < e[1]]
        random.shuffle(edges)                                   # twitch/data.py:53, seed 42
        m = len(edges)
        e = np.asarray(edges, dtype=np.int64)
        train, valid, test = e[: int(0.8 * m)], e[int(0.8 * m): int(0.9 * m)], e[int(0.9 * m):]
        if os.path.exists(spec["feats"]):
            with open(spec["feats"]) as f:
                j = json.load(f)
            feats = np.zeros((n, max(max(v) for v in j.values() if v) + 1), dtype=np.float32)
            for node, fl in j.items():
                if int(node) < n:
                    feats[int(node), np.asarray(fl, dtype=int)] = 1
            x = feats[:, feats.sum(0) != 0]                     # drop all-zero columns (data.py:84)
    else:
        z = np.load(os.path.join(_REPO, "tests", "golden", f"{name}.npz"))
        train, valid, test = (z[k].astype(np.int64) for k in ("train_edges", "valid_edges", "test_edges"))
    if x is None:
        x = _fixture_features(name, n)
    return LinkDataset(name, n, train, valid, test, x=x)


def _fixture_features(name: str, n: int):
    """The dataset's binary feature matrix from the committed CSR fixture (oracle/make_golden_features.py:
    the public MUSAE feature file after the reference's processing, twitch/data.py:73-84), or None."""
    path = os.path.join(_REPO, "tests", "golden", f"{name}_features.npz")
    if not os.path.exists(path):
        return None
    z = np.load(path)
    x = np.zeros((n, int(z["width"])), dtype=np.float32)
    rows = np.repeat(np.arange(n), np.diff(z["indptr"]))
    x[rows, z["indices"].astype(np.int64)] = 1
    return x


def _load_shape(name: str) -> LinkDataset:
    base = name.split("-shape")[0]
    s = synth.make_shape(base)
    e = s["train_edges"]
    rng = np.random.default_rng(s["spec"]["seed"] + 7)
    perm = rng.permutation(e.shape[0])
    m = e.shape[0]
    tr, va, te = perm[: int(0.8 * m)], perm[int(0.8 * m): int(0.9 * m)], perm[int(0.9 * m):]
    w = None if s["edge_weight"] is None else s["edge_weight"][tr]
    return LinkDataset(name, s["n"], e[np.sort(tr)], e[va], e[te], x=s["x"],
                       edge_weight=None if w is None else s["edge_weight"][np.sort(tr)])


def get_dataset(dataset: str):
    if dataset.endswith("-shape"):
        return _load_shape(dataset)
    if dataset in _LOCAL:
        return _load_local(dataset)
    if dataset in ("ddi", "ppa", "collab"):
        try:
            from ogb.linkproppred import PygLinkPropPredDataset  # noqa: F401
        except Exception as exc:
            raise RuntimeError(f"dataset {dataset!r} needs the ogb package and its download; offline use "
                               f"'{dataset}-shape' (seeded synthetic graph of the same shape)") from exc
        ds = PygLinkPropPredDataset(name=f"ogbl-{dataset}")
        d = ds[0]
        split = ds.get_edge_split()
        out = LinkDataset.__new__(LinkDataset)
        out.name, out.num_nodes = dataset, d.num_nodes
        out.data = SimpleNamespace(num_nodes=d.num_nodes, edge_index=d.edge_index, x=getattr(d, "x", None),
                                   edge_weight=getattr(d, "edge_weight", None))
        out.get_edge_split = lambda: split
        return out
    raise NotImplementedError(dataset)


def get_data(args, device="cpu"):
    """rank.get_data (/root/reference/rank.py:83-126): (edge_index, edge_weight, split_edge, data)."""
    dataset = get_dataset(args.dataset)
    d = dataset[0]
    edge_index = d.edge_index
    edge_weight = torch.ones(edge_index.size(1))
    if getattr(d, "edge_weight", None) is not None:
        edge_weight = d.edge_weight.view(-1).float()
    split_edge = dataset.get_edge_split()
    idx = torch.randperm(split_edge["train"]["edge"].size(0))[: split_edge["valid"]["edge"].size(0)]
    split_edge["eval_train"] = {"edge": split_edge["train"]["edge"][idx]}
    name = "collab" if args.dataset.startswith("collab") else args.dataset
    if args.use_feature and d.x is None:
        raise RuntimeError(f"dataset {args.dataset!r} has no node features here (feature file and fixture missing); "
                           "pass --use_feature '' to run on the learnable embedding alone")
    data = SimpleNamespace(num_nodes=d.num_nodes, x=d.x if args.use_feature else None, edge_index=edge_index)
    data.adj_t = add_edges(name, edge_index.to(device), edge_weight.to(device),
                           torch.zeros([2, 0], dtype=torch.long, device=device), d.num_nodes)
    if data.x is not None:
        data.x = data.x.to(device)
    return edge_index, edge_weight, split_edge, data
