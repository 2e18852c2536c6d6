"""Seeded synthetic graphs of the reference's dataset shapes (no OGB download is possible offline).

Shapes follow SURVEY.md §8d: node / train-edge counts of ogbl-ddi, ogbl-collab, ogbl-ppa and the
email graph (/root/reference/email_data/data.py:22-24), Chung-Lu degree sequences with a
power-law tail.  Pure numpy; used by bench.py, the tests and the ``*-shape`` datasets of
``data.get_data``.
"""
from __future__ import annotations

import os

import numpy as np

# name -> n, undirected train edges, power-law exponent, max expected degree, feature width, seed
SHAPES = {
    "email": dict(n=986, m=9971, gamma=2.3, dmax=250, feat=0, seed=0),
    "ddi": dict(n=4267, m=1067911, gamma=None, dmax=3600, feat=0, seed=1),
    "collab": dict(n=235868, m=967632, gamma=2.8, dmax=700, feat=128, seed=2, weighted=True),
    "ppa": dict(n=576289, m=21231931, gamma=2.9, dmax=3200, feat=58, seed=3),
    # small stand-ins with the same generators, for tests
    "tiny": dict(n=300, m=1500, gamma=2.3, dmax=60, feat=16, seed=5),
    "small": dict(n=3000, m=30000, gamma=2.5, dmax=300, feat=32, seed=6),
}


def _expected_degrees(n: int, m: int, gamma, dmax: int, rng) -> np.ndarray:
    if gamma is None:
        # ddi-like: broad, dense degree distribution (lognormal), mean 2m/n
        w = rng.lognormal(mean=0.0, sigma=0.9, size=n)
    else:
        i = np.arange(1, n + 1, dtype=np.float64)
        w = i ** (-1.0 / (gamma - 1.0))
        rng.shuffle(w)
    w = w / w.sum() * (2.0 * m)
    for _ in range(8):                      # cap the tail, keep the mean
        w = np.minimum(w, dmax)
        w *= (2.0 * m) / w.sum()
    return np.minimum(w, dmax)


def chung_lu(n: int, m: int, gamma, dmax: int, seed: int) -> np.ndarray:
    """``m`` distinct undirected edges (u < v) with P(u,v) ~ w_u w_v; int64 [m,2], seeded."""
    rng = np.random.default_rng(seed)
    w = _expected_degrees(n, m, gamma, dmax, rng)
    p = w / w.sum()
    cdf = np.cumsum(p)
    cdf[-1] = 1.0
    keys = np.zeros(0, dtype=np.int64)
    need = m
    while need > 0:
        draw = int(need * 1.25) + 1024
        a = np.searchsorted(cdf, rng.random(draw), side="right")
        b = np.searchsorted(cdf, rng.random(draw), side="right")
        lo, hi = np.minimum(a, b), np.maximum(a, b)
        ok = lo != hi
        k = lo[ok].astype(np.int64) * n + hi[ok]
        keys = np.unique(np.concatenate([keys, k]))
        need = m - keys.size
    if keys.size > m:
        keys = rng.permutation(keys)[:m]
        keys.sort()
    return np.stack([keys // n, keys % n], axis=1)


def make_shape(name: str, scale: float = 1.0):
    """Returns dict(n, train_edges [m,2] (u<v), edge_weight or None, x or None, spec)."""
    spec = dict(SHAPES[name])
    n = max(int(spec["n"] * scale), 16)
    m = max(int(spec["m"] * scale * (scale if spec["gamma"] is None else 1.0)), 16)
    dmax = max(int(min(spec["dmax"], n - 1)), 4)
    # the generator is deterministic, so big shapes are cached on local disk between runs
    cache = os.path.join(os.environ.get("EPS_SYNTH_CACHE", "/tmp/eps_synth_cache"),
                         f"{name}_{n}_{m}_{dmax}_{spec['seed']}.npy")
    edges = None
    if m >= 1_000_000 and os.path.exists(cache):
        try:
            edges = np.load(cache)
        except Exception:
            edges = None
    if edges is None:
        edges = chung_lu(n, m, spec["gamma"], dmax, spec["seed"])
        if m >= 1_000_000:
            try:
                os.makedirs(os.path.dirname(cache), exist_ok=True)
                tmp = f"{cache}.{os.getpid()}.tmp.npy"
                np.save(tmp, edges)
                os.replace(tmp, cache)
            except Exception:
                pass
    rng = np.random.default_rng(spec["seed"] + 1000)
    weight = None
    if spec.get("weighted"):
        # collab: integer multi-edge weights 1..k (number of co-authored papers)
        weight = rng.geometric(0.6, size=edges.shape[0]).astype(np.float32)
    x = None
    if spec["feat"]:
        x = rng.standard_normal((n, spec["feat"])).astype(np.float32)
    return dict(n=n, train_edges=edges, edge_weight=weight, x=x, spec=spec, name=name)


def undirected_edge_index(train_edges: np.ndarray) -> np.ndarray:
    """Both directions, as OGB / ``to_undirected`` store ``edge_index`` — [2, 2m] int64."""
    e = np.asarray(train_edges, dtype=np.int64)
    return np.concatenate([e.T, e[:, ::-1].T], axis=1)
