"""Rank-side helpers (prefix, sweep, --valid_proposal surgery) vs the oracle; CPU only."""
import numpy as np
import torch

from edge_proposal_sets_b200 import rank_step
from oracle import ranking as orank


def test_sweep_and_prefix_match_oracle():
    for args in [(4, 100, 500, None), (None, None, None, 7), (None, None, None, None), (3, None, None, None),
                 (10, 510000, 550000, None)]:
        assert rank_step.sweep_index_ends(*args) == orank.sweep_index_ends(*args)
    se = torch.tensor([[3., 4., .9], [1., 2., .8], [5., 6., .7]])
    assert rank_step.prefix_edges(se, 2).tolist() == orank.prefix_edges(se.numpy(), 2).tolist() == [[3, 1], [4, 2]]
    assert rank_step.prefix_edges(se, 0).shape == (2, 0)


def test_valid_proposal_surgery_matches_oracle():
    rng = np.random.default_rng(0)
    n = 50
    pairs = rng.integers(0, n, size=(400, 2))
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    _, first = np.unique(pairs[:, 0] * n + pairs[:, 1], return_index=True)
    pairs = pairs[np.sort(first)]
    score = np.sort(rng.random(len(pairs)).astype(np.float32))[::-1]
    se = np.concatenate([pairs.astype(np.float32), score[:, None]], 1)
    valid = pairs[rng.choice(len(pairs), 25, replace=False)]
    valid = np.concatenate([valid, rng.integers(0, n, size=(10, 2))])       # some not in the list at all
    valid = valid[valid[:, 0] != valid[:, 1]]
    got = rank_step.valid_proposal(torch.from_numpy(se), torch.from_numpy(valid)).numpy()
    want = orank.valid_proposal_surgery(se, valid)
    nv = len({(int(a), int(b)) for a, b in valid} | {(int(b), int(a)) for a, b in valid})
    assert got.shape == want.shape
    # the reference only fixes the SET of the first nv rows (python set iteration order); the body keeps its order
    assert {tuple(r) for r in got[:nv].tolist()} == {tuple(r) for r in want[:nv].tolist()}
    assert np.array_equal(got[nv:], want[nv:])
    assert np.all(got[:nv, 2] == 100000.0)
