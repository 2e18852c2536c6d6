"""K4 top-k vs the stable-sort oracle: bit-exact index lists, including tie-heavy inputs where the
k-th boundary falls inside a huge equal-score group."""
import numpy as np
import pytest
import torch

from oracle import ranking as orank

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _check(score_np, k):
    from edge_proposal_sets_b200 import ops
    idx, sc = ops.topk(torch.from_numpy(score_np).to(DEV), k)
    want_idx, want_sc = orank.topk_desc(score_np, k)
    assert np.array_equal(idx.cpu().numpy(), want_idx)
    assert np.array_equal(sc.cpu().numpy().view(np.uint32), want_sc.view(np.uint32))


@pytest.mark.parametrize("M,k", [(1, 1), (7, 3), (1000, 1000), (4097, 100), (100000, 1), (262144, 50000), (3000001, 530000)])
def test_random_scores(M, k):
    rng = np.random.default_rng(M + k)
    _check(rng.standard_normal(M).astype(np.float32), k)


@pytest.mark.parametrize("M,k", [(5000, 2500), (1 << 20, 300000), (2000003, 999999)])
def test_tie_heavy_integer_scores(M, k):
    rng = np.random.default_rng(k)
    s = rng.geometric(0.55, size=M).astype(np.float32)          # CN-like: mostly 1.0, 2.0, 3.0 ...
    _check(s, k)


def test_all_equal_and_saturated():
    _check(np.full(100000, 1.0, np.float32), 12345)              # sigmoid-saturated adamic scores
    _check(np.zeros(70000, np.float32), 70000)
    s = np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, -0.0] * 1000, np.float32)
    _check(s, 4500)                                              # -0.0 == +0.0 -> ties by position
    s = np.concatenate([np.full(5000, -np.inf, np.float32), np.full(10, np.inf, np.float32),
                        np.full(5000, 3.0e38, np.float32), np.full(5000, -3.0e38, np.float32)])
    _check(s, 7000)


def test_topk_edges_packing_and_prefix_property():
    from edge_proposal_sets_b200 import ops
    rng = np.random.default_rng(11)
    M = 400000
    e = rng.integers(0, 60000, size=(2, M)).astype(np.int32)
    s = rng.geometric(0.4, size=M).astype(np.float32)
    full = ops.topk_edges(torch.from_numpy(e).to(DEV), torch.from_numpy(s).to(DEV), M).cpu().numpy()
    assert np.array_equal(full, orank.sorted_edges(e, s))        # the reference's [N,3] float32 file
    for k in (1, 1000, 123457):                                  # every sweep point is a prefix
        top = ops.topk_edges(torch.from_numpy(e).to(DEV), torch.from_numpy(s).to(DEV), k).cpu().numpy()
        assert np.array_equal(top, full[:k])


@pytest.mark.parametrize("Ma,Mb,k", [(0, 5000, 700), (700, 5000, 700), (3, 1, 2), (40000, 2000003, 40000),
                                     (250000, 250000, 499999)])
def test_select2_is_the_position_ordered_topk_of_the_concatenation(Ma, Mb, k):
    """eps_topk_select2_f32: the same SET as the stable sort's first k, in ascending position order;
    tie-heavy scores so the boundary falls inside an equal-score group spanning both segments."""
    from edge_proposal_sets_b200 import ops
    rng = np.random.default_rng(Ma + Mb + k)
    a = rng.geometric(0.5, size=Ma).astype(np.float32)
    b = rng.geometric(0.5, size=Mb).astype(np.float32)
    idx, sc = ops.topk_select2(None if Ma == 0 else torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV), k)
    cat = np.concatenate([a, b])
    want_idx, _ = orank.topk_desc(cat, k)
    want_idx = np.sort(want_idx)
    assert np.array_equal(idx.cpu().numpy().astype(np.int64), want_idx)
    assert np.array_equal(sc.cpu().numpy(), cat[want_idx])


@pytest.mark.parametrize("k", [None, 1, 5000, 123457, 10**7])
def test_running_topk_over_slabs_equals_one_global_stable_sort(k):
    """RunningTopK over ragged slabs (incl. empty and 1-element ones) == sort of everything at once."""
    from edge_proposal_sets_b200.filter_step import RunningTopK
    rng = np.random.default_rng(5)
    sizes = [70000, 0, 1, 300000, 12345, 2, 90000]
    M = sum(sizes)
    e = rng.integers(0, 60000, size=(2, M)).astype(np.int32)
    s = rng.geometric(0.4, size=M).astype(np.float32)
    s[rng.integers(0, M, 2000)] = rng.random(2000).astype(np.float32) * 20
    run = RunningTopK(k)
    lo = 0
    for m in sizes:
        run.update(torch.from_numpy(e[:, lo:lo + m]).to(DEV), torch.from_numpy(s[lo:lo + m]).to(DEV))
        lo += m
    assert run.seen == M
    got = run.result(DEV).cpu().numpy()
    want = orank.sorted_edges(e, s, None if k is None else min(k, M))
    assert got.shape == want.shape and np.array_equal(got, want)
