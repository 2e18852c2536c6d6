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


@pytest.mark.parametrize("pushdown", [True, False])
@pytest.mark.parametrize("k", [5000, 123457])
def test_running_topk_pushdown_and_legacy_select_agree(k, pushdown):
    """K4b push-down (strictly-better survivors + select over list ++ survivors) == the plain K4 select over
    every slab == one global stable sort; many small slabs so the k-th score rises while ties stream by."""
    from edge_proposal_sets_b200.filter_step import RunningTopK
    rng = np.random.default_rng(9)
    sizes = [150000] + [40000] * 12 + [0, 3, 99999]
    M = sum(sizes)
    e = rng.integers(0, 60000, size=(2, M)).astype(np.int32)
    s = rng.geometric(0.35, size=M).astype(np.float32)          # CN-like: the boundary sits inside a tie group
    s[rng.integers(0, M, 5000)] += rng.random(5000).astype(np.float32)
    run = RunningTopK(k, pushdown=pushdown)
    lo = 0
    for m in sizes:
        run.update(torch.from_numpy(e[:, lo:lo + m]).to(DEV), torch.from_numpy(s[lo:lo + m]).to(DEV))
        lo += m
    got = run.result(DEV).cpu().numpy()
    assert np.array_equal(got, orank.sorted_edges(e, s, k))
    if pushdown:
        assert 0 < run.survivors < M - sizes[0]                  # the push-down did prune


@pytest.mark.parametrize("inclusive,margin", [(False, 0.0), (True, 0.0), (True, 0.25), (True, 1e-3)])
def test_threshold_compact_vs_numpy(inclusive, margin):
    """K4b: survivors of a slab under the running k-th score, in position order."""
    from edge_proposal_sets_b200 import ops
    rng = np.random.default_rng(3)
    M = 300007
    s = np.round(rng.standard_normal(M), 2).astype(np.float32)   # coarse values: plenty of exact ties
    e = rng.integers(0, 1 << 20, size=(2, M)).astype(np.int32)
    sd, ed = torch.from_numpy(s).to(DEV), torch.from_numpy(e).to(DEV)
    for kth_score in (0.5, -0.0, 2.13, -7.0, 9.0):
        key = ops.key_tensor(ops.score_to_key(kth_score), DEV)
        u, v, sc, pos = ops.threshold_compact(sd, ed, key, margin, inclusive, want_pos=True)
        if inclusive:
            thr = np.float32(kth_score) - np.float32(margin)
            keep = s >= (np.nextafter(thr, np.float32(-np.inf)) if margin > 0 else thr)
            # one ulp of slack below the rounded difference is allowed (never fewer survivors than the band)
            assert np.all(keep[pos.cpu().numpy().astype(np.int64)])
            must = np.nonzero(s >= thr)[0]
            assert np.all(np.isin(must, pos.cpu().numpy().astype(np.int64)))
            want = np.nonzero(keep)[0]
        else:
            want = np.nonzero(s > np.float32(kth_score))[0]
        assert np.array_equal(pos.cpu().numpy().astype(np.int64), want)
        assert np.array_equal(sc.cpu().numpy(), s[want])
        assert np.array_equal(u.cpu().numpy(), e[0, want]) and np.array_equal(v.cpu().numpy(), e[1, want])


def test_kth_key_only_and_band_pool():
    """Band mode of RunningTopK (the bf16 prefilter): the pool is exactly {score >= T_k - margin} of everything
    seen, in candidate order, where T_k is the k-th best score."""
    from edge_proposal_sets_b200 import ops
    from edge_proposal_sets_b200.filter_step import RunningTopK
    rng = np.random.default_rng(21)
    sizes = [50000, 20000, 0, 130000, 7, 60000, 90000]
    M, k, margin = sum(sizes), 30000, 4e-3
    s = (1.0 / (1.0 + np.exp(-rng.standard_normal(M) * 0.4))).astype(np.float32)
    e = np.stack([rng.integers(0, 50000, M), np.sort(rng.integers(0, 50000, M))]).astype(np.int32)
    kth = ops.kth_key(torch.from_numpy(s).to(DEV), k)
    T = np.sort(s)[::-1][k - 1]
    assert ops.key_to_score(kth) == float(T)
    run = RunningTopK(k, margin)
    lo = 0
    for m in sizes:
        run.update(torch.from_numpy(e[:, lo:lo + m]).to(DEV), torch.from_numpy(s[lo:lo + m]).to(DEV))
        lo += m
    edges, sc = run.pool()
    thr = np.float32(T) - np.float32(margin)
    got_pos_ok = np.nonzero(s >= np.nextafter(thr, np.float32(-np.inf)))[0]
    must = np.nonzero(s >= thr)[0]
    got = sc.cpu().numpy()
    assert must.size <= got.size <= got_pos_ok.size
    # candidate order kept, and every element of the band is present
    pool_set = set(map(tuple, np.stack([edges[0].cpu().numpy(), edges[1].cpu().numpy(), got.view(np.int32)]).T))
    assert all((int(e[0, i]), int(e[1, i]), int(s[i:i + 1].view(np.int32)[0])) in pool_set for i in must[::37])
    assert np.array_equal(got, s[got_pos_ok]) or np.array_equal(got, s[must])
