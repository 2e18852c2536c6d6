"""Host-side logic (torch ops on the CPU) against the oracle: graph construction, derived tables,
configs, sharding arithmetic.  No kernels run here."""
import argparse

import numpy as np
import pytest
import torch

from edge_proposal_sets_b200 import graph as pg, models as pm, parallel, synth
from oracle import gnn as ognn, graph as og, heuristics as oh
from util import synth_graph


def _same_csr(adj, g):
    assert adj.n == g.n
    assert np.array_equal(adj.rowptr.numpy().astype(np.int64), g.rowptr)
    assert np.array_equal(adj.col.numpy().astype(np.int64), g.col)
    assert np.array_equal(adj.values().numpy(), g.val)


@pytest.mark.parametrize("dataset", ["ddi", "collab"])
def test_add_edges_matches_oracle(dataset):
    rng = np.random.default_rng(0)
    n = 200
    ei = rng.integers(0, n, size=(2, 3000))
    ei = ei[:, ei[0] != ei[1]]
    w = rng.integers(1, 5, size=ei.shape[1]).astype(np.float32)
    extra = rng.integers(0, n, size=(2, 400))
    extra = extra[:, extra[0] != extra[1]]
    g = og.add_edges(dataset, ei, w, extra, n)
    adj = pg.add_edges(dataset, torch.from_numpy(ei), torch.from_numpy(w), torch.from_numpy(extra), n)
    _same_csr(adj, g)
    empty = pg.add_edges(dataset, torch.from_numpy(ei), torch.from_numpy(w), torch.zeros([2, 0], dtype=int), n)
    _same_csr(empty, og.add_edges(dataset, ei, w, np.zeros((2, 0), np.int64), n))


def test_derived_tables_match_oracle():
    s, ei, w, g = synth_graph("small")
    adj = pg.add_edges("small", torch.from_numpy(ei), torch.from_numpy(w), torch.zeros([2, 0], dtype=int), s["n"])
    rp, c, v = ognn.gcn_norm(g)
    prp, pc, pv = adj.gcn_norm()
    assert np.array_equal(prp.numpy(), rp) and np.array_equal(pc.numpy(), c)
    assert np.allclose(pv.numpy(), v, rtol=1e-6, atol=0)
    assert np.allclose(adj.aa_ogb_weights().numpy(), oh.aa_ogb_weights(g), rtol=1e-6)
    assert np.allclose(adj.adamic_weights().numpy(), oh.adamic_gpu_weights(g), rtol=1e-6)
    assert np.allclose(adj.ra_weights().numpy(), oh.ra_weights(g).astype(np.float32), rtol=1e-6)
    # weighted (collab) graph keeps summed weights
    s, ei, w, g = synth_graph("tiny", dataset="collab")
    wts = np.random.default_rng(1).integers(1, 4, size=ei.shape[1] // 2).astype(np.float32)
    w2 = np.concatenate([wts, wts])
    g = og.add_edges("collab", ei, w2, np.zeros((2, 0), np.int64), s["n"])
    adj = pg.add_edges("collab", torch.from_numpy(ei), torch.from_numpy(w2), torch.zeros([2, 0], dtype=int), s["n"])
    _same_csr(adj, g)
    assert np.allclose(adj.sum(-1).numpy(), np.asarray(g.to_scipy().sum(1)).ravel())


def _args(dataset, model, **kw):
    ns = argparse.Namespace(dataset=dataset, model=model, num_layers=None, hidden_channels=None, dropout=None,
                            batch_size=None, lr=None, epochs=None, use_feature=None, use_learnable_embedding=None)
    for k, v in kw.items():
        setattr(ns, k, v)
    return pm.default_model_configs(ns)


def test_default_model_configs_table():
    a = _args("ddi", "gcn")
    assert (a.num_layers, a.hidden_channels, a.batch_size, a.use_feature, a.use_learnable_embedding) == (2, 256, 65536, False, True)
    assert _args("ddi", "simple").batch_size == 1024
    a = _args("collab", "sage")
    assert (a.num_layers, a.hidden_channels, a.batch_size, a.use_feature) == (3, 256, 16384, True)
    a = _args("email", "gcn")
    assert (a.num_layers, a.hidden_channels, a.batch_size) == (3, 300, 16384)
    a = _args("twitch", "sage")
    assert (a.num_layers, a.hidden_channels, a.batch_size, a.lr) == (3, 256, 65536, 0.005)
    a = _args("collab", "adamic_ogb")
    assert a.use_feature is False and a.use_learnable_embedding is False and a.num_layers is None
    assert _args("ddi", "gcn", num_layers=5).num_layers == 5         # CLI value wins


def test_state_dict_keys_match_reference_layout():
    class D:
        num_nodes = 50
        x = torch.zeros(50, 7)
    a = _args("collab", "gcn")
    m = pm.build_model(a, D, "cpu")
    keys = set(m.state_dict())
    assert {"emb.weight", "gnn.convs.0.weight", "gnn.convs.0.bias", "gnn.convs.2.weight",
            "linkpred.lins.0.weight", "linkpred.lins.2.bias"} <= keys
    assert m.gnn.convs[0].weight.shape == (256 + 7, 256)              # [in,out], PyG 1.7
    assert m.linkpred.lins[2].weight.shape == (1, 256)
    a = _args("twitch", "sage")
    m = pm.build_model(a, D, "cpu")
    keys = set(m.state_dict())
    assert {"gnn.convs.0.lin_l.weight", "gnn.convs.0.lin_l.bias", "gnn.convs.0.lin_r.weight"} <= keys
    assert "gnn.convs.0.lin_r.bias" not in keys
    sd = ognn.random_state_dict("sage", 50, 7, 256, 3)
    m.load_state_dict(sd)                                             # oracle layout == product layout
    cn = pm.build_model(_args("ddi", "simple"), D, "cpu")
    assert sum(p.numel() for p in cn.parameters()) == 0
    with pytest.raises(ValueError):
        pm.build_model(_args("ddi", "katz_not_a_model"), D, "cpu")


def test_partition_by_work():
    work = torch.tensor([5, 1, 1, 1, 8, 2, 2, 4], dtype=torch.int64)
    b = parallel.partition_by_work(work, 2)
    assert b[0] == 0 and b[-1] == 8 and b == sorted(b)
    halves = [int(work[b[i]:b[i + 1]].sum()) for i in range(2)]
    assert abs(halves[0] - halves[1]) <= int(work.max())
    assert parallel.partition_by_work(work, 1) == [0, 8]
    b4 = parallel.partition_by_work(work, 4)
    assert len(b4) == 5 and b4 == sorted(b4)


def test_synth_shapes_are_seeded():
    a = synth.make_shape("tiny")
    b = synth.make_shape("tiny")
    assert np.array_equal(a["train_edges"], b["train_edges"])
    e = a["train_edges"]
    assert e.shape == (1500, 2) and np.all(e[:, 0] < e[:, 1]) and e.max() < a["n"]
    assert np.unique(e[:, 0] * a["n"] + e[:, 1]).size == 1500


@pytest.mark.parametrize("shape", ["tiny", "small"])
def test_owner_bounds_and_slab_plan(shape):
    """candidates.owner_bounds is an upper bound of every owner's candidate count (it sizes the padded slots
    of the one-pass kernel) and filter_step.iter_slabs cuts the owners into consecutive ranges whose
    capacities respect slab_pairs and add up to the bound of the whole range."""
    from edge_proposal_sets_b200 import candidates, filter_step
    from util import to_adj
    s, ei, w, g = synth_graph(shape)
    adj = to_adj(g, "cpu")
    cand = og.two_hop_candidates(g)
    counts = np.bincount(cand[1], minlength=g.n)
    bounds = candidates.owner_bounds(adj).numpy()
    assert bounds.shape == (g.n,) and np.all(bounds >= counts)
    deg = np.diff(g.rowptr)
    assert np.all(bounds <= np.maximum(g.n - 1 - deg, 0))
    assert np.all(bounds[deg == 0] == 0)
    for lo_hi in [(0, g.n), (g.n // 3, g.n // 2), (5, 5)]:
        for slab_pairs in (1, 977, 50_000, 10**9):
            plan = list(filter_step.iter_slabs(adj, lo_hi[0], lo_hi[1], slab_pairs))
            if lo_hi[0] == lo_hi[1]:
                assert plan == []
                continue
            assert plan[0][0] == lo_hi[0] and plan[-1][1] == lo_hi[1]
            assert all(a[1] == b[0] for a, b in zip(plan, plan[1:]))           # consecutive, no gaps
            assert sum(p[2] for p in plan) == int(bounds[lo_hi[0]:lo_hi[1]].sum())
            for lo, hi, cap in plan:
                assert cap == int(bounds[lo:hi].sum())
                assert cap <= slab_pairs or hi - lo == 1                       # a hub gets a slab of its own
                assert int(counts[lo:hi].sum()) <= cap


def test_values_symmetric_detects_asymmetric_weights():
    from edge_proposal_sets_b200 import candidates
    from util import to_adj
    s, ei, w, g = synth_graph("tiny")
    assert candidates.values_symmetric(to_adj(g, "cpu"))                       # unweighted
    key = np.minimum(np.repeat(np.arange(g.n), np.diff(g.rowptr)), g.col) * g.n + \
        np.maximum(np.repeat(np.arange(g.n), np.diff(g.rowptr)), g.col)
    uk, inv = np.unique(key, return_inverse=True)
    g.val = (np.random.default_rng(1).random(uk.size).astype(np.float32) + 0.5)[inv]
    assert candidates.values_symmetric(to_adj(g, "cpu", keep_values=True))
    g.val = g.val.copy()
    g.val[0] += 1.0
    assert not candidates.values_symmetric(to_adj(g, "cpu", keep_values=True))


def test_owner_cost_balances_scoring_and_enumeration():
    """filter_step.owner_cost: 2-paths for the heuristic jobs, 2-paths + 8 x slot size as soon as a GNN model scores;
    the ranges cut on it are contiguous, cover every owner and carry near-equal cost."""
    import torch
    from edge_proposal_sets_b200 import candidates, filter_step, parallel
    from util import synth_graph, to_adj
    s, ei, w, g = synth_graph("small")
    adj = to_adj(g, "cpu")
    two = candidates.two_path_work(adj).double()
    heur = filter_step.owner_cost(adj, [filter_step.FilterJob("adamic_ogb", None)])
    assert torch.equal(heur, two)
    gnn = filter_step.owner_cost(adj, [filter_step.FilterJob("adamic_ogb", None), ("gcn", None)])
    assert torch.equal(gnn, two + 8.0 * candidates.owner_bounds(adj).double())
    for parts in (2, 3, 8):
        b = parallel.partition_by_work(gnn, parts)
        assert b[0] == 0 and b[-1] == adj.n and all(b[i] <= b[i + 1] for i in range(parts))
        loads = [float(gnn[b[i]:b[i + 1]].sum()) for i in range(parts)]
        assert max(loads) <= float(gnn.sum()) / parts + float(gnn.max()) + 1e-9
