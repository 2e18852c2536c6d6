"""K1 SpMM, K2 LinkPredictor MLP and the GCN / SAGE + LinkPredictor forward vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import gnn as ognn, graph as og, ranking as orank
from util import synth_graph, to_adj

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("F", [256, 300, 128, 58, 2770 // 10 * 10 + 6])
@pytest.mark.parametrize("reduce", ["sum", "mean"])
def test_spmm_vs_fp64(F, reduce):
    from edge_proposal_sets_b200 import ops
    s, ei, w, g = synth_graph("small")
    adj = to_adj(g, DEV)
    rng = np.random.default_rng(F)
    x = rng.standard_normal((g.n, F)).astype(np.float32)
    val = rng.random(g.nnz).astype(np.float32) if reduce == "sum" else None
    bias = rng.standard_normal(F).astype(np.float32)
    y = ops.spmm_csr(adj.rowptr, adj.col, None if val is None else torch.from_numpy(val).to(DEV),
                     torch.from_numpy(x).to(DEV), reduce, torch.from_numpy(bias).to(DEV), relu=True)
    want = ognn.spmm(g.rowptr, g.col, val, torch.from_numpy(x).double(), reduce) + torch.from_numpy(bias).double()
    want = torch.relu(want).numpy()
    err = np.abs(y.cpu().numpy() - want)
    assert err.max() <= 2e-6 * max(1.0, np.abs(want).max())      # fp32 sequential sum vs fp64
    # isolated rows: mean of empty -> 0 (+bias)
    iso = np.flatnonzero(np.diff(g.rowptr) == 0)
    if iso.size:
        assert np.allclose(y.cpu().numpy()[iso], np.maximum(bias, 0)[None, :].repeat(iso.size, 0))


def test_spmm_keeps_sequential_order_bit_exact():
    """A.3 order-sensitive note: ascending-column left fold with fma == the warp-per-row kernel."""
    from edge_proposal_sets_b200 import ops
    s, ei, w, g = synth_graph("tiny")
    adj = to_adj(g, DEV)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((g.n, 64)).astype(np.float32)
    val = rng.random(g.nnz).astype(np.float32)
    y = ops.spmm_csr(adj.rowptr, adj.col, torch.from_numpy(val).to(DEV), torch.from_numpy(x).to(DEV), "sum")
    assert np.array_equal(y.cpu().numpy(), ognn.spmm_sequential(g.rowptr, g.col, val, x, "sum"))
    y = ops.spmm_csr(adj.rowptr, adj.col, None, torch.from_numpy(x).to(DEV), "mean")
    assert np.array_equal(y.cpu().numpy(), ognn.spmm_sequential(g.rowptr, g.col, None, x, "mean"))


@pytest.mark.parametrize("H,L,M", [(256, 2, 70001), (256, 3, 33), (300, 3, 5000), (64, 1, 1000)])
def test_linkpred_fp32_vs_oracle(H, L, M):
    from edge_proposal_sets_b200 import ops
    n = 3000
    sd = ognn.random_state_dict("gcn", n, 0, H, L)
    rng = np.random.default_rng(H + L)
    h = rng.standard_normal((n, H)).astype(np.float32) * 0.5
    e = rng.integers(0, n, size=(2, M))
    Ws = [sd[f"linkpred.lins.{i}.weight"].to(DEV) for i in range(L)]
    bs = [sd[f"linkpred.lins.{i}.bias"].to(DEV) for i in range(L)]
    got = ops.linkpred_mlp(torch.from_numpy(h).to(DEV), torch.from_numpy(e).to(DEV), Ws, bs, "fp32").cpu().numpy()
    want64 = ognn.linkpred_forward(torch.from_numpy(h), e, sd, L, torch.float64).numpy()
    want32 = ognn.linkpred_forward(torch.from_numpy(h), e, sd, L, torch.float32).numpy()
    assert np.max(np.abs(got - want64)) <= 1e-5                  # stated tolerance, fp32 arm (App. B)
    assert np.max(np.abs(got - want32)) <= 1e-5
    logit = ops.linkpred_mlp(torch.from_numpy(h).to(DEV), torch.from_numpy(e).to(DEV), Ws, bs, "fp32", sigmoid=False)
    want_l = ognn.linkpred_forward(torch.from_numpy(h), e, sd, L, torch.float64, return_logit=True).numpy()
    assert np.max(np.abs(logit.cpu().numpy() - want_l)) <= 2e-5 * max(1.0, np.abs(want_l).max())


@pytest.mark.parametrize("model,L,feat", [("gcn", 2, 0), ("gcn", 3, 32), ("sage", 3, 32)])
def test_link_gnn_forward_vs_oracle(model, L, feat):
    """Full LinkGNN forward (embedding concat, L conv layers, pair MLP) with seeded weights."""
    import argparse
    from edge_proposal_sets_b200 import models
    s, ei, w, g = synth_graph("small")
    n, H = g.n, 256
    adj = to_adj(g, DEV)
    sd = ognn.random_state_dict(model, n, feat, H, L)
    x = None if feat == 0 else torch.from_numpy(np.random.default_rng(1).standard_normal((n, feat)).astype(np.float32))
    args = argparse.Namespace(model=model, dataset="x", num_layers=L, hidden_channels=H, dropout=0.0,
                              use_feature=feat > 0, use_learnable_embedding=True)

    class D:
        num_nodes = n
    D.x = x
    m = models.build_model(args, D, DEV)
    m.load_state_dict(sd)
    m.eval()
    xin = ognn.link_gnn_input(sd, x)
    fwd = ognn.gcn_forward if model == "gcn" else ognn.sage_forward
    h64 = fwd(g, xin, sd, L, torch.float64)
    h = m.embed(None if x is None else x.to(DEV), adj)
    herr = (h.cpu().double() - h64).abs().max().item()
    assert herr <= 1e-5 * max(1.0, h64.abs().max().item())       # embeddings, per App. B
    e = np.random.default_rng(2).integers(0, n, size=(2, 20000))
    got = m(None if x is None else x.to(DEV), torch.from_numpy(e).to(DEV), adj)
    assert got.shape == (20000, 1)                               # [B,1] like the reference
    want = ognn.linkpred_forward(h64, e, sd, L, torch.float64).numpy()
    assert np.max(np.abs(got.squeeze(1).cpu().numpy() - want)) <= 1e-5
    # second call hits the embedding cache (same graph, same weights)
    assert m.embed(None if x is None else x.to(DEV), adj) is h


def test_filter_topk_end_to_end_cn_and_gcn():
    """filter_step.filter_topk == oracle pipeline (candidates -> scores -> stable sort -> [k,3])."""
    import argparse
    from edge_proposal_sets_b200 import filter_step, models
    from oracle import heuristics as oh
    s, ei, w, g = synth_graph("small")
    adj = to_adj(g, DEV)
    cand = og.two_hop_candidates(g)
    # CN filter: bit-exact list, whole and sliced into many slabs
    cn = oh.cn_scores_pairs(g, cand)
    want = orank.sorted_edges(cand, cn, 5000)
    m = models.CommonNeighborsPredictor(None, 0, None, None, None, None, model_type="simple")
    for slab in (1 << 27, 20000):
        got = filter_step.filter_topk("simple", m, None, adj, k=5000, slab_pairs=slab)
        assert np.array_equal(got.cpu().numpy(), want)
    full = filter_step.filter_topk("simple", m, None, adj, k=None, slab_pairs=50000)
    assert np.array_equal(full.cpu().numpy(), orank.sorted_edges(cand, cn))
    # GCN filter: scores within tolerance, identical ordering wherever the oracle's gaps exceed it
    n, H, L = g.n, 256, 2
    sd = ognn.random_state_dict("gcn", n, 0, H, L)
    args = argparse.Namespace(model="gcn", dataset="x", num_layers=L, hidden_channels=H, dropout=0.0,
                              use_feature=False, use_learnable_embedding=True)

    class D:
        num_nodes = n
        x = None
    mg = models.build_model(args, D, DEV)
    mg.load_state_dict(sd)
    got = filter_step.filter_topk("gcn", mg, None, adj, k=2000, slab_pairs=30000).cpu().numpy()
    h64 = ognn.gcn_forward(g, sd["emb.weight"], sd, L, torch.float64)
    sc64 = ognn.linkpred_forward(h64, cand, sd, L, torch.float64).numpy()
    lookup = {(int(a), int(b)): i for i, (a, b) in enumerate(cand.T)}
    idx = np.array([lookup[(int(a), int(b))] for a, b in got[:, :2]])
    assert np.max(np.abs(got[:, 2] - sc64[idx])) <= 1e-5
    kth = np.sort(sc64)[::-1][1999]
    assert np.all(sc64[idx] >= kth - 2e-5)                       # nothing outside the tolerance band got in
    assert np.all(np.diff(got[:, 2]) <= 0)


@pytest.mark.parametrize("weighted", [False, True])
def test_gcn_norm_kernels_vs_oracle(weighted):
    """K1b: diagonal SET to 1 / inserted where missing, dinv = deg^-1/2, (w*dinv_i)*dinv_j."""
    s, ei, w, g = synth_graph("small", dataset="collab" if weighted else None)
    if weighted:
        rng = np.random.default_rng(5)
        wts = rng.integers(1, 5, size=ei.shape[1] // 2).astype(np.float32)
        # add a few self loops with odd weights: fill_diag must overwrite them with 1
        loops = np.stack([np.arange(0, 60, 3), np.arange(0, 60, 3)])
        ei2 = np.concatenate([ei, loops], 1)
        w2 = np.concatenate([wts, wts, np.full(loops.shape[1], 7.0, np.float32)])
        g = og.add_edges("collab", ei2, w2, np.zeros((2, 0), np.int64), s["n"])
    adj = to_adj(g, DEV)
    rp, c, v = ognn.gcn_norm(g)
    rp2, c2, v2 = adj.gcn_norm()
    assert np.array_equal(rp2.cpu().numpy(), rp)
    nnz2 = int(rp[-1])
    assert np.array_equal(c2.cpu().numpy()[:nnz2], c)
    assert np.allclose(v2.cpu().numpy()[:nnz2], v, rtol=1e-6, atol=0)
