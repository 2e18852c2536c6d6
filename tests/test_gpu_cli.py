"""The drop-in CLI (filter.py / rank.py) end to end on the GPU, against the oracle pipeline:
same argv as submit_job.py emits, same output file name and format."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import graph as og, heuristics as oh, ranking as orank
from util import golden_graph

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(280)
def test_filter_cli_adamic_ogb_then_rank_eval(tmp_path):
    env = dict(os.environ, PYTHONPATH=ROOT)
    k = 30000
    r = subprocess.run([sys.executable, "-u", os.path.join(ROOT, "filter.py"), "--dataset", "fb", "--model", "adamic_ogb",
                        "--checkpoint", "fb_adamic_ogb||0|0.pt", "--topk", str(k)], cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=250)
    assert r.returncode == 0, r.stderr[-2000:]
    out = tmp_path / "filtered_edges" / "fb_adamic_ogb__0_0_sorted_edges.pt"     # filter.py:164 naming
    assert out.exists()
    got = torch.load(out).numpy()
    assert got.dtype == np.float32 and got.shape == (k, 3)
    z, ei, g = golden_graph("fb")
    cand = og.two_hop_candidates(g)
    assert f"using {cand.shape[1]} edges" in r.stdout
    aa_seq = oh.aa_ogb_pairs(g, cand, order="exact")
    # device weight table (CUDA logf) vs numpy logf: scores agree to 1e-5 relative, so compare the
    # list through the oracle's scores of the SAME pairs and require a consistent ordering
    lookup = {(int(a), int(b)): i for i, (a, b) in enumerate(cand.T)}
    idx = np.array([lookup[(int(a), int(b))] for a, b in got[:, :2]])
    assert np.all(np.abs(got[:, 2] - aa_seq[idx]) <= 1e-5 * np.abs(aa_seq[idx]))
    assert np.all(np.diff(got[:, 2]) <= 0)
    kth = np.sort(aa_seq)[::-1][k - 1]
    assert np.all(aa_seq[idx] >= kth * (1 - 2e-5))
    # rank.py consumes the file: CN rank model ('simple'), two sweep points, Hits@K printed
    r2 = subprocess.run([sys.executable, "-u", os.path.join(ROOT, "rank.py"), "--dataset", "fb", "--model", "simple",
                         "--sorted_edge_path", "fb_adamic_ogb__0_0_sorted_edges.pt", "--sweep_min", "0",
                         "--sweep_max", "20000", "--sweep_num", "2", "--runs", "1"], cwd=tmp_path, env=env,
                        capture_output=True, text=True, timeout=250)
    assert r2.returncode == 0, r2.stderr[-2000:]
    assert "Scheduled extra edges sweep: [0, 10000, 20000] x 1" in r2.stdout
    assert r2.stdout.count("Hits@20") >= 3
    assert len(list((tmp_path / "curves").glob("*.pt"))) == 3


@pytest.mark.timeout(200)
def test_rank_side_hits_match_oracle():
    """evaluate(): Hits@K of the CN rank model on the proposal-augmented graph == oracle (identical)."""
    from edge_proposal_sets_b200 import models, rank_step
    from edge_proposal_sets_b200.data import get_data
    import argparse
    dev = torch.device("cuda:0")
    a = argparse.Namespace(dataset="fb", use_feature=False)
    edge_index, edge_weight, split_edge, data = get_data(a, dev)
    z, ei, g0 = golden_graph("fb")
    cand, cn = og.two_hop_candidates(g0, return_values=True)
    prop = orank.sorted_edges(cand, cn.astype(np.float32), 15000)
    extra = torch.from_numpy(orank.prefix_edges(prop, 15000))
    adj, full = rank_step.augmented_graphs("fb", edge_index, edge_weight, extra, split_edge, data.num_nodes, dev)
    m = models.CommonNeighborsPredictor(None, 0, None, None, None, None, model_type="simple")
    got = rank_step.evaluate("simple", m, None, adj, full, split_edge, "fb")
    g = og.add_edges("fb", edge_index.numpy(), edge_weight.numpy(), extra.numpy(), data.num_nodes)
    sc = lambda e: oh.cn_scores_pairs(g, e.t().numpy())
    pv, nv = sc(split_edge["valid"]["edge"]), sc(split_edge["valid"]["edge_neg"])
    pt, nt = sc(split_edge["test"]["edge"]), sc(split_edge["test"]["edge_neg"])
    ptr = sc(split_edge["eval_train"]["edge"])
    for K in rank_step.HITS["fb"]:
        want = (orank.hits_at_k(ptr, nv, K), orank.hits_at_k(pv, nv, K), orank.hits_at_k(pt, nt, K))
        assert got[f"Hits@{K}"] == pytest.approx(want, abs=0)


@pytest.mark.timeout(200)
def test_email_shape_cn_filter_then_cn_rank(tmp_path):
    """BASELINE configs[0]: Common Neighbours filter + CN rank on the email graph (here its seeded
    synthetic shape: the bundled email-Eu-core file is one of the reference's missing blobs).  The saved
    proposal list must be the oracle's stable sort of ALL 2-hop candidates, bit for bit."""
    import argparse
    from edge_proposal_sets_b200.data import get_data
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-u", os.path.join(ROOT, "filter.py"), "--dataset", "email-shape", "--model", "simple",
                        "--checkpoint", "email-shape_simple||0|0.pt"], cwd=tmp_path, env=env, capture_output=True,
                       text=True, timeout=150)
    assert r.returncode == 0, r.stderr[-2000:]
    got = torch.load(tmp_path / "filtered_edges" / "email-shape_simple__0_0_sorted_edges.pt").numpy()
    torch.manual_seed(0)
    edge_index, edge_weight, split_edge, data = get_data(argparse.Namespace(dataset="email-shape", use_feature=False), "cpu")
    g = og.add_edges("email-shape", edge_index.numpy(), edge_weight.numpy(), np.zeros((2, 0), np.int64), data.num_nodes)
    cand, cn = og.two_hop_candidates(g, return_values=True)
    want = orank.sorted_edges(cand, cn.astype(np.float32))
    assert got.shape == want.shape and np.array_equal(got, want)        # k = None keeps every candidate, like the reference
    r2 = subprocess.run([sys.executable, "-u", os.path.join(ROOT, "rank.py"), "--dataset", "email-shape", "--model", "simple",
                         "--sorted_edge_path", "email-shape_simple__0_0_sorted_edges.pt", "--num_sorted_edge", "1500",
                         "--runs", "1"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=150)
    assert r2.returncode == 0, r2.stderr[-2000:]
    assert "Using 1500 highest scoring edges" in r2.stdout and "Hits@20" in r2.stdout


@pytest.mark.timeout(280)
def test_twitch_sage_filter_with_real_features_then_aa_rank_sweep(tmp_path):
    """BASELINE configs[2]: SAGE filter on twitch-DE with the real 2,514-column binary features (layer-1
    neighbour mean over 2,770-wide rows) -> proposal list -> Adamic-Adar rank with a proposal-size sweep.
    The checkpoint is written in the reference's state-dict layout (models/{dataset}_{model}||0|0.pt)."""
    import argparse
    from oracle import gnn as ognn
    from edge_proposal_sets_b200.data import get_data
    env = dict(os.environ, PYTHONPATH=ROOT)
    torch.manual_seed(0)
    edge_index, edge_weight, split_edge, data = get_data(argparse.Namespace(dataset="twitch", use_feature=True), "cpu")
    assert tuple(data.x.shape) == (9498, 2514) and float(data.x.sum()) == 193132.0
    n, H, L = data.num_nodes, 256, 3
    sd = ognn.random_state_dict("sage", n, data.x.shape[1], H, L, seed=4)
    (tmp_path / "models").mkdir()
    torch.save(sd, tmp_path / "models" / "twitch_sage||0|0.pt")
    k = 40000
    r = subprocess.run([sys.executable, "-u", os.path.join(ROOT, "filter.py"), "--dataset", "twitch", "--model", "sage",
                        "--checkpoint", "twitch_sage||0|0.pt", "--topk", str(k)], cwd=tmp_path, env=env,
                       capture_output=True, text=True, timeout=250)
    assert r.returncode == 0, r.stderr[-3000:]
    got = torch.load(tmp_path / "filtered_edges" / "twitch_sage__0_0_sorted_edges.pt").numpy()
    assert got.dtype == np.float32 and got.shape == (k, 3)
    z, ei, g = golden_graph("twitch")
    assert f"using {int(z['num_candidates'])} edges" in r.stdout
    # fp64 oracle of the same pairs: scores within tolerance, descending, nothing from outside the band
    xin = ognn.link_gnn_input(sd, data.x)
    h64 = ognn.sage_forward(g, xin, sd, L, torch.float64)
    uv = got[:, :2].astype(np.int64).T
    sc64 = ognn.linkpred_forward(h64, uv, sd, L, torch.float64).numpy()
    assert np.max(np.abs(got[:, 2] - sc64)) <= 2e-5
    assert np.all(np.diff(got[:, 2]) <= 0)
    cand = og.two_hop_candidates(g)
    all64 = ognn.linkpred_forward(h64, cand, sd, L, torch.float64).numpy()
    kth = np.sort(all64)[::-1][k - 1]
    assert np.all(sc64 >= kth - 4e-5)
    # the fp32 arm on every candidate gives the same file, bit for bit (prefilter == fp32)
    os.link(tmp_path / "models" / "twitch_sage||0|0.pt", tmp_path / "models" / "twitch_sage||0|1.pt")
    r32 = subprocess.run([sys.executable, "-u", os.path.join(ROOT, "filter.py"), "--dataset", "twitch", "--model", "sage",
                          "--checkpoint", "twitch_sage||0|1.pt", "--topk", str(k), "--mlp_precision", "fp32"],
                         cwd=tmp_path, env=env, capture_output=True, text=True, timeout=250)
    assert r32.returncode == 0, r32.stderr[-3000:]
    got32 = torch.load(tmp_path / "filtered_edges" / "twitch_sage__0_1_sorted_edges.pt").numpy()
    assert np.array_equal(got, got32)
    # rank: Adamic-Adar on the proposal-augmented graph, three sweep points (rank.py:260-272)
    r2 = subprocess.run([sys.executable, "-u", os.path.join(ROOT, "rank.py"), "--dataset", "twitch", "--model", "adamic_ogb",
                         "--sorted_edge_path", "twitch_sage__0_0_sorted_edges.pt", "--sweep_min", "0",
                         "--sweep_max", "40000", "--sweep_num", "2", "--runs", "1"], cwd=tmp_path, env=env,
                        capture_output=True, text=True, timeout=250)
    assert r2.returncode == 0, r2.stderr[-3000:]
    assert "Scheduled extra edges sweep: [0, 20000, 40000] x 1" in r2.stdout
    assert r2.stdout.count("Hits@50") >= 3                       # twitch evaluates Hits@10/50/100 (train_and_eval.py:20-29)
