"""The training path on the GPU: gradients of one train_and_eval.train step through
autograd.spmm (K1 forward + K1 backward) against fp64 torch-CPU autograd of the oracle restatement,
a short training run, and the submit_job.py flow  rank.py --save_models -> filter.py -> rank.py."""
import argparse
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import gnn as ognn
from util import synth_graph, to_adj

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(model_name, n, feat, H, L, sd):
    from edge_proposal_sets_b200 import models
    args = argparse.Namespace(model=model_name, dataset="tiny", num_layers=L, hidden_channels=H, dropout=0.0,
                              use_feature=feat > 0, use_learnable_embedding=True)

    class D:
        num_nodes = n
        x = torch.zeros(n, feat)
    m = models.build_model(args, D, torch.device(DEV))
    m.load_state_dict(sd)
    return m


@pytest.mark.parametrize("model_name,weighted", [("gcn", False), ("gcn", True), ("sage", False)])
def test_train_step_gradients_match_fp64_autograd(model_name, weighted):
    from edge_proposal_sets_b200 import train_step
    s, ei, w, g = synth_graph("tiny", dataset="collab" if weighted else None)
    if weighted:
        rng = np.random.default_rng(3)
        # symmetric integer weights on the stored entries (collab keeps values)
        key = np.minimum(np.repeat(np.arange(g.n), np.diff(g.rowptr)), g.col) * g.n + \
            np.maximum(np.repeat(np.arange(g.n), np.diff(g.rowptr)), g.col)
        uk, inv = np.unique(key, return_inverse=True)
        g.val = rng.integers(1, 5, uk.size).astype(np.float32)[inv]
    adj = to_adj(g, DEV, keep_values=weighted)
    n, H, L = g.n, 64, 3
    feat = s["x"].shape[1]
    sd = ognn.random_state_dict(model_name, n, feat, H, L)
    model = _build(model_name, n, feat, H, L, sd)
    model.train()
    x = torch.from_numpy(s["x"]).to(DEV)
    rng = np.random.default_rng(7)
    pos = torch.from_numpy(ei[:, rng.permutation(ei.shape[1])[:400]])
    neg = torch.from_numpy(rng.integers(0, n, size=(2, 400)))
    out = model(x, torch.cat([pos, neg], 1).to(DEV), adj).reshape(-1)
    assert out.requires_grad
    loss = train_step.link_loss(out[:400], out[400:])
    loss.backward()

    sd64 = {k: v.double().clone().requires_grad_(True) for k, v in sd.items()}
    xin = torch.cat([sd64["emb.weight"], torch.from_numpy(s["x"]).double()], 1)
    fwd = ognn.gcn_forward if model_name == "gcn" else ognn.sage_forward
    h64 = fwd(g, xin, sd64, L, torch.float64)
    o64 = ognn.linkpred_forward(h64, torch.cat([pos, neg], 1).numpy(), sd64, L, torch.float64)
    l64 = train_step.link_loss(o64[:400], o64[400:])
    l64.backward()
    assert float(loss.detach()) == pytest.approx(float(l64.detach()), rel=2e-6)
    for k, p in model.named_parameters():
        gw, gg = sd64[k].grad.numpy(), p.grad.double().cpu().numpy()
        scale = max(np.abs(gw).max(), 1e-12)
        assert np.abs(gg - gw).max() <= 2e-5 * scale, (k, np.abs(gg - gw).max(), scale)
        assert np.abs(gw).max() > 0, k


def test_eval_mode_unchanged_after_training_step():
    """model.eval() after an optimizer step takes the fused kernels again and sees the NEW weights
    (the embedding cache is keyed on parameter versions)."""
    from edge_proposal_sets_b200 import train_step
    s, ei, w, g = synth_graph("tiny")
    adj = to_adj(g, DEV)
    n, H, L = g.n, 64, 2
    sd = ognn.random_state_dict("gcn", n, 0, H, L)
    model = _build("gcn", n, 0, H, L, sd)
    edges = torch.from_numpy(ei[:, :300]).to(DEV)
    model.eval()
    before = model(None, edges, adj).clone()
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    data = argparse.Namespace(adj_t=adj, num_nodes=n, x=None)
    split = {"train": {"edge": torch.from_numpy(s["train_edges"])}}
    gen = torch.Generator(device=DEV).manual_seed(0)
    l0 = train_step.train(model, data, "tiny", split, opt, 512, True, "gcn", DEV, generator=gen)
    for _ in range(8):
        l1 = train_step.train(model, data, "tiny", split, opt, 512, True, "gcn", DEV, generator=gen)
    assert l1 < l0, (l0, l1)
    model.eval()
    after = model(None, edges, adj)
    assert not after.requires_grad and after.shape == before.shape
    assert float((after - before).abs().max()) > 1e-4
    sd_now = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    h64 = ognn.gcn_forward(g, sd_now["emb.weight"], sd_now, L, torch.float64)
    want = ognn.linkpred_forward(h64, ei[:, :300], sd_now, L, torch.float64).numpy()
    assert np.abs(after.reshape(-1).cpu().numpy() - want).max() <= 1e-5
    assert float(after.mean()) > float(before.mean())            # positives scored higher after training


@pytest.mark.timeout(600)
def test_submit_job_flow_train_filter_rank(tmp_path):
    """submit_job.py:15-21: train the filter model (rank.py --save_models --runs 1), score the 2-hop
    candidates with it (filter.py --checkpoint), rank with the proposal prefix (rank.py
    --sorted_edge_path --num_sorted_edge)."""
    env = dict(os.environ, PYTHONPATH=ROOT)
    run = lambda *a: subprocess.run([sys.executable, "-u", *a], cwd=tmp_path, env=env, capture_output=True,
                                    text=True, timeout=280)
    r = run(os.path.join(ROOT, "rank.py"), "--dataset", "email-shape", "--model", "gcn", "--runs", "1",
            "--epochs", "4", "--save_models")
    assert r.returncode == 0, r.stderr[-3000:]
    ckpt = tmp_path / "models" / "email-shape_gcn||0|0.pt"
    assert ckpt.exists(), r.stdout[-2000:]
    sd = torch.load(ckpt)
    assert sd["gnn.convs.0.weight"].shape == (300, 300) and "linkpred.lins.2.bias" in sd
    assert "Epoch: 04" in r.stdout and "Highest Valid" in r.stdout and "All runs:" in r.stdout
    by_epoch = {}
    for line in r.stdout.splitlines():
        if "Loss: " in line and "Epoch: " in line:
            by_epoch[int(line.split("Epoch: ")[1].split(",")[0])] = float(line.split("Loss: ")[1].split(",")[0])
    losses = [by_epoch[e] for e in sorted(by_epoch)]
    assert len(losses) == 4 and min(losses[1:]) < losses[0], losses
    r = run(os.path.join(ROOT, "filter.py"), "--dataset", "email-shape", "--model", "gcn",
            "--checkpoint", "email-shape_gcn||0|0.pt")
    assert r.returncode == 0, r.stderr[-3000:]
    out = tmp_path / "filtered_edges" / "email-shape_gcn__0_0_sorted_edges.pt"
    assert out.exists(), os.listdir(tmp_path / "filtered_edges")
    t = torch.load(out)
    assert t.dtype == torch.float32 and t.shape[1] == 3 and bool((t[1:, 2] <= t[:-1, 2]).all())
    r = run(os.path.join(ROOT, "rank.py"), "--dataset", "email-shape", "--model", "simple", "--runs", "1",
            "--sorted_edge_path", out.name, "--num_sorted_edge", "2000")
    assert r.returncode == 0, r.stderr[-3000:]
    assert "Using 2000 highest scoring edges" in r.stdout and r.stdout.count("Hits@20") >= 1


@pytest.mark.timeout(120)
def test_pair_hadamard_kernel_forward_bit_exact_backward_vs_autograd():
    """K7 (eps_pair_hadamard_f32 / _bwd_f32), the LinkPredictor's training-mode input: the forward equals
    h[u] * h[v] bit for bit, the backward equals autograd's scatter of the index_select / mul graph up to
    fp32 summation order (repeated nodes: many pairs add into one row)."""
    from edge_proposal_sets_b200 import autograd as ag
    torch.manual_seed(0)
    n, H, B = 500, 256, 20000
    h = torch.randn(n, H, device=DEV, requires_grad=True)
    e = torch.randint(0, n, (2, B), device=DEV)
    e[:, :50] = 7                                              # self pairs and a hot node
    z = ag.pair_hadamard(h, e)
    zr = h[e[0]] * h[e[1]]
    assert torch.equal(z, zr)
    g = torch.randn(B, H, device=DEV)
    (dh,) = torch.autograd.grad(z, h, g)
    h64 = h.detach().double().requires_grad_(True)
    (dr,) = torch.autograd.grad(h64[e[0]] * h64[e[1]], h64, g.double())
    assert float((dh.double() - dr).abs().max()) <= 1e-5 * float(dr.abs().max())
    # empty batch
    assert ag.pair_hadamard(h, e[:, :0]).shape == (0, H)
