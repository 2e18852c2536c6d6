#!/usr/bin/env python
"""Score statistics of the GCN+LinkPredictor filter on one slab of a synthetic shape: spread of the fp32 scores,
deviation of the tensor-core arm, and how many candidates a band of +-margin around the top-q boundary holds.
usage: tools/prefilter_stats.py [workload=ppa] [slab_log2=26]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from edge_proposal_sets_b200 import candidates  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "ppa"
    slab = 1 << (int(sys.argv[2]) if len(sys.argv) > 2 else 26)
    dev = torch.device("cuda:0")
    args = argparse.Namespace(scale=1.0, mlp="prefilter", owners_frac=1.0, slab_pairs=slab, no_pushdown=False)
    wl = bench.Workload(name, args, dev, 1, pin=False)
    adj, x = wl.upload()
    h = wl.model.embed(x, adj)
    print("h: abs max %.3e  mean |h| %.3e" % (h.abs().max().item(), h.abs().mean().item()))
    bounds = torch.cumsum(candidates.owner_bounds(adj), 0)
    v_hi = int(torch.searchsorted(bounds, torch.tensor(slab, device=dev)).item())
    edges = candidates.two_hop(adj, 0, max(v_hi, 1))
    M = edges.shape[1]
    s16 = wl.model.linkpred.score_pairs(h, edges, "f16")
    idx = torch.randperm(M, device=dev)[: 1 << 22]
    e = edges[:, idx].contiguous()
    s32 = wl.model.linkpred.score_pairs(h, e, "fp32")
    l32 = torch.special.logit(s32.double())
    d = (s32 - s16[idx]).abs()
    print(f"{name}: slab candidates {M}; fp32 scores: mean {s32.mean().item():.6f} std {s32.std().item():.3e} "
          f"min {s32.min().item():.6f} max {s32.max().item():.6f}; logit std {l32.std().item():.3e}")
    qs = torch.tensor([0.5, 0.9, 0.99, 0.999, 0.9999, 1.0], device=dev, dtype=torch.float64)
    print("|tc - fp32| quantiles (50/90/99/99.9/99.99/max):", [f"{v:.2e}" for v in torch.quantile(d.double(), qs).tolist()])
    srt = torch.sort(s16, descending=True)[0]
    for frac in (0.0005, 0.005, 0.05):
        kk = max(int(M * frac), 1)
        T = srt[kk - 1].item()
        for mult in (1, 2, 4, 8):
            margin = 2 * mult * d.max().item()
            pool = int((s16 >= T - margin).sum().item())
            print(f"  top {frac:.2%} (k={kk}): margin = 2 x {mult} x max dev = {margin:.2e} -> pool {pool} = {pool / kk:.2f} k")


if __name__ == "__main__":
    main()
