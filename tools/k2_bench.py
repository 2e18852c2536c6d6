#!/usr/bin/env python
"""K2 micro-benchmark on a ppa-like pair list without generating the graph: n = 576,289 nodes, H = 256, L = 3,
owner runs of ~14,300 pairs with ascending random u (what a candidate slab looks like).  Prints kernel time,
TFLOP/s and checks the scores against the fp32 arm on a sample.   usage: tools/k2_bench.py [M_log2=25] [reps=5]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from edge_proposal_sets_b200 import ops  # noqa: E402


def main():
    mlog = int(sys.argv[1]) if len(sys.argv) > 1 else 25
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    dev = torch.device("cuda:0")
    n, H, L, run = int(os.environ.get('K2_N', 576289)), 256, 3, 14336
    M = 1 << mlog
    g = torch.Generator(device=dev).manual_seed(0)
    h = torch.randn(n, H, device=dev, generator=g) * 0.5
    owners = M // run
    v = torch.arange(owners, device=dev, dtype=torch.int32).repeat_interleave(run)
    u = torch.randint(0, n, (owners, run), device=dev, generator=g, dtype=torch.int32).sort(dim=1)[0].reshape(-1)
    edges = torch.stack([u, v])[:, :M].contiguous()
    M = edges.shape[1]
    Ws = [(torch.rand(H if i < L - 1 else 1, H, device=dev, generator=g) * 2 - 1) / 16 for i in range(L)]
    bs = [(torch.rand(H if i < L - 1 else 1, device=dev, generator=g) * 2 - 1) / 16 for i in range(L)]
    ctx = ops.LinkpredTC(h, Ws, bs)
    s = ctx.score(edges)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(reps):
        s = ctx.score(edges)
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    flops = M * (2 * H * H * (L - 1) + 3 * H)
    idx = torch.randint(0, M, (1 << 18,), device=dev, generator=g)
    ref = ops.linkpred_mlp(h, edges[:, idx].contiguous(), Ws, bs, "fp32")
    err = (ref - s[idx]).abs().max().item()
    print(f"K2 tcgen05: M={M} {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s  {M / ms / 1e6:.2f} G pairs/s  "
          f"max|bf16-fp32| on 2^18 = {err:.2e}  env={ {k: v for k, v in os.environ.items() if k.startswith('EPS_')} }")


if __name__ == "__main__":
    main()
