"""world_size-2 gloo run of the multi-GPU host logic on the CPU: owner partitioning by 2-path work,
per-rank candidate shards, per-rank sorted proposal lists, the all-gather plumbing
(parallel.allgather_rows / pad_rows / merge_topk) and the property the design rests on — merging by
"score desc, position-in-gathered-array asc" reproduces the global stable order bit for bit.
The K4 select itself is a CUDA kernel; here the oracle's stable sort stands in for it through
merge_topk's `select` hook (the GPU merge is covered by tests/test_gpu_multi.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from edge_proposal_sets_b200 import candidates, parallel
        from oracle import graph as og, heuristics as oh, ranking as orank
        from util import synth_graph, to_adj
        s, ei, w, g = synth_graph("small")
        adj = to_adj(g, "cpu")
        k = 3000
        cand, cn = og.two_hop_candidates(g, return_values=True)
        aa = oh.aa_ogb_pairs(g, cand)
        bounds = parallel.partition_by_work(candidates.two_path_work(adj), world)
        lo, hi = bounds[rank], bounds[rank + 1]
        mine = (cand[1] >= lo) & (cand[1] < hi)            # owners are all_edges[1]
        results = {}
        for name, score in (("cn", cn.astype(np.float32)), ("aa", aa)):
            local = orank.sorted_edges(cand[:, mine], score[mine], k)       # this rank's proposal list
            select = lambda sc, kk: torch.from_numpy(orank.stable_order_desc(sc.numpy())[:kk])
            merged = parallel.merge_topk(torch.from_numpy(local), k, select=select)
            results[name] = (merged.numpy(), orank.sorted_edges(cand, score, k))
        ok = all(np.array_equal(a, b) for a, b in results.values())
        # shards are contiguous, disjoint and cover every candidate
        cnt = torch.tensor([int(mine.sum())])
        dist.all_reduce(cnt)
        ok = ok and int(cnt) == cand.shape[1] and bounds[0] == 0 and bounds[-1] == g.n
        # a rank owning fewer than k candidates pads with -inf rows that never surface
        short = parallel.pad_rows(torch.from_numpy(results["cn"][1][:5]), 8)
        ok = ok and short.shape == (8, 3) and bool(torch.isinf(short[5:, 2]).all())
        # ---- global k-th score exchange: two all-reduced histograms == k-th key of the concatenation ----
        rng = np.random.default_rng(7 + rank)
        for scores, kk in ((rng.standard_normal(5000).astype(np.float32), 1234),
                           (rng.integers(0, 6, 4000).astype(np.float32), 3999),          # tie-heavy CN-like scores
                           (np.array([0.5, -0.0, 0.0, np.inf, -np.inf, 1.0], np.float32), 7),
                           (rng.random(10).astype(np.float32), 500)):                     # fewer than k in total
            t = torch.from_numpy(scores)
            both = [torch.empty_like(t) for _ in range(world)]
            dist.all_gather(both, t)
            allk = np.sort(parallel.order_keys(torch.cat(both)).numpy())
            want = int(allk[kk - 1]) if kk <= allk.size else 0xFFFFFFFF
            got = int(parallel.global_kth_key(t, kk).item()) & 0xFFFFFFFF
            ok = ok and got == want
        # order_keys is the device kernel's key map (ops.score_to_key restates csrc/topk.cu score_key)
        from edge_proposal_sets_b200 import ops
        probe = np.array([1.5, 0.0, -0.0, -3.25, 1e-30, np.inf], np.float32)
        ok = ok and [int(x) for x in parallel.order_keys(torch.from_numpy(probe))] == [ops.score_to_key(float(x)) for x in probe]
        ok = ok and all(ops.key_to_score(ops.score_to_key(float(x))) == float(x) + 0.0 for x in probe)
        # ---- row-sharded embeddings plumbing: block-aligned nnz partition, uneven slabs gathered in place ----
        n = g.n
        R = parallel.row_block(n)
        rb = parallel.row_partition(adj.rowptr, n, world)
        ok = ok and rb[0] == 0 and rb[-1] == n and all(b % R == 0 or b == n for b in rb) and rb == sorted(rb)
        full = torch.arange(n * 3, dtype=torch.float32).reshape(n, 3)
        for bnds in (rb, [0, R, n], [0, n, n]):
            got = parallel.allgather_row_slabs(full[bnds[rank]:bnds[rank + 1]].clone(), bnds)
            ok = ok and torch.equal(got, full)
        xw = torch.randn(n, 8, generator=torch.Generator().manual_seed(1))
        wmat = torch.randn(8, 5, generator=torch.Generator().manual_seed(2))
        mine_rows = parallel.block_matmul(xw[rb[rank]:rb[rank + 1]], wmat, rb[rank], rb[rank + 1], n)
        whole = parallel.block_matmul(xw, wmat, 0, n, n)
        ok = ok and torch.equal(mine_rows, whole[rb[rank]:rb[rank + 1]])       # a row's bits do not depend on the shard
        out[rank] = ok
    finally:
        dist.destroy_process_group()


def test_two_rank_shard_and_merge():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert dict(out) == {0: True, 1: True}
