"""Sharded filter step on 2 GPUs (skipped with fewer): torchrun-style spawn, NCCL all-gather merge
(parallel.merge_topk) and the raw-NCCL C-ABI merge (eps_topk_merge_allgather) must both reproduce
the single-GPU proposal list bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import ctypes as C
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from edge_proposal_sets_b200 import _lib, filter_step, models, ops, parallel
        from util import synth_graph, to_adj
        s, ei, w, g = synth_graph("small")
        adj = to_adj(g, dev)
        k = 4000
        m = models.CommonNeighborsPredictor(None, 0, None, None, None, None, model_type="simple")
        single = filter_step.filter_topk("simple", m, None, adj, k=k)                   # whole graph, this GPU
        sharded = filter_step.filter_topk("simple", m, None, adj, k=k, distributed=True, slab_pairs=100000)
        ok = torch.equal(single, sharded)
        aa_single = filter_step.filter_topk("adamic_ogb", None, None, adj, k=k)
        aa_sharded = filter_step.filter_topk("adamic_ogb", None, None, adj, k=k, distributed=True)
        ok = ok and torch.equal(aa_single, aa_sharded)
        # GCN filter: embeddings row-sharded + all-gathered == single-GPU bits; prefilter list (bf16 tcgen05 scores,
        # global k-th exchange, fp32 re-scoring) == the single-GPU fp32-arm list
        import argparse
        from oracle import gnn as ognn
        feat = s["x"].shape[1]
        sd = ognn.random_state_dict("gcn", g.n, feat, 256, 3, seed=3)
        args = argparse.Namespace(model="gcn", dataset="x", num_layers=3, hidden_channels=256, dropout=0.0,
                                  use_feature=True, use_learnable_embedding=True)

        class D:
            num_nodes = g.n
            x = torch.zeros(1, feat)
        mg = models.build_model(args, D, dev)
        mg.load_state_dict(sd)
        mg.eval()
        x = torch.from_numpy(s["x"]).to(dev)
        h1 = mg.embed(x, adj).clone()
        mg._h_key = None
        h2 = mg.embed(x, adj, distributed=True)
        ok = ok and torch.equal(h1, h2)
        g32 = filter_step.filter_topk("gcn", mg, x, adj, k=k, precision="fp32")
        st = {}
        g16 = filter_step.filter_topk("gcn", mg, x, adj, k=k, precision="prefilter", distributed=True,
                                      slab_pairs=150000, stats=st)
        ok = ok and torch.equal(g32, g16) and "prefilter_fallback" not in st
        both = filter_step.filter_topk_multi([filter_step.FilterJob("adamic_ogb", None), filter_step.FilterJob("gcn", mg)],
                                             x, adj, k=k, distributed=True, slab_pairs=200000)
        ok = ok and torch.equal(both[0], aa_single) and torch.equal(both[1], g32)
        # raw-NCCL C-ABI merge on the same local lists
        lib = _lib.load()
        idbuf = (C.c_char * 128)()
        if rank == 0:
            _lib.check(lib.eps_comm_unique_id(idbuf), "eps_comm_unique_id")
        idt = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device=dev)
        dist.broadcast(idt, 0)
        idbuf = (C.c_char * 128).from_buffer_copy(bytes(idt.cpu().tolist()))
        comm = C.c_void_p()
        _lib.check(lib.eps_comm_init(idbuf, world, rank, C.byref(comm)), "eps_comm_init")
        bounds = parallel.partition_by_work(__import__("edge_proposal_sets_b200.candidates", fromlist=["x"]).two_path_work(adj), world)
        from edge_proposal_sets_b200 import candidates
        edges = candidates.two_hop(adj, bounds[rank], bounds[rank + 1])
        sc = ops.cn_aa(adj, edges, None, grouped_by_v=True)
        local = parallel.pad_rows(ops.topk_edges(edges, sc, k), k)
        outk3 = torch.empty((k, 3), dtype=torch.float32, device=dev)
        ws = torch.empty(lib.eps_topk_merge_workspace_bytes(world, k, k), dtype=torch.uint8, device=dev)
        _lib.check(lib.eps_topk_merge_allgather(comm, local.data_ptr(), k, k, outk3.data_ptr(), ws.data_ptr(),
                                                ws.numel(), torch.cuda.current_stream().cuda_stream), "merge")
        torch.cuda.synchronize()
        ok = ok and torch.equal(outk3, single)
        lib.eps_comm_destroy(comm)
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_gpu_sharded_filter_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=280)
        assert p.exitcode == 0
    assert dict(out) == {0: True, 1: True}
