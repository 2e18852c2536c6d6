#!/usr/bin/env python
"""rank.py — the hot-path side of /root/reference/rank.py behind the same argv and file names.

What is kept (the consumer of the filter step, SURVEY §8 a1/a12 and "next" row 3):
  * ``from rank import get_data, add_edges, get_dataset`` for filter.py / submit_job.py
    (/root/reference/filter.py:17, submit_job.py:8);
  * ``--sorted_edge_path / --num_sorted_edge / --sweep_min / --sweep_max / --sweep_num /
    --valid_proposal``: load ``filtered_edges/{path}``, take each scheduled prefix (rank.py:260-294),
    rebuild ``adj_t`` / ``full_adj_t`` with ``add_edges`` (rank.py:299-314) and evaluate Hits@K
    (train_and_eval.py:98-270) on the GPU kernels; results are printed like the reference and the
    curve point ``[index_end, valid, test]`` is saved under ``curves/``.
  * the run / epoch loop of rank.py:317-385: ``model.reset_parameters()``, Adam, one
    ``train_step.train`` epoch (train_and_eval.py:31-96; K1 forward AND backward through
    ``autograd.spmm``), evaluation every ``--eval_steps``, ``--save_models`` writes
    ``models/{out_name}|{stem}|{index_end}|{run}.pt`` at every new best validation Hits@K[1] — the
    checkpoint ``filter.py`` loads (``{dataset}_{model}||0|0.pt`` for the filter model of
    submit_job.py:15-17).  Heuristic rank models (simple / adamic / adamic_ogb / resource_allocation)
    have no parameters and are evaluated once, as in the reference.
Extension: ``--epochs 0`` evaluates an existing checkpoint of that name without training.
"""
from __future__ import annotations

import argparse
import os
from datetime import datetime
from pathlib import Path

import torch

from edge_proposal_sets_b200.data import get_data, get_dataset  # noqa: F401  (re-exported)
from edge_proposal_sets_b200.graph import add_edges  # noqa: F401  (re-exported)


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="rank step, evaluation side (B200-native)")
    p.add_argument("--dataset", type=str, required=True)
    p.add_argument("--model", type=str)
    p.add_argument("--runs", type=int, default=10)
    p.add_argument("--sorted_edge_path", type=str, default="")
    p.add_argument("--num_sorted_edge", type=int)
    p.add_argument("--sweep_max", type=int)
    p.add_argument("--sweep_min", type=int)
    p.add_argument("--sweep_num", type=int)
    p.add_argument("--only_supervision", action="store_true", default=False)
    p.add_argument("--also_supervision", action="store_true", default=False)
    p.add_argument("--gen_dataset_only", action="store_true", default=False)
    p.add_argument("--valid_proposal", action="store_true", default=False)
    p.add_argument("--out_name", type=str)
    p.add_argument("--save_models", action="store_true", default=False)
    p.add_argument("--num_layers", type=int)
    p.add_argument("--hidden_channels", type=int)
    p.add_argument("--dropout", type=float)
    p.add_argument("--batch_size", type=int)
    p.add_argument("--lr", type=float)
    p.add_argument("--epochs", type=int)
    p.add_argument("--use_feature", type=bool)
    p.add_argument("--use_learnable_embedding", type=bool)
    p.add_argument("--device", type=int, default=0)
    p.add_argument("--log_steps", type=int, default=1)
    p.add_argument("--eval_steps", type=int, default=1)
    return p.parse_args(argv)


def main(argv=None):
    from edge_proposal_sets_b200 import _lib, rank_step, train_step
    from edge_proposal_sets_b200.models import build_model, default_model_configs

    args = parse_args(argv)
    if args.model is None and not args.gen_dataset_only:
        raise SystemExit("Model not specified")
    if args.model is not None:
        args = default_model_configs(args)
    print(args)
    Path("curves").mkdir(exist_ok=True)
    Path("models").mkdir(exist_ok=True)
    assert not (args.only_supervision and args.also_supervision)
    if args.out_name is None and args.model is not None:
        args.out_name = args.dataset + "_" + args.model
        if args.only_supervision:
            args.out_name += "_onlys"
        elif args.also_supervision:
            args.out_name += "_alsos"
        elif args.valid_proposal:
            args.out_name += "_validproposal"
    if args.gen_dataset_only:
        args.use_feature = bool(args.use_feature)
        get_data(args, "cpu")
        return
    if not torch.cuda.is_available():
        raise SystemExit("rank.py: no CUDA device (this build has no CPU fallback)")
    _lib.load()
    device = torch.device("cuda", args.device)
    torch.cuda.set_device(device)

    edge_index, edge_weight, split_edge, data = get_data(args, device)
    model = build_model(args, data, device)
    model.eval()
    print(f"using model {model}")
    name = "collab" if args.dataset.startswith("collab") else args.dataset
    base = args.dataset.split("-shape")[0]
    K = rank_step.HITS[base]
    print("Evaluating at hits: ", K)

    if args.sorted_edge_path:
        sorted_test_edges = torch.load(f"filtered_edges/{args.sorted_edge_path}")
        print("sorted test edges", sorted_test_edges.size())
        if args.valid_proposal:
            sorted_test_edges = rank_step.valid_proposal(sorted_test_edges, split_edge["valid"]["edge"])
    else:
        sorted_test_edges = torch.zeros(42, 2)
    index_ends = rank_step.sweep_index_ends(args.sweep_num, args.sweep_min, args.sweep_max, args.num_sorted_edge)
    print(f"Scheduled extra edges sweep: {index_ends} x {args.runs}")

    use_params = sum(p.numel() for p in model.parameters() if p.requires_grad) > 0
    stem = args.sorted_edge_path.split(".")[0]
    for index_end in index_ends:
        loggers = {f"Hits@{k}": rank_step.RunLog(args.runs) for k in K}
        print("---------------------")
        print(f"Using {index_end} highest scoring edges")
        print("---------------------")
        extra = rank_step.prefix_edges(sorted_test_edges, index_end)
        if args.only_supervision:                                       # rank.py:299-300: graph untouched
            no_extra = torch.zeros([2, 0], dtype=torch.long)
            adj, full_adj = rank_step.augmented_graphs(name, edge_index, edge_weight, no_extra, split_edge,
                                                       data.num_nodes, device, eval_extra=extra)
        else:
            adj, full_adj = rank_step.augmented_graphs(name, edge_index, edge_weight, extra, split_edge,
                                                       data.num_nodes, device)
        data.adj_t, data.full_adj_t = adj, full_adj
        if args.only_supervision or args.also_supervision:               # rank.py:304-305
            split_edge["train"]["edge"] = torch.cat((split_edge["train"]["edge"], extra.t().cpu()))
        for run in range(args.runs):
            curve_point = None
            model.reset_parameters()
            print(sum(p.numel() for p in model.parameters() if p.requires_grad))
            optimizer = torch.optim.Adam(model.parameters(), lr=args.lr) if use_params else None
            ckpt = os.path.join("models", f"{args.out_name}|{stem}|{index_end}|{run}.pt")
            epochs = args.epochs if use_params else 1
            if use_params and args.epochs == 0:
                # evaluation only: score a checkpoint trained elsewhere (reference layout)
                model.load_state_dict(torch.load(ckpt, map_location=device))
                epochs = 1
            highest_eval = 0
            for epoch in range(1, 1 + epochs):
                loss = -1
                if use_params and args.epochs > 0:
                    loss = train_step.train(model, data, name, split_edge, optimizer, args.batch_size, use_params,
                                            args.model, device)
                if epoch % args.eval_steps != 0:
                    continue
                model.eval()
                results = rank_step.evaluate(args.model, model, data.x, adj, full_adj, split_edge, name)
                for key, result in results.items():
                    loggers[key].add_result(run, result)
                if epoch % args.log_steps == 0:
                    for key, (tr, va, te) in results.items():
                        if key == f"Hits@{K[1]}" and va >= highest_eval:
                            highest_eval = va
                            if args.save_models and use_params and args.epochs > 0:
                                torch.save(model.state_dict(), ckpt)
                        print(key)
                        print(f"Run: {run + 1:02d}, Epoch: {epoch:02d}, Loss: {loss:.4f}, Train: {100 * tr:.2f}%, "
                              f"Valid: {100 * va:.2f}%, Test: {100 * te:.2f}%")
                    print("---")
            for key in loggers:
                print(key)
                loggers[key].print_statistics(run)
                if key == f"Hits@{K[1]}":
                    curve_point = loggers[key].curve_point(run, index_end)
            time = datetime.now().strftime("%Y-%m-%d-%H:%M:%S")
            filename = f"{args.out_name}|{stem}|{index_end}|{time}.pt"
            print(curve_point)
            print("Saving curve to ", filename)
            torch.save(curve_point, os.path.join("curves", filename))
        for key in loggers:
            print(key)
            loggers[key].print_statistics()


if __name__ == "__main__":
    main()
