"""CPU oracle for the Edge-Proposal-Sets filter-and-rank scoring path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``edge_proposal_sets_b200`` may import
this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the timed CPU baseline, never as the product path.

Every function restates, in numpy / scipy / torch-CPU, one piece of the
reference (``/root/reference``, cited file:line in each docstring).

Pinning status (see DESIGN.md "Oracle pinning"):

* ``heuristics.aa_ogb`` / ``heuristics.cn_scores`` / ``graph.two_hop_candidates``
  are PINNED against the reference's own ``adamic_utils.AA`` and scipy ``A@A``
  executed in the build container on the bundled twitch / fb graphs
  (``oracle/make_golden.py`` -> ``tests/golden/*.npz``).
* ``gnn.*`` (GCNConv / SAGEConv / SparseTensor semantics) restates third-party
  torch_geometric 1.7.0 / torch_sparse behaviour whose source is not under
  /root/reference and which cannot be installed here: **parity unpinned** for
  that part (the reference holds no test or golden vector at that boundary).
  ``LinkPredictor`` is plain torch and is pinned by construction.
"""
