"""Import adapters that let the UNMODIFIED reference functions run in the build container.

TEST INFRASTRUCTURE — see ``oracle/__init__.py``.  Used only by
``oracle/make_golden.py`` (to generate tests/golden) and by CPU tests that are
skipped when /root/reference is absent (it does not exist on the GPU box).

* ``adamic_utils.AA`` (/root/reference/adamic_utils.py:13-25) indexes a scipy
  matrix with a torch tensor, which scipy >= 1.14 rejects; ``NumpyIndexed``
  wraps ``edge_index`` so ``edge_index[0, ind]`` yields numpy.
* ``train_and_eval.resource_allocation`` (/root/reference/train_and_eval.py:195-216)
  needs ``ogb`` / ``torch_geometric`` at import; inert stubs are injected.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("EPS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "adamic_utils.py"))


class NumpyIndexed:
    """Quacks like the ``edge_index`` tensor as far as AA / resource_allocation touch it."""

    def __init__(self, t):
        self.t = torch.as_tensor(np.asarray(t)) if not torch.is_tensor(t) else t

    def size(self, d=None):
        return self.t.size() if d is None else self.t.size(d)

    def t_(self):
        return NumpyIndexed(self.t.t())

    def __getitem__(self, key):
        r, ind = key
        ind = torch.as_tensor(ind)
        return self.t[r, ind].numpy()


class _LinkList:
    """``link_list`` for resource_allocation: only ``.t()`` is called on it."""

    def __init__(self, edges_2xN):
        self.e = edges_2xN

    def t(self):
        return NumpyIndexed(self.e)


def _install_stubs():
    if "ogb.linkproppred" not in sys.modules:
        ogb = types.ModuleType("ogb")
        lp = types.ModuleType("ogb.linkproppred")

        class Evaluator:  # never evaluated by the functions we call
            def __init__(self, name=None):
                self.name = name

        class PygLinkPropPredDataset:  # pragma: no cover
            pass

        lp.Evaluator = Evaluator
        lp.PygLinkPropPredDataset = PygLinkPropPredDataset
        ogb.linkproppred = lp
        sys.modules["ogb"] = ogb
        sys.modules["ogb.linkproppred"] = lp
    if "torch_geometric.utils" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        tu = types.ModuleType("torch_geometric.utils")
        tu.negative_sampling = lambda *a, **k: None
        tu.to_undirected = lambda *a, **k: None
        tu.add_self_loops = lambda *a, **k: None
        tg.utils = tu
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.utils"] = tu


def reference_AA(A_scipy, edges_2xN: np.ndarray, batch_size: int = 2000) -> np.ndarray:
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import adamic_utils  # the reference module, unmodified
    pred, _ = adamic_utils.AA(A_scipy, NumpyIndexed(torch.as_tensor(np.asarray(edges_2xN))), batch_size)
    return pred.numpy()


def reference_RA(A_scipy, edges_2xN: np.ndarray, batch_size: int = 8192) -> np.ndarray:
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import train_and_eval  # the reference module, unmodified
    pred = train_and_eval.resource_allocation(A_scipy, _LinkList(torch.as_tensor(np.asarray(edges_2xN))), batch_size)
    return pred.numpy()


def reference_candidates(A_scipy_csr):
    """filter.py:96-109 with scipy standing in for torch_sparse: CSC product,
    masked assignment of explicit zeros, CSC->COO walk, keep non-zeros."""
    import warnings
    A = A_scipy_csr
    import scipy.sparse as ssp
    A2 = (A @ A).tocsc()
    A2 = (A2 - ssp.diags(A2.diagonal(), format="csc")).tocsc()   # remove_diag
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        A2[A.tocsc() > 0] = 0
    coo = A2.tocoo()
    sel = coo.data != 0
    return np.stack([coo.row[sel], coo.col[sel]]).astype(np.int64), coo.data[sel]


def _install_model_stubs():
    """Inert stand-ins for the third-party names /root/reference/models.py imports at module level
    (models.py:9-29), so that the module itself can be imported and its PLAIN-TORCH members
    (``LinkPredictor``, ``default_model_configs``) executed unmodified.  The stubs carry no
    arithmetic: anything that would need torch_geometric / torch_sparse raises when touched."""
    _install_stubs()

    class _Absent:
        def __init__(self, *a, **k):
            raise RuntimeError("torch_geometric / torch_sparse are not installed (inert import stub)")

    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
        for k, v in attrs.items():
            if not hasattr(m, k):
                setattr(m, k, v)
        return m

    tg = mod("torch_geometric")
    tg.transforms = mod("torch_geometric.transforms")
    tg.nn = mod("torch_geometric.nn", GCNConv=_Absent, SAGEConv=_Absent, TAGConv=_Absent, JumpingKnowledge=_Absent)
    tg.nn.conv = mod("torch_geometric.nn.conv", MessagePassing=torch.nn.Module)
    tg.typing = mod("torch_geometric.typing", OptPairTensor=object, Adj=object, Size=object)
    mod("torch_sparse", SparseTensor=_Absent, sum=_Absent, matmul=_Absent)


def reference_models_module():
    """The reference's ``models`` module, imported unmodified over the inert stubs above."""
    _install_model_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    m = importlib.import_module("models")
    assert os.path.abspath(m.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), m.__file__
    return m
