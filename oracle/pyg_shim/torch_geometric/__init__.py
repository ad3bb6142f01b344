"""Minimal pure-torch stand-in for the parts of torch-geometric 2.7.0 that the
reference's hot path touches.

TEST INFRASTRUCTURE ONLY.  This package is part of ``oracle/`` -- the CPU
checker.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
CPU-baseline / ``--impl reference`` legs may import it.  The product package
(``gnn_elasticity_predictor_b200``) never does.

Why it exists: the reference (``/root/reference/scripts/train.py:25-27``)
imports ``torch_geometric.loader.DataLoader``,
``torch_geometric.data.Dataset`` and ``torch_geometric.nn.{TransformerConv,
global_mean_pool}``.  ``torch-geometric==2.7.0`` (``requirements.txt:9``) is not
vendored under ``/root/reference``, is not installed in this image and cannot
be installed (no network).  Putting this directory first on ``sys.path`` lets
the reference's *own* ``AlignnRegressor`` / ``HeteroAlignnRegressor`` classes
run unmodified on CPU, so they -- not a re-typed copy -- are the oracle.

PARITY UNPINNED: the reference holds no golden vectors or numeric assertions
for this path (``tests/smoke.py`` only checks the exit code), and real PyG is
unavailable, so this restatement of PyG's published algorithm is pinned only
by the known-answer tests in ``tests/test_oracle_kat.py`` (single-edge,
duplicate-edge, permutation invariance, dense masked-attention cross-check,
fp64 gradcheck) and by an independent plain-C fp64 restatement
(``oracle/conv_ref.c``).
"""

__version__ = "2.7.0+shim"

from . import data, loader, nn, utils  # noqa: F401
