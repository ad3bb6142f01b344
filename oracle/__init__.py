"""``oracle/`` -- the CPU checker for the ALIGNN hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import, call or execute anything in here.  The
product (``gnn_elasticity_predictor_b200``) never does and fails loudly when
its CUDA library is missing.

PARITY UNPINNED for the third-party part: the reference ships no golden
vectors / numeric assertions for this path and the conv arithmetic lives in
``torch-geometric==2.7.0`` (``/root/reference/requirements.txt:9``), absent
from the mount and from this image.  See
``oracle/pyg_shim/torch_geometric/__init__.py`` for how that restatement is
self-pinned.  Everything the reference's OWN code computes (model classes,
loss, featuriser loops, ensemble post-processing) is pinned by fixtures made by
running that code here (``gen_golden*.py`` -> ``tests/golden/``).

Contents
--------
``pyg_shim/``     pure-torch stand-in for the PyG leaves the reference imports
``model_ref.py``  own-code restatement of the four model classes + loss
``conv_ref.c``    independent plain-C fp64 restatement of the conv core and
                  of the stable dst-sort (CSR) -- cross-checks the shim
``gen_golden.py`` imports the reference's OWN classes (``/root/reference``)
                  under the shim and writes ``tests/golden/*.pt``
``Makefile``      builds ``oracle/_ref/libconv_ref.so`` from ``conv_ref.c``
``linegraph_ref.py`` / ``gen_golden_linegraph.py``   featuriser loops of
                  ``scripts/fetch.py`` restated / golden vectors from the
                  reference's own ``build_graph_from_structure``
``ensemble_ref.py`` / ``gen_golden_ensemble.py``     ensemble post-processing
                  restated / golden vectors from the reference's own
                  ``ensemble_collect``, ``conformal_*``, ``LogTransformer``
"""
from __future__ import annotations

import importlib.util
import os
import sys

ORACLE_DIR = os.path.dirname(os.path.abspath(__file__))
SHIM_DIR = os.path.join(ORACLE_DIR, "pyg_shim")
REFERENCE_ROOT = "/root/reference"


def install_shim() -> None:
    """Put the PyG shim first on ``sys.path`` (idempotent)."""
    if SHIM_DIR not in sys.path:
        sys.path.insert(0, SHIM_DIR)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "scripts", "train.py"))


def load_reference_train_module():
    """Import the reference's ``scripts/train.py`` unmodified, under the shim.

    Returns the module, or ``None`` when ``/root/reference`` is not mounted (the
    GPU box).  Nothing is copied: the file is executed from where it lies.
    """
    if not reference_available():
        return None
    install_shim()
    name = "_reference_scripts_train"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, "scripts", "train.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
