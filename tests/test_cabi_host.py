"""C-ABI library loads and exports every symbol include/alignn_b200.h declares; host-side module logic
(state_dict layout, error behaviour, loud failure without a GPU).  No kernel is launched here."""
import os
import re

import pytest
import torch

from conftest import ROOT, load_golden
from oracle import model_ref
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import _lib, build


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "alignn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(alignn_[a-z0-9_]+)\s*\(", text)))


def test_build_is_current_and_sm100a_only():
    path = build.build()
    assert os.path.exists(path) and build.is_current()
    assert "arch=compute_100a,code=sm_100a" in " ".join(build.NVCC_FLAGS)


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 11
    assert set(declared) == set(_lib.SIGNATURES), "binding table and header disagree"
    for name in declared:
        assert hasattr(lib, name), f"{name} not exported"
    assert lib.alignn_abi_version() == _lib.ABI_VERSION
    assert lib.alignn_error_string(0) == b"ok"
    assert b"workspace" in lib.alignn_error_string(3)
    assert lib.alignn_plan_workspace_bytes(1000, 10) > 6 * 4 * 1000
    assert lib.alignn_gate_ln_bwd_partial_rows() % 148 == 0      # persistent grid: multiple of the SM count


def test_sass_has_vectorised_loads_and_no_local_memory():
    """128-bit global loads in the conv kernels and zero spills (checked from the ptxas logs of the build)."""
    log = open(os.path.join(build.BUILD_DIR, "conv.ptxas.log")).read()
    assert "conv_fwd_kernel" in log
    assert not re.search(r"[1-9]\d* bytes spill", log)


def test_module_state_dict_matches_reference_layout():
    ctor = dict(node_dim=206, edge_dim=36, angle_dim=11, global_dim=289, target_dim=2, hidden=64, layers=2, heads=4,
                dropout=0.15)
    ours = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(**ctor), 2)
    ref = model_ref.HeteroAlignnRegressor(model_ref.AlignnRegressor(**ctor), 2)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    assert [tuple(v.shape) for v in ours.state_dict().values()] == [tuple(v.shape) for v in ref.state_dict().values()]
    ours.load_state_dict(ref.state_dict(), strict=True)
    # golden checkpoints (made by the reference's own classes) load strictly
    g = load_golden("model_default_dims_h32.pt")
    m = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(**g["ctor"]), 2)
    m.load_state_dict(g["state_dict"], strict=True)
    # attributes the reference's trainer reads (train.py:534,1516-1517)
    assert m.base.feat_proj[0].out_features == 32 and m.base.hidden == 32 and m.base.heads == 4
    assert len(list(m.mean_heads.parameters())) == 4 and len(list(m.logvar_heads.parameters())) == 4
    # evaluate.py:161 parses layer count from keys with ^base\.edge_blocks\.(\d+)\.
    layers = {int(re.match(r"^base\.edge_blocks\.(\d+)\.", k).group(1)) for k in m.state_dict() if k.startswith("base.edge_blocks.")}
    assert layers == {0, 1}


def test_constructor_errors_match_reference():
    ctor = dict(node_dim=6, edge_dim=8, angle_dim=7, global_dim=289, target_dim=2, hidden=32, layers=1, heads=1, dropout=0.0)
    for bad in ({"heads": 0}, {"target_dim": 0}, {"hidden": 30, "heads": 4}):
        with pytest.raises(ValueError):
            pkg.AlignnRegressor(**{**ctor, **bad})
    with pytest.raises(ValueError):
        pkg.EdgeUpdateBlock(30, 4, 0.0)
    with pytest.raises(ValueError):
        pkg.NodeUpdateBlock(30, 30, 4, 0.0)
    assert pkg.AlignnRegressor(**{**ctor, "angle_dim": 0}).angle_encoder is None


def test_cpu_tensors_fail_loudly():
    ctor = dict(node_dim=6, edge_dim=8, angle_dim=7, global_dim=289, target_dim=2, hidden=32, layers=1, heads=1, dropout=0.0)
    m = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(**ctor), 2)
    batch = pkg.synthetic_batch(2, 6, 4, seed=0, node_dim=6, edge_dim=8, angle_dim=7)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(batch)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.build_plan(batch.edge_index, 12)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        pkg.conv_core(torch.zeros(2, 8), torch.zeros(2, 8), torch.zeros(2, 8), torch.zeros(1, 8), None, 1)


def test_product_never_imports_oracle():
    pkg_dir = os.path.join(ROOT, "gnn_elasticity_predictor_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "pyg_shim" not in src and "torch_geometric" not in src.replace("torch_geometric.", "PYG."), f
