"""The own-code oracle (oracle/model_ref.py) against the committed golden vectors and, where the
reference tree is mounted, against the reference's own classes run under the PyG shim."""
import glob
import os

import pytest
import torch

import oracle
from oracle import model_ref
from conftest import GOLDEN_DIR, Bag, load_golden
from gnn_elasticity_predictor_b200.synthetic import synthetic_batch, zscore_targets

MODEL_GOLDENS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN_DIR, "model_*.pt")))
BLOCK_GOLDENS = sorted(os.path.basename(p) for p in glob.glob(os.path.join(GOLDEN_DIR, "blocks_*.pt")))


def test_goldens_present():
    assert len(MODEL_GOLDENS) >= 5 and len(BLOCK_GOLDENS) >= 4


@pytest.mark.parametrize("name", MODEL_GOLDENS)
def test_model_ref_reproduces_golden(name):
    g = load_golden(name)
    ctor = g["ctor"]
    model = model_ref.HeteroAlignnRegressor(model_ref.AlignnRegressor(**ctor), ctor["target_dim"])
    model.load_state_dict(g["state_dict"], strict=True)
    model.train()
    batch = Bag(g["batch"], g["num_graphs"])
    mean, logvar = model(batch)
    # identical op sequence on the same machine class: bit-exact in practice; 1e-6 guards BLAS variance
    assert torch.allclose(mean, g["mean"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(logvar, g["logvar"], rtol=1e-6, atol=1e-7)
    loss = model_ref.gaussian_nll_loss(mean, logvar, zscore_targets(batch.y, batch.num_graphs))
    assert torch.allclose(loss, g["loss"], rtol=1e-6, atol=1e-7)
    loss.backward()
    grads = {k: p.grad for k, p in model.named_parameters() if p.grad is not None}
    assert set(grads) == set(g["grads"])
    for k, v in g["grads"].items():
        assert torch.allclose(grads[k], v, rtol=1e-5, atol=1e-7), k
    assert torch.allclose(model.embed(batch), g["embed"], rtol=1e-6, atol=1e-7)
    assert torch.allclose(model.base(batch), g["plain_output"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("name", BLOCK_GOLDENS)
def test_block_ref_reproduces_golden(name):
    g = load_golden(name)
    hidden, heads = g["hidden"], g["heads"]
    for tag, blk in (("edge_block", model_ref.EdgeUpdateBlock(hidden, heads, 0.0)),
                     ("node_block", model_ref.NodeUpdateBlock(hidden, hidden, heads, 0.0))):
        blk.load_state_dict(g[tag]["state_dict"], strict=True)
        x = g["x"].clone().requires_grad_(True)
        ea = g["edge_attr"].clone().requires_grad_(True)
        y = blk(x, g["index"], ea)
        y.backward(g["gout"])
        assert torch.allclose(y, g[tag]["y"], rtol=1e-6, atol=1e-6)
        assert torch.allclose(x.grad, g[tag]["dx"], rtol=1e-5, atol=1e-6)
        assert torch.allclose(ea.grad, g[tag]["dedge"], rtol=1e-5, atol=1e-6)


def test_state_dict_layout_counts():
    # 12 + 12*layers + 14*layers + 2 + 12 tensors
    for layers in (1, 4):
        m = model_ref.build_hetero(hidden=16, layers=layers, heads=2)
        assert len(m.state_dict()) == 12 + 12 * layers + 14 * layers + 2 + 12
    sd = model_ref.build_hetero(hidden=16, layers=1, heads=2).state_dict()
    assert sd["base.edge_blocks.0.conv.lin_beta.weight"].shape == (1, 48)
    assert "base.edge_blocks.0.conv.lin_edge.bias" not in sd and "base.node_blocks.0.edge_proj.bias" in sd
    assert sd["base.feat_proj.0.weight"].shape == (16, 16 + 289)


@pytest.mark.skipif(not oracle.reference_available(), reason="/root/reference not mounted (GPU box)")
def test_model_ref_bit_identical_to_reference_classes():
    ref = oracle.load_reference_train_module()
    ctor = dict(node_dim=206, edge_dim=36, angle_dim=11, global_dim=289, target_dim=2, hidden=64, layers=2, heads=4,
                dropout=0.0)
    torch.manual_seed(7)
    theirs = ref.HeteroAlignnRegressor(ref.AlignnRegressor(**ctor), 2)
    mine = model_ref.HeteroAlignnRegressor(model_ref.AlignnRegressor(**ctor), 2)
    assert list(theirs.state_dict().keys()) == list(mine.state_dict().keys())
    assert [tuple(v.shape) for v in theirs.state_dict().values()] == [tuple(v.shape) for v in mine.state_dict().values()]
    mine.load_state_dict(theirs.state_dict(), strict=True)
    for lg_inc in ("pyg", "bonds"):
        batch = synthetic_batch(5, 10, 6, seed=3, lg_inc=lg_inc)
        a, b = theirs(batch), mine(batch)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
        assert torch.equal(theirs.embed(batch), mine.embed(batch))
    # error behaviour (train.py:342-347)
    for cls in (ref.AlignnRegressor, model_ref.AlignnRegressor):
        with pytest.raises(ValueError):
            cls(**{**ctor, "heads": 0})
        with pytest.raises(ValueError):
            cls(**{**ctor, "target_dim": 0})
        with pytest.raises(ValueError):
            cls(**{**ctor, "hidden": 30})
