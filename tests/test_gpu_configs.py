"""BASELINE.json configs 4 and 5 as GPU parity / property cases (they are not bench lines):

* config 4 -- large-cell stress (200-atom supercells, 16 neighbours: tens of millions of line-graph angle edges at the full
  batch): a scaled batch through the bf16 training path, checked through size-independent properties (finite, bit-reproducible,
  graphs in a batch do not influence each other);
* config 5 -- ensemble inference (5 members, forward only, fp32 as ``predict.py`` runs it): mixture moments of
  ``ensemble.ensemble_forward`` against the oracle's members combined with the reference's formulas
  (``scripts/predict.py:604-623``)."""
import pytest
import torch

from conftest import rel_err
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import ensemble
from oracle import model_ref

pytestmark = pytest.mark.gpu
DEV = "cuda"
ARCH = dict(node_dim=206, edge_dim=36, angle_dim=11, global_dim=289, target_dim=2, hidden=256, layers=4, heads=4)


def _step(model, batch, tz):
    model.zero_grad(set_to_none=True)
    mean, logvar = model(batch)
    loss = pkg.gaussian_nll_loss(mean.float(), logvar.float(), tz)
    loss.backward()
    return mean.detach(), logvar.detach(), loss.detach(), {k: p.grad.clone() for k, p in model.named_parameters()
                                                           if p.grad is not None}


@pytest.mark.parametrize("lg_inc", ["pyg", "bonds"])
def test_config4_large_cells_scaled(lg_inc):
    torch.manual_seed(0)
    model = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.0, **ARCH), 2).to(DEV)
    model.base.compute_dtype = torch.bfloat16
    model.train()
    host = pkg.synthetic_batch(12, 200, 16, seed=3, lg_inc=lg_inc)
    assert host.sizes == {"B": 12, "N": 2400, "E": 38400, "L": 576000}
    batch = host.to(DEV)
    tz = pkg.zscore_targets(batch.y, batch.num_graphs)
    m1, v1, l1, g1 = _step(model, batch, tz)
    m2, v2, l2, g2 = _step(model, batch, tz)
    assert torch.isfinite(l1) and all(torch.isfinite(g).all() for g in g1.values())
    assert torch.equal(m1, m2) and torch.equal(v1, v2) and torch.equal(l1, l2)       # atomics-free kernels
    if lg_inc == "bonds":
        # with geometrically correct offsets the adjacency is block diagonal: a graph's prediction does not depend on
        # its batch mates (the pyg offsets mix bonds across crystals by construction)
        solo = pkg.synthetic_batch(12, 200, 16, seed=3, lg_inc="bonds")
        n, e, l = 200, 200 * 16, 200 * 16 * 15
        sub = pkg.GraphBatch(x=solo.x[:n], edge_index=solo.edge_index[:, :e], edge_attr=solo.edge_attr[:e],
                             lg_edge_index=solo.lg_edge_index[:, :l], lg_edge_attr=solo.lg_edge_attr[:l],
                             global_x=solo.global_x[:59], sg_one_hot=solo.sg_one_hot[:230], batch=solo.batch[:n],
                             y=solo.y[:2], train_idx=solo.train_idx[:1], num_graphs=1, lg_inc="bonds").to(DEV)
        with torch.no_grad():
            ms, vs = model(sub)
        assert rel_err(ms, m1[:1]) < 2e-2 and rel_err(vs, v1[:1]) < 2e-2


def test_config5_ensemble_inference_fp32_vs_oracle():
    members, refs = [], []
    for i in range(5):
        torch.manual_seed(42 + 1007 * i)                    # member seeds of the reference (train.py:2053)
        ref = model_ref.HeteroAlignnRegressor(model_ref.AlignnRegressor(dropout=0.15, **{**ARCH, "layers": 2}), 2).eval()
        ours = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.15, **{**ARCH, "layers": 2}), 2).to(DEV).eval()
        ours.load_state_dict(ref.state_dict(), strict=True)
        members.append(ours); refs.append(ref)
    host = pkg.synthetic_batch(10, 16, 12, seed=11)
    with torch.no_grad():
        want_m = torch.stack([r(host)[0] for r in refs]).double()
        want_v = torch.stack([r(host)[1] for r in refs]).double()
    mu = want_m.mean(0)
    var = torch.exp(want_v.clamp(min=-2.9)).mean(0) + want_m.pow(2).mean(0) - mu.pow(2)
    mean_z, var_z, std_z = ensemble.ensemble_forward(members, host.to(DEV))
    assert rel_err(mean_z, mu) < 1e-5 and rel_err(var_z, var) < 1e-5
    assert rel_err(std_z, var.clamp(min=1e-12).sqrt()) < 1e-5


def test_ensemble_predictor_graph_replay_matches_eager_members():
    from gnn_elasticity_predictor_b200 import engine
    members = []
    for i in range(3):
        torch.manual_seed(7 + i)
        members.append(pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.15, **{**ARCH, "layers": 2}), 2).to(DEV))
    pred = engine.EnsemblePredictor(members, compute_dtype=torch.bfloat16, graph=True, graph_warmup=1)
    batches = [pkg.synthetic_batch(12, 16, 12, seed=s).to(DEV) for s in (0, 1, 2)]
    for i, b in enumerate(batches + batches):
        got = pred.predict(b)
        want = ensemble.ensemble_forward(members, b)
        for x, y in zip(got, want):
            assert torch.equal(x, y), i                       # eval mode: no dropout, same kernels -> identical
    assert pred.replays == 5 and pred.eager_calls == 1
    # ragged batches through the bucket padding: real graphs' moments unchanged within bf16 tolerance
    from test_batching import ragged_batch
    ragged = engine.EnsemblePredictor(members, compute_dtype=torch.bfloat16, graph=True, graph_warmup=0, pad_to_buckets=True)
    for seed, sizes in ((1, (8, 12, 10, 16, 9)), (2, (9, 11, 10, 16, 9))):
        b = ragged_batch(seed, sizes=sizes).to(DEV)
        got = ragged.predict(b)
        want = ensemble.ensemble_forward(members, b)
        assert got[0].shape == (5, 2)
        for x, y in zip(got, want):
            assert rel_err(x, y) < 2e-2
    assert ragged.replays == 2 and len(ragged._captured) == 1
