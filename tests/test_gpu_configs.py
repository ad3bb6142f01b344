"""BASELINE.json configs 4 and 5 as GPU parity / property cases (they are not bench lines):

* config 4 -- large-cell stress (200-atom supercells, 16 neighbours: tens of millions of line-graph angle edges at the full
  batch): a scaled batch through the bf16 training path, checked through size-independent properties (finite, bit-reproducible,
  graphs in a batch do not influence each other);
* config 5 -- ensemble inference (5 members, forward only, fp32 as ``predict.py`` runs it): mixture moments of
  ``ensemble.ensemble_forward`` against the oracle's members combined with the reference's formulas
  (``scripts/predict.py:604-623``)."""
import pytest
import torch

from conftest import check_per_tensor, reference_amp_grads, rel_err
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


@pytest.mark.parametrize("lg_inc", ["pyg", "bonds"])
def test_config4_large_cells_vs_fp64_oracle(lg_inc):
    """Config 4's regime (200-atom cells, 16 neighbours: 240 in-edges per bond row with geometric offsets, thousands with
    PyG's) against the ORACLE: forward, loss and every parameter gradient of the default arch in bf16 vs the oracle in
    fp64 on a scaled batch (12 cells: N=2 400, E=38 400, L=576 000 -- the full config is the same per-graph shape), with
    the per-tensor policy of the config-2 test (2e-2 of each tensor's own scale, named exceptions bounded by the
    reference's own bf16-autocast error)."""
    import copy
    import os
    torch.set_num_threads(max(torch.get_num_threads(), (os.cpu_count() or 1)))
    ref = model_ref.build_hetero(hidden=256, layers=4, heads=4, seed=42)
    ours = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.0, **ARCH), 2).to(DEV)
    ours.load_state_dict(ref.state_dict(), strict=True)
    ours.train()
    host = pkg.synthetic_batch(12, 200, 16, seed=3, lg_inc=lg_inc)
    tz = pkg.zscore_targets(host.y, host.num_graphs)
    ref64 = copy.deepcopy(ref).double()
    b64 = copy.copy(host)
    for k in ("x", "edge_attr", "lg_edge_attr", "global_x", "sg_one_hot"):
        setattr(b64, k, getattr(host, k).double())
    r_mean, r_logvar = ref64(b64)
    r_loss = model_ref.gaussian_nll_loss(r_mean, r_logvar, tz.double())
    r_loss.backward()
    want = {k: p.grad.clone() for k, p in ref64.named_parameters() if p.grad is not None}
    amp = reference_amp_grads(ref, host, tz, model_ref.gaussian_nll_loss)
    ours.zero_grad(set_to_none=True)
    batch = host.to(DEV)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        mean, logvar = ours(batch)
        loss = pkg.gaussian_nll_loss(mean.float(), logvar.float(), pkg.zscore_targets(batch.y, batch.num_graphs))
    loss.backward()
    grads = {k: p.grad.detach().float().cpu() for k, p in ours.named_parameters() if p.grad is not None}
    assert rel_err(mean, r_mean) < 2e-2 and rel_err(logvar, r_logvar) < 2e-2 and rel_err(loss, r_loss) < 2e-2
    check_per_tensor(grads, want, 2e-2, {"conv.lin_key.bias": (("abs", 2e-3), "true gradient is exactly zero "
                                                                "(softmax shift invariance)")},
                     label=f"config4_scaled_bf16_{lg_inc}", amp=amp)


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
