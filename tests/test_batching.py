"""Shape-bucket padding (``batching.py``): host-side structure checks (CPU) and, on the GPU, that padding changes no real
graph's prediction or any parameter gradient, and that ragged batches replay one captured graph per bucket."""
import pytest
import torch

from conftest import rel_err
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import batching, engine

CTOR = dict(node_dim=206, edge_dim=36, angle_dim=11, global_dim=289, target_dim=2, hidden=256, layers=2, heads=4)


def ragged_batch(seed, lg_inc="pyg", sizes=(8, 12, 10, 16, 9)):
    gen = torch.Generator().manual_seed(seed)
    return pkg.collate([pkg.make_crystal(a, 6, gen) for a in sizes], lg_inc=lg_inc)


def test_round_up_bucket_grid():
    for x in (1, 255, 256, 257, 8193, 98304, 1081344, 12288000):
        y = batching.round_up_bucket(x)
        assert y >= x and y % 256 == 0 and (y - x) <= max(256, x // 8)
    assert batching.round_up_bucket(0) == 0


def test_pad_batch_structure():
    b = ragged_batch(0)
    s = b.sizes
    p, mask = batching.pad_batch(b)
    shape = batching.bucket_shape(b)
    assert p.sizes == shape and p.padded and not b.padded
    assert p.sizes["N"] > s["N"] and p.sizes["B"] > s["B"]
    assert torch.equal(p.x[:s["N"]], b.x) and float(p.x[s["N"]:].abs().max()) == 0.0
    assert torch.equal(p.edge_index[:, :s["E"]], b.edge_index) and bool((p.edge_index[:, s["E"]:] == -1).all())
    assert torch.equal(p.lg_edge_index[:, :s["L"]], b.lg_edge_index) and bool((p.lg_edge_index[:, s["L"]:] == -1).all())
    assert bool((p.batch[s["N"]:] == s["B"]).all())                    # padded atoms -> dummy graph
    assert p.global_x.shape == (59 * shape["B"], 1) and p.sg_one_hot.shape == (230 * shape["B"], 1)
    assert mask.tolist() == [1.0] * s["B"] + [0.0] * (shape["B"] - s["B"])
    assert b.lg_active_rows <= p.lg_active_rows <= shape["E"]
    with pytest.raises(ValueError):
        batching.pad_batch(b, {"N": s["N"] - 1, "E": s["E"], "L": s["L"], "B": s["B"] + 1})
    # two different ragged batches of similar size land in the same bucket
    assert batching.bucket_shape(ragged_batch(1, sizes=(9, 11, 10, 16, 9))) == shape


def test_masked_loss_ignores_dummy_graphs():
    g = torch.Generator().manual_seed(0)
    mean, logvar, tz = (torch.randn(5, 2, generator=g) for _ in range(3))
    want = pkg.gaussian_nll_loss(mean[:3], logvar[:3], tz[:3])
    mask = torch.tensor([1.0, 1.0, 1.0, 0.0, 0.0])
    got = pkg.gaussian_nll_loss(mean, logvar, tz, mask=mask)
    assert abs(float(got) - float(want)) < 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("lg_inc", ["pyg", "bonds"])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_padding_changes_no_real_result(dtype, lg_inc):
    torch.manual_seed(3)
    m = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.0, **CTOR), 2).cuda()
    m.base.compute_dtype = dtype
    m.train()
    host = ragged_batch(5, lg_inc)
    padded, mask = batching.pad_batch(host)
    nb = host.num_graphs
    tz = pkg.zscore_targets(host.y, nb)
    tzp = pkg.zscore_targets(padded.y, padded.num_graphs)
    out = {}
    for name, b, t, mk in (("plain", host, tz, None), ("padded", padded, tzp, mask)):
        b = b.to("cuda")
        m.zero_grad(set_to_none=True)
        mean, logvar = m(b)
        loss = pkg.gaussian_nll_loss(mean.float(), logvar.float(), t.cuda(), mask=None if mk is None else mk.cuda())
        loss.backward()
        out[name] = (mean[:nb], logvar[:nb], loss, {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
    a, b = out["padded"], out["plain"]
    tol, gtol = (1e-5, 1e-4) if dtype == torch.float32 else (2e-2, 5e-2)
    assert rel_err(a[0], b[0]) < tol and rel_err(a[1], b[1]) < tol and rel_err(a[2], b[2]) < tol
    gmax = max(float(g.abs().max()) for g in b[3].values())
    assert all(torch.isfinite(g).all() for g in a[3].values())
    for k, g in b[3].items():
        assert float((a[3][k] - g).abs().max()) < gtol * max(float(g.abs().max()), 1e-2 * gmax), k


@pytest.mark.gpu
def test_ragged_batches_replay_one_graph_per_bucket():
    torch.manual_seed(4)
    m = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=0.0, **CTOR), 2).cuda()
    m.base.compute_dtype = torch.bfloat16
    m.train()
    ts = engine.TrainStep(m, graph=True, graph_warmup=1, pad_to_buckets=True, lr=0.0, weight_decay=0.0)
    ref = engine.TrainStep(m, graph=False, lr=0.0, weight_decay=0.0, optimizer=False)
    variants = [(8, 12, 10, 16, 9), (9, 11, 10, 16, 9), (10, 10, 10, 15, 10), (8, 12, 11, 15, 9)]
    for i, sizes in enumerate(variants * 2):
        host = ragged_batch(20 + i, sizes=sizes)
        assert batching.bucket_shape(host) == batching.bucket_shape(ragged_batch(0))
        b = host.to("cuda")
        tz = pkg.zscore_targets(b.y, b.num_graphs)
        loss, mean, logvar = ts.step(b, tz)
        want, wm, _ = ref.step(b, tz)
        assert rel_err(loss, want) < 2e-2 and rel_err(mean[:b.num_graphs], wm) < 2e-2, i
    assert ts.eager_steps == 1 and ts.replays == 7 and len(ts._captured) == 1
