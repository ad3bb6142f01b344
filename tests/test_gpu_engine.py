"""GPU tests of the one-call training step (``engine.TrainStep``): the fused clip + AdamW kernel against
``torch.nn.utils.clip_grad_norm_`` + ``torch.optim.AdamW`` (what the reference runs after the hot path,
scripts/train.py:690-699, optimizer built at :1516-1540), CUDA-graph replay against the eager step, fresh dropout masks
across replays, and the in-kernel-feature line-graph family against the stored-feature family on the whole model."""
import copy

import pytest
import torch

from conftest import rel_err
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import engine, ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
CTOR = dict(node_dim=206, edge_dim=36, angle_dim=11, global_dim=289, target_dim=2, hidden=256, layers=2, heads=4)


def _model(dropout=0.0, seed=0, **over):
    torch.manual_seed(seed)
    m = pkg.HeteroAlignnRegressor(pkg.AlignnRegressor(dropout=dropout, **{**CTOR, **over}), 2).to(DEV)
    m.train()
    return m


def test_fused_clip_adamw_matches_torch_two_lr_groups():
    """fp32 regime, dropout 0: three steps of TrainStep == clip_grad_norm_(5.0) + torch AdamW with the reference's two
    parameter groups (base + mean heads | log-variance heads)."""
    a = _model()
    b = copy.deepcopy(a)
    batch = pkg.synthetic_batch(6, 8, 4, seed=3).to(DEV)
    tz = pkg.zscore_targets(batch.y, batch.num_graphs)
    ts = engine.TrainStep(a, lr=2e-3, lr_sigma=5e-4, weight_decay=1e-2, max_norm=0.05, graph=False)
    base = list(b.base.parameters()) + list(b.mean_heads.parameters())
    opt = torch.optim.AdamW([{"params": base, "lr": 2e-3}, {"params": list(b.logvar_heads.parameters()), "lr": 5e-4}],
                            weight_decay=1e-2)
    for _ in range(3):
        loss_a, _, _ = ts.step(batch, tz)
        opt.zero_grad(set_to_none=True)
        mean, logvar = b(batch)
        loss_b = pkg.gaussian_nll_loss(mean, logvar, tz)
        loss_b.backward()
        norm_b = torch.nn.utils.clip_grad_norm_(b.parameters(), max_norm=0.05)
        opt.step()
        assert rel_err(loss_a, loss_b) < 1e-5
        assert rel_err(ts.opt.norm, norm_b) < 1e-4 and float(norm_b) > 0.05      # the clip is active
    pa, pb = dict(a.named_parameters()), dict(b.named_parameters())
    # Adam's m / sqrt(v) turns last-bit differences of tiny gradients into O(lr) steps: the two sides round the loss
    # gradient differently (one fused kernel vs autograd's chain of fp32 ops), hence 5e-5 on the parameters after 3 steps
    for k in pb:
        assert rel_err(pa[k], pb[k]) < 5e-5, k
    assert float(ts.opt.step_count) == 3.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_graph_replay_matches_eager_step(dtype):
    """dropout 0: the replayed graph and the eager step are the same kernels on the same data -> identical results,
    including on a second batch copied into the graph's static buffers."""
    a = _model(seed=1)
    b = copy.deepcopy(a)
    a.base.compute_dtype = b.base.compute_dtype = dtype
    batches = [pkg.synthetic_batch(5, 8, 4, seed=s).to(DEV) for s in (0, 1)]
    tzs = [pkg.zscore_targets(x.y, x.num_graphs) for x in batches]
    tg = engine.TrainStep(a, graph=True, graph_warmup=1)
    te = engine.TrainStep(b, graph=False)
    for i in range(5):
        la, ma, va = tg.step(batches[i % 2], tzs[i % 2])
        lb, mb, vb = te.step(batches[i % 2], tzs[i % 2])
        assert torch.equal(la, lb) and torch.equal(ma, mb) and torch.equal(va, vb), i
    assert tg.replays == 4 and tg.eager_steps == 1
    for (k, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert torch.equal(pa, pb), k
    # a new signature falls back to eager steps until captured
    other = pkg.synthetic_batch(3, 8, 4, seed=5).to(DEV)
    tg.step(other, pkg.zscore_targets(other.y, 3))
    assert tg.eager_steps == 2


def test_graph_replays_draw_fresh_dropout_masks():
    a = _model(dropout=0.3, seed=2)
    a.base.compute_dtype = torch.bfloat16
    batch = pkg.synthetic_batch(8, 8, 4, seed=0).to(DEV)
    tz = pkg.zscore_targets(batch.y, batch.num_graphs)
    ts = engine.TrainStep(a, graph=True, graph_warmup=1, lr=0.0, weight_decay=0.0)     # frozen weights
    losses = [float(ts.step(batch, tz)[0]) for _ in range(6)]
    assert ts.replays == 5
    assert len({round(x, 6) for x in losses[1:]}) >= 4, losses      # same weights, same batch: only the masks differ
    assert all(torch.isfinite(torch.tensor(losses)))


def test_lg_family_matches_stored_feature_family_on_the_model(monkeypatch):
    """bf16, default arch: in-kernel angle features (csrc/lgattn.cu) vs stored h1 (csrc/edgeattn_mma.cu)."""
    m = _model(seed=4, layers=4)
    m.base.compute_dtype = torch.bfloat16
    batch = pkg.synthetic_batch(16, 16, 12, seed=2).to(DEV)
    tz = pkg.zscore_targets(batch.y, batch.num_graphs)
    out = {}
    m.base.elide_isolated = False      # the stored-feature family has no elision: compare like with like
    for use_lg in (True, False):
        monkeypatch.setattr(ops, "USE_LG", use_lg)
        m.zero_grad(set_to_none=True)
        k0 = ops.STATS.kernels
        mean, logvar = m(batch)
        loss = pkg.gaussian_nll_loss(mean.float(), logvar.float(), tz)
        loss.backward()
        out[use_lg] = (mean, logvar, {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None},
                       ops.STATS.kernels - k0)
    a, b = out[True], out[False]
    assert rel_err(a[0], b[0]) < 2e-2 and rel_err(a[1], b[1]) < 2e-2
    gmax = max(float(g.abs().max()) for g in b[2].values())
    assert set(a[2]) == set(b[2])
    # two valid bf16 evaluations: a one-ulp difference in a bf16 projection is amplified by LayerNorm on rows whose
    # pre-norm variance is ~eps (bond rows without neighbours), so small gradients carry a few % of rounding noise --
    # the accuracy claim itself is tests/test_gpu_model.py (fp64 oracle, reference-AMP-relative tolerance)
    for k, g in b[2].items():
        assert float((a[2][k] - g).abs().max()) < 5e-2 * max(float(g.abs().max()), 1e-2 * gmax), k


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("shape", [(98304, 2048), (8544, 1792), (1000, 1024), (37, 256), (5, 8), (12, 48), (333, 24)])
def test_colsum_matches_fp64(shape, dtype):
    g = torch.Generator().manual_seed(shape[0])
    big = torch.randn(shape[0], shape[1] + 64, generator=g).to(dtype).to(DEV)
    x = big[:, 32:32 + shape[1]] if (shape[1] % 8 == 0) else big[:, :shape[1]]      # strided rows, 16-byte aligned
    got = ops.colsum(x)
    want = x.double().sum(0)
    assert got.dtype == torch.float32 and rel_err(got, want) < 2e-6
    assert torch.equal(ops.colsum(x), got)                                            # deterministic


def test_gate_ln_bwd3_adds_the_folded_bias_gradient():
    n, hid, h = 1237, 256, 4
    g = torch.Generator().manual_seed(7)
    dy, agg = torch.randn(n, hid, generator=g).to(DEV), torch.randn(n, hid, generator=g).to(DEV)
    proj = torch.randn(n, 4 * hid, generator=g).to(torch.bfloat16).to(DEV)
    xr = proj[:, 3 * hid:]
    wb, gm, bl = (torch.randn(k, generator=g).to(DEV) * 0.1 for k in (3 * hid, hid, hid))
    beta, mean, rstd = (torch.rand(n, generator=g).to(DEV) for _ in range(3))
    s = torch.rand(n, h, generator=g).to(DEV)
    d2 = torch.zeros_like(proj); d3 = torch.zeros_like(proj)
    a2 = ops.raw_gate_ln_bwd2(dy, agg, xr, wb, gm, bl, beta, mean, rstd, d2[:, 3 * hid:], True, 0.1, 5, 9)
    a3 = ops.raw_gate_ln_bwd3(dy, agg, xr, wb, gm, bl, beta, mean, rstd, s, h, d3[:, 3 * hid:], 0.1, 5, 9)
    assert torch.equal(a2[0], a3[0]) and torch.equal(a2[1], a3[1]) and torch.equal(d2, d3)
    assert torch.equal(a2[2], a3[2][:5 * hid])
    want = (a3[0].double().view(n, h, hid // h) * s.double().unsqueeze(-1)).sum(0).reshape(hid)
    assert rel_err(a3[2][5 * hid:], want) < 1e-5


@pytest.mark.parametrize("lg_inc", ["pyg", "bonds"])
def test_fused_trunk_matches_per_block_autograd_path(lg_inc):
    """The explicit trunk program (trunk.py: folded gradient sums, stacked weight folds) against one autograd node per
    block (fused.py) on the same bf16 kernels: outputs and every parameter gradient."""
    m = _model(seed=6, layers=3)
    m.base.compute_dtype = torch.bfloat16
    batch = pkg.synthetic_batch(12, 16, 12, seed=4, lg_inc=lg_inc).to(DEV)
    tz = pkg.zscore_targets(batch.y, batch.num_graphs)
    out = {}
    m.base.elide_isolated = False      # elision is checked on its own (bit-identical forward) below
    for fused_trunk in (True, False):
        m.base.fused_trunk = fused_trunk
        m.zero_grad(set_to_none=True)
        mean, logvar = m(batch)
        loss = pkg.gaussian_nll_loss(mean.float(), logvar.float(), tz)
        loss.backward()
        out[fused_trunk] = (mean, logvar, {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
    a, b = out[True], out[False]
    assert rel_err(a[0], b[0]) < 1e-2 and rel_err(a[1], b[1]) < 1e-2
    assert set(a[2]) == set(b[2])
    gmax = max(float(g.abs().max()) for g in b[2].values())
    for k, g in b[2].items():      # bf16 rounding noise, see the comment in the test above
        assert float((a[2][k] - g).abs().max()) < 5e-2 * max(float(g.abs().max()), 1e-2 * gmax), k


def test_isolated_row_elision_changes_nothing():
    """PyG-collated batches leave most bond rows without line-graph neighbours.  Skipping the q/k/v/qt
    projections, the attention kernels and the weight-gradient GEMM rows of those rows must not change any result."""
    m = _model(seed=8, layers=2)
    m.base.compute_dtype = torch.bfloat16
    host = pkg.synthetic_batch(24, 16, 12, seed=7, lg_inc="pyg")
    n_bonds = host.edge_index.size(1)
    assert host.lg_active_rows == int(host.lg_edge_index.max()) + 1 and host.lg_active_rows < n_bonds // 4
    batch = host.to(DEV)
    assert batch.lg_active_rows == host.lg_active_rows
    tz = pkg.zscore_targets(batch.y, batch.num_graphs)
    out = {}
    for elide in (True, False):
        m.base.elide_isolated = elide
        m.zero_grad(set_to_none=True)
        mean, logvar = m(batch)
        pkg.gaussian_nll_loss(mean.float(), logvar.float(), tz).backward()
        out[elide] = (mean, logvar, {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
    a, b = out[True], out[False]
    # same GEMM shapes and kernels on every row that is computed: the forward is bit-identical, gradients differ only by
    # the fp32 summation order of the (shorter) weight-gradient reductions
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    gmax = max(float(g.abs().max()) for g in b[2].values())
    for k, g in b[2].items():
        assert float((a[2][k] - g).abs().max()) < 1e-4 * max(float(g.abs().max()), 1e-2 * gmax), k
    # a "bonds"-collated batch has no isolated prefix: the bound equals the number of bonds
    assert pkg.synthetic_batch(4, 16, 12, seed=1, lg_inc="bonds").lg_active_rows == 4 * 16 * 12


def test_train_step_sample_weights_match_reference_weighting():
    """Per-sample loss weights (the reference's KNN weighting, train.py:661-675: nll * w before the means) through the fused
    loss, eager and replayed: same loss as the torch composition with the weights applied."""
    m = _model(seed=4)
    batch = pkg.synthetic_batch(6, 8, 4, seed=5).to(DEV)
    tz = pkg.zscore_targets(batch.y, batch.num_graphs)
    w = torch.tensor([0.5, 2.0, 1.0, 1.5, 0.25, 3.0], device=DEV)
    with torch.no_grad():
        mean, logvar = copy.deepcopy(m)(batch)
    lv = torch.clamp(logvar.float(), min=-2.9)
    want = ((0.5 * (lv + (mean.float() - tz) ** 2 / torch.exp(lv))) * w[:, None]).mean(1).mean() + 0.1 * (0.5 * lv).pow(2).mean()
    ts = engine.TrainStep(m, lr=0.0, weight_decay=0.0, graph=True, graph_warmup=1)
    got = [float(ts.step(batch, tz, sample_weight=w)[0]) for _ in range(3)]          # eager, then captured + replayed
    assert ts.replays >= 1
    for g in got:
        assert abs(g - float(want)) <= 1e-5 * max(1.0, abs(float(want)))
    plain = float(engine.TrainStep(copy.deepcopy(m), lr=0.0, weight_decay=0.0, graph=False).step(batch, tz)[0])
    assert abs(plain - got[0]) > 1e-4                                               # the weights do change the loss


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
def test_two_gpu_train_step_equals_single_gpu(dtype, tmp_path):
    """N-GPU == 1-GPU on the real engine (captured graph + NCCL all-reduce + fused clip/AdamW): launches
    scripts/check_dp_equality.py on 2 GPUs of this box.  Skipped on a single-GPU box (kept as a script + committed output
    under profiles/ for that case)."""
    import json
    import os
    import socket
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(root, "scripts", "check_dp_equality.py"), "--dtype", dtype]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert line["ok"] and line["graph_replays"] >= 2


def test_generic_width_model_with_dropout_is_not_frozen_by_graph_capture():
    """hidden != 256 runs on kernels whose dropout keys are host integers; capturing them would replay
    one mask for ever.  TrainStep must notice, fall back to eager steps, and keep drawing fresh masks."""
    a = _model(dropout=0.3, seed=2, hidden=64)
    batch = pkg.synthetic_batch(8, 8, 4, seed=0).to(DEV)
    tz = pkg.zscore_targets(batch.y, batch.num_graphs)
    ts = engine.TrainStep(a, graph=True, graph_warmup=1, lr=0.0, weight_decay=0.0)     # frozen weights
    with pytest.warns(UserWarning, match="CUDA-graph replay disabled"):
        losses = [float(ts.step(batch, tz)[0]) for _ in range(6)]
    assert ts.replays == 0 and ts.eager_steps == 6 and ts.use_graph is False
    assert len({round(x, 6) for x in losses}) >= 5, losses


def test_fp32_inference_elides_isolated_rows_without_changing_results():
    """predict.py / evaluate.py regime (fp32, no autocast, no autograd): with a PyG-collated batch the line-graph blocks run
    q|k|v|qt, the attention kernel and the abar products on the active prefix of the bond rows only.  Same outputs as with
    the elision switched off (1e-6: the library GEMMs may tile a shorter M differently) and the elided path refuses
    autograd."""
    m = _model(seed=9, layers=2)
    m.base.compute_dtype = torch.float32
    m.eval()
    batch = pkg.synthetic_batch(24, 16, 12, seed=11, lg_inc="pyg").to(DEV)
    assert batch.lg_active_rows < batch.edge_index.size(1) // 4
    out = {}
    with torch.no_grad():
        for elide in (True, False):
            m.base.elide_isolated = elide
            k0 = ops.STATS.kernels
            out[elide] = m(batch)
            assert ops.STATS.kernels > k0
    for a, b in zip(out[True], out[False]):
        assert float((a - b).abs().max()) <= 1e-6 * max(float(b.abs().max()), 1.0)
    # with autograd on, the per-block path keeps every row (its backward needs full-size saved tensors)
    m.base.elide_isolated = True
    m.train()
    mean, logvar = m(batch)
    (mean.sum() + logvar.sum()).backward()
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)
