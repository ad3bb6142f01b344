"""GPU parity tests of the individual C-ABI operators against the oracle (run with -m gpu on a B200).

Tolerances (BASELINE.json north_star): graph plan bit-exact; fp32 forward rel 1e-5, gradients rel 1e-4;
bf16 rel 2e-2.  `rel` = max|a-b| / max|b| (conftest.rel_err)."""
import math

import pytest
import torch

import oracle
from conftest import rel_err

oracle.install_shim()
from torch_geometric.utils import scatter, softmax  # noqa: E402  (oracle leaf ops)

import gnn_elasticity_predictor_b200 as pkg  # noqa: E402
from gnn_elasticity_predictor_b200 import ops  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda"


def ref_conv_core(q, k, v, e, index, heads):
    """oracle: PyG message + softmax + add-aggregation (fp64 on CPU)"""
    n, hidden = q.shape
    c = hidden // heads
    src, dst = index[0], index[1]
    qi = q.view(n, heads, c)[dst]
    kj = k.view(n, heads, c)[src] + e.view(-1, heads, c)
    alpha = softmax((qi * kj).sum(-1) / math.sqrt(c), dst, None, n)
    msg = (v.view(n, heads, c)[src] + e.view(-1, heads, c)) * alpha.unsqueeze(-1)
    return scatter(msg, dst, 0, dim_size=n, reduce="sum").view(n, hidden)


def random_multigraph(n, e, seed, empty_tail=True):
    g = torch.Generator().manual_seed(seed)
    hi = max(1, (3 * n) // 4) if empty_tail else n
    dst = torch.randint(0, hi, (e,), generator=g)
    src = torch.randint(0, n, (e,), generator=g)
    if e >= 6:
        src[:3] = dst[:3]
        src[3:6], dst[3:6] = src[0:3].clone(), dst[0:3].clone()
    return torch.stack([src, dst])


# ---- graph plan: bit-exact ----------------------------------------------------------------------------
@pytest.mark.parametrize("n,e,seed", [(1, 0, 0), (5, 1, 1), (7, 50, 2), (300, 5000, 3), (98304, 1081344, 4),
                                      (2_000_000, 300_000, 5), (70000, 4097, 6)])
def test_plan_matches_stable_sort(n, e, seed):
    index = random_multigraph(n, e, seed)
    plan = pkg.build_plan(index.to(DEV), n, validate=True)
    for key_row, other_row, rowptr, col, eid in ((1, 0, plan.rowptr, plan.col, plan.eid),
                                                 (0, 1, plan.rowptr_t, plan.col_t, plan.eid_t)):
        _, perm = torch.sort(index[key_row], stable=True)
        counts = torch.bincount(index[key_row], minlength=n)
        want_ptr = torch.cat([torch.zeros(1, dtype=torch.long), counts.cumsum(0)])
        assert torch.equal(rowptr.cpu().long(), want_ptr)
        assert torch.equal(eid.cpu().long(), perm)
        assert torch.equal(col.cpu().long(), index[other_row][perm])
    assert plan.rowptr.dtype == torch.int32 and plan.eid.dtype == torch.int32


def test_plan_source_sorted_input_gives_identity_csc():
    b = pkg.synthetic_batch(4, 16, 12, seed=0, lg_inc="bonds")
    plan = pkg.build_plan(b.lg_edge_index.to(DEV), b.edge_index.size(1))
    assert torch.equal(plan.eid_t.cpu().long(), torch.arange(b.lg_edge_index.size(1)))


@pytest.mark.parametrize("lg_inc", ["pyg", "bonds"])
def test_plan_source_sorted_hint_gives_the_same_plan(lg_inc):
    """GraphBatch.source_sorted (host fact, fetch.py emits source-major lists) skips the CSC radix passes: identical plan."""
    from gnn_elasticity_predictor_b200 import batching
    b = pkg.synthetic_batch(5, 16, 12, seed=3, lg_inc=lg_inc, dups=True)
    # under PyG's default collate the line-graph ids of consecutive crystals overlap: not source-sorted
    assert b.source_sorted == (True, lg_inc == "bonds")
    padded, _ = batching.pad_batch(b, align=64)
    assert padded.source_sorted == b.source_sorted
    for batch in (b, padded):
        for index, n, hint in ((batch.edge_index, batch.x.size(0), batch.source_sorted[0]),
                               (batch.lg_edge_index, batch.edge_index.size(1), batch.source_sorted[1])):
            if not hint:
                assert int(pkg.build_plan(index.to(DEV), n, source_sorted=True).status.item()) & 2
                continue
            full = pkg.build_plan(index.to(DEV), n)
            fast = pkg.build_plan(index.to(DEV), n, source_sorted=True)
            assert int(fast.status.item()) & 2 == 0
            for name in ("rowptr", "col", "eid", "rowptr_t", "col_t", "eid_t"):
                assert torch.equal(getattr(full, name), getattr(fast, name)), name


@pytest.mark.parametrize("n_graphs,atoms,k", [(6, 16, 12), (256, 32, 12)])
def test_plan_key_bound_gives_the_same_plan_with_fewer_passes(n_graphs, atoms, k):
    """GraphBatch.lg_active_rows bounds every line-graph index under PyG's collate: the bounded plan is the
    same plan (rows at / beyond the bound are empty either way) built with fewer radix passes."""
    from gnn_elasticity_predictor_b200 import batching
    b = pkg.synthetic_batch(n_graphs, atoms, k, seed=2, lg_inc="pyg")
    n = b.edge_index.size(1)
    assert 0 < b.lg_active_rows < n
    for batch in (b, batching.pad_batch(b, align=64)[0]):
        n = batch.edge_index.size(1)
        idx = batch.lg_edge_index.to(DEV)
        full = pkg.build_plan(idx, n)
        fast = pkg.build_plan(idx, n, key_bound=batch.lg_active_rows)
        assert int(fast.status.item()) & 4 == 0
        for name in ("rowptr", "col", "eid", "rowptr_t", "col_t", "eid_t"):
            assert torch.equal(getattr(full, name), getattr(fast, name)), name
        assert ops._plan_kernels(idx.size(1), n, False, batch.lg_active_rows) <= ops._plan_kernels(idx.size(1), n)
    wrong = pkg.build_plan(b.lg_edge_index.to(DEV), n, key_bound=b.lg_active_rows - 1)      # one index is at the bound
    assert int(wrong.status.item()) & 4
    with pytest.raises(ValueError):
        wrong.check()


def test_plan_wrong_source_sorted_hint_is_flagged():
    index = torch.tensor([[0, 2, 1], [1, 0, 2]])
    plan = pkg.build_plan(index.to(DEV), 3, source_sorted=True)
    assert int(plan.status.item()) == 2
    with pytest.raises(ValueError):
        plan.check()
    assert pkg.GraphBatch(num_graphs=1, edge_index=index, lg_edge_index=index[:, :1]).source_sorted == (False, True)


def test_plan_flags_out_of_range_and_drops_those_edges():
    index = torch.tensor([[0, 1, 9, 2, -1], [1, 2, 0, 7, 0]])
    plan = pkg.build_plan(index.to(DEV), 3)
    assert int(plan.status.item()) == 1
    with pytest.raises(IndexError):
        plan.check()
    assert plan.rowptr.cpu().tolist() == [0, 0, 1, 2]          # only edges 0 and 1 are in range
    assert plan.eid.cpu().tolist()[:2] == [0, 1]
    # dropped edges stay in the permutation (after the live ones, in input order) with a safe endpoint
    assert plan.eid.cpu().tolist() == [0, 1, 2, 3, 4] and plan.col.cpu().tolist()[2:] == [0, 0, 0]
    assert sorted(plan.eid_t.cpu().tolist()) == [0, 1, 2, 3, 4]


# ---- conv core ------------------------------------------------------------------------------------------
SHAPES = [(256, 4), (256, 1), (256, 8), (128, 4), (64, 4), (32, 1), (32, 4), (16, 2), (8, 1),
          (48, 3), (96, 4), (40, 5), (320, 4), (512, 4)]


@pytest.mark.parametrize("hidden,heads", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conv_core_forward_backward(hidden, heads, dtype):
    n, e = 37, 400
    index = random_multigraph(n, e, hidden + heads)
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(n, hidden, generator=g) for _ in range(3))
    ee = torch.randn(e, hidden, generator=g)
    gout = torch.randn(n, hidden, generator=g)
    # round operands to the storage dtype first so that both sides see identical inputs
    q, k, v, ee = (t.to(dtype).float() for t in (q, k, v, ee))
    ref_in = [t.double().requires_grad_(True) for t in (q, k, v, ee)]
    ref = ref_conv_core(*ref_in, index, heads)
    ref.backward(gout.double())

    plan = pkg.build_plan(index.to(DEV), n)
    dev_in = [t.to(DEV, dtype).requires_grad_(True) for t in (q, k, v, ee)]
    out = pkg.conv_core(*dev_in, plan, heads)
    assert out.dtype == torch.float32
    out.backward(gout.to(DEV))
    fwd_tol, bwd_tol = (1e-5, 1e-4) if dtype == torch.float32 else (2e-2, 2e-2)
    if dtype == torch.float32:
        assert rel_err(out.cpu(), ref) < fwd_tol
    else:
        assert rel_err(out.cpu(), ref) < 1e-5          # operands pre-rounded: forward accumulates in fp32
    for got, want, name in zip(dev_in, ref_in, "qkve"):
        assert rel_err(got.grad.float().cpu(), want.grad) < bwd_tol, name
    # rows without in-edges aggregate to exactly zero
    empty = torch.bincount(index[1], minlength=n) == 0
    assert bool(empty.any()) and float(out.cpu()[empty].abs().max()) == 0.0


def test_conv_core_long_rows_and_degree_skew():
    """one hub row with thousands of in-edges next to empty rows (the 'pyg' batching quirk regime)"""
    n, hidden, heads = 64, 256, 4
    g = torch.Generator().manual_seed(3)
    dst = torch.cat([torch.zeros(5000, dtype=torch.long), torch.randint(1, 8, (3000,), generator=g)])
    src = torch.randint(0, n, (8000,), generator=g)
    index = torch.stack([src, dst])
    q, k, v = (torch.randn(n, hidden, generator=g) for _ in range(3))
    ee = torch.randn(8000, hidden, generator=g)
    ref = ref_conv_core(q.double(), k.double(), v.double(), ee.double(), index, heads)
    plan = pkg.build_plan(index.to(DEV), n)
    out = pkg.conv_core(q.to(DEV), k.to(DEV), v.to(DEV), ee.to(DEV), plan, heads)
    assert rel_err(out.cpu(), ref) < 1e-5


def test_conv_core_properties_at_config2_size():
    """size-independent properties at BASELINE config 2 line-graph size (N=98304, L=1081344, H=256)."""
    b = pkg.synthetic_batch(256, 32, 12, seed=0, lg_inc="pyg")
    n, e, hidden, heads = b.edge_index.size(1), b.lg_edge_index.size(1), 256, 4
    assert (n, e) == (98304, 1081344)
    index = b.lg_edge_index.to(DEV)
    plan = pkg.build_plan(index, n, validate=True)
    gen = torch.Generator(device=DEV).manual_seed(0)
    q, k = (torch.randn(n, hidden, device=DEV, generator=gen).bfloat16() for _ in range(2))
    ones = torch.ones(n, hidden, device=DEV, dtype=torch.bfloat16)
    zeros_e = torch.zeros(e, hidden, device=DEV, dtype=torch.bfloat16)
    # (1) attention weights of every non-empty row sum to one: v = 1, e = 0 -> agg in {0, 1}
    agg = pkg.conv_core(q, k, ones, zeros_e, plan, heads)
    deg = torch.bincount(index[1], minlength=n)
    assert float((agg[deg > 0] - 1).abs().max()) < 1e-5 and float(agg[deg == 0].abs().max()) == 0.0
    # (2) deterministic: bitwise identical on a second run
    v = torch.randn(n, hidden, device=DEV, generator=gen).bfloat16()
    ee = torch.randn(e, hidden, device=DEV, generator=gen).bfloat16()
    a1 = pkg.conv_core(q, k, v, ee, plan, heads)
    a2 = pkg.conv_core(q, k, v, ee, plan, heads)
    assert torch.equal(a1, a2)
    # (3) invariance to a permutation of the edge list (up to fp32 summation order)
    perm = torch.randperm(e, device=DEV, generator=gen)
    plan_p = pkg.build_plan(index[:, perm].contiguous(), n)
    a3 = pkg.conv_core(q, k, v, ee[perm].contiguous(), plan_p, heads)
    assert rel_err(a3, a1) < 1e-5
    # (4) linear in (v, e) jointly for fixed logits is not separable (e enters the logits); but v alone is linear
    a4 = pkg.conv_core(q, k, (2 * v.float()).bfloat16(), ee, plan, heads)
    a5 = pkg.conv_core(q, k, torch.zeros_like(v), ee, plan, heads)
    assert rel_err(a4 - a5, 2 * (a1 - a5)) < 1e-4


def test_conv_core_dropout_statistics_and_backward_mask_consistency():
    n, hidden, heads, e = 200, 64, 4, 20000
    index = random_multigraph(n, e, 9, empty_tail=False)
    plan = pkg.build_plan(index.to(DEV), n)
    g = torch.Generator().manual_seed(2)
    q = torch.zeros(n, hidden)                        # uniform attention: alpha = 1/deg
    k = torch.randn(n, hidden, generator=g)
    v = torch.ones(n, hidden)
    ee = torch.zeros(e, hidden)
    p = 0.25
    args = [t.to(DEV).requires_grad_(True) for t in (q, k, v, ee)]
    out = pkg.conv_core(*args, plan, heads, p_drop=p, seed=123, offset=7)
    # E[agg] = 1 ; per head the kept fraction ~ 1-p
    assert abs(float(out.mean()) - 1.0) < 0.02
    out2 = pkg.conv_core(*args, plan, heads, p_drop=p, seed=123, offset=7)
    assert torch.equal(out, out2)                     # same key -> same mask
    out3 = pkg.conv_core(*args, plan, heads, p_drop=p, seed=124, offset=7)
    assert not torch.equal(out, out3)
    # backward regenerates the same mask: d(sum agg)/dv_j = sum_i a~_ij, and sum_j of that = sum(agg) when v = 1
    out.sum().backward()
    assert abs(float(args[2].grad.sum()) - float(out.sum())) / float(out.sum()) < 1e-4


# ---- epilogue -----------------------------------------------------------------------------------------
@pytest.mark.parametrize("hidden", [256, 128, 64, 32, 8, 48, 96, 320])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gate_ln_forward_backward(hidden, dtype):
    n = 333
    g = torch.Generator().manual_seed(hidden)
    agg = torch.randn(n, hidden, generator=g)
    xr = torch.randn(n, hidden, generator=g).to(dtype).float()
    x = torch.randn(n, hidden, generator=g)
    wb = torch.randn(1, 3 * hidden, generator=g) * 0.2
    gamma = torch.rand(hidden, generator=g) + 0.5
    bias = torch.randn(hidden, generator=g) * 0.3
    gout = torch.randn(n, hidden, generator=g)

    ref_in = [t.double().requires_grad_(True) for t in (agg, xr, x, wb, gamma, bias)]
    a, s, xx, w, gm, bs = ref_in
    beta = torch.sigmoid(torch.cat([a, s, a - s], -1) @ w.t())
    o = beta * s + (1 - beta) * a
    ref = xx + torch.relu(torch.nn.functional.layer_norm(o, (hidden,), gm, bs, 1e-5))
    ref.backward(gout.double())

    dev_in = [agg.to(DEV), xr.to(DEV, dtype), x.to(DEV), wb.to(DEV), gamma.to(DEV), bias.to(DEV)]
    dev_in = [t.requires_grad_(True) for t in dev_in]
    y, y_lp = pkg.gate_ln_relu_residual(*dev_in, 1e-5, want_lp=True)
    assert y_lp.dtype == dtype and rel_err(y_lp.float().cpu(), ref) < (1e-5 if dtype == torch.float32 else 1e-2)
    y.backward(gout.to(DEV))
    assert rel_err(y.cpu(), ref) < 1e-5
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    for got, want, name in zip(dev_in, ref_in, ("agg", "xr", "x", "wbeta", "gamma", "bias")):
        assert rel_err(got.grad.float().cpu(), want.grad) < tol, name


@pytest.mark.parametrize("hidden,dtype", [(256, torch.bfloat16), (256, torch.float32), (64, torch.bfloat16), (8, torch.float32)])
def test_gate_ln_forward_many_rows_active_prefix(hidden, dtype):
    """Enough rows that every warp of the persistent forward wraps its cp.async ring several times; only a prefix of the
    rows has an aggregate (line-graph elision), the aggregate is assembled in-kernel, the residual enters in either dtype."""
    n, na, heads = 30011, 2600, 4 if hidden >= 32 else 1
    c = hidden // heads
    g = torch.Generator().manual_seed(hidden + 1)
    r = lambda *s: torch.randn(*s, generator=g)                                       # noqa: E731
    aggv, agge, cvec, stat = r(na, hidden), r(heads, na, c).to(dtype), r(hidden), r(na, heads).abs()
    xr, x32, xlp = r(n, hidden).to(dtype), r(n, hidden), r(n, hidden).to(dtype)
    wb, gm, bl = r(3 * hidden) * 0.2, torch.rand(hidden, generator=g) + 0.5, r(hidden) * 0.3
    d = lambda t: t.to(DEV)                                                           # noqa: E731
    for x_in, x_lp in ((x32, None), (None, xlp)):
        y, y_lp, agg, beta, mean, rstd = ops.raw_gate_ln_fwd2(d(aggv), d(agge), d(cvec), d(stat), heads, d(xr), None if x_in is None
                                                              else d(x_in), d(wb), d(gm), d(bl), 1e-5, 0.0, 1, 0, True, None,
                                                              agg_rows=na, x_lp=None if x_lp is None else d(x_lp))
        a = torch.zeros(n, hidden, dtype=torch.float64)
        a[:na] = (aggv.double() + agge.double().permute(1, 0, 2).reshape(na, hidden)
                  + cvec.double() * stat.double().repeat_interleave(c, dim=1))
        s = xr.double()
        w1, w2, w3 = wb.double().split(hidden)
        bt = torch.sigmoid(a @ w1 + s @ w2 + (a - s) @ w3)
        o = bt[:, None] * s + (1 - bt[:, None]) * a
        res = (x32 if x_in is not None else xlp).double()
        ref = res + torch.relu(torch.nn.functional.layer_norm(o, (hidden,), gm.double(), bl.double(), 1e-5))
        assert rel_err(y.cpu(), ref) < 1e-5
        assert rel_err(y_lp.float().cpu(), ref) < (1e-5 if dtype == torch.float32 else 1e-2)
        assert rel_err(agg[:na].cpu(), a[:na]) < 1e-6 and rel_err(beta.cpu(), bt) < 1e-5
        assert rel_err(mean.cpu(), o.mean(1)) < 1e-5 and rel_err(rstd.cpu(), 1 / torch.sqrt(o.var(1, unbiased=False) + 1e-5)) < 1e-5


def test_gate_ln_dropout_mask_shared_by_forward_and_backward():
    n, hidden = 500, 256
    g = torch.Generator().manual_seed(0)
    agg, xr = torch.randn(n, hidden, generator=g), torch.randn(n, hidden, generator=g)
    x = torch.zeros(n, hidden)
    wb = torch.zeros(1, 3 * hidden)
    gamma, bias = torch.ones(hidden), torch.full((hidden,), 5.0)      # relu always active
    args = [t.to(DEV).requires_grad_(True) for t in (agg, xr, x, wb, gamma, bias)]
    p = 0.4
    y, _ = pkg.gate_ln_relu_residual(*args, 1e-5, p_drop=p, seed=5, offset=9)
    kept = (y != 0)
    assert abs(float(kept.float().mean()) - (1 - p)) < 0.01
    y.sum().backward()
    # d/dbias of sum(y) = sum over rows of keep/(1-p): zero exactly where the forward dropped
    expect = kept.float().sum(0) / (1 - p)
    assert rel_err(args[5].grad, expect) < 1e-5


# ---- pooling ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("hidden", [256, 32, 48])
def test_segment_mean(hidden):
    counts = torch.tensor([3, 0, 5, 1, 7])
    batch = torch.repeat_interleave(torch.arange(5), counts)
    x = torch.randn(int(counts.sum()), hidden)
    ref_x = x.double().requires_grad_(True)
    ref = scatter(ref_x, batch, 0, dim_size=5, reduce="mean")
    gout = torch.randn(5, hidden)
    ref.backward(gout.double())
    xd = x.to(DEV).requires_grad_(True)
    plan = pkg.build_pool_plan(batch.to(DEV), 5)
    out = pkg.segment_mean(xd, plan)
    out.backward(gout.to(DEV))
    assert rel_err(out.cpu(), ref) < 1e-6 and float(out[1].abs().max()) == 0.0
    assert rel_err(xd.grad.cpu(), ref_x.grad) < 1e-6
    # unsorted batch vector
    perm = torch.randperm(batch.numel())
    plan2 = pkg.build_pool_plan(batch[perm].to(DEV), 5)
    assert rel_err(pkg.segment_mean(x[perm].to(DEV), plan2).cpu(), ref) < 1e-6


# ---- fused training loss (csrc/loss.cu) ---------------------------------------------------------------------------------
@pytest.mark.parametrize("b,t,l2,use_mask,use_w", [(256, 2, 0.1, False, False), (37, 2, 0.0, True, False),
                                                    (1000, 3, 0.1, True, True), (1, 2, 0.1, False, True)])
def test_fused_gaussian_nll_value_and_gradients(b, t, l2, use_mask, use_w):
    """alignn_gaussian_nll against the reference's loss arithmetic (train.py:655-681, incl. sample weights :661-675) in fp64."""
    g = torch.Generator().manual_seed(b)
    mean, tz = torch.randn(b, t, generator=g), torch.randn(b, t, generator=g)
    logvar = torch.randn(b, t, generator=g) * 2.0 - 2.0                      # a good share below the -2.9 floor
    mask = (torch.rand(b, generator=g) < 0.7).float() if use_mask else None
    w = torch.rand(b, generator=g) + 0.5 if use_w else None
    md, ld = mean.double().requires_grad_(True), logvar.double().requires_grad_(True)
    lv = torch.clamp(ld, min=-2.9)
    nll = 0.5 * (lv + (md - tz.double()).pow(2) / torch.exp(lv))
    if w is not None:
        nll = nll * w.double().view(-1, 1)
    if mask is None:
        want = nll.mean(dim=1).mean() + l2 * (0.5 * lv).pow(2).mean()
    else:
        mk = mask.double().view(-1, 1)
        n_real = mk.sum().clamp(min=1.0)
        want = (nll.mean(dim=1, keepdim=True) * mk).sum() / n_real + l2 * ((0.5 * lv).pow(2) * mk).sum() / (n_real * t)
    want.backward()
    mg, lg = mean.to(DEV).requires_grad_(True), logvar.to(DEV).requires_grad_(True)
    got = pkg.fused_gaussian_nll(mg, lg, tz.to(DEV), l2, -2.9, mask=None if mask is None else mask.to(DEV),
                                 sample_weight=None if w is None else w.to(DEV))
    (got * 3.0).backward()
    assert abs(float(got) - float(want)) <= 2e-6 * max(1.0, abs(float(want)))
    assert rel_err(mg.grad.cpu() / 3.0, md.grad) < 1e-5 and rel_err(lg.grad.cpu() / 3.0, ld.grad) < 1e-5
    assert bool((lg.grad.cpu()[logvar < -2.9] == 0).all())                 # clamp passes no gradient below the floor
    if not use_w:                                                            # same numbers as the torch composition in the package
        ref = pkg.gaussian_nll_loss(mean.to(DEV), logvar.to(DEV), tz.to(DEV), l2, -2.9, mask=None if mask is None else mask.to(DEV))
        assert abs(float(got) - float(ref)) <= 2e-6 * max(1.0, abs(float(ref)))
    with pytest.raises(RuntimeError, match="no CPU"):
        pkg.fused_gaussian_nll(mean, logvar, tz)


# ---- tcgen05 weight + bias gradient (csrc/wgrad_tc.cu) ------------------------------------------------------------------
@pytest.mark.parametrize("k,m,strided", [(98304, 256, True), (8544, 1792, False), (8192, 256, False), (1000, 1792, False),
                                         (63, 256, False), (1, 8, False), (12345, 136, True)])
def test_wgrad_tcgen05_matches_fp64(k, m, strided, monkeypatch):
    """alignn_wgrad (UMMA with both operands MN-major, TMEM accumulators, split over K) against the fp64 product and
    column sums of the same bf16 operands; also bit-identical across two runs (fixed-order reduction)."""
    g = torch.Generator().manual_seed(k + m)
    a_full = (torch.randn(k, 2 * m if strided else m, generator=g) * 0.5).to(torch.bfloat16).to(DEV)
    a = a_full[:, :m]                                         # row stride 2m when strided (the trunk's [dx_r | df] buffer)
    b = torch.randn(k, 256, generator=g).to(torch.bfloat16).to(DEV)
    monkeypatch.setattr(ops, "WGRAD_TC", True)
    monkeypatch.setattr(ops, "WGRAD_TC_MAX_M", -1)            # exercise the kernel for every M (the default gates it to M <= 256)
    w1, s1 = torch.empty(m, 256, device=DEV), torch.empty(m, device=DEV)
    ops.STATS.reset(); ops.STATS.events = True
    ops.wgrad(a, b, w1, s1)
    torch.cuda.synchronize()
    assert "wgrad" in ops.STATS.durations_ms(), "the tcgen05 kernel must have been taken for this shape"
    ops.STATS.events = False
    w2, s2 = torch.empty_like(w1), torch.empty_like(s1)
    ops.wgrad(a, b, w2, s2)
    assert torch.equal(w1, w2) and torch.equal(s1, s2)
    want_w = a.double().t() @ b.double()
    want_s = a.double().sum(0)
    assert rel_err(w1, want_w) < 2e-6 and rel_err(s1, want_s) < 2e-6
    monkeypatch.setattr(ops, "WGRAD_TC", False)               # the library path computes the same thing
    w3, s3 = torch.empty_like(w1), torch.empty_like(s1)
    ops.wgrad(a, b, w3, s3)
    assert rel_err(w3, want_w) < 5e-5 and rel_err(s3, want_s) < 2e-5          # cuBLAS split-K: looser than the UMMA kernel


@pytest.mark.parametrize("m,n,ld,with_bias", [(98304, 256, 256, True), (8544, 1792, 256, True), (1, 256, 256, False),
                                               (129, 256, 2048, True), (1000, 512, 512, False)])
def test_proj_tc_matches_fp32_reference(m, n, ld, with_bias):
    """csrc/proj_tc.cu (tcgen05 + TMA node projection, ``nn.Linear`` of TransformerConv: reference train.py:308, 326) against
    a plain PyTorch fp32 GEMM of the same bf16 operands: fp32 accumulation, one rounding to bf16 at the end -> within one
    bf16 ulp (2^-8 relative) of the fp32 result; rows past the last full 128-row tile, strided A rows, no bias."""
    from gnn_elasticity_predictor_b200 import _lib
    lib = _lib.load()
    assert lib.alignn_proj_tc_supported(256, n, ops.BF16_CODE)
    g = torch.Generator(device=DEV).manual_seed(m + n)
    big = (torch.randn(m, ld, device=DEV, generator=g) * 0.7).to(torch.bfloat16)
    x = big[:, ld - 256:] if ld > 256 else big                   # strided rows when ld > 256
    w = (torch.randn(n, 256, device=DEV, generator=g) * 0.06).to(torch.bfloat16)
    b = (torch.randn(n, device=DEV, generator=g) * 0.5).to(torch.bfloat16) if with_bias else None
    k0 = ops.STATS.kernels
    out = ops.linear_lp(x, w, b)
    assert ops.STATS.kernels == k0 + 1, "the tcgen05 kernel must be the one that ran"
    want = x.float() @ w.float().t()
    if b is not None:
        want = want + b.float()
    err = (out.float() - want).abs()
    bound = want.abs() * 2.0 ** -8 + 1e-6
    assert bool((err <= bound).all()), float((err / bound).max())
    # and the same numbers as the library GEMM it replaces, up to the last bf16 bit of rare ties
    lib_out = torch.addmm(b, x, w.t()) if b is not None else torch.mm(x, w.t())
    assert float((out.float() - lib_out.float()).abs().max()) <= float(want.abs().max()) * 2.0 ** -7


@pytest.mark.parametrize("n,na", [(98304, 8544), (8192, 8192), (300, 0), (1000, 129)])
def test_block_projections_one_launch_matches_two_library_gemms(n, na):
    """``ops.block_projections``: x_r over all rows + q|k|v|qt over the active prefix from one tcgen05 launch
    (``alignn_proj_tc2``) against the two ``torch.addmm`` calls it replaces and an fp32 reference."""
    g = torch.Generator(device=DEV).manual_seed(n + na)
    x = (torch.randn(n, 256, device=DEV, generator=g) * 0.7).to(torch.bfloat16)
    w8 = (torch.randn(2048, 256, device=DEV, generator=g) * 0.06).to(torch.bfloat16)
    b8 = (torch.randn(2048, device=DEV, generator=g) * 0.5).to(torch.bfloat16)
    k0 = ops.STATS.kernels
    xr, proj = ops.block_projections(x, w8, b8, na)
    assert ops.STATS.kernels == k0 + 1
    assert xr.shape == (n, 256) and proj.shape == (na, 1792)
    for got, rows, w, b in ((xr, x, w8[1792:], b8[1792:]), (proj, x[:na], w8[:1792], b8[:1792])):
        if rows.size(0) == 0:
            continue
        want = rows.float() @ w.float().t() + b.float()
        err = (got.float() - want).abs()
        bound = want.abs() * 2.0 ** -8 + 1e-6
        assert bool((err <= bound).all()), float((err / bound).max())
