"""GPU parity of the streaming edge-attention kernels (CUDA-core and tensor-core families) through the C ABI against an
fp64 torch restatement of the same arithmetic (PyG TransformerConv.message + softmax + 'add' aggregation at
reference scripts/train.py:315,334 with the edge projection folded into per-node operands, see csrc/edgeattn.cu).

The backward kernels are checked as the exact adjoint of the forward: with upstream gradients (dagg, gt, Gc) the scalar
    L = sum_i <dagg_i, aggv_i> + sum_{i,t} <gt_{i,t}, abar_{i,t}> + sum_{i,t} Gc_{i,t} S_{i,t}
has dL/dq = dq, dL/dqt = bbar, dL/dk = dk, dL/dv = dv, dL/df = df (autograd of the fp64 restatement)."""
import math

import pytest
import torch

from conftest import rel_err
import gnn_elasticity_predictor_b200 as pkg
from gnn_elasticity_predictor_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
H, HEADS, C = 256, 4, 64


def skewed_graph(n, e, seed, hub):
    """duplicates, self loops, rows without in-edges, one hub row with `hub` in-edges, degrees around 16"""
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, n, (e,), generator=g)
    dst = torch.randint(0, max(n - 5, 1), (e,), generator=g)     # last rows: no in-edges
    dst[:hub] = 3
    dst[hub:hub + 16] = 5                                          # exactly one full chunk
    dst[hub + 16:hub + 33] = 6                                     # 17 edges: chunk + 1
    src[-4:] = dst[-4:]
    return torch.stack([src, dst])


def reference(q, k, v, qt, f, index, dagg, gt, cvec):
    """fp64 restatement; returns outputs and the adjoint gradients."""
    q, k, v, qt, f = (t.detach().double().cpu().requires_grad_(True) for t in (q, k, v, qt, f))
    dagg, gt = dagg.double().cpu(), gt.double().cpu()
    n = q.size(0)
    j, i = index[0], index[1]
    s = ((q[i].view(-1, HEADS, C) * k[j].view(-1, HEADS, C)).sum(-1)
         + torch.einsum("tec,ec->et", qt[:, i, :], f)) / math.sqrt(C)                       # [E, h]
    m = torch.full((n, HEADS), -float("inf"), dtype=torch.float64).scatter_reduce(0, i[:, None].expand(-1, HEADS), s,
                                                                                  "amax", include_self=True)
    ex = torch.exp(s - m[i])
    z = torch.zeros(n, HEADS, dtype=torch.float64).index_add_(0, i, ex)
    a = ex / (z[i] + 1e-16)
    aggv = torch.zeros(n, HEADS, C, dtype=torch.float64).index_add_(0, i, a[:, :, None] * v[j].view(-1, HEADS, C))
    abar = torch.stack([torch.zeros(n, H, dtype=torch.float64).index_add_(0, i, a[:, t:t + 1] * f) for t in range(HEADS)])
    S = torch.zeros(n, HEADS, dtype=torch.float64).index_add_(0, i, a)
    out = dict(aggv=aggv.reshape(n, H), abar=abar, S=S)
    gc = (dagg.view(n, HEADS, C) * cvec.double().cpu().view(HEADS, C)).sum(-1)
    loss = (dagg * aggv.reshape(n, H)).sum() + (gt * abar).sum() + (gc * S).sum()
    loss.backward()
    out.update(dq=q.grad, dk=k.grad, dv=v.grad, bbar=qt.grad, df=f.grad)
    return {kk: vv.detach() for kk, vv in out.items()}


def make_case(n, e, seed, hub, dtype):
    index = skewed_graph(n, e, seed, hub)
    g = torch.Generator().manual_seed(seed + 1)
    proj = (torch.randn(n, 4 * H, generator=g) * 0.7).to(dtype).to(DEV)
    q, k, v = (proj[:, t * H:(t + 1) * H] for t in range(3))          # strided rows, like the fused projection
    qt = (torch.randn(HEADS, n, H, generator=g) * 0.15).to(dtype).to(DEV)
    f = torch.relu(torch.randn(e, H, generator=g)).to(dtype).to(DEV)
    dagg = torch.randn(n, H, generator=g).to(DEV)
    wc = torch.randn(HEADS, C, H, generator=g) * 0.06
    cvec = (torch.randn(H, generator=g) * 0.3).to(DEV)
    gt = torch.bmm(dagg.cpu().to(dtype).float().view(n, HEADS, C).transpose(0, 1), wc).to(dtype).to(DEV)   # [h, n, H]
    return index, q, k, v, qt, f, dagg, gt, cvec, wc


def run_kernels(index, q, k, v, qt, f, dagg, gt, cvec, wc, p=0.0, seed=0, df_in=None, relu_mask=False):
    n = q.size(0)
    plan = pkg.build_plan(index.to(DEV), n)
    aggv, abar, m, z, s = ops.raw_edgeattn_fwd(q, k, v, qt, f, plan, HEADS, p, seed, 5)
    # the assembled aggregate, as the block epilogue builds it: aggv + Wc[t] abar_t + c_t S_t
    agge = torch.bmm(abar.float(), wc.to(DEV).transpose(1, 2)).transpose(0, 1).reshape(n, H)
    agg = aggv + agge + (cvec.view(HEADS, C) * s.unsqueeze(-1)).reshape(n, H)
    dproj = torch.zeros(n, 4 * H, dtype=q.dtype, device=DEV)
    dq, dk, dv = (dproj[:, t * H:(t + 1) * H] for t in range(3))
    df = torch.empty_like(f)
    bbar = ops.raw_edgeattn_bwd(dagg, dagg.to(q.dtype), agg, q, k, v, qt, gt, cvec, f, m, z, plan, HEADS, dq, dk, dv,
                                df_in, df, relu_mask, p, seed, 5)
    torch.cuda.synchronize()
    return dict(aggv=aggv, abar=abar, S=s, dq=dq, dk=dk, dv=dv, bbar=bbar, df=df, m=m, z=z)


@pytest.mark.parametrize("mma", [False, True])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_edgeattn_forward_backward_vs_fp64(mma, dtype, monkeypatch):
    if mma and dtype != torch.bfloat16:
        pytest.skip("tensor-core family is bf16 only")
    monkeypatch.setattr(ops, "USE_MMA", mma)
    n, e = 211, 4000
    case = make_case(n, e, 3, 700, dtype)
    index, q, k, v, qt, f, dagg, gt, cvec, wc = case
    want = reference(q, k, v, qt, f, index, dagg.to(dtype).float() if mma else dagg, gt, cvec)
    got = run_kernels(*case)
    # the same bf16-rounded operands go into both sides: only accumulation order / coefficient rounding differs
    tol = 1e-5 if dtype == torch.float32 else (1.2e-2 if mma else 6e-3)
    for key in ("aggv", "abar", "S"):
        assert rel_err(got[key], want[key]) < tol, key
    gtol = 1e-4 if dtype == torch.float32 else 2e-2
    for key in ("dq", "dk", "dv", "bbar", "df"):
        assert rel_err(got[key], want[key]) < gtol, key
    # rows without in-edges
    assert float(got["aggv"][-5:].abs().max()) == 0.0 and float(got["abar"][:, -5:].abs().max()) == 0.0


def test_mma_family_matches_cuda_core_family():
    case = make_case(300, 6000, 9, 100, torch.bfloat16)
    ops.USE_MMA = False
    try:
        a = run_kernels(*case)
    finally:
        ops.USE_MMA = True
    b = run_kernels(*case)
    assert rel_err(b["m"], a["m"]) < 1e-5 and rel_err(b["z"], a["z"]) < 1e-4
    for key in ("aggv", "abar", "S", "dq", "dk", "dv", "bbar", "df"):
        assert rel_err(b[key], a[key]) < 1.5e-2, key


@pytest.mark.parametrize("mma", [False, True])
def test_edgeattn_accumulate_and_relu_mask(mma, monkeypatch):
    monkeypatch.setattr(ops, "USE_MMA", mma)
    case = make_case(120, 2500, 5, 64, torch.bfloat16)
    f = case[5]
    base = run_kernels(*case)["df"].float()
    prev = (torch.randn_like(base) * 0.5).to(torch.bfloat16)
    acc = run_kernels(*case, df_in=prev.clone(), relu_mask=True)["df"].float()
    want = (base + prev.float()) * (f.float() > 0)
    assert rel_err(acc, want) < 1e-2


@pytest.mark.parametrize("mma", [False, True])
def test_edgeattn_dropout_forward_backward_masks_agree(mma, monkeypatch):
    """aggv is linear in v for a fixed mask, so <dagg, aggv(v)> = <dv, v> holds iff backward regenerates the
    forward's dropout mask; also E[S] = 1 and a fraction ~p of the coefficients is dropped."""
    monkeypatch.setattr(ops, "USE_MMA", mma)
    case = list(make_case(256, 8000, 13, 300, torch.bfloat16))
    index, q, k, v, qt, f, dagg, gt, cvec, wc = case
    case[4] = qt * 0         # no feature term: S, aggv depend on (q, k, v) only
    case[7] = gt * 0
    case[8] = cvec * 0
    got = run_kernels(*case, p=0.25, seed=1234)
    lhs = float((dagg.double() * got["aggv"].double()).sum())
    rhs = float((got["dv"].double() * v.double()).sum())
    assert abs(lhs - rhs) < 2e-2 * max(abs(lhs), 1.0), (lhs, rhs)
    s = got["S"][: 256 - 5]
    assert abs(float(s.mean()) - 1.0) < 0.05
    other = run_kernels(*case, p=0.25, seed=99)
    assert rel_err(other["aggv"], got["aggv"]) > 1e-3          # a different seed gives a different mask
    again = run_kernels(*case, p=0.25, seed=1234)
    assert torch.equal(again["aggv"], got["aggv"]) and torch.equal(again["dq"], got["dq"])   # bit-reproducible


# ---- line-graph kernels with the angle embedding recomputed in-kernel (csrc/lgattn.cu) ---------------------------------
def lg_case(n, e, seed, hub, in_dim=11):
    index, q, k, v, qt, f, dagg, gt, cvec, wc = make_case(n, e, seed, hub, torch.bfloat16)
    g = torch.Generator().manual_seed(seed + 7)
    a = torch.rand(e, in_dim, generator=g).to(DEV)
    w1 = (torch.randn(H, in_dim, generator=g) * 0.5).to(DEV)
    b1 = (torch.randn(H, generator=g) * 0.2).to(DEV)
    # what the reference computes under bf16 autocast: Linear(bf16 a, bf16 W1, bf16 b1) -> bf16 -> relu
    pre = a.to(torch.bfloat16).float() @ w1.to(torch.bfloat16).float().t() + b1.to(torch.bfloat16).float()
    feat = torch.relu(pre).to(torch.bfloat16)
    return index, q, k, v, qt, feat, dagg, gt, cvec, wc, a, w1, b1


@pytest.mark.parametrize("p", [0.0, 0.2])
def test_lgattn_forward_matches_stored_feature_kernel(p):
    n, e = 301, 7000
    index, q, k, v, qt, feat, dagg, gt, cvec, wc, a, w1, b1 = lg_case(n, e, 21, 500)
    plan = pkg.build_plan(index.to(DEV), n)
    a_csr = ops.pack_angles(a, plan)
    assert a_csr.shape == (e, 16) and float(a_csr[:, 11].min()) == 1.0 and float(a_csr[:, 12:].abs().max()) == 0.0
    assert torch.equal(a_csr[:, :11], a[plan.eid.long()].to(torch.bfloat16))
    got = ops.raw_lgattn_fwd(q, k, v, qt, a_csr, w1, b1, plan, HEADS, 0.0, 7, 3)
    want = ops.raw_edgeattn_fwd(q, k, v, qt, feat, plan, HEADS, 0.0, 7, 3)
    for name, x, y in zip(("aggv", "abar", "m", "z", "s"), got, want):
        assert rel_err(x, y) < (1e-2 if name != "m" else 1e-3), name
    if p > 0:   # row-interleaved qt/abar strides + dropout statistics
        qt_i = qt.permute(1, 0, 2).contiguous().permute(1, 0, 2)           # [h, n, 256] view of an [n, h, 256] buffer
        d1 = ops.raw_lgattn_fwd(q, k, v, qt_i, a_csr, w1, b1, plan, HEADS, p, 7, 3)
        d2 = ops.raw_lgattn_fwd(q, k, v, qt, a_csr, w1, b1, plan, HEADS, p, 7, 3)
        assert torch.equal(d1[0], d2[0]) and torch.equal(d1[1], d2[1])
        step = torch.tensor([5], dtype=torch.int64, device=DEV)
        d3 = ops.raw_lgattn_fwd(q, k, v, qt, a_csr, w1, b1, plan, HEADS, p, 7, 3, rng_step=step)
        d4 = ops.raw_lgattn_fwd(q, k, v, qt, a_csr, w1, b1, plan, HEADS, p, 7, 8)
        assert torch.equal(d3[0], d4[0]) and not torch.equal(d3[0], d2[0])   # device counter == host offset
        s = d2[4][: n - 5]
        assert abs(float(s.mean()) - 1.0) < 0.05


@pytest.mark.parametrize("n,e,hub,p", [(9, 40, 4, 0.0), (301, 7000, 500, 0.0), (301, 7000, 500, 0.2), (64, 9000, 3000, 0.0),
                                       (2000, 22000, 10, 0.15)])
def test_lgattn_tcgen05_forward_matches_mma_sync_kernel(n, e, hub, p, monkeypatch):
    """csrc/lgattn_tc.cu (tcgen05.mma + TMEM, 128-angle tiles) against csrc/lgattn.cu (mma.sync, 16-angle chunks): same
    statistics to fp32 rounding, same dropout masks, outputs within bf16 rounding of the probabilities; multi-row tiles,
    rows split over several tiles (online rescale in TMEM), rows without in-edges."""
    index, q, k, v, qt, feat, dagg, gt, cvec, wc, a, w1, b1 = lg_case(n, e, 31, hub)
    plan = pkg.build_plan(index.to(DEV), n)
    a_csr = ops.pack_angles(a, plan)
    monkeypatch.setattr(ops, "LGATTN_TC", False)
    want = ops.raw_lgattn_fwd(q, k, v, qt, a_csr, w1, b1, plan, HEADS, p, 7, 3)
    monkeypatch.setattr(ops, "LGATTN_TC", True)
    got = ops.raw_lgattn_fwd(q, k, v, qt, a_csr, w1, b1, plan, HEADS, p, 7, 3)
    again = ops.raw_lgattn_fwd(q, k, v, qt, a_csr, w1, b1, plan, HEADS, p, 7, 3)
    for name, x, y, z in zip(("aggv", "abar", "m", "z", "s"), got, want, again):
        assert not torch.isnan(x).any(), name
        assert rel_err(x, y) < (2e-2 if name in ("aggv", "abar") else 1e-5), name
        assert torch.equal(x, z), name                                   # deterministic
    # against the stored-feature reference path too (independent of both in-kernel-feature kernels)
    ref = ops.raw_edgeattn_fwd(q, k, v, qt, feat, plan, HEADS, 0.0, 7, 3) if p == 0.0 else None
    if ref is not None:
        for name, x, y in zip(("aggv", "abar"), got, ref):
            assert rel_err(x, y) < 2e-2, name


def test_lgattn_backward_and_angle_gradient_match_stored_feature_kernels():
    """dq / bbar / dk / dv of the in-kernel-feature backward == the stored-feature tensor-core backward fed with the
    same h1; the fused angle-encoder gradient == (sum_l df_l * [h1 > 0])^T a from the stored-feature kernels' df."""
    n, e, layers = 257, 6000, 3
    index, q, k, v, qt, feat, dagg, gt, cvec, wc, a, w1, b1 = lg_case(n, e, 31, 300)
    plan = pkg.build_plan(index.to(DEV), n)
    a_csr = ops.pack_angles(a, plan)
    g = torch.Generator().manual_seed(5)
    coefs, qts, gts = [], [], []
    df_total = torch.zeros(e, H, device=DEV)
    for l in range(layers):
        qt_l = (qt.float() * (1.0 + 0.3 * l)).to(torch.bfloat16)
        gt_l = (gt.float() * (1.0 - 0.2 * l)).to(torch.bfloat16)
        dagg_l = dagg * (1.0 + 0.1 * l)
        p_drop, seed = (0.0, 0) if l != 1 else (0.1, 77)
        aggv, abar, m, z, s = ops.raw_lgattn_fwd(q, k, v, qt_l, a_csr, w1, b1, plan, HEADS, p_drop, seed, 3)
        agge = torch.bmm(abar.float(), wc.to(DEV).transpose(1, 2)).transpose(0, 1).reshape(n, H)
        agg = aggv + agge + (cvec.view(HEADS, C) * s.unsqueeze(-1)).reshape(n, H)
        dlp = dagg_l.to(torch.bfloat16)
        outs = []
        for jit in (True, False):
            dproj = torch.zeros(n, 4 * H, dtype=torch.bfloat16, device=DEV)
            dq, dk, dv = (dproj[:, t * H:(t + 1) * H] for t in range(3))
            if jit:
                bbar = torch.empty(HEADS, n, H, dtype=torch.bfloat16, device=DEV)
                coef = ops.raw_lgattn_bwd(dagg_l, dlp, agg, q, k, v, qt_l, gt_l, cvec, a_csr, w1, b1, m, z, plan, HEADS,
                                          dq, dk, dv, bbar, p_drop, seed, 3)
                coefs.append(coef); qts.append(qt_l); gts.append(gt_l)
                outs.append((dq.clone(), dk.clone(), dv.clone(), bbar))
            elif p_drop == 0.0:    # (dropout streams differ between the families: eid- vs position-keyed)
                df = torch.empty_like(feat)
                bbar = ops.raw_edgeattn_bwd(dagg_l, dlp, agg, q, k, v, qt_l, gt_l, cvec, feat, m, z, plan, HEADS, dq, dk,
                                            dv, None, df, False, 0.0, 0, 3)
                outs.append((dq.clone(), dk.clone(), dv.clone(), bbar))
                df_total += df.float()
        if len(outs) == 2:
            for name, x, y in zip(("dq", "dk", "dv", "bbar"), outs[0], outs[1]):
                assert rel_err(x, y) < 1.5e-2, (l, name)
    # angle gradient over the dropout-free layers, against the stored-feature df
    keep = [0, 2]
    dw1, db1 = ops.raw_lg_angle_grad(a_csr, w1, b1, plan, [coefs[i] for i in keep], [qts[i] for i in keep],
                                     [gts[i] for i in keep])
    masked = df_total * (feat.float() > 0)
    want_w = masked.t() @ a.to(torch.bfloat16).float()
    want_b = masked.sum(0)
    assert rel_err(dw1, want_w) < 2e-2 and rel_err(db1, want_b) < 2e-2
    # determinism + all three layers (incl. the dropout one) run
    d1 = ops.raw_lg_angle_grad(a_csr, w1, b1, plan, coefs, qts, gts)
    d2 = ops.raw_lg_angle_grad(a_csr, w1, b1, plan, coefs, qts, gts)
    assert torch.equal(d1[0], d2[0]) and torch.equal(d1[1], d2[1]) and bool(torch.isfinite(d1[0]).all())


# ---- the in-kernel-feature line-graph kernels DIRECTLY against the oracle (not against sibling CUDA kernels) -----------
def _round_bf16_(module):
    with torch.no_grad():
        for p_ in module.parameters():
            p_.copy_(p_.to(torch.bfloat16).to(p_.dtype))


@pytest.mark.parametrize("case", ["pyg", "bonds", "skewed"])
def test_lgattn_fwd_bwd_and_angle_grad_directly_vs_fp64_oracle(case):
    """``angle_encoder`` + ``EdgeUpdateBlock`` of the oracle (reference ``train.py:303-317, 360-364``; PyG arithmetic from
    the shim) in fp64 on bf16-representable parameters and inputs, against ONE line-graph block on the kernels the
    benchmark times: ``alignn_lgattn_fwd`` / ``alignn_lgattn_bwd_dst`` / ``alignn_edgeattn_bwd_src_lp`` /
    ``alignn_lg_angle_grad`` (+ the gate/LayerNorm kernels and the stacked projection).  Output and EVERY gradient --
    bond states, both angle-encoder layers, all conv parameters, LayerNorm -- per tensor at 2e-2 of its own scale."""
    from oracle import model_ref
    from conftest import check_per_tensor
    from gnn_elasticity_predictor_b200 import modules
    from gnn_elasticity_predictor_b200.fused import LgShared
    torch.manual_seed(11)
    if case == "skewed":
        n_b, n_l = 301, 7000
        index = skewed_graph(n_b, n_l, 41, 500)
        ang = torch.rand(n_l, 11)
    else:
        b = pkg.synthetic_batch(6, 10, 6, seed=5, lg_inc=case)
        index, ang, n_b = b.lg_edge_index, b.lg_edge_attr, b.edge_index.size(1)
        n_l = index.size(1)
    enc = torch.nn.Sequential(torch.nn.Linear(11, H), torch.nn.ReLU(), torch.nn.Linear(H, H)).double()
    blk = model_ref.EdgeUpdateBlock(H, HEADS, 0.0).double()
    with torch.no_grad():
        blk.norm.weight.uniform_(0.5, 1.5); blk.norm.bias.uniform_(-0.3, 0.3)
    _round_bf16_(enc); _round_bf16_(blk)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(n_b, H, generator=g).to(torch.bfloat16).float()
    ang = ang.to(torch.bfloat16).float()
    gout = torch.randn(n_b, H, generator=g)
    # oracle, fp64
    xr = x.double().requires_grad_(True)
    yr = blk(xr, index, enc(ang.double()))
    yr.backward(gout.double())
    want = {f"blk.{k}": p_.grad for k, p_ in blk.named_parameters()}
    want.update({f"enc.{k}": p_.grad for k, p_ in enc.named_parameters()})
    want["x"] = xr.grad
    # the reference's OWN bf16-autocast run of the same block (oracle on CPU): the per-tensor yardstick for exceptions
    import copy
    enc_a, blk_a = copy.deepcopy(enc).float(), copy.deepcopy(blk).float()
    for m_ in (enc_a, blk_a):
        m_.zero_grad(set_to_none=True)
    xa = x.clone().requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16):
        ya = blk_a(xa, index, enc_a(ang))
    ya.float().backward(gout)
    amp = {f"blk.{k}": p_.grad for k, p_ in blk_a.named_parameters()}
    amp.update({f"enc.{k}": p_.grad for k, p_ in enc_a.named_parameters()})
    amp["x"] = xa.grad
    # ours: the lg family, one block, exactly as AlignnRegressor._trunk_streaming drives it
    ours = pkg.EdgeUpdateBlock(H, HEADS, 0.0).to(DEV)
    ours.load_state_dict({k: v.float() for k, v in blk.state_dict().items()}, strict=True)
    enc_o = torch.nn.Sequential(torch.nn.Linear(11, H), torch.nn.ReLU(), torch.nn.Linear(H, H)).to(DEV)
    enc_o.load_state_dict({k: v.float() for k, v in enc.state_dict().items()})
    cd = torch.bfloat16
    assert ops.lgattn_enabled(H, HEADS, 11, cd)
    plan = pkg.build_plan(index.to(DEV), n_b)
    w1p, b1p = enc_o[0].weight, enc_o[0].bias
    lg = LgShared(ops.pack_angles(ang.to(DEV), plan), w1p.detach().contiguous().float(), b1p.detach().contiguous().float(), 1)
    we = ours.conv.lin_edge.weight.float()
    wc, cvec = we @ enc_o[2].weight.float(), we @ enc_o[2].bias.float()
    xo = x.to(DEV).requires_grad_(True)
    k0 = ops.STATS.kernels
    yo, _ = modules._stream_block(ours.conv, ours.norm, 0.0, True, xo, None, None, None, wc, cvec, plan, cd,
                                  is_last_visitor=True, lg=lg, w1=w1p, b1=b1p)
    yo.backward(gout.to(DEV))
    assert ops.STATS.kernels > k0
    assert rel_err(yo, yr) < 2e-2
    got = {f"blk.{k}": p_.grad for k, p_ in ours.named_parameters()}
    got.update({f"enc.{k}": p_.grad for k, p_ in enc_o.named_parameters()})
    got["x"] = xo.grad
    check_per_tensor(got, want, 2e-2, {"conv.lin_key.bias": (("abs", 2e-3), "true gradient is exactly zero")},
                     label=f"lg_block_vs_oracle_{case}", amp=amp)
