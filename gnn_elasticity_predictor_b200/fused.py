"""Block-level autograd for the streaming path (hidden = 256): one ``torch.autograd.Function`` per conv block,
hand-written backward, no per-edge dense projection anywhere.

Algebra (see ``csrc/edgeattn.cu``): with per-edge features ``f`` and the folded edge projection
``e = Wc f + c`` (line graph: ``f = h1 = relu(W1 a + b1)``, ``Wc = W_e W2``, ``c = W_e b2``, folding the second
Linear of ``angle_encoder`` -- reference ``train.py:360-364`` -- into PyG's ``lin_edge``; atom graph: ``f`` = bond state,
``Wc = W_e W_p``, ``c = W_e b_p``, folding ``edge_proj`` -- ``train.py:324,333``), every ``E x H x H`` contraction of the
reference becomes an ``N x H x H`` one on the target nodes:

    forward :  P = x [Wq;Wk;Wv;Ws]^T + b      QT_t = q_t Wc[t]        (per node)
               (aggv, abar, stats) = edgeattn_fwd(q, k, v, QT, f)      (per edge, streams f once)
               agg = aggv + abar_t Wc[t]^T + c_t S_t ;  y = x + drop(relu(LN(beta xr + (1-beta) agg)))
    backward:  gate/LN backward -> dagg ;  GT_t = dagg_t Wc[t]
               (dq_direct, bbar, coef, df) = edgeattn_bwd_dst(...) ;  (dk, dv) = edgeattn_bwd_src(coef, ...)
               dq += bbar_t Wc[t]^T ;  dWc[t] = dagg_t^T abar_t + q_t^T bbar_t ;  dc_t = sum_i dagg_i,t S_i,t
               dx = dy + dP [Wq;Wk;Wv;Ws] ;  dW4 = dP^T x ;  db4 = colsum(dP)

The feature gradient ``df`` of the line-graph layers, which all share ``h1``, is accumulated in place across layers
inside the backward kernels (:class:`FeatGradAccumulator`), ReLU-masked by the last one, and handed to
:class:`_AngleH1` whose backward reduces it to ``dW1, db1`` -- no ``[L, H]`` gradient is ever summed by autograd.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import torch
from torch import Tensor

from . import ops
from .ops import GraphPlan


class FeatGradAccumulator:
    """Shared ``[L, H]`` buffer that the line-graph blocks' backward kernels accumulate ``df`` into.

    Backward visits layers in reverse order: the first visitor (``order == 0``) writes, later ones
    read-modify-write in place, the last one (``order == n - 1``) also applies the ReLU mask of ``h1``.
    """

    def __init__(self, n_blocks: int):
        self.n_blocks = n_blocks
        self.buf: Optional[Tensor] = None
        self.visits = 0


class LgShared:
    """Per-forward state of the in-kernel-feature line-graph family (``csrc/lgattn.cu``): the packed angle rows, the first
    angle-encoder layer, and -- filled during backward -- every layer's (coef, qt, gt), from which the LAST visited
    block (layer 0) forms ``dW1, db1`` in one fused pass (no ``[L, H]`` gradient ever exists)."""

    def __init__(self, a_csr: Tensor, w1: Tensor, b1: Tensor, n_blocks: int):
        self.a_csr, self.w1, self.b1, self.n_blocks = a_csr, w1, b1, n_blocks
        self.coefs, self.qts, self.gts = [], [], []


@dataclass
class BlockCfg:
    heads: int
    eps: float
    p_attn: float
    p_out: float
    seed_attn: int
    off_attn: int
    seed_out: int
    off_out: int
    cd: torch.dtype
    want_lp: bool                      # also emit a compute-dtype copy of y
    accum: Optional[FeatGradAccumulator] = None   # line-graph blocks sharing h1
    is_last_visitor: bool = False      # this block's backward applies the ReLU mask and returns df to the anchor
    anchor_dtype: Optional[torch.dtype] = None
    lg: Optional[LgShared] = None      # in-kernel-feature line-graph family (bf16, hidden 256, 4 heads)
    strided: bool = False              # stored-feature tensor-core family with strided operands + device RNG counter
    active_rows: int = -1              # inference only (no autograd): rows >= active_rows are isolated in this block's graph --
                                       # q | k | v | qt, the attention kernel and the abar products run on the prefix only


class _LinearCS(torch.autograd.Function):
    """``x W^T + b`` (cuBLAS) whose backward forms the bias gradient with the deterministic ``colsum`` kernel instead of
    torch's generic reduction (the encoders' second Linears see ``[E, 256]`` gradients: reference ``train.py:350-359``)."""

    @staticmethod
    def forward(ctx, x: Tensor, w: Tensor, b: Tensor, relu: bool):
        y = torch.addmm(b, x, w.t())
        if relu:
            y = torch.relu_(y)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, dy: Tensor):
        x, w, y = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.relu:
            dy = torch.ops.aten.threshold_backward(dy, y, 0)
        dx = dy @ w if ctx.needs_input_grad[0] else None
        dw = torch.mm(dy.t(), x, out_dtype=torch.float32) if dy.dtype != torch.float32 else dy.t() @ x
        db = ops.colsum(dy)
        return dx, dw.to(w.dtype), db.to(w.dtype), None


def mlp2(x: Tensor, w0: Tensor, b0: Tensor, w2: Tensor, b2: Tensor, cd: torch.dtype) -> Tensor:
    """``Linear -> ReLU -> Linear`` encoder in compute dtype ``cd`` (reference ``train.py:350-359``)."""
    k = x.size(1)
    if cd != torch.float32 and k % 8:
        # pad the input width to a multiple of 8 elements (16 bytes): cuBLAS otherwise falls back to its unaligned
        # legacy kernels for the [n, 36] / [n, 206] first-layer GEMMs (5x slower); zero columns change nothing
        kp = (k + 7) // 8 * 8
        xp = torch.zeros(x.size(0), kp, dtype=cd, device=x.device)
        xp[:, :k] = x
        w0p = torch.nn.functional.pad(w0.to(cd), (0, kp - k))
        h = _LinearCS.apply(xp, w0p, b0.to(cd), True)
    else:
        h = _LinearCS.apply(x.to(cd), w0.to(cd), b0.to(cd), True)
    return _LinearCS.apply(h, w2.to(cd), b2.to(cd), False)


class _AngleH1(torch.autograd.Function):
    """``h1 = relu(W1 a + b1)`` (first angle-encoder layer).  Its backward expects the ALREADY ReLU-masked gradient
    (the last line-graph block's backward kernel applies the mask while accumulating)."""

    @staticmethod
    def forward(ctx, a: Tensor, w1: Tensor, b1: Tensor, cd: torch.dtype):
        a = a.contiguous().float()
        h1 = ops.raw_angle_h1_fwd(a, w1.detach().contiguous().float(), b1.detach().contiguous().float(), cd)
        ctx.save_for_backward(a)
        ctx.hidden = int(w1.size(0))
        ctx.param_dtypes = (w1.dtype, b1.dtype)
        return h1

    @staticmethod
    def backward(ctx, dpre: Tensor):
        (a,) = ctx.saved_tensors
        dw1, db1 = ops.raw_angle_h1_bwd(dpre.contiguous(), a, ctx.hidden)
        return None, dw1.to(ctx.param_dtypes[0]), db1.to(ctx.param_dtypes[1]), None


def angle_h1(a: Tensor, w1: Tensor, b1: Tensor, cd: torch.dtype) -> Tensor:
    return _AngleH1.apply(a, w1, b1, cd)


class _AttnBlock(torch.autograd.Function):
    """One EdgeUpdateBlock / NodeUpdateBlock (reference ``train.py:303-336``) on the streaming kernels.

    Kernel family per block: ``cfg.lg`` -> in-kernel angle features (``feat`` unused, ``w1 / b1`` differentiable at the
    block that backward visits last); ``cfg.strided`` -> stored-feature tensor-core kernels; else the generic
    stored-feature kernels (fp32 / shared-h1 accumulation)."""

    @staticmethod
    def forward(ctx, x32: Tensor, xb: Optional[Tensor], feat: Optional[Tensor], anchor: Optional[Tensor], w4: Tensor,
                b4: Tensor, wc: Tensor, cvec: Optional[Tensor], wbeta: Tensor, gamma: Tensor, beta_ln: Tensor,
                w1: Optional[Tensor], b1: Optional[Tensor], plan: GraphPlan, cfg: BlockCfg):
        cd, h = cfg.cd, cfg.heads
        n, hid = x32.shape
        c = hid // h
        rs = ops.RNG_STEP
        x32 = x32.contiguous()
        if xb is None or xb.dtype != cd:
            xb = x32.to(cd)
        if cfg.lg is None:
            feat = feat.contiguous()
            if feat.dtype != cd:
                feat = feat.to(cd)
        else:
            feat = None
        w4c, b4c = w4.to(cd), b4.to(cd)
        if 0 <= cfg.active_rows < n:
            # Forward-only elision of the isolated rows (PyG-collated line graphs; the fp32 regime of
            # predict.py / evaluate.py): x_r for every row, everything attention-related for the active prefix.  Nothing
            # is saved: the caller guarantees that autograd is off.
            na = int(cfg.active_rows)
            xr = torch.addmm(b4c[3 * hid:], xb, w4c[3 * hid:].t())            # [n, H]
            p3 = torch.addmm(b4c[:3 * hid], xb[:na], w4c[:3 * hid].t())       # [na, 3H]: q | k | v
            wc3 = wc.to(cd).view(h, c, hid)
            q, k, v = (p3[:, i * hid:(i + 1) * hid] for i in range(3))
            qt = torch.bmm(q.unflatten(1, (h, c)).transpose(0, 1), wc3)       # [h, na, H]
            if cfg.lg is not None:
                lg = cfg.lg
                aggv, abar, m, z, s = ops.raw_lgattn_fwd(q, k, v, qt, lg.a_csr, lg.w1, lg.b1, plan, h, cfg.p_attn,
                                                         cfg.seed_attn, cfg.off_attn, rs)
            elif cfg.strided:
                aggv, abar, m, z, s = ops.raw_attn_fwd_s(q, k, v, qt, feat, plan, h, cfg.p_attn, cfg.seed_attn,
                                                         cfg.off_attn, rs)
            else:
                aggv, abar, m, z, s = ops.raw_edgeattn_fwd(q, k, v, qt, feat, plan, h, cfg.p_attn, cfg.seed_attn,
                                                           cfg.off_attn)
            agge = torch.bmm(abar, wc3.transpose(1, 2))                       # [h, na, C]
            cv = cvec.detach().contiguous().float() if cvec is not None else None
            y, y_lp, _, _, _, _ = ops.raw_gate_ln_fwd2(aggv, agge, cv, s if cv is not None else None, h, xr, x32,
                                                       wbeta.detach().reshape(-1).contiguous().float(),
                                                       gamma.detach().contiguous().float(),
                                                       beta_ln.detach().contiguous().float(), cfg.eps, cfg.p_out,
                                                       cfg.seed_out, cfg.off_out, cfg.want_lp and cd != torch.float32, rs,
                                                       agg_rows=na)
            ctx.elided = True
            if y_lp is not None:
                ctx.mark_non_differentiable(y_lp)
            return y, y_lp
        proj = torch.addmm(b4c, xb, w4c.t())                                  # [n, 4H]: q | k | v | skip
        wc3 = wc.to(cd).view(h, c, hid)                                       # Wc[t] : [C, H]
        q, k, v, xr = (proj[:, i * hid:(i + 1) * hid] for i in range(4))
        q3 = q.unflatten(1, (h, c)).transpose(0, 1)                           # [h, n, C] view
        qt = torch.bmm(q3, wc3)                                               # [h, n, H]
        if cfg.lg is not None:
            lg = cfg.lg
            aggv, abar, m, z, s = ops.raw_lgattn_fwd(q, k, v, qt, lg.a_csr, lg.w1, lg.b1, plan, h, cfg.p_attn,
                                                     cfg.seed_attn, cfg.off_attn, rs)
        elif cfg.strided:
            aggv, abar, m, z, s = ops.raw_attn_fwd_s(q, k, v, qt, feat, plan, h, cfg.p_attn, cfg.seed_attn,
                                                     cfg.off_attn, rs)
        else:
            aggv, abar, m, z, s = ops.raw_edgeattn_fwd(q, k, v, qt, feat, plan, h, cfg.p_attn, cfg.seed_attn,
                                                       cfg.off_attn)
        agge = torch.bmm(abar, wc3.transpose(1, 2))                           # [h, n, C]
        cv = cvec.detach().contiguous().float() if cvec is not None else None
        wb = wbeta.detach().reshape(-1).contiguous().float()
        gm, bl = gamma.detach().contiguous().float(), beta_ln.detach().contiguous().float()
        y, y_lp, agg, beta, mean, rstd = ops.raw_gate_ln_fwd2(aggv, agge, cv, s if cv is not None else None, h, xr, x32,
                                                              wb, gm, bl, cfg.eps, cfg.p_out, cfg.seed_out, cfg.off_out,
                                                              cfg.want_lp and cd != torch.float32, rs)
        ctx.save_for_backward(xb, feat, proj, qt, abar, agg, m, z, s, beta, mean, rstd, w4c, wc3, cv, wb, gm, bl)
        ctx.plan, ctx.cfg, ctx.rs = plan, cfg, rs
        ctx.shapes = (wbeta.shape, wbeta.dtype, gamma.dtype, beta_ln.dtype, w4.dtype, b4.dtype, wc.dtype,
                      None if cvec is None else cvec.dtype, None if w1 is None else (w1.dtype, b1.dtype))
        if y_lp is not None:
            ctx.mark_non_differentiable(y_lp)
        ctx.set_materialize_grads(False)
        return y, y_lp                                                        # y_lp is None in the fp32 regime

    @staticmethod
    def backward(ctx, dy: Optional[Tensor], _dy_lp):
        if getattr(ctx, "elided", False):
            raise RuntimeError("this block ran the forward-only (isolated-row elision) path: BlockCfg.active_rows is for "
                               "inference without autograd")
        xb, feat, proj, qt, abar, agg, m, z, s, beta, mean, rstd, w4c, wc3, cv, wb, gm, bl = ctx.saved_tensors
        plan, cfg, rs = ctx.plan, ctx.cfg, ctx.rs
        cd, h = cfg.cd, cfg.heads
        n, hid = agg.shape
        c = hid // h
        if dy is None:
            dy = torch.zeros_like(agg)
        dy = dy.contiguous().float()
        q, k, v, xr = (proj[:, i * hid:(i + 1) * hid] for i in range(4))
        dproj = torch.empty_like(proj)
        dq, dk, dv, dxr = (dproj[:, i * hid:(i + 1) * hid] for i in range(4))

        dagg, dagg_lp, dparams = ops.raw_gate_ln_bwd2(dy, agg, xr, wb, gm, bl, beta, mean, rstd, dxr,
                                                      cd != torch.float32, cfg.p_out, cfg.seed_out, cfg.off_out, rs)
        if dagg_lp is None:
            dagg_lp = dagg
        g3 = dagg_lp.unflatten(1, (h, c)).transpose(0, 1)                     # [h, n, C]
        gt = torch.bmm(g3, wc3)                                               # [h, n, H]

        acc = cfg.accum
        df_out = dw1 = db1 = None
        if cfg.lg is not None:
            lg = cfg.lg
            bbar = torch.empty(h, n, hid, dtype=cd, device=agg.device)
            coef = ops.raw_lgattn_bwd(dagg, dagg_lp, agg, q, k, v, qt, gt, cv, lg.a_csr, lg.w1, lg.b1, m, z, plan, h,
                                      dq, dk, dv, bbar, cfg.p_attn, cfg.seed_attn, cfg.off_attn, rs)
            lg.coefs.append(coef); lg.qts.append(qt); lg.gts.append(gt)
            if cfg.is_last_visitor:
                dw1, db1 = ops.raw_lg_angle_grad(lg.a_csr, lg.w1, lg.b1, plan, lg.coefs, lg.qts, lg.gts)
                lg.coefs, lg.qts, lg.gts = [], [], []
        elif cfg.strided:
            bbar = torch.empty(h, n, hid, dtype=cd, device=agg.device)
            df_out = torch.zeros_like(feat)       # rows of edges the plan dropped (padding) are never written
            ops.raw_attn_bwd_s(dagg, dagg_lp, agg, q, k, v, qt, gt, cv, feat, m, z, plan, h, dq, dk, dv, bbar, df_out,
                               cfg.p_attn, cfg.seed_attn, cfg.off_attn, rs)
        else:
            # feature gradient: shared accumulator (line graph) or a fresh buffer (atom graph / standalone)
            relu_mask = False
            if acc is not None:
                first = acc.visits == 0
                acc.visits += 1
                if acc.buf is None:
                    acc.buf = torch.zeros_like(feat)
                df_in, df_out = (None if first else acc.buf), acc.buf
                relu_mask = cfg.is_last_visitor
            else:
                df_in, df_out = None, torch.zeros_like(feat)
            bbar = ops.raw_edgeattn_bwd(dagg, dagg_lp, agg, q, k, v, qt, gt, cv, feat, m, z, plan, h, dq, dk, dv, df_in,
                                        df_out, relu_mask, cfg.p_attn, cfg.seed_attn, cfg.off_attn)

        q3 = q.unflatten(1, (h, c)).transpose(0, 1)
        dq.unflatten(1, (h, c)).add_(torch.bmm(bbar, wc3.transpose(1, 2)).transpose(0, 1))   # dq += bbar_t Wc[t]^T
        dwc = (torch.bmm(g3.transpose(1, 2), abar) + torch.bmm(q3.transpose(1, 2), bbar)).reshape(hid, hid)
        dcvec = None
        if cv is not None:
            dcvec = (dagg.view(n, h, c) * s.unsqueeze(-1)).sum(0).reshape(hid)

        dxb = dproj @ w4c                                                     # [n, H]
        dx32 = dy + dxb                                                       # promotes to fp32
        dw4 = dproj.t() @ xb
        db4 = dproj.sum(0, dtype=torch.float32)

        wshape, wdt, gdt, bdt, w4dt, b4dt, wcdt, cdt, w1dts = ctx.shapes
        d_anchor = None
        if df_out is not None and ctx.needs_input_grad[3]:
            if acc is None or cfg.is_last_visitor:
                d_anchor = df_out if cfg.anchor_dtype in (None, df_out.dtype) else df_out.to(cfg.anchor_dtype)
        if dw1 is not None and w1dts is not None:
            dw1, db1 = dw1.to(w1dts[0]), db1.to(w1dts[1])
        else:
            dw1 = db1 = None
        return (dx32, None, None, d_anchor, dw4.to(w4dt), db4.to(b4dt), dwc.to(wcdt),
                None if dcvec is None else dcvec.to(cdt), dparams[:3 * hid].reshape(wshape).to(wdt),
                dparams[3 * hid:4 * hid].to(gdt), dparams[4 * hid:].to(bdt), dw1, db1, None, None)


class _AttnBlock8(torch.autograd.Function):
    """bf16 tensor-core families (``cfg.lg`` / ``cfg.strided``) with the query-side fold moved INTO the node projection:

        proj8 = x [Wq; Wk; Wv; Ws; Wc[0]^T Wq_0; ..; Wc[3]^T Wq_3]^T + b8        ->  q | k | v | x_r | qt_0 .. qt_3

    one ``[n, H] x [H, 8H]`` GEMM instead of the 4H projection + a batched ``q_t Wc[t]`` product, and in backward the
    kernels write ``dq | dk | dv | dx_r | bbar_0..3`` straight into one ``[n, 8H]`` buffer, so that ``dx = dy + dproj8 W8``
    (fp32 accumulate-into GEMM), ``dW8 = dproj8^T x`` and ``db8 = colsum(dproj8)`` replace five GEMMs, a strided add and
    two reductions.  ``W8 / b8`` are built from the parameters with autograd (tiny [H, H] products), which carries the chain
    rule back to ``lin_query`` and the folded edge projection."""

    @staticmethod
    def forward(ctx, x32: Tensor, xb: Optional[Tensor], feat: Optional[Tensor], anchor: Optional[Tensor], w8: Tensor,
                b8: Tensor, wc: Tensor, cvec: Optional[Tensor], wbeta: Tensor, gamma: Tensor, beta_ln: Tensor,
                w1: Optional[Tensor], b1: Optional[Tensor], plan: GraphPlan, cfg: BlockCfg):
        cd, h = cfg.cd, cfg.heads
        n, hid = x32.shape
        c = hid // h
        rs = ops.RNG_STEP
        x32 = x32.contiguous()
        if xb is None or xb.dtype != cd:
            xb = x32.to(cd)
        if cfg.lg is None:
            feat = feat.contiguous()
            if feat.dtype != cd:
                feat = feat.to(cd)
        else:
            feat = None
        w8c, b8c = w8.to(cd), b8.to(cd)
        proj = torch.addmm(b8c, xb, w8c.t())                                  # [n, 8H]
        wc3 = wc.to(cd).view(h, c, hid)                                       # Wc[t] : [C, H]
        q, k, v, xr = (proj[:, i * hid:(i + 1) * hid] for i in range(4))
        qt = proj[:, 4 * hid:].unflatten(1, (h, hid)).transpose(0, 1)         # [h, n, H] view (row stride 8H)
        if cfg.lg is not None:
            lg = cfg.lg
            aggv, abar, m, z, s = ops.raw_lgattn_fwd(q, k, v, qt, lg.a_csr, lg.w1, lg.b1, plan, h, cfg.p_attn,
                                                     cfg.seed_attn, cfg.off_attn, rs)
        else:
            aggv, abar, m, z, s = ops.raw_attn_fwd_s(q, k, v, qt, feat, plan, h, cfg.p_attn, cfg.seed_attn,
                                                     cfg.off_attn, rs)
        agge = torch.bmm(abar, wc3.transpose(1, 2))                           # [h, n, C]
        cv = cvec.detach().contiguous().float() if cvec is not None else None
        wb = wbeta.detach().reshape(-1).contiguous().float()
        gm, bl = gamma.detach().contiguous().float(), beta_ln.detach().contiguous().float()
        y, y_lp, agg, beta, mean, rstd = ops.raw_gate_ln_fwd2(aggv, agge, cv, s if cv is not None else None, h, xr, x32,
                                                              wb, gm, bl, cfg.eps, cfg.p_out, cfg.seed_out, cfg.off_out,
                                                              cfg.want_lp, rs)
        ctx.save_for_backward(xb, feat, proj, abar, agg, m, z, s, beta, mean, rstd, w8c, wc3, cv, wb, gm, bl)
        ctx.plan, ctx.cfg, ctx.rs = plan, cfg, rs
        ctx.shapes = (wbeta.shape, wbeta.dtype, gamma.dtype, beta_ln.dtype, w8.dtype, b8.dtype, wc.dtype,
                      None if cvec is None else cvec.dtype, None if w1 is None else (w1.dtype, b1.dtype))
        if y_lp is not None:
            ctx.mark_non_differentiable(y_lp)
        ctx.set_materialize_grads(False)
        return y, y_lp

    @staticmethod
    def backward(ctx, dy: Optional[Tensor], _dy_lp):
        xb, feat, proj, abar, agg, m, z, s, beta, mean, rstd, w8c, wc3, cv, wb, gm, bl = ctx.saved_tensors
        plan, cfg, rs = ctx.plan, ctx.cfg, ctx.rs
        cd, h = cfg.cd, cfg.heads
        n, hid = agg.shape
        c = hid // h
        if dy is None:
            dy = torch.zeros_like(agg)
        dy = dy.contiguous().float()
        q, k, v, xr = (proj[:, i * hid:(i + 1) * hid] for i in range(4))
        qt = proj[:, 4 * hid:].unflatten(1, (h, hid)).transpose(0, 1)
        dproj = torch.empty_like(proj)
        dq, dk, dv, dxr = (dproj[:, i * hid:(i + 1) * hid] for i in range(4))
        bbar = dproj[:, 4 * hid:].unflatten(1, (h, hid)).transpose(0, 1)      # [h, n, H] view of the same buffer

        if cv is not None:
            dagg, dagg_lp, dparams = ops.raw_gate_ln_bwd3(dy, agg, xr, wb, gm, bl, beta, mean, rstd, s, h, dxr,
                                                          cfg.p_out, cfg.seed_out, cfg.off_out, rs)
            dcvec = dparams[5 * hid:]
        else:
            dagg, dagg_lp, dparams = ops.raw_gate_ln_bwd2(dy, agg, xr, wb, gm, bl, beta, mean, rstd, dxr, True,
                                                          cfg.p_out, cfg.seed_out, cfg.off_out, rs)
            dcvec = None
        g3 = dagg_lp.unflatten(1, (h, c)).transpose(0, 1)                     # [h, n, C]
        gt = torch.bmm(g3, wc3)                                               # [h, n, H]

        df_out = dw1 = db1 = None
        if cfg.lg is not None:
            lg = cfg.lg
            coef = ops.raw_lgattn_bwd(dagg, dagg_lp, agg, q, k, v, qt, gt, cv, lg.a_csr, lg.w1, lg.b1, m, z, plan, h,
                                      dq, dk, dv, bbar, cfg.p_attn, cfg.seed_attn, cfg.off_attn, rs)
            lg.coefs.append(coef); lg.qts.append(qt); lg.gts.append(gt)
            if cfg.is_last_visitor:
                dw1, db1 = ops.raw_lg_angle_grad(lg.a_csr, lg.w1, lg.b1, plan, lg.coefs, lg.qts, lg.gts)
                lg.coefs, lg.qts, lg.gts = [], [], []
        else:
            df_out = torch.zeros_like(feat)       # rows of edges the plan dropped (padding) are never written
            ops.raw_attn_bwd_s(dagg, dagg_lp, agg, q, k, v, qt, gt, cv, feat, m, z, plan, h, dq, dk, dv, bbar, df_out,
                               cfg.p_attn, cfg.seed_attn, cfg.off_attn, rs)

        dwc = torch.bmm(g3.transpose(1, 2), abar).reshape(hid, hid)           # dWc[t] = dagg_t^T abar_t  (+ via W8)
        dx32 = torch.addmm(dy, dproj, w8c, out_dtype=torch.float32)           # dy + dproj8 W8, fp32 accumulate
        dw8 = torch.mm(dproj.t(), xb, out_dtype=torch.float32)                # [8H, H]
        db8 = ops.colsum(dproj)

        wshape, wdt, gdt, bdt, w8dt, b8dt, wcdt, cdt, w1dts = ctx.shapes
        d_anchor = None
        if df_out is not None and ctx.needs_input_grad[3]:
            d_anchor = df_out if cfg.anchor_dtype in (None, df_out.dtype) else df_out.to(cfg.anchor_dtype)
        if dw1 is not None and w1dts is not None:
            dw1, db1 = dw1.to(w1dts[0]), db1.to(w1dts[1])
        else:
            dw1 = db1 = None
        return (dx32, None, None, d_anchor, dw8.to(w8dt), db8.to(b8dt), dwc.to(wcdt),
                None if dcvec is None else dcvec.to(cdt), dparams[:3 * hid].reshape(wshape).to(wdt),
                dparams[3 * hid:4 * hid].to(gdt), dparams[4 * hid:5 * hid].to(bdt), dw1, db1, None, None)


def attn_block8(x32: Tensor, xb: Optional[Tensor], feat: Optional[Tensor], anchor: Optional[Tensor], w8: Tensor,
                b8: Tensor, wc: Tensor, cvec: Optional[Tensor], wbeta: Tensor, gamma: Tensor, beta_ln: Tensor,
                plan: GraphPlan, cfg: BlockCfg, w1: Optional[Tensor] = None, b1: Optional[Tensor] = None):
    """bf16 tensor-core families with the 8H stacked projection (see :class:`_AttnBlock8`)."""
    return _AttnBlock8.apply(x32, xb, feat, anchor, w8, b8, wc, cvec, wbeta, gamma, beta_ln, w1, b1, plan, cfg)


def attn_block(x32: Tensor, xb: Optional[Tensor], feat: Optional[Tensor], anchor: Optional[Tensor], w4: Tensor,
               b4: Tensor, wc: Tensor, cvec: Optional[Tensor], wbeta: Tensor, gamma: Tensor, beta_ln: Tensor,
               plan: GraphPlan, cfg: BlockCfg, w1: Optional[Tensor] = None, b1: Optional[Tensor] = None):
    """Returns ``(y fp32 [N, H], y in compute dtype)``; ``anchor`` is the tensor that receives ``df``; ``w1 / b1`` (lg
    family, the block backward visits last) receive the fused angle-encoder gradient."""
    return _AttnBlock.apply(x32, xb, feat, anchor, w4, b4, wc, cvec, wbeta, gamma, beta_ln, w1, b1, plan, cfg)
