"""The whole message-passing trunk (``layers`` x [EdgeUpdateBlock, NodeUpdateBlock], reference ``scripts/train.py:558-560``)
as ONE explicit forward / backward program over the C-ABI kernels -- the bf16 fast path (hidden 256, 4 heads).

Why one program instead of one autograd node per block (``fused.py``): the backward of the trunk has two gradient streams
that autograd would sum with separate elementwise passes -- every bond state ``edge_state_l`` is consumed by the next
line-graph block (residual stream) AND by the atom-graph block of the same layer (as its edge feature).  Here the backward
is scheduled by hand:

    for l = L-1 .. 0:
        NodeUpdateBlock_l backward   -> writes the bond-feature gradient df_l (bf16) into the LAST H columns of the
                                        [E, 9H] buffer that the line-graph block of the same layer uses for its
                                        projection gradients
        EdgeUpdateBlock_l backward   -> gate/LayerNorm backward reads dy = d_edge (fp32, from block l+1) + df_l on load;
                                        kernels fill columns [0, 8H) with dq | dk | dv | dx_r | bbar_0..3;
                                        d_edge <- d_edge + [dproj8 | df_l] . [W8 ; I]       (one fp32-accumulating GEMM)

so the gradient sum is folded into a GEMM that runs anyway (an identity block appended to the stacked weights), nothing
of size [E, H] is added, copied or converted on its own, and the angle-encoder gradient is formed once at the end from
all layers' coefficients.  Parameters arrive STACKED over blocks (index 2l = EdgeUpdateBlock_l, 2l+1 =
NodeUpdateBlock_l) so that the weight folds of all blocks are a handful of batched [H, H] products (``modules.py``).

Column layout of the stacked projection: ``q | k | v | qt_0..3 | x_r`` (x_r last, next to the df columns of the gradient
buffer).  ISOLATED bond rows: with PyG's default collate the reference offsets ``lg_edge_index`` by atoms,
so in a batch of B > 1 crystals only the first ``lg_active`` bond rows have line-graph neighbours at all -- 8 544 of
98 304 at BASELINE config 2.  For the rows beyond that bound a block is just ``x + drop(relu(LN(beta x_r)))``: only the
``x_r`` slice of the projection (and of every gradient GEMM) is computed for them, and the attention kernels, the
``abar`` products and the column sums run on the active prefix only.  The bound is a host integer the batch carries
(``GraphBatch.lg_active_rows``, the largest line-graph index + 1, known at collate time); without it all rows are active.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import torch
from torch import Tensor

from . import ops
from .ops import GraphPlan


@dataclass
class TrunkCfg:
    heads: int
    n_layers: int
    eps: List[float]                     # LayerNorm eps per block (2 * n_layers)
    p_attn: List[float]                  # attention dropout per block (0 in eval mode)
    p_out: List[float]                   # block-output dropout per block
    keys: List[Tuple[int, int, int, int]]  # (seed_attn, off_attn, seed_out, off_out) per block
    lg_plan: GraphPlan = None
    g_plan: GraphPlan = None
    a_csr: Tensor = None                 # [L, 16] bf16 packed angle rows (target-sorted)
    w1: Tensor = None                    # fp32 [H, angle_dim], detached, contiguous
    b1: Tensor = None
    want_node_lp: bool = False
    lg_active: int = -1                  # bond rows >= lg_active are isolated in the line graph (< 0: unknown, all active)


def _block_forward(idx: int, is_lg: bool, x32: Tensor, xb: Tensor, feat: Optional[Tensor], w8c: Tensor, b8c: Tensor,
                   wc3: Tensor, cv: Tensor, wb: Tensor, gm: Tensor, bl: Tensor, cfg: TrunkCfg, want_lp: bool, rs):
    """One conv block forward on raw kernels.  Returns (y32, y_lp, saved-state dict)."""
    h = cfg.heads
    n, hid = x32.shape
    sa, oa, so, oo = cfg.keys[idx]
    na = cfg.lg_active if (is_lg and 0 <= cfg.lg_active < n) else n            # active prefix
    # [n, 8H]: q | k | v | qt_0..3 | x_r.  Two GEMMs with the same shapes whether or not rows are elided (x_r over all rows,
    # the other seven slices over the active prefix), so elision changes no rounding anywhere.
    proj = torch.empty(n, 8 * hid, dtype=xb.dtype, device=xb.device)
    torch.addmm(b8c[7 * hid:], xb, w8c[7 * hid:].t(), out=proj[:, 7 * hid:])
    torch.addmm(b8c[:7 * hid], xb[:na], w8c[:7 * hid].t(), out=proj[:na, :7 * hid])
    q, k, v = (proj[:na, i * hid:(i + 1) * hid] for i in range(3))
    xr = proj[:, 7 * hid:]
    qt = proj[:na, 3 * hid:7 * hid].unflatten(1, (h, hid)).transpose(0, 1)    # [h, na, H] view, row stride 8H
    abar_rows = torch.empty(na, h, hid, dtype=xb.dtype, device=xb.device)     # row-interleaved: [na, 4H] for the dWc GEMM
    abar = abar_rows.transpose(0, 1)                                          # [h, na, H] view
    if is_lg:
        aggv, _, m, z, s = ops.raw_lgattn_fwd(q, k, v, qt, cfg.a_csr, cfg.w1, cfg.b1, cfg.lg_plan, h, cfg.p_attn[idx],
                                              sa, oa, rs, abar=abar)
    else:
        aggv, _, m, z, s = ops.raw_attn_fwd_s(q, k, v, qt, feat, cfg.g_plan, h, cfg.p_attn[idx], sa, oa, rs, abar=abar)
    agge = torch.bmm(abar, wc3.transpose(1, 2))                               # [h, na, C]
    y, y_lp, agg, beta, mean, rstd = ops.raw_gate_ln_fwd2(aggv, agge, cv, s, h, xr, x32, wb, gm, bl, cfg.eps[idx],
                                                          cfg.p_out[idx], so, oo, want_lp, rs, agg_rows=na)
    st = dict(xb=xb, feat=feat, proj=proj, abar_rows=abar_rows, agg=agg, m=m, z=z, s=s, beta=beta, mean=mean, rstd=rstd,
              na=na)
    return y, y_lp, st


class _Trunk(torch.autograd.Function):
    @staticmethod
    def forward(ctx, node32: Tensor, node_b: Tensor, edge32: Tensor, edge_b: Tensor, w8: Tensor, b8: Tensor, wc: Tensor,
                cvec: Tensor, wbeta: Tensor, gamma: Tensor, beta_ln: Tensor, w1: Tensor, b1: Tensor, cfg: TrunkCfg):
        cd = torch.bfloat16
        h, nl = cfg.heads, cfg.n_layers
        hid = node32.size(1)
        c = hid // h
        rs = ops.RNG_STEP
        w8c, b8c = w8.detach().to(cd), b8.detach().to(cd)                      # [2L, 8H, H], [2L, 8H]
        wc3 = wc.detach().to(cd).view(2 * nl, h, c, hid)
        cvf = cvec.detach().float().contiguous()
        wbf = wbeta.detach().float().contiguous()
        gmf, blf = gamma.detach().float().contiguous(), beta_ln.detach().float().contiguous()
        n32, nb = node32.contiguous().float(), node_b.contiguous()
        e32, eb = edge32.contiguous().float(), edge_b.contiguous()
        if nb.dtype != cd:
            nb = nb.to(cd)
        if eb.dtype != cd:
            eb = eb.to(cd)
        saved = []
        for l in range(nl):
            i = 2 * l
            e32, eb, st = _block_forward(i, True, e32, eb, None, w8c[i], b8c[i], wc3[i], cvf[i], wbf[i], gmf[i], blf[i],
                                         cfg, True, rs)
            saved.append(st)
            i = 2 * l + 1
            last = l == nl - 1
            n32, nb, st = _block_forward(i, False, n32, nb, eb, w8c[i], b8c[i], wc3[i], cvf[i], wbf[i], gmf[i], blf[i],
                                         cfg, (not last) or cfg.want_node_lp, rs)
            saved.append(st)
        ctx.saved, ctx.cfg, ctx.rs = saved, cfg, rs
        ctx.weights = (w8c, wc3, cvf, wbf, gmf, blf)
        ctx.dtypes = (w8.dtype, b8.dtype, wc.dtype, cvec.dtype, wbeta.dtype, gamma.dtype, beta_ln.dtype, w1.dtype, b1.dtype)
        ctx.sizes = (int(node32.size(0)), int(edge32.size(0)), hid)
        return n32

    @staticmethod
    def backward(ctx, dn: Tensor):
        cfg, rs, saved = ctx.cfg, ctx.rs, ctx.saved
        w8c, wc3, cvf, wbf, gmf, blf = ctx.weights
        h, nl = cfg.heads, cfg.n_layers
        n_atoms, n_bonds, hid = ctx.sizes
        c = hid // h
        cd = torch.bfloat16
        dev = dn.device
        f32 = dict(dtype=torch.float32, device=dev)
        nb2 = 2 * nl
        d_w8 = torch.empty(nb2, 8 * hid, hid, **f32)
        d_b8 = torch.empty(nb2, 8 * hid, **f32)
        d_wc = torch.empty(nb2, hid, hid, **f32)
        d_par = torch.empty(nb2, 6 * hid, **f32)       # per block: dw_beta x3 | dgamma | dbias | dcvec
        eye = torch.eye(hid, dtype=cd, device=dev)
        dn = dn.contiguous().float()
        de: Optional[Tensor] = None                    # fp32 gradient of the bond residual stream (None above the top)
        coefs, qts, gts = [], [], []

        def block_backward(idx: int, is_lg: bool, dy: Optional[Tensor], dy2: Optional[Tensor], dbuf: Tensor,
                           df_out: Optional[Tensor]) -> Tensor:
            st = saved[idx]
            sa, oa, so, oo = cfg.keys[idx]
            proj, xb, agg, s, na = st["proj"], st["xb"], st["agg"], st["s"], st["na"]
            n = proj.size(0)
            q, k, v = (proj[:na, i * hid:(i + 1) * hid] for i in range(3))
            xr = proj[:, 7 * hid:]
            qt = proj[:na, 3 * hid:7 * hid].unflatten(1, (h, hid)).transpose(0, 1)
            dproj = dbuf[:, :8 * hid]
            dq, dk, dv = (dbuf[:na, i * hid:(i + 1) * hid] for i in range(3))
            dxr = dbuf[:, 7 * hid:8 * hid]
            bbar = dbuf[:na, 3 * hid:7 * hid].unflatten(1, (h, hid)).transpose(0, 1)
            dagg, dagg_lp, dparams = ops.raw_gate_ln_bwd3(dy, agg, xr, wbf[idx], gmf[idx], blf[idx], st["beta"],
                                                          st["mean"], st["rstd"], s, h, dxr, cfg.p_out[idx], so, oo, rs,
                                                          dy2=dy2, agg_rows=na)
            d_par[idx].copy_(dparams)
            dagg, dagg_lp, agg = dagg[:na], dagg_lp[:na], agg[:na]
            g3 = dagg_lp.unflatten(1, (h, c)).transpose(0, 1)                 # [h, na, C]
            gt = torch.bmm(g3, wc3[idx])                                      # [h, na, H]
            if is_lg:
                coef = ops.raw_lgattn_bwd(dagg, dagg_lp, agg, q, k, v, qt, gt, cvf[idx], cfg.a_csr, cfg.w1, cfg.b1,
                                          st["m"], st["z"], cfg.lg_plan, h, dq, dk, dv, bbar, cfg.p_attn[idx], sa, oa, rs)
                coefs.append(coef); qts.append(qt); gts.append(gt)
            else:
                ops.raw_attn_bwd_s(dagg, dagg_lp, agg, q, k, v, qt, gt, cvf[idx], st["feat"], st["m"], st["z"],
                                   cfg.g_plan, h, dq, dk, dv, bbar, df_out, cfg.p_attn[idx], sa, oa, rs)
            # dWc[t] = dagg_t^T abar_t: diagonal blocks of one dense [H, na] x [na, 4H] product (K = na is what costs)
            gfull = torch.mm(dagg_lp.t(), st["abar_rows"].view(-1, h * hid), out_dtype=torch.float32)   # [H, 4H]
            d_wc[idx].copy_(torch.stack([gfull[t * c:(t + 1) * c, t * hid:(t + 1) * hid] for t in range(h)]).view(hid, hid))
            ext = dbuf.size(1) == 9 * hid
            wext = torch.cat([w8c[idx], eye], dim=0) if ext else w8c[idx]     # [9H | 8H, H]: identity block adds df
            if na == n:
                if dy is not None:
                    dx = torch.addmm(dy, dbuf, wext, out_dtype=torch.float32)
                else:
                    dx = torch.mm(dbuf, wext, out_dtype=torch.float32)
                torch.mm(dproj.t(), xb, out_dtype=torch.float32, out=d_w8[idx])
                d_b8[idx].copy_(ops.colsum(dproj))
            else:
                dx = torch.empty(n, hid, **f32)
                tail, wtail = dbuf[na:, 7 * hid:], wext[7 * hid:]             # isolated rows: dx_r (| df) only
                if dy is not None:
                    torch.addmm(dy[:na], dbuf[:na], wext, out_dtype=torch.float32, out=dx[:na])
                    torch.addmm(dy[na:], tail, wtail, out_dtype=torch.float32, out=dx[na:])
                else:
                    torch.mm(dbuf[:na], wext, out_dtype=torch.float32, out=dx[:na])
                    torch.mm(tail, wtail, out_dtype=torch.float32, out=dx[na:])
                torch.mm(dproj[:na].t(), xb[:na], out_dtype=torch.float32, out=d_w8[idx])
                ws_tail = torch.mm(dxr[na:].t(), xb[na:], out_dtype=torch.float32)
                d_w8[idx, 7 * hid:].add_(ws_tail)
                d_b8[idx].copy_(ops.colsum(dproj[:na]))
                d_b8[idx, 7 * hid:].add_(ops.colsum(dxr[na:]))
            st.clear()
            return dx

        for l in reversed(range(nl)):
            dpe = torch.empty(n_bonds, 9 * hid, dtype=cd, device=dev)         # LG block l: dproj8 | df_l
            dpa = torch.empty(n_atoms, 8 * hid, dtype=cd, device=dev)
            df = dpe[:, 8 * hid:]
            dn = block_backward(2 * l + 1, False, dn, None, dpa, df)
            de = block_backward(2 * l, True, de, df, dpe, None)
        dw1, db1 = ops.raw_lg_angle_grad(cfg.a_csr, cfg.w1, cfg.b1, cfg.lg_plan, coefs, qts, gts)

        t = ctx.dtypes
        return (dn, None, de, None, d_w8.to(t[0]), d_b8.to(t[1]), d_wc.to(t[2]), d_par[:, 5 * hid:].to(t[3]),
                d_par[:, :3 * hid].to(t[4]), d_par[:, 3 * hid:4 * hid].to(t[5]), d_par[:, 4 * hid:5 * hid].to(t[6]),
                dw1.to(t[7]), db1.to(t[8]), None)


def run_trunk(node32: Tensor, node_b: Tensor, edge32: Tensor, edge_b: Tensor, w8: Tensor, b8: Tensor, wc: Tensor,
              cvec: Tensor, wbeta: Tensor, gamma: Tensor, beta_ln: Tensor, w1: Tensor, b1: Tensor, cfg: TrunkCfg) -> Tensor:
    """fp32 ``[N, H]`` atom states after all blocks.  Stacked parameter layouts: ``w8 [2L, 8H, H]``, ``b8 [2L, 8H]``,
    ``wc [2L, H, H]`` (folded edge projections), ``cvec [2L, H]``, ``wbeta [2L, 3H]``, ``gamma / beta_ln [2L, H]``."""
    return _Trunk.apply(node32, node_b, edge32, edge_b, w8, b8, wc, cvec, wbeta, gamma, beta_ln, w1, b1, cfg)
