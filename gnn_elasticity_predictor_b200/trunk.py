"""The whole message-passing trunk (``layers`` x [EdgeUpdateBlock, NodeUpdateBlock], reference ``scripts/train.py:558-560``)
as ONE explicit forward / backward program over the C-ABI kernels -- the bf16 fast path (hidden 256, 4 heads).

Why one program instead of one autograd node per block (``fused.py``): the backward of the trunk has two gradient streams
that autograd would sum with separate elementwise passes -- every bond state ``edge_state_l`` is consumed by the next
line-graph block (residual stream) AND by the atom-graph block of the same layer (as its edge feature).  Here the backward
is scheduled by hand:

    for l = L-1 .. 0:
        NodeUpdateBlock_l backward   -> writes the bond-feature gradient df_l (bf16) into the right half of the [E, 2H]
                                        buffer whose left half will hold the line-graph block's dx_r
        EdgeUpdateBlock_l backward   -> gate/LayerNorm backward reads dy = d_edge (fp32, from block l+1) + df_l on load
                                        and writes dx_r next to df_l; the attention kernels fill dq | dk | dv | bbar_0..3;
                                        d_edge <- d_edge + [dx_r | df_l] . [Ws ; I] + [dq|dk|dv|bbar] . [Wq;Wk;Wv;WQT]
                                        (fp32-accumulating GEMMs)

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

import os
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
    overlap: bool = True                 # run the atom-graph chain on a second stream (fork/join; capturable)
    zero_df: bool = False                # padded batches: bond rows that are no atom-graph edge get no df write
    dp_group: object = None              # data-parallel process group: the gradients of the FOLDED stacked weights are
                                         # all-reduced inside the backward, under the angle-encoder gradient kernel
    dp_done: object = None               # callback() -> None: tells the engine that this backward has reduced them
    defer_state: object = None           # dict the backward marks with {"deferred": True} when it left the side stream unjoined
    defer_angle: int = 0                 # > 0: run the angle-encoder gradient kernel on the side stream with at most this
                                         # many CTAs and do NOT join it here -- the caller (engine.TrainStep) joins the side
                                         # stream after backward(), so the rest of the backward runs beside the kernel


DX_INPLACE = os.environ.get("ALIGNN_DX_INPLACE", "1") == "1"      # dx GEMM accumulates into the incoming gradient buffer
_SIDE_STREAMS = {}
_AUX_STREAMS = {}


def _aux_stream(device: torch.device) -> torch.cuda.Stream:
    """A third stream per device for work that only feeds PARAMETER gradients or is a function of the parameters alone
    (the weight folds at the head of the step; ``dWc``, the weight / bias gradients of the stacked projections in the
    backward): none of it is on the activation-gradient chain, so it fills the holes beside the gate / LayerNorm and GEMM
    launches of the next block instead of lengthening the main stream."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _AUX_STREAMS.get(key)
    if st is None:
        st = _AUX_STREAMS[key] = torch.cuda.Stream(device=device)
    return st


def _side_stream(device: torch.device) -> torch.cuda.Stream:
    """One extra stream per device for the atom-graph chain.  NodeUpdateBlock_l only depends on EdgeUpdateBlock_l and
    NodeUpdateBlock_{l-1}, so it can run beside EdgeUpdateBlock_{l+1} (forward) / its backward beside EdgeUpdateBlock_{l+1}'s
    (backward); under CUDA-graph capture the fork/join events become graph edges."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return st


def _block_forward(idx: int, is_lg: bool, x32: Tensor, xb: Tensor, feat: Optional[Tensor], w8c: Tensor, b8c: Tensor,
                   wc3: Tensor, cv: Tensor, wb: Tensor, gm: Tensor, bl: Tensor, cfg: TrunkCfg, want_lp: bool, rs):
    """One conv block forward on raw kernels.  Returns (y32, y_lp, saved-state dict)."""
    h = cfg.heads
    n, hid = xb.shape                    # x32 is None for the first block of a chain: the residual input is read as xb
    sa, oa, so, oo = cfg.keys[idx]
    na = cfg.lg_active if (is_lg and 0 <= cfg.lg_active < n) else n            # active prefix
    # q | k | v | qt_0..3 over the ACTIVE prefix ([na, 7H]) and x_r over all rows ([n, H]): two contiguous GEMM outputs
    # with the same shapes per row whether or not rows are elided, so elision changes no rounding anywhere.
    xr, proj = ops.block_projections(xb, w8c, b8c, na)       # [n, H], [na, 7H]: ONE tcgen05 + TMA launch (csrc/proj_tc.cu)
    q, k, v = (proj[:, i * hid:(i + 1) * hid] for i in range(3))
    qt = proj[:, 3 * hid:].unflatten(1, (h, hid)).transpose(0, 1)             # [h, na, H] view, row stride 7H
    abar_rows = torch.empty(na, h, hid, dtype=xb.dtype, device=xb.device)     # row-interleaved: [na, 4H] for the dWc GEMM
    abar = abar_rows.transpose(0, 1)                                          # [h, na, H] view
    if is_lg:
        aggv, _, m, z, s = ops.raw_lgattn_fwd(q, k, v, qt, cfg.a_csr, cfg.w1, cfg.b1, cfg.lg_plan, h, cfg.p_attn[idx],
                                              sa, oa, rs, abar=abar)
    else:
        aggv, _, m, z, s = ops.raw_attn_fwd_s(q, k, v, qt, feat, cfg.g_plan, h, cfg.p_attn[idx], sa, oa, rs, abar=abar)
    agge = torch.bmm(abar, wc3.transpose(1, 2))                               # [h, na, C]
    y, y_lp, agg, beta, mean, rstd = ops.raw_gate_ln_fwd2(aggv, agge, cv, s, h, xr, x32, wb, gm, bl, cfg.eps[idx],
                                                          cfg.p_out[idx], so, oo, want_lp, rs, agg_rows=na,
                                                          x_lp=xb if x32 is None else None)
    st = dict(xb=xb, feat=feat, proj=proj, xr=xr, abar_rows=abar_rows, agg=agg, m=m, z=z, s=s, beta=beta, mean=mean,
              rstd=rstd, na=na)
    return y, y_lp, st


class _Trunk(torch.autograd.Function):
    @staticmethod
    def forward(ctx, node32: Tensor, node_b: Tensor, edge32: Tensor, edge_b: Tensor, w8: Tensor, b8: Tensor, wc: Tensor,
                cvec: Tensor, wbeta: Tensor, gamma: Tensor, beta_ln: Tensor, w1: Tensor, b1: Tensor, cfg: TrunkCfg):
        cd = torch.bfloat16
        h, nl = cfg.heads, cfg.n_layers
        hid = node_b.size(1)
        c = hid // h
        rs = ops.RNG_STEP
        w8c, b8c = w8.detach().to(cd), b8.detach().to(cd)                      # [2L, 8H, H], [2L, 8H]
        wc3 = wc.detach().to(cd).view(2 * nl, h, c, hid)
        cvf = cvec.detach().float().contiguous()
        wbf = wbeta.detach().float().contiguous()
        gmf, blf = gamma.detach().float().contiguous(), beta_ln.detach().float().contiguous()
        # node32 / edge32 may be None: the encoder outputs then enter the first blocks in the storage dtype (no fp32 copy)
        n32 = node32.contiguous().float() if node32 is not None else None
        e32 = edge32.contiguous().float() if edge32 is not None else None
        nb, eb = node_b.contiguous(), edge_b.contiguous()
        if nb.dtype != cd:
            nb = nb.to(cd)
        if eb.dtype != cd:
            eb = eb.to(cd)
        saved = [None] * (2 * nl)
        main = torch.cuda.current_stream()
        side = _side_stream(node_b.device) if cfg.overlap else None
        if side is not None:
            side.wait_stream(main)
            for t in (n32, nb):
                if t is not None:
                    t.record_stream(side)
        for l in range(nl):
            i = 2 * l
            e32, eb, saved[i] = _block_forward(i, True, e32, eb, None, w8c[i], b8c[i], wc3[i], cvf[i], wbf[i], gmf[i],
                                               blf[i], cfg, True, rs)
            i = 2 * l + 1
            want_lp = (l < nl - 1) or cfg.want_node_lp
            if side is None:
                n32, nb, saved[i] = _block_forward(i, False, n32, nb, eb, w8c[i], b8c[i], wc3[i], cvf[i], wbf[i], gmf[i],
                                                   blf[i], cfg, want_lp, rs)
            else:
                ev = torch.cuda.Event()
                ev.record(main)
                eb.record_stream(side)
                with torch.cuda.stream(side):
                    side.wait_event(ev)
                    n32, nb, saved[i] = _block_forward(i, False, n32, nb, eb, w8c[i], b8c[i], wc3[i], cvf[i], wbf[i],
                                                       gmf[i], blf[i], cfg, want_lp, rs)
        if side is not None:
            main.wait_stream(side)
            n32.record_stream(main)
        ctx.saved, ctx.cfg, ctx.rs = saved, cfg, rs
        ctx.weights = (w8c, wc3, cvf, wbf, gmf, blf)
        ctx.dtypes = (w8.dtype, b8.dtype, wc.dtype, cvec.dtype, wbeta.dtype, gamma.dtype, beta_ln.dtype, w1.dtype, b1.dtype)
        ctx.sizes = (int(node_b.size(0)), int(edge_b.size(0)), hid)
        ctx.lp_inputs = (node32 is None, edge32 is None)
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
        # the gradients of the folded, stacked weights live in ONE flat buffer: under data parallelism it is all-reduced
        # here, as soon as the last block has written it (the fold backward that turns it into parameter gradients is
        # linear, so reducing before or after it is the same sum) -- the collective then runs beside the angle-encoder
        # gradient kernel instead of after the whole backward
        sizes = (nb2 * 8 * hid * hid, nb2 * 8 * hid, nb2 * hid * hid, nb2 * 6 * hid)
        gflat = torch.empty(sum(sizes), **f32)
        o1, o2, o3 = sizes[0], sizes[0] + sizes[1], sizes[0] + sizes[1] + sizes[2]
        d_w8 = gflat[:o1].view(nb2, 8 * hid, hid)
        d_b8 = gflat[o1:o2].view(nb2, 8 * hid)
        d_wc = gflat[o2:o3].view(nb2, hid, hid)
        d_par = gflat[o3:].view(nb2, 6 * hid)          # per block: dw_beta x3 | dgamma | dbias | dcvec
        # [Ws ; I] of every block at once (the identity block adds the atom-graph feature gradient inside the dx GEMM)
        wtail_all = torch.cat([w8c[:, 7 * hid:], torch.eye(hid, dtype=cd, device=dev).expand(nb2, hid, hid)], dim=1)
        dn = dn.contiguous().float()
        de: Optional[Tensor] = None                    # fp32 gradient of the bond residual stream (None above the top)
        coefs, qts, gts = [], [], []

        def block_backward(idx: int, is_lg: bool, dy: Optional[Tensor], dtail: Tensor, df_out: Optional[Tensor],
                           own_dy: bool = False) -> Tensor:
            """``dtail``: [n, H] (dx_r) or [n, 2H] (dx_r | df from the atom-graph block of the same layer, already written);
            ``df_out``: where THIS block writes its edge-feature gradient (atom-graph blocks)."""
            st = saved[idx]
            sa, oa, so, oo = cfg.keys[idx]
            proj, xr, xb, agg, s, na = st["proj"], st["xr"], st["xb"], st["agg"], st["s"], st["na"]
            n = xr.size(0)
            q, k, v = (proj[:, i * hid:(i + 1) * hid] for i in range(3))
            qt = proj[:, 3 * hid:].unflatten(1, (h, hid)).transpose(0, 1)
            dbuf = torch.empty(na, 7 * hid, dtype=cd, device=dev)             # dq | dk | dv | bbar_0..3 (active prefix)
            dq, dk, dv = (dbuf[:, i * hid:(i + 1) * hid] for i in range(3))
            bbar = dbuf[:, 3 * hid:].unflatten(1, (h, hid)).transpose(0, 1)
            dxr = dtail[:, :hid]
            dy2 = dtail[:, hid:] if dtail.size(1) == 2 * hid else None
            dagg, dagg_lp, dparams = ops.raw_gate_ln_bwd3(dy, agg, xr, wbf[idx], gmf[idx], blf[idx], st["beta"],
                                                          st["mean"], st["rstd"], s, h, dxr, cfg.p_out[idx], so, oo, rs,
                                                          dy2=dy2, agg_rows=na, dparams=d_par[idx])
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
            def param_grads():
                # dWc[t] = dagg_t^T abar_t: one batched [C, na] x [na, H] product per head, written straight into its slot
                torch.bmm(dagg_lp.view(na, h, c).permute(1, 2, 0), st["abar_rows"].transpose(0, 1), out_dtype=torch.float32,
                          out=d_wc[idx].view(h, c, hid))
                # weight + bias gradients of the stacked projection: one streaming pass per operand pair (tcgen05, wgrad_tc.cu)
                ops.wgrad(dbuf, xb[:na], d_w8[idx, :7 * hid], d_b8[idx, :7 * hid])
                ops.wgrad(dxr, xb, d_w8[idx, 7 * hid:], d_b8[idx, 7 * hid:])

            cur = torch.cuda.current_stream()
            if aux is not None and is_lg:
                # parameter gradients leave the activation-gradient chain: third stream, joined once at the very end.  Every
                # operand stays referenced until that join (``keep``), so no block is recycled under the aux stream.
                ev = torch.cuda.Event()
                ev.record(cur)
                keep.append((dbuf, dagg_lp, st["abar_rows"], xb, dtail))
                with torch.cuda.stream(aux):
                    aux.wait_event(ev)
                    param_grads()
            else:
                param_grads()
            # dx = dy + [dx_r | df] [Ws ; I]  (all rows; the identity block adds df)  +  dbuf [Wq; Wk; Wv; WQT]  (active rows)
            w7, ws = w8c[idx, :7 * hid], w8c[idx, 7 * hid:]
            wtail = wtail_all[idx] if dy2 is not None else ws
            if dy is not None and own_dy and DX_INPLACE:
                dx = torch.addmm(dy, dtail, wtail, out_dtype=torch.float32, out=dy)     # in place: dy is dead after this block
            elif dy is not None:
                dx = torch.addmm(dy, dtail, wtail, out_dtype=torch.float32)
            else:
                dx = torch.mm(dtail, wtail, out_dtype=torch.float32)
            torch.addmm(dx[:na], dbuf, w7, out_dtype=torch.float32, out=dx[:na])
            st.clear()
            return dx

        main = torch.cuda.current_stream()
        side = _side_stream(dev) if cfg.overlap else None
        aux = _aux_stream(dev) if cfg.overlap else None
        keep: list = []
        if aux is not None:
            aux.wait_stream(main)
            gflat.record_stream(aux)
        mk = torch.zeros if cfg.zero_df else torch.empty
        tails = [mk(n_bonds, 2 * hid, dtype=cd, device=dev) for _ in range(nl)]            # LG block l: dx_r | df_l
        if side is not None:
            side.wait_stream(main)
            for t in tails + [dn, gflat, wtail_all]:
                t.record_stream(side)
        for l in reversed(range(nl)):
            # the incoming dn of the top layer belongs to autograd (never written in place); below it dn / de are this
            # function's own buffers and the dx GEMM accumulates straight into them
            own = l < nl - 1
            if side is None:
                dn = block_backward(2 * l + 1, False, dn, torch.empty(n_atoms, hid, dtype=cd, device=dev), tails[l][:, hid:],
                                    own_dy=own)
            else:
                with torch.cuda.stream(side):
                    dn = block_backward(2 * l + 1, False, dn, torch.empty(n_atoms, hid, dtype=cd, device=dev),
                                        tails[l][:, hid:], own_dy=own)
                    ev = torch.cuda.Event()
                    ev.record(side)
                main.wait_event(ev)
            de = block_backward(2 * l, True, de, tails[l], None, own_dy=own)
        if side is not None:
            main.wait_stream(side)
            dn.record_stream(main)
        if cfg.dp_group is not None:
            import torch.distributed as dist
            if aux is not None:
                ev = torch.cuda.Event()
                ev.record(main)                       # every block of both chains has written its slice
                with torch.cuda.stream(aux):          # ... and the aux stream its weight / bias gradients
                    aux.wait_event(ev)
                    dist.all_reduce(gflat, op=dist.ReduceOp.SUM, group=cfg.dp_group)
            else:
                dist.all_reduce(gflat, op=dist.ReduceOp.SUM, group=cfg.dp_group)
            if cfg.dp_done is not None:
                cfg.dp_done()
        if cfg.defer_angle > 0 and side is not None and cfg.defer_state is not None:
            # the kernel is persistent and would hold every SM: capped to `defer_angle` CTAs on the side stream it leaves
            # the other SMs to the encoder / fold backward and the optimizer prologue that follow on this stream
            ev = torch.cuda.Event()
            ev.record(main)
            for t in coefs + qts + gts + [cfg.a_csr, cfg.w1, cfg.b1]:
                t.record_stream(side)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                dw1, db1 = ops.raw_lg_angle_grad(cfg.a_csr, cfg.w1, cfg.b1, cfg.lg_plan, coefs, qts, gts,
                                                 max_blocks=cfg.defer_angle)
            dw1.record_stream(main)
            db1.record_stream(main)
            if cfg.defer_state is not None:
                cfg.defer_state["deferred"] = True
        else:
            dw1, db1 = ops.raw_lg_angle_grad(cfg.a_csr, cfg.w1, cfg.b1, cfg.lg_plan, coefs, qts, gts)
        if aux is not None:
            main.wait_stream(aux)
        keep.clear()

        t = ctx.dtypes
        lp_n, lp_e = ctx.lp_inputs            # gradients go to whichever copy of the encoder output was the input
        return (None if lp_n else dn, dn if lp_n else None, None if lp_e else de, de if lp_e else None, d_w8.to(t[0]), d_b8.to(t[1]), d_wc.to(t[2]), d_par[:, 5 * hid:].to(t[3]),
                d_par[:, :3 * hid].to(t[4]), d_par[:, 3 * hid:4 * hid].to(t[5]), d_par[:, 4 * hid:5 * hid].to(t[6]),
                dw1.to(t[7]), db1.to(t[8]), None)


def run_trunk(node32: Tensor, node_b: Tensor, edge32: Tensor, edge_b: Tensor, w8: Tensor, b8: Tensor, wc: Tensor,
              cvec: Tensor, wbeta: Tensor, gamma: Tensor, beta_ln: Tensor, w1: Tensor, b1: Tensor, cfg: TrunkCfg) -> Tensor:
    """fp32 ``[N, H]`` atom states after all blocks.  Stacked parameter layouts: ``w8 [2L, 8H, H]``, ``b8 [2L, 8H]``,
    ``wc [2L, H, H]`` (folded edge projections), ``cvec [2L, H]``, ``wbeta [2L, 3H]``, ``gamma / beta_ln [2L, H]``."""
    return _Trunk.apply(node32, node_b, edge32, edge_b, w8, b8, wc, cvec, wbeta, gamma, beta_ln, w1, b1, cfg)
