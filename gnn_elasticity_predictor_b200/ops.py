"""Autograd operators over the C ABI (``include/alignn_b200.h``).

PyTorch is plumbing here: it owns device memory and the stream; every operator below is a thin
``torch.autograd.Function`` that hands raw device pointers to ``libalignn_b200.so``.  No operator
has a CPU, eager-PyTorch or Triton fallback -- non-CUDA inputs raise ``RuntimeError``.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Optional, Tuple

import torch
from torch import Tensor

from . import _lib

_DT = {torch.float32: _lib.F32, torch.bfloat16: _lib.BF16}


def _p(t: Optional[Tensor]):
    if t is None or t.numel() == 0:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(*tensors: Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "gnn_elasticity_predictor_b200 runs on CUDA (sm_100a) only: got a "
                f"{t.device.type} tensor and there is no CPU fallback path")


def _dtype_code(t: Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise RuntimeError(f"unsupported operand dtype {t.dtype}: use float32 or bfloat16") from None


# --------------------------------------------------------------------------------------------------
# launch accounting + optional CUDA-event profiler (used by bench.py; zero cost when disabled)
# --------------------------------------------------------------------------------------------------
class LaunchStats:
    """Counts kernels launched through the C ABI and, when ``events`` is on, brackets every call with
    CUDA events on the launching stream so that per-kernel durations can be read after a sync.

    Inside a CUDA-graph capture the calls are bracketed with EXTERNAL timing events instead (``graph_events`` on;
    names in ``graph_filter`` if set): they become event-record nodes of the graph, are re-recorded by every replay,
    and ``graph_durations_ms()`` reads the most recent replay's per-kernel durations."""

    def __init__(self):
        self.kernels = 0
        self.calls = 0
        self.events = False
        self.records = []   # (name, meta, start_event, end_event)
        self.graph_events = False
        self.graph_filter = None
        self.graph_records = []

    def reset(self):
        self.kernels = self.calls = 0
        self.records = []

    @staticmethod
    def _durations(records):
        out = {}
        for name, meta, a, b in records:
            out.setdefault(name, []).append((a.elapsed_time(b), meta))
        return out

    def durations_ms(self):
        """name -> list of (ms, meta); call after torch.cuda.synchronize()."""
        return self._durations(self.records)

    def graph_durations_ms(self):
        """Same for the event nodes captured into CUDA graphs (last replay of each graph)."""
        return self._durations(self.graph_records)


STATS = LaunchStats()


class _Launch:
    __slots__ = ("name", "n", "meta", "start", "ext")

    def __init__(self, name: str, n_kernels: int, meta=None):
        self.name, self.n, self.meta, self.start, self.ext = name, n_kernels, meta, None, False

    def __enter__(self):
        STATS.kernels += self.n
        STATS.calls += 1
        if STATS.events or STATS.graph_events:
            capturing = torch.cuda.is_current_stream_capturing()
            if capturing:
                if STATS.graph_events and (STATS.graph_filter is None or self.name in STATS.graph_filter):
                    self.ext = True
                    self.start = torch.cuda.Event(enable_timing=True, external=True)
                    self.start.record()
            elif STATS.events:
                self.start = torch.cuda.Event(enable_timing=True)
                self.start.record()
        return self

    def __exit__(self, *exc):
        if self.start is not None:
            end = torch.cuda.Event(enable_timing=True, external=self.ext)
            end.record()
            (STATS.graph_records if self.ext else STATS.records).append((self.name, self.meta, self.start, end))
        return False


def _plan_kernels(n_edges: int, n_nodes: int, source_sorted: bool = False, key_bound: int = -1) -> int:
    if n_edges == 0:
        return 0
    top = n_nodes if key_bound < 0 or key_bound > n_nodes else key_bound
    passes = (max(int(top), 1).bit_length() + 7) // 8
    return 1 + (1 if source_sorted else 2) * 4 * passes + 4


class StaticDropoutUnderCapture(RuntimeError):
    """A dropout mask keyed by HOST integers only is about to be baked into a CUDA graph (every replay would reuse it)."""


def _static_dropout_guard(p_drop: float, what: str) -> None:
    """The generic kernels (hidden != 256, fp32 regime, standalone blocks) take the Philox ``(seed, offset)`` as host
    integers and do not read the device step counter ``RNG_STEP``.  Capturing them with ``p > 0`` would freeze the mask
    across replays, so refuse: ``engine.TrainStep`` catches this and runs such models eagerly (fresh keys per step)."""
    if p_drop > 0.0 and RNG_STEP is not None and torch.cuda.is_current_stream_capturing():
        raise StaticDropoutUnderCapture(
            f"{what}: dropout p={p_drop} on a kernel family without a device-side RNG counter cannot be captured into a "
            "CUDA graph (replays would reuse one mask); run this model eagerly (TrainStep(graph=False))")


def next_dropout_key() -> Tuple[int, int]:
    """(seed, offset) for one dropout mask, drawn from torch's CPU generator (no device sync), so
    ``torch.manual_seed`` makes training runs reproducible."""
    r = torch.empty(2, dtype=torch.int64).random_()
    return int(r[0]) & 0x7FFFFFFFFFFFFFFF, int(r[1]) & 0x7FFFFFFFFFFFFFFF


# --------------------------------------------------------------------------------------------------
# graph plan
# --------------------------------------------------------------------------------------------------
@dataclass
class GraphPlan:
    """CSR (by target) + CSC (by source) views of one edge list; built once per batch."""
    rowptr: Tensor
    col: Tensor
    eid: Tensor
    rowptr_t: Tensor
    col_t: Tensor
    eid_t: Tensor
    status: Tensor
    n_nodes: int
    n_edges: int

    def check(self) -> None:
        """Synchronising validation: raises if any edge referenced a node outside ``[0, n_nodes)`` or a
        ``source_sorted`` hint was wrong."""
        st = int(self.status.item())
        if st & 4:
            raise ValueError("build_plan(key_bound=...): an index at or beyond the stated bound")
        if st & 2:
            raise ValueError("build_plan(source_sorted=True): edge_index[0] is not non-decreasing")
        if st & 1:
            raise IndexError(f"edge_index contains indices outside [0, {self.n_nodes})")


def build_plan(edge_index: Tensor, n_nodes: int, validate: bool = False, source_sorted: bool = False,
               key_bound: int = -1) -> GraphPlan:
    """Stable target-sort / source-sort of ``edge_index`` (int64 ``[2, E]``) on the device.  ``source_sorted``: the
    caller knows ``edge_index[0]`` is non-decreasing (``GraphBatch.source_sorted``), so the source sort is skipped; the
    device verifies it (``plan.check()``).  ``key_bound``: the caller knows every index is below it (e.g.
    ``GraphBatch.lg_active_rows``): fewer radix passes; verified on the device as well."""
    _require_cuda(edge_index)
    if edge_index.dtype != torch.int64 or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise RuntimeError("edge_index must be an int64 tensor of shape [2, E]")
    lib = _lib.load()
    ei = edge_index.contiguous()
    n_edges = int(ei.size(1))
    n_nodes = int(n_nodes)
    dev = ei.device
    i32 = dict(dtype=torch.int32, device=dev)
    rowptr = torch.empty(n_nodes + 1, **i32)
    rowptr_t = torch.empty(n_nodes + 1, **i32)
    col, eid, col_t, eid_t = (torch.empty(max(n_edges, 1), **i32) for _ in range(4))
    status = torch.empty(1, **i32)
    ws_bytes = int(lib.alignn_plan_workspace_bytes(n_edges, n_nodes))
    ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev), _Launch("build_plan", _plan_kernels(n_edges, n_nodes, source_sorted, key_bound), (n_nodes, n_edges)):
        rc = lib.alignn_build_plan_bounded(_p(ei), n_edges, n_nodes, int(key_bound), _p(rowptr), _p(col), _p(eid),
                                           _p(rowptr_t), _p(col_t), _p(eid_t), _p(status), _p(ws), ws_bytes,
                                           1 if source_sorted else 0, _stream())
    _lib.check(rc, "alignn_build_plan_bounded")
    plan = GraphPlan(rowptr, col[:n_edges], eid[:n_edges], rowptr_t, col_t[:n_edges], eid_t[:n_edges], status,
                     n_nodes, n_edges)
    if validate:
        plan.check()
    return plan


# --------------------------------------------------------------------------------------------------
# The generic operators as TORCH CUSTOM OPS (namespace ``alignn_b200``; SURVEY.md 8(b) "native boundary").
#
# Each op is a thin shim over one ``extern "C"`` launcher of ``libalignn_b200.so``: torch allocates every tensor, the C
# side gets raw pointers + sizes + the current stream.  Registered with ``torch.library.custom_op`` (schema inferred
# from the annotations), a fake (meta) implementation for tracing / ``torch.compile`` / FakeTensor shape propagation,
# and ``register_autograd`` formulas that call the matching ``*_backward`` op -- so ``torch.ops.alignn_b200.*`` are
# first-class dispatcher citizens (``torch.library.opcheck`` passes on them, tests/test_gpu_ops.py) instead of opaque
# ctypes calls inside ``autograd.Function``s.  No op has a CPU kernel: CPU tensors raise ``RuntimeError``.
# --------------------------------------------------------------------------------------------------
_lib_def = torch.library.custom_op


@_lib_def("alignn_b200::conv_core", mutates_args=(), device_types="cuda")
def _op_conv_core(q: Tensor, k: Tensor, v: Tensor, e: Tensor, rowptr: Tensor, col: Tensor, eid: Tensor,
                  rowptr_t: Tensor, col_t: Tensor, eid_t: Tensor, heads: int, p_drop: float, seed: int, offset: int
                  ) -> Tuple[Tensor, Tensor, Tensor]:
    """(agg, stat_m, stat_z).  ``rowptr_t / col_t / eid_t`` (the CSC half of the plan) are only read by the backward."""
    _require_cuda(q, k, v, e)
    _static_dropout_guard(p_drop, "conv_core")
    lib = _lib.load()
    q, k, v, e = (t.contiguous() for t in (q, k, v, e))
    if not (q.dtype == k.dtype == v.dtype == e.dtype):
        raise RuntimeError("q, k, v, e must share one dtype")
    n_nodes, hidden = q.shape
    n_edges = int(e.size(0))
    f32 = dict(dtype=torch.float32, device=q.device)
    agg = torch.empty(n_nodes, hidden, **f32)
    stat_m = torch.empty(n_nodes, heads, **f32)
    stat_z = torch.empty(n_nodes, heads, **f32)
    with torch.cuda.device(q.device), _Launch("conv_fwd", 1, (n_nodes, n_edges, hidden, heads, q.element_size())):
        rc = lib.alignn_conv_fwd(_p(q), _p(k), _p(v), _p(e), _p(rowptr), _p(col), _p(eid), _p(agg), _p(stat_m),
                                 _p(stat_z), n_nodes, n_edges, hidden, heads, _dtype_code(q), float(p_drop), seed, offset,
                                 _stream())
    _lib.check(rc, "alignn_conv_fwd")
    return agg, stat_m, stat_z


@_op_conv_core.register_fake
def _(q, k, v, e, rowptr, col, eid, rowptr_t, col_t, eid_t, heads, p_drop, seed, offset):
    f32 = dict(dtype=torch.float32, device=q.device)
    return (torch.empty(q.shape, **f32), torch.empty(q.size(0), heads, **f32), torch.empty(q.size(0), heads, **f32))


@_lib_def("alignn_b200::conv_core_backward", mutates_args=(), device_types="cuda")
def _op_conv_core_backward(dagg: Tensor, agg: Tensor, q: Tensor, k: Tensor, v: Tensor, e: Tensor, stat_m: Tensor,
                           stat_z: Tensor, rowptr: Tensor, col: Tensor, eid: Tensor, rowptr_t: Tensor, col_t: Tensor,
                           eid_t: Tensor, heads: int, p_drop: float, seed: int, offset: int
                           ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    lib = _lib.load()
    n_nodes, hidden = q.shape
    n_edges = int(e.size(0))
    dagg = dagg.contiguous().float()
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    de = torch.zeros_like(e)                 # rows of edges the plan dropped (padding) are never written
    coef = torch.empty(max(n_edges, 1), 2 * heads, dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device), _Launch("conv_bwd", 2, (n_nodes, n_edges, hidden, heads, q.element_size())):
        rc = lib.alignn_conv_bwd(_p(dagg), _p(agg), _p(q), _p(k), _p(v), _p(e), _p(stat_m), _p(stat_z), _p(rowptr),
                                 _p(col), _p(eid), _p(rowptr_t), _p(col_t), _p(eid_t), _p(dq), _p(dk), _p(dv), _p(de),
                                 _p(coef), n_nodes, n_edges, hidden, heads, _dtype_code(q), float(p_drop), seed, offset,
                                 _stream())
    _lib.check(rc, "alignn_conv_bwd")
    return dq, dk, dv, de


@_op_conv_core_backward.register_fake
def _(dagg, agg, q, k, v, e, stat_m, stat_z, rowptr, col, eid, rowptr_t, col_t, eid_t, heads, p_drop, seed, offset):
    return torch.empty_like(q), torch.empty_like(k), torch.empty_like(v), torch.empty_like(e)


def _conv_core_setup(ctx, inputs, output):
    q, k, v, e, rowptr, col, eid, rowptr_t, col_t, eid_t, heads, p_drop, seed, offset = inputs
    agg, stat_m, stat_z = output
    ctx.save_for_backward(q.contiguous(), k.contiguous(), v.contiguous(), e.contiguous(), agg, stat_m, stat_z, rowptr,
                          col, eid, rowptr_t, col_t, eid_t)
    ctx.args = (heads, float(p_drop), seed, offset)


def _conv_core_backward(ctx, dagg, _dm, _dz):
    q, k, v, e, agg, stat_m, stat_z, rowptr, col, eid, rowptr_t, col_t, eid_t = ctx.saved_tensors
    heads, p_drop, seed, offset = ctx.args
    dq, dk, dv, de = torch.ops.alignn_b200.conv_core_backward(dagg, agg, q, k, v, e, stat_m, stat_z, rowptr, col, eid,
                                                              rowptr_t, col_t, eid_t, heads, p_drop, seed, offset)
    return (dq, dk, dv, de) + (None,) * 10


_op_conv_core.register_autograd(_conv_core_backward, setup_context=_conv_core_setup)


def conv_core(q: Tensor, k: Tensor, v: Tensor, e: Tensor, plan: GraphPlan, heads: int, p_drop: float = 0.0,
              seed: int = 0, offset: int = 0) -> Tensor:
    """``agg[i] = sum_j softmax_i(<q_i, k_j + e_ij>/sqrt(C)) (v_j + e_ij)`` per head; returns fp32 ``[N, H]``
    (custom op ``alignn_b200::conv_core``)."""
    _require_cuda(q, k, v, e)
    n_nodes, n_edges = int(q.size(0)), int(e.size(0))
    if n_nodes != plan.n_nodes or n_edges != plan.n_edges:
        raise RuntimeError(f"plan is for {plan.n_nodes} nodes / {plan.n_edges} edges, operands have "
                           f"{n_nodes} / {n_edges}")
    return torch.ops.alignn_b200.conv_core(q, k, v, e, plan.rowptr, plan.col, plan.eid, plan.rowptr_t, plan.col_t,
                                           plan.eid_t, int(heads), float(p_drop), int(seed), int(offset))[0]


# --------------------------------------------------------------------------------------------------
# beta gate + LayerNorm + ReLU + dropout + residual
# --------------------------------------------------------------------------------------------------
@_lib_def("alignn_b200::gate_ln", mutates_args=(), device_types="cuda")
def _op_gate_ln(agg: Tensor, xr: Tensor, x: Tensor, wbeta: Tensor, gamma: Tensor, bias: Tensor, eps: float,
                p_drop: float, seed: int, offset: int, want_lp: bool
                ) -> Tuple[Tensor, Tensor, Tensor, Tensor, Tensor]:
    """Returns (y f32, y_lp (empty when not wanted), beta, mean, rstd)."""
    _require_cuda(agg, xr, x, wbeta, gamma, bias)
    _static_dropout_guard(p_drop, "gate_ln_relu_residual")
    lib = _lib.load()
    agg = agg.contiguous().float()
    xr = xr.contiguous()
    x = x.contiguous().float()
    wb = wbeta.detach().reshape(-1).contiguous().float()
    gm, bs = gamma.detach().contiguous().float(), bias.detach().contiguous().float()
    n_rows, hidden = agg.shape
    f32 = dict(dtype=torch.float32, device=agg.device)
    y = torch.empty(n_rows, hidden, **f32)
    y_lp = torch.empty((n_rows, hidden) if want_lp else (0,), dtype=xr.dtype, device=agg.device)
    beta, mean, rstd = (torch.empty(n_rows, **f32) for _ in range(3))
    with torch.cuda.device(agg.device), _Launch("gate_ln_fwd", 1, (n_rows, hidden, xr.element_size())):
        rc = lib.alignn_gate_ln_fwd(_p(agg), _p(xr), _p(x), _p(wb), _p(gm), _p(bs), _p(y), _p(y_lp) if want_lp else None,
                                    _p(beta), _p(mean), _p(rstd), n_rows, hidden, _dtype_code(xr), float(eps),
                                    float(p_drop), seed, offset, _stream())
    _lib.check(rc, "alignn_gate_ln_fwd")
    return y, y_lp, beta, mean, rstd


@_op_gate_ln.register_fake
def _(agg, xr, x, wbeta, gamma, bias, eps, p_drop, seed, offset, want_lp):
    n_rows, hidden = agg.shape
    f32 = dict(dtype=torch.float32, device=agg.device)
    return (torch.empty(n_rows, hidden, **f32),
            torch.empty((n_rows, hidden) if want_lp else (0,), dtype=xr.dtype, device=agg.device),
            torch.empty(n_rows, **f32), torch.empty(n_rows, **f32), torch.empty(n_rows, **f32))


@_lib_def("alignn_b200::gate_ln_backward", mutates_args=(), device_types="cuda")
def _op_gate_ln_backward(dy: Tensor, agg: Tensor, xr: Tensor, wbeta: Tensor, gamma: Tensor, bias: Tensor, beta: Tensor,
                         mean: Tensor, rstd: Tensor, p_drop: float, seed: int, offset: int
                         ) -> Tuple[Tensor, Tensor, Tensor]:
    """Returns (dagg f32, dxr, dparams f32 [5 * hidden] = dw_beta x3 | dgamma | dbias)."""
    lib = _lib.load()
    agg = agg.contiguous().float()
    xr = xr.contiguous()
    wb = wbeta.detach().reshape(-1).contiguous().float()
    gm, bs = gamma.detach().contiguous().float(), bias.detach().contiguous().float()
    n_rows, hidden = agg.shape
    dy = dy.contiguous().float()
    f32 = dict(dtype=torch.float32, device=agg.device)
    dagg = torch.empty(n_rows, hidden, **f32)
    dxr = torch.empty_like(xr)
    partials = torch.empty(int(lib.alignn_gate_ln_bwd_partial_rows()) * 5 * hidden, **f32)
    dparams = torch.empty(5 * hidden, **f32)
    with torch.cuda.device(agg.device), _Launch("gate_ln_bwd", 2, (n_rows, hidden, xr.element_size())):
        rc = lib.alignn_gate_ln_bwd(_p(dy), _p(agg), _p(xr), _p(wb), _p(gm), _p(bs), _p(beta), _p(mean), _p(rstd),
                                    _p(dagg), _p(dxr), _p(partials), _p(dparams), n_rows, hidden, _dtype_code(xr),
                                    float(p_drop), seed, offset, _stream())
    _lib.check(rc, "alignn_gate_ln_bwd")
    return dagg, dxr, dparams


@_op_gate_ln_backward.register_fake
def _(dy, agg, xr, wbeta, gamma, bias, beta, mean, rstd, p_drop, seed, offset):
    return (torch.empty(agg.shape, dtype=torch.float32, device=agg.device), torch.empty_like(xr),
            torch.empty(5 * agg.size(1), dtype=torch.float32, device=agg.device))


def _gate_ln_setup(ctx, inputs, output):
    agg, xr, x, wbeta, gamma, bias, eps, p_drop, seed, offset, want_lp = inputs
    _, _, beta, mean, rstd = output
    ctx.save_for_backward(agg, xr, wbeta, gamma, bias, beta, mean, rstd)
    ctx.args = (float(p_drop), seed, offset)


def _gate_ln_backward(ctx, dy, dy_lp, _db, _dm, _dr):
    agg, xr, wbeta, gamma, bias, beta, mean, rstd = ctx.saved_tensors
    p_drop, seed, offset = ctx.args
    hidden = agg.size(1)
    if dy is None:
        dy = torch.zeros(agg.shape, dtype=torch.float32, device=agg.device)
    dy = dy.float()
    if dy_lp is not None and dy_lp.numel() > 0:
        dy = dy + dy_lp.float()
    dagg, dxr, dparams = torch.ops.alignn_b200.gate_ln_backward(dy, agg, xr, wbeta, gamma, bias, beta, mean, rstd,
                                                                p_drop, seed, offset)
    return (dagg.to(agg.dtype), dxr, dy, dparams[:3 * hidden].reshape(wbeta.shape).to(wbeta.dtype),
            dparams[3 * hidden:4 * hidden].to(gamma.dtype), dparams[4 * hidden:].to(bias.dtype), None, None, None, None,
            None)


_op_gate_ln.register_autograd(_gate_ln_backward, setup_context=_gate_ln_setup)


def gate_ln_relu_residual(agg: Tensor, xr: Tensor, x: Tensor, wbeta: Tensor, gamma: Tensor, bias: Tensor,
                          eps: float = 1e-5, p_drop: float = 0.0, seed: int = 0, offset: int = 0,
                          want_lp: bool = False):
    """``x + dropout(relu(LayerNorm(beta*xr + (1-beta)*agg)))`` with ``beta = sigmoid(wbeta . [agg, xr, agg-xr])``
    (custom op ``alignn_b200::gate_ln``).  Returns ``(y_fp32, y_lowprecision_or_None)``."""
    y, y_lp, _, _, _ = torch.ops.alignn_b200.gate_ln(agg, xr, x, wbeta, gamma, bias, float(eps), float(p_drop),
                                                     int(seed), int(offset), bool(want_lp))
    return y, (y_lp if want_lp else None)


# --------------------------------------------------------------------------------------------------
# per-graph mean pooling
# --------------------------------------------------------------------------------------------------
@_lib_def("alignn_b200::segment_mean", mutates_args=(), device_types="cuda")
def _op_segment_mean(x: Tensor, rowptr: Tensor, eid: Tensor, n_graphs: int) -> Tensor:
    _require_cuda(x)
    lib = _lib.load()
    x = x.contiguous().float()
    hidden = int(x.size(1))
    pooled = torch.empty(n_graphs, hidden, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device), _Launch("segment_mean_fwd", 1):
        rc = lib.alignn_segment_mean_fwd(_p(x), _p(rowptr), _p(eid), _p(pooled), n_graphs, hidden, _stream())
    _lib.check(rc, "alignn_segment_mean_fwd")
    return pooled


@_op_segment_mean.register_fake
def _(x, rowptr, eid, n_graphs):
    return torch.empty(n_graphs, x.size(1), dtype=torch.float32, device=x.device)


@_lib_def("alignn_b200::segment_mean_backward", mutates_args=(), device_types="cuda")
def _op_segment_mean_backward(dpooled: Tensor, rowptr: Tensor, eid: Tensor, n_rows: int) -> Tensor:
    lib = _lib.load()
    dpooled = dpooled.contiguous().float()
    n_graphs, hidden = dpooled.shape
    dx = torch.zeros(n_rows, hidden, dtype=torch.float32, device=dpooled.device)
    with torch.cuda.device(dpooled.device), _Launch("segment_mean_bwd", 1):
        rc = lib.alignn_segment_mean_bwd(_p(dpooled), _p(rowptr), _p(eid), _p(dx), n_graphs, hidden, _stream())
    _lib.check(rc, "alignn_segment_mean_bwd")
    return dx


@_op_segment_mean_backward.register_fake
def _(dpooled, rowptr, eid, n_rows):
    return torch.empty(n_rows, dpooled.size(1), dtype=torch.float32, device=dpooled.device)


def _segment_mean_setup(ctx, inputs, output):
    x, rowptr, eid, n_graphs = inputs
    ctx.save_for_backward(rowptr, eid)
    ctx.n_rows, ctx.x_dtype = int(x.size(0)), x.dtype


def _segment_mean_backward(ctx, dpooled):
    rowptr, eid = ctx.saved_tensors
    dx = torch.ops.alignn_b200.segment_mean_backward(dpooled, rowptr, eid, ctx.n_rows)
    return dx.to(ctx.x_dtype), None, None, None


_op_segment_mean.register_autograd(_segment_mean_backward, setup_context=_segment_mean_setup)


def build_pool_plan(batch: Tensor, num_graphs: int) -> GraphPlan:
    """Plan over the ``batch`` vector (key = graph id) for :func:`segment_mean`."""
    _require_cuda(batch)
    n = int(batch.numel())
    idx = torch.stack([torch.arange(n, device=batch.device, dtype=torch.int64), batch.to(torch.int64)])
    # keys (row 1) live in [0, num_graphs); the "source" row holds node ids in [0, n): plan over max(n, B)
    plan = build_plan(idx, max(n, int(num_graphs)), source_sorted=True)      # row 0 is arange: its sort is the identity
    plan.n_nodes = int(num_graphs)
    return plan


def segment_mean(x: Tensor, pool_plan: GraphPlan) -> Tensor:
    """``global_mean_pool``: per-graph mean of node rows (fp32 ``[B, H]``; custom op ``alignn_b200::segment_mean``)."""
    _require_cuda(x)
    return torch.ops.alignn_b200.segment_mean(x, pool_plan.rowptr, pool_plan.eid, int(pool_plan.n_nodes))


# --------------------------------------------------------------------------------------------------
# streaming path (hidden = 256): raw wrappers, no autograd -- used by fused.py
# --------------------------------------------------------------------------------------------------
def edgeattn_supported(hidden: int, heads: int) -> bool:
    return bool(_lib.load().alignn_edgeattn_supported(int(hidden), int(heads)))


def angle_supported(in_dim: int, hidden: int) -> bool:
    return bool(_lib.load().alignn_angle_supported(int(in_dim), int(hidden)))


USE_LG = True    # in-kernel-feature line-graph kernels (csrc/lgattn.cu) where supported (H=256, 4 heads, bf16)
RNG_STEP: Optional[Tensor] = None   # device uint64 counter added to every dropout offset (set by engine.GraphedTrainStep
                                    # so that CUDA-graph replays draw fresh masks); None in eager mode


def lgattn_enabled(hidden: int, heads: int, in_dim: int, dtype: torch.dtype) -> bool:
    return USE_LG and USE_MMA and dtype in _DT and bool(
        _lib.load().alignn_lgattn_supported(int(hidden), int(heads), int(in_dim), _DT[dtype]))


USE_MMA = True   # tensor-core (mma.sync) variants of the streaming kernels where supported (H=256, 4 heads, bf16)


def mma_enabled(hidden: int, heads: int, dtype: torch.dtype) -> bool:
    return USE_MMA and dtype in _DT and bool(
        _lib.load().alignn_edgeattn_mma_supported(int(hidden), int(heads), _DT[dtype]))


def _ld(t: Tensor) -> int:
    """row stride (elements) of a 2-D tensor whose rows are contiguous"""
    if t.dim() != 2 or t.stride(1) != 1:
        raise RuntimeError("expected a 2-D tensor with unit column stride")
    return int(t.stride(0))


def raw_edgeattn_fwd(q: Tensor, k: Tensor, v: Tensor, qt: Tensor, feat: Tensor, plan: GraphPlan, heads: int,
                     p_drop: float, seed: int, offset: int):
    _static_dropout_guard(p_drop, "edgeattn_fwd")
    lib = _lib.load()
    n_nodes, hidden = q.shape
    n_edges = plan.n_edges
    dev = q.device
    f32 = dict(dtype=torch.float32, device=dev)
    aggv = torch.empty(n_nodes, hidden, **f32)
    abar = torch.empty(heads, n_nodes, hidden, dtype=q.dtype, device=dev)
    m, z, s = (torch.empty(n_nodes, heads, **f32) for _ in range(3))
    fn = lib.alignn_edgeattn_mma_fwd if mma_enabled(hidden, heads, q.dtype) else lib.alignn_edgeattn_fwd
    with torch.cuda.device(dev), _Launch("edgeattn_fwd", 1, (n_nodes, n_edges, hidden, heads, q.element_size())):
        rc = fn(_p(q), _p(k), _p(v), _ld(q), _ld(k), _ld(v), _p(qt), _p(feat), _p(plan.rowptr),
                _p(plan.col), _p(plan.eid), _p(aggv), _p(abar), _p(m), _p(z), _p(s), n_nodes,
                n_edges, hidden, heads, _dtype_code(q), float(p_drop), seed, offset, _stream())
    _lib.check(rc, "alignn_edgeattn_fwd")
    return aggv, abar, m, z, s


def raw_attn_fwd_s(q: Tensor, k: Tensor, v: Tensor, qt: Tensor, feat: Tensor, plan: GraphPlan, heads: int,
                   p_drop: float, seed: int, offset: int, rng_step: Optional[Tensor] = None,
                   abar: Optional[Tensor] = None):
    """Stored-feature tensor-core forward; ``qt`` (and ``abar`` when given) are [heads, n, 256] views with arbitrary row /
    head strides."""
    lib = _lib.load()
    n_nodes, hidden = q.shape
    n_edges = plan.n_edges
    dev = q.device
    f32 = dict(dtype=torch.float32, device=dev)
    aggv = torch.empty(n_nodes, hidden, **f32)
    if abar is None:
        abar = torch.empty(heads, n_nodes, hidden, dtype=q.dtype, device=dev)
    m, z, s = (torch.empty(n_nodes, heads, **f32) for _ in range(3))
    with torch.cuda.device(dev), _Launch("edgeattn_fwd", 1, (n_nodes, n_edges, hidden, heads, q.element_size())):
        rc = lib.alignn_edgeattn_mma_fwd_s(_p(q), _p(k), _p(v), _ld(q), _ld(k), _ld(v), _p(qt), int(qt.stride(1)),
                                           int(qt.stride(0)), _p(feat), _p(plan.rowptr), _p(plan.col), _p(plan.eid),
                                           _p(aggv), _p(abar), int(abar.stride(1)), int(abar.stride(0)), _p(m), _p(z),
                                           _p(s), n_nodes, n_edges, hidden, heads, _dtype_code(q), float(p_drop), seed,
                                           offset, _p(rng_step), _stream())
    _lib.check(rc, "alignn_edgeattn_mma_fwd_s")
    return aggv, abar, m, z, s


def raw_attn_bwd_s(dagg: Tensor, dagg_lp: Tensor, agg: Tensor, q: Tensor, k: Tensor, v: Tensor, qt: Tensor, gt: Tensor,
                   cvec: Optional[Tensor], feat: Tensor, m: Tensor, z: Tensor, plan: GraphPlan, heads: int, dq: Tensor,
                   dk: Tensor, dv: Tensor, bbar: Tensor, df_out: Optional[Tensor], p_drop: float, seed: int, offset: int,
                   rng_step: Optional[Tensor] = None) -> None:
    """Stored-feature tensor-core backward (both passes); qt / gt / bbar are [heads, n, 256] views."""
    lib = _lib.load()
    n_nodes, hidden = q.shape
    n_edges = plan.n_edges
    dev = q.device
    coef = torch.empty(max(n_edges, 1), 2 * heads, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        with _Launch("edgeattn_bwd_dst", 1, (n_nodes, n_edges, hidden, heads, q.element_size(), False)):
            rc = lib.alignn_edgeattn_mma_bwd_dst_s(
                _p(dagg), _p(dagg_lp), _p(agg), _p(q), _p(k), _p(v), _ld(q), _ld(k), _ld(v), _p(qt), int(qt.stride(1)),
                int(qt.stride(0)), _p(gt), int(gt.stride(1)), int(gt.stride(0)), _p(cvec), _p(feat), _p(m), _p(z),
                _p(plan.rowptr), _p(plan.col), _p(plan.eid), _p(dq), _ld(dq), _p(bbar), int(bbar.stride(1)),
                int(bbar.stride(0)), _p(coef), None, _p(df_out), _ld(df_out) if df_out is not None else hidden, 0,
                n_nodes, n_edges, hidden, heads, _dtype_code(q),
                float(p_drop), seed, offset, _p(rng_step), _stream())
        _lib.check(rc, "alignn_edgeattn_mma_bwd_dst_s")
        with _Launch("edgeattn_bwd_src", 1, (n_nodes, n_edges, hidden, heads, q.element_size())):
            rc = lib.alignn_edgeattn_bwd_src_lp(_p(dagg_lp), _p(q), _ld(q), _p(coef), _p(plan.rowptr_t), _p(plan.col_t),
                                                _p(plan.eid_t), _p(dk), _p(dv), _ld(dk), n_nodes, n_edges, hidden, heads,
                                                _dtype_code(q), _stream())
        _lib.check(rc, "alignn_edgeattn_bwd_src_lp")


def _work(plan: GraphPlan) -> Tensor:
    """Zeroed scheduler words of the dynamically scheduled line-graph kernels (they leave them zero); one per plan."""
    w = getattr(plan, "_work", None)
    if w is None:
        w = plan._work = torch.zeros(4, dtype=torch.int32, device=plan.rowptr.device)
    return w


def pack_angles(a: Tensor, plan: GraphPlan) -> Tensor:
    """``[L, 16]`` bf16 angle features in the plan's target-sorted order, bias column set to 1 (``csrc/lgattn.cu``)."""
    lib = _lib.load()
    a = a.contiguous().float()
    n_edges, in_dim = a.shape
    out = torch.empty(max(n_edges, 1), 16, dtype=torch.bfloat16, device=a.device)
    with torch.cuda.device(a.device), _Launch("lg_pack_angles", 1, (n_edges, in_dim)):
        rc = lib.alignn_lg_pack_angles(_p(a), _p(plan.eid), _p(out), n_edges, in_dim, _stream())
    _lib.check(rc, "alignn_lg_pack_angles")
    return out[:n_edges]


# line-graph attention forward on tcgen05 / TMEM (csrc/lgattn_tc.cu) instead of the mma.sync kernel (csrc/lgattn.cu)
BF16_CODE = 1
LGATTN_TC = os.environ.get("ALIGNN_LGATTN_TC", "0") == "1"


def raw_lgattn_fwd(q: Tensor, k: Tensor, v: Tensor, qt: Tensor, a_csr: Tensor, w1: Tensor, b1: Tensor, plan: GraphPlan,
                   heads: int, p_drop: float, seed: int, offset: int, rng_step: Optional[Tensor] = None,
                   abar: Optional[Tensor] = None):
    """``qt`` (and ``abar`` when given): [heads, n, 256] views (any row / head strides, unit channel stride).
    Returns (aggv, abar, m, z, s)."""
    lib = _lib.load()
    n_nodes, hidden = q.shape
    n_edges = plan.n_edges
    dev = q.device
    f32 = dict(dtype=torch.float32, device=dev)
    aggv = torch.empty(n_nodes, hidden, **f32)
    if abar is None:
        abar = torch.empty(heads, n_nodes, hidden, dtype=q.dtype, device=dev)
    m, z, s = (torch.empty(n_nodes, heads, **f32) for _ in range(3))
    if LGATTN_TC:
        with torch.cuda.device(dev), _Launch("lgattn_fwd", 1, (n_nodes, n_edges, hidden, heads, q.element_size())):
            rc = lib.alignn_lgattn_fwd_tc(_p(q), _p(k), _p(v), _ld(q), _ld(k), _ld(v), _p(qt), int(qt.stride(1)),
                                          int(qt.stride(0)), _p(a_csr), _p(w1), _p(b1), int(w1.size(1)), _p(plan.rowptr),
                                          _p(plan.col), _p(aggv), _p(abar), int(abar.stride(1)), int(abar.stride(0)),
                                          _p(m), _p(z), _p(s), n_nodes, n_edges, hidden, heads, _dtype_code(q),
                                          float(p_drop), seed, offset, _p(rng_step), _stream())
        _lib.check(rc, "alignn_lgattn_fwd_tc")
        return aggv, abar, m, z, s
    with torch.cuda.device(dev), _Launch("lgattn_fwd", 1, (n_nodes, n_edges, hidden, heads, q.element_size())):
        rc = lib.alignn_lgattn_fwd(_p(q), _p(k), _p(v), _ld(q), _ld(k), _ld(v), _p(qt), int(qt.stride(1)),
                                   int(qt.stride(0)), _p(a_csr), _p(w1), _p(b1), int(w1.size(1)), _p(plan.rowptr),
                                   _p(plan.col), _p(aggv), _p(abar), int(abar.stride(1)), int(abar.stride(0)),
                                   _p(m), _p(z), _p(s), n_nodes, n_edges, hidden, heads, _dtype_code(q), float(p_drop),
                                   seed, offset, _p(rng_step), _p(_work(plan)), _stream())
    _lib.check(rc, "alignn_lgattn_fwd")
    return aggv, abar, m, z, s


def csc_positions(plan: GraphPlan) -> Tensor:
    """int32 [Ne]: CSR (target-sorted) position of the p-th edge of the CSC (source-sorted) order; cached on the plan."""
    pos_t = getattr(plan, "_pos_t", None)
    if pos_t is None:
        lib = _lib.load()
        dev = plan.eid.device
        scratch = torch.empty(max(plan.n_edges, 1), dtype=torch.int32, device=dev)
        pos_t = torch.empty(max(plan.n_edges, 1), dtype=torch.int32, device=dev)
        with torch.cuda.device(dev), _Launch("build_plan", 2, (plan.n_nodes, plan.n_edges)):
            rc = lib.alignn_plan_csc_positions(_p(plan.eid), _p(plan.eid_t), plan.n_edges, _p(scratch), _p(pos_t), _stream())
        _lib.check(rc, "alignn_plan_csc_positions")
        pos_t = pos_t[:plan.n_edges]
        plan._pos_t = pos_t
    return pos_t


def raw_lgattn_bwd(dagg: Tensor, dagg_lp: Tensor, agg: Tensor, q: Tensor, k: Tensor, v: Tensor, qt: Tensor, gt: Tensor,
                   cvec: Optional[Tensor], a_csr: Tensor, w1: Tensor, b1: Tensor, m: Tensor, z: Tensor, plan: GraphPlan,
                   heads: int, dq: Tensor, dk: Tensor, dv: Tensor, bbar: Tensor, p_drop: float, seed: int, offset: int,
                   rng_step: Optional[Tensor] = None) -> Tensor:
    """Both backward passes of the in-kernel-feature line-graph conv.  qt / gt / bbar: [heads, n, 256] views (any row /
    head strides).  Returns coef [Ne, 8] (CSR order) for :func:`raw_lg_angle_grad`."""
    lib = _lib.load()
    n_nodes, hidden = q.shape
    n_edges = plan.n_edges
    dev = q.device
    coef = torch.empty(max(n_edges, 1), 2 * heads, dtype=torch.float32, device=dev)
    pos_t = csc_positions(plan)
    with torch.cuda.device(dev):
        with _Launch("lgattn_bwd_dst", 1, (n_nodes, n_edges, hidden, heads, q.element_size())):
            rc = lib.alignn_lgattn_bwd_dst(
                _p(dagg), _p(dagg_lp), _p(agg), _p(q), _p(k), _p(v), _ld(q), _ld(k), _ld(v),
                _p(qt), int(qt.stride(1)), int(qt.stride(0)), _p(gt), int(gt.stride(1)), int(gt.stride(0)),
                _p(cvec), _p(a_csr), _p(w1), _p(b1), int(w1.size(1)), _p(m), _p(z), _p(plan.rowptr), _p(plan.col),
                _p(dq), _ld(dq), _p(bbar), int(bbar.stride(1)), int(bbar.stride(0)), _p(coef), n_nodes, n_edges, hidden,
                heads, _dtype_code(q), float(p_drop), seed, offset, _p(rng_step), _p(_work(plan)), _stream())
        _lib.check(rc, "alignn_lgattn_bwd_dst")
        with _Launch("edgeattn_bwd_src", 1, (n_nodes, n_edges, hidden, heads, q.element_size())):
            rc = lib.alignn_edgeattn_bwd_src_lp(_p(dagg_lp), _p(q), _ld(q), _p(coef), _p(plan.rowptr_t), _p(plan.col_t),
                                                _p(pos_t), _p(dk), _p(dv), _ld(dk), n_nodes, n_edges, hidden, heads,
                                                _dtype_code(q), _stream())
        _lib.check(rc, "alignn_edgeattn_bwd_src_lp")
    return coef


def raw_lg_angle_grad(a_csr: Tensor, w1: Tensor, b1: Tensor, plan: GraphPlan, coefs, qts, gts, max_blocks: int = 0):
    """(dW1 [256, in_dim], db1 [256]) in fp32 from the coefficient tensors of all line-graph layers.  ``max_blocks`` > 0 caps
    the grid (one CTA per SM) so that work on other streams keeps the remaining SMs."""
    lib = _lib.load()
    in_dim = int(w1.size(1))
    dev = w1.device
    n_nodes, n_edges = plan.n_nodes, plan.n_edges
    f32 = dict(dtype=torch.float32, device=dev)
    total = torch.zeros((in_dim + 1) * 256, **f32)
    partials = torch.empty(int(lib.alignn_lg_angle_grad_partial_floats(n_nodes, n_edges)), **f32)
    for lo in range(0, len(coefs), 4):
        c, qq, gg = coefs[lo:lo + 4], qts[lo:lo + 4], gts[lo:lo + 4]
        n = len(c)
        out = torch.empty((in_dim + 1) * 256, **f32)
        arr = lambda ts: (ctypes.c_void_p * n)(*[t.data_ptr() for t in ts])   # noqa: E731
        with torch.cuda.device(dev), _Launch("lg_angle_grad", 2, (n_nodes, n_edges, n)):
            rc = lib.alignn_lg_angle_grad2(_p(a_csr), _p(w1), _p(b1), in_dim, _p(plan.rowptr), n, arr(c), arr(qq), arr(gg),
                                           int(qq[0].stride(1)), int(qq[0].stride(0)), int(gg[0].stride(1)),
                                           int(gg[0].stride(0)), _p(partials), _p(out), n_nodes, n_edges, int(max_blocks),
                                           _stream())
        _lib.check(rc, "alignn_lg_angle_grad")
        total = out if lo == 0 and len(coefs) <= 4 else total + out
    dw1 = total[:in_dim * 256].view(in_dim, 256).t().contiguous()
    return dw1, total[in_dim * 256:].clone()


def raw_edgeattn_bwd(dagg: Tensor, dagg_lp: Optional[Tensor], agg: Tensor, q: Tensor, k: Tensor, v: Tensor, qt: Tensor,
                     gt: Tensor, cvec: Optional[Tensor], feat: Tensor, m: Tensor, z: Tensor, plan: GraphPlan, heads: int,
                     dq: Tensor, dk: Tensor, dv: Tensor, df_in: Optional[Tensor], df_out: Optional[Tensor],
                     relu_mask: bool, p_drop: float, seed: int, offset: int) -> Tensor:
    """Both backward passes; dq/dk/dv are (strided) outputs; returns bbar [heads, Nn, 256].
    ``dagg_lp``: storage-dtype copy of ``dagg`` (tensor-core path; made here when missing)."""
    lib = _lib.load()
    n_nodes, hidden = q.shape
    n_edges = plan.n_edges
    dev = q.device
    bbar = torch.empty(heads, n_nodes, hidden, dtype=q.dtype, device=dev)
    coef = torch.empty(max(n_edges, 1), 2 * heads, dtype=torch.float32, device=dev)
    use_mma = mma_enabled(hidden, heads, q.dtype)
    with torch.cuda.device(dev):
        with _Launch("edgeattn_bwd_dst", 1, (n_nodes, n_edges, hidden, heads, q.element_size(), df_in is not None)):
            if use_mma:
                if dagg_lp is None or dagg_lp.dtype != q.dtype:
                    dagg_lp = dagg.to(q.dtype)
                rc = lib.alignn_edgeattn_mma_bwd_dst(
                    _p(dagg), _p(dagg_lp), _p(agg), _p(q), _p(k), _p(v), _ld(q), _ld(k), _ld(v), _p(qt), _p(gt),
                    _p(cvec), _p(feat), _p(m), _p(z), _p(plan.rowptr), _p(plan.col), _p(plan.eid), _p(dq), _ld(dq),
                    _p(bbar), _p(coef), _p(df_in), _p(df_out), int(bool(relu_mask)), n_nodes, n_edges, hidden, heads,
                    _dtype_code(q), float(p_drop), seed, offset, _stream())
            else:
                rc = lib.alignn_edgeattn_bwd_dst(_p(dagg), _p(agg), _p(q), _p(k), _p(v), _ld(q), _ld(k), _ld(v), _p(qt),
                                                 _p(gt), _p(cvec), _p(feat), _p(m), _p(z), _p(plan.rowptr), _p(plan.col),
                                                 _p(plan.eid), _p(dq), _ld(dq), _p(bbar), _p(coef), _p(df_in), _p(df_out),
                                                 int(bool(relu_mask)), n_nodes, n_edges, hidden, heads, _dtype_code(q),
                                                 float(p_drop), seed, offset, _stream())
        _lib.check(rc, "alignn_edgeattn_bwd_dst")
        with _Launch("edgeattn_bwd_src", 1, (n_nodes, n_edges, hidden, heads, q.element_size())):
            rc = lib.alignn_edgeattn_bwd_src(_p(dagg), _p(q), _ld(q), _p(coef), _p(plan.rowptr_t), _p(plan.col_t),
                                             _p(plan.eid_t), _p(dk), _p(dv), _ld(dk), n_nodes, n_edges, hidden, heads,
                                             _dtype_code(q), _stream())
        _lib.check(rc, "alignn_edgeattn_bwd_src")
    return bbar


def raw_gate_ln_fwd2(aggv: Tensor, agge: Optional[Tensor], cvec: Optional[Tensor], stat_s: Optional[Tensor],
                     heads: int, xr: Tensor, x: Tensor, wbeta: Tensor, gamma: Tensor, bias: Tensor, eps: float,
                     p_drop: float, seed: int, offset: int, want_lp: bool, rng_step: Optional[Tensor] = None,
                     agg_rows: int = -1, x_lp: Optional[Tensor] = None):
    """``agg_rows >= 0``: only the first ``agg_rows`` rows have an aggregate (``agge`` is ``[heads, agg_rows, C]``).
    ``x = None`` + ``x_lp``: the residual input is read in the storage dtype (first block of a chain)."""
    lib = _lib.load()
    n_rows, hidden = (x if x is not None else x_lp).shape
    if x is None:
        if agg_rows < 0:
            agg_rows = n_rows
        x_lp = x_lp.contiguous()
    dev = aggv.device
    f32 = dict(dtype=torch.float32, device=dev)
    agg = torch.empty(n_rows, hidden, **f32)
    y = torch.empty(n_rows, hidden, **f32)
    y_lp = torch.empty(n_rows, hidden, dtype=xr.dtype, device=dev) if want_lp else None
    beta, mean, rstd = (torch.empty(n_rows, **f32) for _ in range(3))
    with torch.cuda.device(dev), _Launch("gate_ln_fwd", 1, (n_rows, hidden, xr.element_size())):
        if agg_rows >= 0:
            rc = lib.alignn_gate_ln_fwd4(_p(aggv), _p(agge), _p(cvec), _p(stat_s), heads, int(agg_rows), _p(xr), _ld(xr),
                                         _p(x), _p(x_lp) if x is None else None, _p(wbeta), _p(gamma), _p(bias), _p(agg),
                                         _p(y), _p(y_lp), _p(beta),
                                         _p(mean), _p(rstd), n_rows, hidden, _dtype_code(xr), float(eps), float(p_drop),
                                         seed, offset, _p(rng_step), _stream())
        else:
            rc = lib.alignn_gate_ln_fwd2(_p(aggv), _p(agge), _p(cvec), _p(stat_s), heads, _p(xr), _ld(xr), _p(x),
                                         _p(wbeta), _p(gamma), _p(bias), _p(agg), _p(y), _p(y_lp), _p(beta), _p(mean),
                                         _p(rstd), n_rows, hidden, _dtype_code(xr), float(eps), float(p_drop), seed,
                                         offset, _p(rng_step), _stream())
    _lib.check(rc, "alignn_gate_ln_fwd2/3")
    return y, y_lp, agg, beta, mean, rstd


def raw_gate_ln_bwd2(dy: Tensor, agg: Tensor, xr: Tensor, wbeta: Tensor, gamma: Tensor, bias: Tensor, beta: Tensor,
                     mean: Tensor, rstd: Tensor, dxr: Tensor, want_lp: bool, p_drop: float, seed: int, offset: int,
                     rng_step: Optional[Tensor] = None):
    """Returns (dagg f32, dagg_lp or None, dparams f32 [5*hidden]); writes dxr (strided) in place."""
    lib = _lib.load()
    n_rows, hidden = agg.shape
    dev = agg.device
    f32 = dict(dtype=torch.float32, device=dev)
    dagg = torch.empty(n_rows, hidden, **f32)
    dagg_lp = torch.empty(n_rows, hidden, dtype=xr.dtype, device=dev) if want_lp else None
    partials = torch.empty(int(lib.alignn_gate_ln_bwd_partial_rows()) * 5 * hidden, **f32)
    dparams = torch.empty(5 * hidden, **f32)
    with torch.cuda.device(dev), _Launch("gate_ln_bwd", 2, (n_rows, hidden, xr.element_size())):
        rc = lib.alignn_gate_ln_bwd2(_p(dy), _p(agg), _p(xr), _ld(xr), _p(wbeta), _p(gamma), _p(bias), _p(beta),
                                     _p(mean), _p(rstd), _p(dagg), _p(dagg_lp), _p(dxr), _ld(dxr), _p(partials),
                                     _p(dparams), n_rows, hidden, _dtype_code(xr), float(p_drop), seed, offset,
                                     _p(rng_step), _stream())
    _lib.check(rc, "alignn_gate_ln_bwd2")
    return dagg, dagg_lp, dparams


def raw_gate_ln_bwd3(dy: Optional[Tensor], agg: Tensor, xr: Tensor, wbeta: Tensor, gamma: Tensor, bias: Tensor, beta: Tensor,
                     mean: Tensor, rstd: Tensor, stat_s: Tensor, heads: int, dxr: Tensor, p_drop: float, seed: int,
                     offset: int, rng_step: Optional[Tensor] = None, dy2: Optional[Tensor] = None, agg_rows: int = -1,
                     dparams: Optional[Tensor] = None):
    """As :func:`raw_gate_ln_bwd2` (always emits the storage-dtype copy of dagg) plus the gradient of the folded
    edge-projection bias: returns (dagg f32, dagg_lp, dparams f32 [6*hidden] = dw_beta x3 | dgamma | dbias | dcvec)."""
    lib = _lib.load()
    n_rows, hidden = agg.shape
    dev = agg.device
    f32 = dict(dtype=torch.float32, device=dev)
    dagg = torch.empty(n_rows, hidden, **f32)
    dagg_lp = torch.empty(n_rows, hidden, dtype=xr.dtype, device=dev)
    partials = torch.empty(int(lib.alignn_gate_ln_bwd_partial_rows()) * 6 * hidden, **f32)
    if dparams is None:
        dparams = torch.empty(6 * hidden, **f32)
    with torch.cuda.device(dev), _Launch("gate_ln_bwd", 2, (n_rows, hidden, xr.element_size())):
        rc = lib.alignn_gate_ln_bwd3(_p(dy), _p(dy2), _ld(dy2) if dy2 is not None else hidden, _p(agg), _p(xr), _ld(xr), _p(wbeta), _p(gamma), _p(bias), _p(beta),
                                     _p(mean), _p(rstd), _p(stat_s), heads, int(agg_rows), _p(dagg), _p(dagg_lp), _p(dxr),
                                     _ld(dxr),
                                     _p(partials), _p(dparams), n_rows, hidden, _dtype_code(xr), float(p_drop), seed,
                                     offset, _p(rng_step), _stream())
    _lib.check(rc, "alignn_gate_ln_bwd3")
    return dagg, dagg_lp, dparams


WGRAD_TC = os.environ.get("ALIGNN_WGRAD_TC", "1") == "1"     # tcgen05 weight + bias gradient (csrc/wgrad_tc.cu) vs cuBLAS + colsum
# the UMMA kernel wins where the reduction is long and the output small (one 256-channel group: dW_s, K = all bond rows);
# for the wide stacked projection (M = 7H over the active rows only) its per-split partials outweigh the operands
WGRAD_TC_MAX_M = int(os.environ.get("ALIGNN_WGRAD_TC_MAX_M", "256"))
WGRAD_TC_MIN_K = 32768            # below this the launch + reduction overhead of the split kernel eats the gain


def wgrad(a: Tensor, b: Tensor, w_out: Tensor, b_out: Optional[Tensor] = None) -> None:
    """``w_out[M, N] = a^T b`` and ``b_out[M] = a.sum(0)`` (fp32) for ``a [K, M]``, ``b [K, N]`` bf16 with unit column
    stride: the weight and bias gradient of a node projection in one pass (``alignn_wgrad``); falls back to cuBLAS + the
    column-sum kernel for shapes the tcgen05 kernel does not take."""
    lib = _lib.load()
    k, m = a.shape
    n = int(b.size(1))
    ok = (WGRAD_TC and a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and bool(lib.alignn_wgrad_supported(n, BF16_CODE))
          and m % 8 == 0 and a.stride(1) == 1 and b.stride(1) == 1 and _ld(a) % 8 == 0 and _ld(b) % 8 == 0
          and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0 and w_out.is_contiguous() and w_out.dtype == torch.float32
          and (b_out is None or (b_out.is_contiguous() and b_out.dtype == torch.float32)) and k > 0
          and (m <= WGRAD_TC_MAX_M or WGRAD_TC_MAX_M < 0) and (k >= WGRAD_TC_MIN_K or WGRAD_TC_MAX_M < 0))
    if not ok:
        torch.mm(a.t(), b, out_dtype=torch.float32, out=w_out)
        if b_out is not None:
            colsum(a, out=b_out)
        return
    partials = torch.empty(int(lib.alignn_wgrad_partial_floats(k, m)), dtype=torch.float32, device=a.device)
    with torch.cuda.device(a.device), _Launch("wgrad", 2, (k, m, n)):
        rc = lib.alignn_wgrad(_p(a), _ld(a), _p(b), _ld(b), k, m, n, BF16_CODE, _p(partials), _p(w_out), _p(b_out), _stream())
    _lib.check(rc, "alignn_wgrad")


PROJ_TC = os.environ.get("ALIGNN_PROJ_TC", "1") == "1"       # tcgen05 + TMA node projections (csrc/proj_tc.cu) vs cuBLAS addmm


def linear_lp(x: Tensor, w: Tensor, bias: Optional[Tensor] = None) -> Tensor:
    """``x w^T + bias`` for the bf16 node projections (``nn.Linear`` of the reference's TransformerConv on the node state):
    the hand-written tcgen05 / TMA GEMM (``alignn_proj_tc``) when the shape is one it takes (K = 256, N a multiple of 128,
    bf16, unit column strides, 16-byte aligned rows), cuBLAS (``torch.addmm``) otherwise."""
    lib = _lib.load()
    m, k = x.shape
    n = int(w.size(0))
    ok = (PROJ_TC and x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and m > 0
          and (bias is None or (bias.dtype == torch.bfloat16 and bias.is_contiguous() and bias.data_ptr() % 16 == 0))
          and bool(lib.alignn_proj_tc_supported(k, n, BF16_CODE)) and x.stride(1) == 1 and w.stride(1) == 1
          and _ld(x) % 8 == 0 and _ld(w) % 8 == 0 and x.data_ptr() % 16 == 0 and w.data_ptr() % 16 == 0)
    if not ok:
        return torch.addmm(bias, x, w.t()) if bias is not None else torch.mm(x, w.t())
    out = torch.empty(m, n, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device), _Launch("proj_tc", 1, (m, n, k)):
        rc = lib.alignn_proj_tc(_p(x), _ld(x), _p(w), _ld(w), _p(bias), _p(out), n, m, n, k, BF16_CODE, _stream())
    _lib.check(rc, "alignn_proj_tc")
    return out


def block_projections(x: Tensor, w8: Tensor, b8: Tensor, n_active: int) -> Tuple[Tensor, Tensor]:
    """The two node projections of one conv block from ONE pass over ``x [n, H]`` (bf16): ``x_r = x Ws^T + bs`` for every
    row and ``q | k | v | qt_0..3 = x[:n_active] W7^T + b7`` for the active prefix, with ``w8 = [W7; Ws]`` (``[8H, H]``) and
    ``b8`` stacked the same way (``trunk.py``).  One launch of the tcgen05 / TMA GEMM (``alignn_proj_tc2``); two cuBLAS
    GEMMs when the shape is not one it takes.  Returns ``(x_r [n, H], proj [n_active, 7H])``."""
    lib = _lib.load()
    n, hid = x.shape
    ok = (PROJ_TC and x.is_cuda and x.dtype == torch.bfloat16 and w8.dtype == torch.bfloat16 and b8.dtype == torch.bfloat16
          and hid == 256 and tuple(w8.shape) == (8 * hid, hid) and w8.is_contiguous() and b8.is_contiguous() and n > 0
          and x.stride(1) == 1 and _ld(x) % 8 == 0 and x.data_ptr() % 16 == 0 and w8.data_ptr() % 16 == 0
          and b8.data_ptr() % 16 == 0)
    if not ok:
        w7, b7, ws, bs = w8[:7 * hid], b8[:7 * hid], w8[7 * hid:], b8[7 * hid:]
        return torch.addmm(bs, x, ws.t()), torch.addmm(b7, x[:n_active], w7.t())
    xr = torch.empty(n, hid, dtype=torch.bfloat16, device=x.device)
    proj = torch.empty(n_active, 7 * hid, dtype=torch.bfloat16, device=x.device)
    with torch.cuda.device(x.device), _Launch("proj_tc", 1, (n, n_active, hid)):
        rc = lib.alignn_proj_tc2(_p(x), _ld(x), _p(w8), hid, 8 * hid, _p(b8), _p(xr), hid, hid, 7 * hid,
                                 _p(proj), 7 * hid, 7 * hid if n_active > 0 else 0, 0, n, n_active, hid, BF16_CODE, _stream())
    _lib.check(rc, "alignn_proj_tc2")
    return xr, proj


def colsum(x: Tensor, out: Optional[Tensor] = None) -> Tensor:
    """fp32 column sums of a 2-D tensor with unit column stride (deterministic hand-written reduction); ``out``: optional
    contiguous fp32 ``[width]`` destination."""
    lib = _lib.load()
    n_rows, width = x.shape
    if n_rows == 0 or not lib.alignn_colsum_supported(width) or x.dtype not in _DT or _ld(x) % 8 or x.data_ptr() % 16:
        res = x.sum(0, dtype=torch.float32)
        return res if out is None else out.copy_(res)
    f32 = dict(dtype=torch.float32, device=x.device)
    partials = torch.empty(int(lib.alignn_colsum_partial_floats(width)), **f32)
    if out is None:
        out = torch.empty(width, **f32)
    with torch.cuda.device(x.device), _Launch("colsum", 2, (n_rows, width, x.element_size())):
        rc = lib.alignn_colsum(_p(x), _ld(x), n_rows, width, _dtype_code(x), _p(partials), _p(out), _stream())
    _lib.check(rc, "alignn_colsum")
    return out


def raw_angle_h1_fwd(a: Tensor, w1: Tensor, b1: Tensor, dtype: torch.dtype) -> Tensor:
    lib = _lib.load()
    n_edges, in_dim = a.shape
    hidden = int(w1.size(0))
    h1 = torch.empty(n_edges, hidden, dtype=dtype, device=a.device)
    with torch.cuda.device(a.device), _Launch("angle_h1_fwd", 1, (n_edges, in_dim, hidden)):
        rc = lib.alignn_angle_h1_fwd(_p(a), _p(w1), _p(b1), _p(h1), n_edges, in_dim, hidden, _DT[dtype], _stream())
    _lib.check(rc, "alignn_angle_h1_fwd")
    return h1


def raw_angle_h1_bwd(dpre: Tensor, a: Tensor, hidden: int):
    """(dW1 [hidden, in_dim], db1 [hidden]) in fp32 from the ReLU-masked feature gradient."""
    lib = _lib.load()
    n_edges, in_dim = a.shape
    f32 = dict(dtype=torch.float32, device=a.device)
    partials = torch.empty(int(lib.alignn_angle_partial_floats(in_dim)), **f32)
    out = torch.empty((in_dim + 1) * hidden, **f32)
    with torch.cuda.device(a.device), _Launch("angle_h1_bwd", 2, (n_edges, in_dim, hidden)):
        rc = lib.alignn_angle_h1_bwd(_p(dpre), _p(a), _p(partials), _p(out), n_edges, in_dim, hidden,
                                     _dtype_code(dpre), _stream())
    _lib.check(rc, "alignn_angle_h1_bwd")
    dw1 = out[:in_dim * hidden].view(in_dim, hidden).t().contiguous()
    return dw1, out[in_dim * hidden:].clone()
