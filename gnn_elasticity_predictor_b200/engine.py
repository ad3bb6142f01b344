"""One-call training step over the hot path: plan + forward + Gaussian NLL + backward (+ gradient all-reduce) + global-norm
clip + AdamW, optionally captured into a CUDA graph per batch signature.

What it replaces in the reference is the body of ``train_epoch_hetero``'s batch loop (``scripts/train.py:639-699``):
``model(batch)`` -> NLL + log-sigma L2 (``:655-681``) -> ``backward()`` (``:691/697``) -> ``clip_grad_norm_(5.0)`` (``:693,698``)
-> ``AdamW.step()`` with the two learning-rate groups built at ``:1516-1540`` (base + mean heads | log-variance heads).

B200 specifics:

* all parameters live in ONE flat fp32 buffer (``model.parameters()`` order, log-variance heads last), their gradients in a
  second flat buffer with identical offsets (:class:`dp.FlatGradBucket`), Adam moments in two more: the optimizer is ONE
  fused kernel pair (``csrc/optim.cu``: deterministic norm -> clip -> decoupled-weight-decay Adam), the DP exchange ONE
  NCCL all-reduce;
* nothing in the step synchronises with the host, so the whole step is captured once per batch signature
  ``(B, N, E, L)`` into a CUDA graph and replayed: ~400 kernel launches collapse into one ``cudaGraphLaunch``.  Dropout
  masks stay fresh across replays through a device-side step counter added to every Philox offset (``ops.RNG_STEP``);
  the learning rates and the Adam step count also live on the device;
* batches whose signature has not been captured (ragged datasets) run the same code eagerly.

The step leaves ``param.grad`` populated (views of the flat bucket), so callers can still inspect gradients.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from . import _lib, ops
from .dp import FlatGradBucket
from .modules import HeteroAlignnRegressor, fused_gaussian_nll, gaussian_nll_loss  # noqa: F401
from .synthetic import GraphBatch
from . import batching

_P = ops._p


class FusedAdamW:
    """Clip + AdamW over the flat parameter bucket (``alignn_clip_adamw_step``); two LR groups split at ``split``."""

    def __init__(self, bucket: FlatGradBucket, split: int, lr: float, lr_sigma: Optional[float], weight_decay: float,
                 betas=(0.9, 0.999), eps: float = 1e-8, max_norm: float = 5.0, n_active: Optional[int] = None):
        self.bucket, self.split = bucket, int(split)
        dev = bucket.flat.device
        n = bucket.flat.numel()
        # parameters past n_active never receive a gradient (torch's AdamW skips such parameters entirely: no decay)
        self.n_active = n if n_active is None else int(n_active)
        self.flat_params = torch.empty(n, dtype=torch.float32, device=dev)
        with torch.no_grad():
            self.flat_params.zero_()
            for p, off in zip(bucket.params, bucket.offsets):
                self.flat_params[off:off + p.numel()].copy_(p.detach().reshape(-1))
                p.data = self.flat_params[off:off + p.numel()].view_as(p)
        self.exp_avg = torch.zeros_like(self.flat_params)
        self.exp_avg_sq = torch.zeros_like(self.flat_params)
        lib = _lib.load()
        self.partials = torch.zeros(int(lib.alignn_adamw_partial_floats()), dtype=torch.float32, device=dev)
        self.step_count = torch.zeros(1, dtype=torch.float32, device=dev)
        self.lr = torch.tensor([lr, lr if lr_sigma is None else lr_sigma], dtype=torch.float32, device=dev)
        self.norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.betas, self.eps, self.weight_decay, self.max_norm = betas, float(eps), float(weight_decay), float(max_norm)

    def set_lr(self, lr: float, lr_sigma: Optional[float] = None) -> None:
        """Scheduler hook (reference ``_cosine_schedule``, ``train.py:1215``): device-side, seen by captured graphs."""
        self.lr.copy_(torch.tensor([lr, lr if lr_sigma is None else lr_sigma], dtype=torch.float32), non_blocking=True)

    def step(self, grad_scale: float = 1.0) -> None:
        lib = _lib.load()
        n = self.n_active
        with torch.cuda.device(self.flat_params.device), ops._Launch("clip_adamw", 2, (n,)):
            rc = lib.alignn_clip_adamw_step(_P(self.flat_params), _P(self.bucket.flat), _P(self.exp_avg),
                                            _P(self.exp_avg_sq), None, _P(self.partials), _P(self.step_count),
                                            _P(self.lr), _P(self.norm), n, self.split, self.betas[0], self.betas[1],
                                            self.eps, self.weight_decay, self.max_norm, float(grad_scale), ops._stream())
        _lib.check(rc, "alignn_clip_adamw_step")

    def state(self):
        return [self.flat_params, self.exp_avg, self.exp_avg_sq, self.step_count]


class _Captured:
    __slots__ = ("graph", "graph_opt", "batch", "tz", "mask", "weight", "loss", "mean", "logvar", "kernels", "trunk_reduced")


class TrainStep:
    """``step(batch, target_z) -> (loss, mean, logvar)`` -- device tensors, valid until the next call.

    ``graph=True`` captures one CUDA graph per batch signature after ``graph_warmup`` eager steps on that signature (the
    warm-up steps are real training steps).  ``loss_scale`` is the DP weight ``B_local / B_global``: a constant, so the
    summed gradient equals the global-batch mean gradient only when every rank holds the SAME number of real graphs per
    step (``dp.shard_ranges`` + equal shard sizes, what ``bench.py`` and the 2-GPU equality check use); with ragged or
    bucket-padded shards pass per-step ``sample_weight = B_global_real / (world * B_local_real)`` instead."""

    def __init__(self, model: HeteroAlignnRegressor, lr: float = 1e-3, lr_sigma: Optional[float] = None,
                 weight_decay: float = 1e-4, max_norm: float = 5.0, log_sigma_l2: float = 0.1,
                 min_logvar_floor: float = -2.9, loss_scale: float = 1.0, graph: bool = True, graph_warmup: int = 2,
                 optimizer: bool = True, group=None, pad_to_buckets: bool = False, bucket_align: int = 256,
                 data_parallel: bool = True, feature_jitter_std: float = 0.0, allreduce_in_graph: bool = True,
                 early_allreduce: bool = False, defer_angle_blocks: int = 132):
        self.model = model
        params = [p for p in model.parameters() if p.requires_grad]
        if not params or not params[0].is_cuda:
            raise RuntimeError("TrainStep needs a model on a CUDA device (there is no CPU path)")
        self.dev = params[0].device
        # bucket order: [base (used) + mean heads | log-variance heads | base.output_heads]: the hetero forward never
        # touches base.output_heads (train.py:579-586), so they get no gradient and the reference's optimizer skips them
        sigma = {id(p) for p in model.logvar_heads.parameters()}
        unused = {id(p) for p in model.base.output_heads.parameters()}
        ordered = ([p for p in params if id(p) not in sigma and id(p) not in unused]
                   + [p for p in params if id(p) in sigma] + [p for p in params if id(p) in unused])
        self.bucket = FlatGradBucket(ordered)
        split = n_active = self.bucket.flat.numel()
        for p, off in zip(self.bucket.params, self.bucket.offsets):
            if id(p) in sigma:
                split = min(split, off)
            if id(p) in unused:
                n_active = min(n_active, off)
        split = min(split, n_active)
        self.opt = FusedAdamW(self.bucket, split, lr, lr_sigma, weight_decay, max_norm=max_norm,
                              n_active=n_active) if optimizer else None
        self.log_sigma_l2, self.floor, self.loss_scale = float(log_sigma_l2), float(min_logvar_floor), float(loss_scale)
        self.use_graph, self.graph_warmup, self.group = bool(graph), int(graph_warmup), group
        # data_parallel=False: this rank trains its own ensemble member (member-per-GPU placement), no gradient exchange
        self.world = dist.get_world_size(group) if (data_parallel and dist.is_available() and dist.is_initialized()) else 1
        self.rng_step = torch.zeros(1, dtype=torch.int64, device=self.dev)
        self.pad_to_buckets, self.bucket_align = bool(pad_to_buckets), int(bucket_align)
        # reference train.py:641-646 (--feature-jitter-std, default 0.1 there): Gaussian noise on x and global_x every
        # training step.  0 here by default so that parity runs are noise-free; torch's CUDA generator is graph-safe.
        self.feature_jitter_std = float(feature_jitter_std)
        # N > 1: capture the NCCL all-reduce of the flat bucket INSIDE the step graph (one cudaGraphLaunch per step, no
        # host round trip between backward, collective and optimizer); falls back to two graphs with the collective
        # between them if the capture is refused
        self.allreduce_in_graph = bool(allreduce_in_graph)
        # early_allreduce (N > 1, fused trunk; OFF by default): the gradients of the trunk's FOLDED weights (~95 % of the
        # bucket) are all-reduced inside the trunk backward, beside the angle-encoder gradient kernel (trunk.py); what is left
        # for the end of the step are the bucket slices of the parameters outside the trunk (~1.4 MB).  Parity-green
        # (scripts/check_dp_equality.py) but not faster on 2 or 8 x B200: the persistent angle-gradient kernel holds every SM,
        # the NCCL CTAs only get in at its wave boundary, and the folded buffer is 17 MB against the bucket's 13 MB
        # (profiles/r02_experiments/dropped_experiments.txt)
        self._trunk_reduced = False
        self._rest_slices = None
        if self.world > 1 and early_allreduce:
            base = model.base
            base._dp_group = group if group is not None else dist.group.WORLD
            base._dp_done = self._mark_trunk_reduced
            trunk_ids = {id(p) for blk in list(base.edge_blocks) + list(base.node_blocks) for p in blk.parameters()}
            if base.angle_encoder is not None:
                trunk_ids |= {id(p) for p in base.angle_encoder[2].parameters()}
            runs, start = [], None
            for p, off in zip(self.bucket.params, self.bucket.offsets):
                if off >= n_active:
                    break
                if id(p) in trunk_ids:
                    if start is not None:
                        runs.append((start, off)); start = None
                elif start is None:
                    start = off
            if start is not None:
                runs.append((start, n_active))
            self._rest_slices = runs
        # the angle-encoder gradient kernel (once per step, persistent, ~0.26 ms) runs on the trunk's side stream with a
        # capped grid and is joined HERE after backward(): the encoder / fold backward and the gradient gather run beside it
        # on the SMs it leaves free.  Only this engine may defer the join (a plain loss.backward() user gets the joined form).
        self.defer_angle_blocks = int(os.environ.get("ALIGNN_DEFER_ANGLE", defer_angle_blocks))
        self._captured: Dict[Tuple, _Captured] = {}
        self._seen: Dict[Tuple, int] = {}
        self.replays = 0
        self.eager_steps = 0

    # -- the step itself (eager; also what gets captured) ------------------------------------------------------
    def _fwd_bwd(self, batch, tz: Tensor, mask: Optional[Tensor] = None, weight: Optional[Tensor] = None):
        self.bucket.detach_grads()
        self._trunk_reduced = False
        self.model.base.build_plans(batch)                 # CSR/CSC sorts of this batch: part of every step
        if self.feature_jitter_std > 0.0 and self.model.training:
            batch = self._jittered(batch)
        base = self.model.base
        state = {"deferred": False}
        base._defer_angle_blocks, base._defer_state = self.defer_angle_blocks, state    # read by the forward when it builds
        try:                                                                          # the trunk program ...
            mean, logvar = self.model(batch)
        finally:
            base._defer_angle_blocks, base._defer_state = 0, None    # ... and by nobody else: a plain backward() joins itself
        loss = fused_gaussian_nll(mean, logvar, tz, self.log_sigma_l2, self.floor, mask=mask, sample_weight=weight)
        (loss * self.loss_scale).backward()
        if state["deferred"]:                              # join the deferred angle-gradient kernel (trunk.py)
            from . import trunk as _trunk
            torch.cuda.current_stream().wait_stream(_trunk._side_stream(self.dev))
        self.bucket.gather()                               # one multi-tensor copy into the flat gradient bucket
        return loss.detach(), mean.detach(), logvar.detach()

    def _jittered(self, batch):
        """``x`` and ``global_x`` plus N(0, std^2) noise (reference ``train.py:641-646``) on a shallow copy of the batch:
        the caller's tensors (and a CUDA graph's static input buffers) are never modified; plans stay cached."""
        import copy
        std = self.feature_jitter_std
        b = copy.copy(batch)
        b.x = batch.x + torch.randn_like(batch.x) * std
        b.global_x = batch.global_x + torch.randn_like(batch.global_x) * std
        return b

    def _mark_trunk_reduced(self) -> None:
        self._trunk_reduced = True

    def _all_reduce(self) -> None:
        """The step's gradient exchange.  When the trunk backward has already reduced its folded gradients (see __init__)
        only the slices of the other parameters are left; otherwise (per-block path, fp32 regime, ...) the whole bucket."""
        if self._trunk_reduced and self._rest_slices is not None:
            for a, b in self._rest_slices:
                dist.all_reduce(self.bucket.flat[a:b], op=dist.ReduceOp.SUM, group=self.group)
        else:
            self.bucket.all_reduce(self.group)
        self._trunk_reduced = False

    def _finish(self):
        if self.opt is not None:
            self.opt.step()
        self.rng_step.add_(1)

    def _eager(self, batch, tz, mask=None, weight=None):
        prev, ops.RNG_STEP = ops.RNG_STEP, self.rng_step
        try:
            out = self._fwd_bwd(batch, tz, mask, weight)
            if self.world > 1:
                self._all_reduce()
            self._finish()
        finally:
            ops.RNG_STEP = prev
        self.eager_steps += 1
        return out

    # -- capture ---------------------------------------------------------------------------------------------------
    @staticmethod
    def signature(batch) -> Tuple:
        return tuple((k, tuple(v.shape), v.dtype) for k, v in sorted(batch.tensors().items())) + (
            batch.num_graphs, getattr(batch, "lg_active_rows", None), getattr(batch, "source_sorted", None),
            bool(getattr(batch, "padded", False)))

    def _capture(self, batch: GraphBatch, tz: Tensor, mask: Optional[Tensor] = None,
                 weight: Optional[Tensor] = None) -> _Captured:
        cap = _Captured()
        cap.batch = batch._like()
        for k in GraphBatch._TENSORS:
            v = getattr(batch, k)
            setattr(cap.batch, k, v.clone() if isinstance(v, Tensor) else v)
        cap.tz = tz.clone()
        cap.mask = None if mask is None else mask.clone()
        cap.weight = None if weight is None else weight.clone()
        prev, ops.RNG_STEP = ops.RNG_STEP, self.rng_step
        torch.cuda.synchronize(self.dev)
        k0 = ops.STATS.kernels
        try:
            one_graph = self.world == 1 or self.allreduce_in_graph
            cap.graph_opt = None
            try:
                cap.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cap.graph):
                    cap.loss, cap.mean, cap.logvar = self._fwd_bwd(cap.batch, cap.tz, cap.mask, cap.weight)
                    if one_graph:
                        if self.world > 1:
                            self._all_reduce()                      # NCCL kernel node(s) inside the graph
                        self._finish()
            except ops.StaticDropoutUnderCapture:
                raise
            except RuntimeError:
                if self.world == 1 or not self.allreduce_in_graph:
                    raise
                self.allreduce_in_graph, one_graph = False, False   # collective refused capture: keep it between two graphs
                torch.cuda.synchronize(self.dev)
                cap.graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cap.graph):
                    cap.loss, cap.mean, cap.logvar = self._fwd_bwd(cap.batch, cap.tz, cap.mask, cap.weight)
            cap.trunk_reduced = self._trunk_reduced
            if not one_graph:
                cap.graph_opt = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cap.graph_opt, pool=cap.graph.pool()):
                    self._finish()
        finally:
            ops.RNG_STEP = prev
        cap.kernels = ops.STATS.kernels - k0       # hand-written kernels inside the graph(s): relaunched by every replay
        ops.STATS.kernels = k0
        return cap

    def static_inputs(self, batch_like, masked: bool = False):
        """``(batch, target_z)`` input buffers of the CUDA graph captured for ``batch_like``'s signature, or ``None`` before
        capture.  A loader may write the next batch straight into them (e.g. the H2D copy on a side stream, once the
        previous replay of this graph has finished) and pass them to :meth:`step`, which then copies nothing."""
        cap = self._captured.get(self.signature(batch_like) + (bool(masked),)) if isinstance(batch_like, GraphBatch) else None
        return None if cap is None else (cap.batch, cap.tz)

    def step(self, batch, target_z: Tensor, mask: Optional[Tensor] = None, sample_weight: Optional[Tensor] = None):
        """``sample_weight`` (``[B]``): the per-sample loss weights of the reference's KNN weighting (``train.py:661-675``).
        ``mask`` (``[B]``, 1 = real graph) restricts the loss to real graphs.  With ``pad_to_buckets`` a ``GraphBatch``
        is padded to its shape bucket first (``batching.pad_batch``), so ragged datasets replay one graph per bucket; the
        returned ``mean / logvar`` then have the bucket's graph count (real graphs first)."""
        if self.pad_to_buckets and isinstance(batch, GraphBatch) and not getattr(batch, "padded", False):
            n_real = batch.num_graphs
            batch, mask = batching.pad_batch(batch, align=self.bucket_align)
            tz = torch.zeros(batch.num_graphs, target_z.size(1), dtype=target_z.dtype, device=target_z.device)
            tz[:n_real] = target_z
            target_z = tz
            if sample_weight is not None:
                w = torch.ones(batch.num_graphs, dtype=sample_weight.dtype, device=sample_weight.device)
                w[:n_real] = sample_weight
                sample_weight = w
        if mask is None and getattr(batch, "padded", False):
            # a batch padded elsewhere (DeviceGraphStore.collate(pad_to_bucket=True)) carries its own real-graph mask:
            # the dummy graph and the empty slots must never enter the loss
            mask = getattr(batch, "loss_mask", None)
            if mask is None:
                raise ValueError("padded batch without a loss mask: pass mask= (1 = real graph) or set batch.loss_mask")
        if not self.use_graph or not isinstance(batch, GraphBatch):
            return self._eager(batch, target_z, mask, sample_weight)
        sig = self.signature(batch) + (mask is not None,) + ((True,) if sample_weight is not None else ())
        cap = self._captured.get(sig)
        if cap is None:
            seen = self._seen.get(sig, 0)
            self._seen[sig] = seen + 1
            if seen < self.graph_warmup:
                return self._eager(batch, target_z, mask, sample_weight)
            try:
                cap = self._captured[sig] = self._capture(batch, target_z, mask, sample_weight)
            except ops.StaticDropoutUnderCapture as exc:
                # this model runs on a kernel family whose dropout keys are host integers (hidden != 256, fp32 regime):
                # a captured graph would replay ONE mask for ever.  Train it eagerly instead (fresh keys every step).
                import warnings
                warnings.warn(f"TrainStep: CUDA-graph replay disabled for this model ({exc})", stacklevel=2)
                self.use_graph = False
                torch.cuda.synchronize(self.dev)
                return self._eager(batch, target_z, mask, sample_weight)
        if batch is not cap.batch:                      # callers may fill the graph's own input buffers (static_inputs)
            for k, v in cap.batch.tensors().items():
                v.copy_(getattr(batch, k), non_blocking=True)
        if target_z is not cap.tz:
            cap.tz.copy_(target_z, non_blocking=True)
        if mask is not None:
            cap.mask.copy_(mask, non_blocking=True)
        if sample_weight is not None:
            cap.weight.copy_(sample_weight, non_blocking=True)
        cap.graph.replay()
        if cap.graph_opt is not None:
            self._trunk_reduced = cap.trunk_reduced
            self._all_reduce()
            cap.graph_opt.replay()
        self.replays += 1
        ops.STATS.kernels += cap.kernels
        return cap.loss, cap.mean, cap.logvar


class EnsemblePredictor:
    """All ensemble members on one batch + mixture moments as ONE replayed CUDA graph per batch signature -- the member
    loops of ``ensemble_collect`` (reference ``scripts/train.py:876-894``), ``predict.ensemble_predict``
    (``scripts/predict.py:604-623``) and ``evaluate.collect_member_predictions`` (``scripts/evaluate.py:244-261``).

    The graph plans are built once per batch and shared by the members.  ``compute_dtype``: ``torch.float32`` is what
    ``predict.py`` / ``evaluate.py`` run; ``torch.bfloat16`` is what ``eval_epoch_hetero`` runs (``train.py:757``) and is
    ~6x faster here (tensor-core kernels).  Returns ``(mean_z, var_z, std_z)``, each ``[B, T]`` fp32 (real graphs only
    when the batch had to be padded)."""

    def __init__(self, models, compute_dtype: torch.dtype = torch.float32, graph: bool = True, graph_warmup: int = 1,
                 pad_to_buckets: bool = False, bucket_align: int = 256, min_logvar_floor: float = -2.9):
        from . import ensemble
        if not models:
            raise ValueError("no ensemble members")
        self._ensemble = ensemble
        self.models = list(models)
        for m in self.models:
            m.eval()
            m.base.compute_dtype = compute_dtype
        self.dev = next(self.models[0].parameters()).device
        self.use_graph, self.graph_warmup = bool(graph), int(graph_warmup)
        self.pad_to_buckets, self.bucket_align, self.floor = bool(pad_to_buckets), int(bucket_align), float(min_logvar_floor)
        self._captured, self._seen = {}, {}
        self.replays = self.eager_calls = 0

    @torch.no_grad()
    def _forward(self, batch):
        self.models[0].base.build_plans(batch)
        return self._ensemble.ensemble_forward(self.models, batch, self.floor)

    @torch.no_grad()
    def predict(self, batch):
        n_real = batch.num_graphs
        if self.pad_to_buckets and isinstance(batch, GraphBatch) and not getattr(batch, "padded", False):
            batch, _ = batching.pad_batch(batch, align=self.bucket_align)
        if not self.use_graph or not isinstance(batch, GraphBatch):
            self.eager_calls += 1
            return tuple(t[:n_real] for t in self._forward(batch))
        sig = TrainStep.signature(batch)
        cap = self._captured.get(sig)
        if cap is None:
            seen = self._seen.get(sig, 0)
            self._seen[sig] = seen + 1
            if seen < self.graph_warmup:
                self.eager_calls += 1
                return tuple(t[:n_real] for t in self._forward(batch))
            static = batch._like()
            for k in GraphBatch._TENSORS:
                v = getattr(batch, k)
                setattr(static, k, v.clone() if isinstance(v, Tensor) else v)
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = self._forward(static)
            cap = self._captured[sig] = (g, static, out)
        g, static, out = cap
        for k, v in static.tensors().items():
            v.copy_(getattr(batch, k), non_blocking=True)
        g.replay()
        self.replays += 1
        return tuple(t[:n_real] for t in out)
