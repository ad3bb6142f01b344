"""Data-parallel sharding of graph batches across the GPUs of one box (one process per GPU).

The reference has no distributed code at all (single device, ``scripts/train.py:1474``; members trained
sequentially, ``:2052-2095``).  Graphs in a batch are independent and LayerNorm has no cross-sample
statistics, so the path shards by graphs with ONE exchange per step: an all-reduce of the flat gradient
bucket (3.3 M parameters at H=256: 13 MB fp32) over NCCL/NVLink.  No activations or graph data cross GPUs.

* :func:`shard_ranges`      contiguous graph ranges per rank, balanced by line-graph edges
* :class:`FlatGradBucket`   all parameter gradients as views of one flat fp32 buffer -> one collective
* :func:`global_grad_clip`  ``clip_grad_norm_(5.0)`` (reference ``train.py:693,698``) on the reduced bucket,
                            computed identically on every rank
* :func:`member_placement`  ensemble members -> ranks (member-per-GPU training / inference)

Parity target: the N-rank result equals the 1-rank result on the same global batch (tests use ``gloo``
with world_size 2 on CPU for the host logic, ``lg_inc='bonds'`` so that sharding does not change the graph).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor, nn


def shard_ranges(weights: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """Split ``len(weights)`` graphs into ``world_size`` contiguous ranges with near-equal total weight
    (weight = line-graph edges of the graph, the quantity the conv kernels stream)."""
    n = len(weights)
    if world_size <= 0:
        raise ValueError("world_size must be positive")
    total = float(sum(weights))
    bounds = [0]
    acc = 0.0
    g = 0
    for r in range(1, world_size):
        target = total * r / world_size
        while g < n and acc + weights[g] / 2.0 <= target:
            acc += weights[g]
            g += 1
        # leave at least one graph for every remaining rank when possible
        g = min(g, n - (world_size - r)) if n >= world_size else min(g, n)
        g = max(g, bounds[-1])
        bounds.append(g)
    bounds.append(n)
    return [(bounds[i], bounds[i + 1]) for i in range(world_size)]


class FlatGradBucket:
    """Points every ``param.grad`` at a slice of one flat fp32 buffer (pre-allocated, reused every step).

    ``zero()`` is one memset, ``all_reduce()`` is one collective, and the optimizer / clip see ordinary
    ``.grad`` tensors.  Parameters unused in a step simply contribute zeros.
    """

    def __init__(self, params: Iterable[nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        # 16-byte aligned slices so vectorised kernels can consume the bucket directly
        self.offsets = []
        off = 0
        for s in sizes:
            self.offsets.append(off)
            off += (s + 3) // 4 * 4
        self.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        self.attach()

    def attach(self) -> None:
        for p, off in zip(self.params, self.offsets):
            p.grad = self.flat[off:off + p.numel()].view_as(p)

    def detach_grads(self) -> None:
        """``param.grad = None`` for every parameter: autograd then hands over its gradient tensors instead of
        launching one accumulate kernel per parameter; :meth:`gather` moves them into the bucket in one fused copy."""
        for p in self.params:
            p.grad = None

    def gather(self) -> None:
        """Copy the gradients autograd produced into the flat bucket (one multi-tensor copy), zero the slices of
        parameters that received none, and re-point ``param.grad`` at the bucket."""
        views = [self.flat[off:off + p.numel()].view_as(p) for p, off in zip(self.params, self.offsets)]
        src, dst = [], []
        written = getattr(self, "_written", None)
        # under CUDA-graph capture the decision below is baked into the graph, and a graph captured for a signature whose
        # forward skips some parameters may later be replayed right after a graph that DID write their slices: there a
        # slice without a gradient is zeroed unconditionally
        capturing = self.flat.is_cuda and torch.cuda.is_current_stream_capturing()
        now = set()
        for i, (p, v) in enumerate(zip(self.params, views)):
            g = p.grad
            if g is not None and g.data_ptr() != v.data_ptr():
                src.append(g if g.dtype == v.dtype else g.to(v.dtype)); dst.append(v); now.add(i)
            elif g is None and (capturing or written is None or i in written):
                v.zero_()
        if src:
            torch._foreach_copy_(dst, src)
        self._written = now
        for p, v in zip(self.params, views):
            p.grad = v

    def zero(self) -> None:
        self.flat.zero_()
        self.attach()   # in case something replaced .grad (e.g. zero_grad(set_to_none=True))

    def all_reduce(self, group=None, async_op: bool = False):
        """Sum over ranks (losses are pre-scaled by B_local / B_global, so the sum is the global-batch gradient)."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)

    def nbytes(self) -> int:
        return self.flat.numel() * 4


def global_grad_clip(bucket: FlatGradBucket, max_norm: float = 5.0) -> Tensor:
    """``torch.nn.utils.clip_grad_norm_`` semantics on the flat bucket (one norm, one scale; no host sync)."""
    total = torch.linalg.vector_norm(bucket.flat)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    bucket.flat.mul_(coef)
    return total


def dp_loss_scale(n_local_graphs: int, n_global_graphs: int) -> float:
    """Local mean-loss weight so that summed gradients equal the global-batch mean gradient."""
    return float(n_local_graphs) / float(max(n_global_graphs, 1))


def member_placement(n_members: int, world_size: int) -> Dict[int, List[int]]:
    """rank -> member ids.  Members are independent (own seed ``seed + 1007*i``, reference ``train.py:2053``):
    with M <= world_size each member gets its own GPU (zero communication); otherwise round-robin."""
    out: Dict[int, List[int]] = {r: [] for r in range(world_size)}
    for m in range(n_members):
        out[m % world_size].append(m)
    return out


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; initialises the process group if needed."""
    import os
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend)
    return rank, local_rank, world


def shutdown(holders=(), grace_s: float = 20.0) -> None:
    """Tear the process group down without hanging.  CUDA graphs that captured an NCCL kernel keep the communicator
    busy: ``destroy_process_group`` (and interpreter exit) then blocks for ever (seen on 2 x B200 with the all-reduce
    captured inside the step graph).  So: drop every captured graph first (``holders``: objects with a ``_captured``
    dict, e.g. ``engine.TrainStep``), synchronise, and arm a watchdog that ends the process with exit code 0 if the
    teardown still does not return within ``grace_s`` seconds -- all results have been printed by then."""
    import gc
    import os
    import sys
    import threading
    for h in holders:
        cap = getattr(h, "_captured", None)
        if isinstance(cap, dict):
            cap.clear()
    gc.collect()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if not (dist.is_available() and dist.is_initialized()):
        return
    sys.stdout.flush()
    sys.stderr.flush()
    timer = threading.Timer(grace_s, lambda: os._exit(0))
    timer.daemon = True
    timer.start()
    try:
        dist.barrier()
        dist.destroy_process_group()
    finally:
        timer.cancel()
