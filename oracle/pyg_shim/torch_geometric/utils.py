"""``torch_geometric.utils.{softmax, scatter}`` restated (oracle; test infrastructure).

Follows the published PyG 2.7.0 algorithm:

* ``scatter(src, index, dim=0, dim_size, reduce)`` -- ``index_add_`` for
  ``sum``; ``scatter_reduce_('amax', include_self=False)`` for ``max`` (rows
  that receive nothing stay 0).
* ``softmax(src, index, num_nodes)`` -- shift by the per-segment ``amax`` of
  ``src.detach()``, ``exp``, segment sum **+ 1e-16**, divide.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor


def maybe_num_nodes(index: Tensor, num_nodes: Optional[int] = None) -> int:
    if num_nodes is not None:
        return int(num_nodes)
    return int(index.max()) + 1 if index.numel() > 0 else 0


def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None,
            reduce: str = "sum") -> Tensor:
    if dim != 0:
        raise NotImplementedError("shim scatter only supports dim=0")
    n = maybe_num_nodes(index, dim_size)
    out_shape = (n,) + tuple(src.shape[1:])
    if reduce in ("sum", "add"):
        out = src.new_zeros(out_shape)
        return out.index_add_(0, index, src)
    if reduce == "mean":
        out = src.new_zeros(out_shape).index_add_(0, index, src)
        cnt = src.new_zeros(n).index_add_(0, index, src.new_ones(index.numel()))
        cnt = cnt.clamp(min=1).view((n,) + (1,) * (src.dim() - 1))
        return out / cnt
    if reduce in ("max", "amax"):
        out = src.new_zeros(out_shape)
        idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
        return out.scatter_reduce_(0, idx, src, reduce="amax", include_self=False)
    raise NotImplementedError(reduce)


def softmax(src: Tensor, index: Tensor, ptr=None, num_nodes: Optional[int] = None, dim: int = 0) -> Tensor:
    n = maybe_num_nodes(index, num_nodes)
    src_max = scatter(src.detach(), index, 0, dim_size=n, reduce="max")
    out = src - src_max.index_select(0, index)
    out = out.exp()
    out_sum = scatter(out, index, 0, dim_size=n, reduce="sum") + 1e-16
    out_sum = out_sum.index_select(0, index)
    return out / out_sum
