"""Shape buckets for ragged batches: pad a collated batch to a bucket shape so that ``engine.TrainStep`` can replay ONE
captured CUDA graph per bucket instead of running real (variable-size) datasets eagerly.

The reference feeds whatever ``torch_geometric.loader.DataLoader`` collates (``scripts/train.py:2037``): every batch has
its own atom / bond / angle counts.  Padding must not change any real graph's result:

* padded ATOMS get zero features and belong to one extra dummy graph (index ``B``); padded BONDS and ANGLES get the index
  ``-1`` on both ends -- the graph plan drops out-of-range edges (``csrc/plan.cu``: sentinel key, never visited by any
  kernel) -- and zero features, so no real row ever sees them; padded bond ROWS are isolated in both graphs;
* the dummy graph (and any further empty graphs up to the bucket's graph count) is masked out of the loss
  (``modules.gaussian_nll_loss(..., mask=...)``), so it contributes no gradient;
* ``lg_active_rows`` (isolated-row bound) is unchanged: ``-1`` never raises the largest line-graph index.

Bucket grid: sizes are rounded up to a multiple of ``max(align, 2^(floor(log2 x) - 3))`` -- at most 12.5 % padding, about
eight buckets per octave.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from .synthetic import GLOBAL_SCALARS, SG_DIM, GraphBatch


def round_up_bucket(x: int, align: int = 256) -> int:
    """Smallest grid point >= x (grid step = 1/8 .. 1/16 of the size, at least ``align``)."""
    x = int(x)
    if x <= 0:
        return 0
    step = max(int(align), 1 << max(x.bit_length() - 4, 0))
    return (x + step - 1) // step * step


def bucket_shape(batch: GraphBatch, align: int = 256) -> Dict[str, int]:
    """Bucket sizes for a batch: one spare atom row / graph slot is always reserved for the dummy graph."""
    s = batch.sizes
    return {"N": round_up_bucket(s["N"] + 1, align), "E": round_up_bucket(s["E"], align),
            "L": round_up_bucket(s["L"], align), "B": round_up_bucket(s["B"] + 1, 8)}


def _pad_rows(t: Tensor, rows: int, value: float = 0.0) -> Tensor:
    if t.size(0) == rows:
        return t
    pad = torch.full((rows - t.size(0),) + tuple(t.shape[1:]), value, dtype=t.dtype, device=t.device)
    return torch.cat([t, pad], dim=0)


def _pad_index(t: Tensor, cols: int) -> Tensor:
    if t.size(1) == cols:
        return t
    pad = torch.full((2, cols - t.size(1)), -1, dtype=t.dtype, device=t.device)
    return torch.cat([t, pad], dim=1)


def pad_batch(batch: GraphBatch, shape: Optional[Dict[str, int]] = None, align: int = 256) -> Tuple[GraphBatch, Tensor]:
    """Returns ``(padded batch, mask)``; ``mask`` is ``[B_pad]`` fp32 with 1 for the real graphs.  Works on host or device
    tensors (the padded tensors live where the inputs live)."""
    s = batch.sizes
    shape = dict(shape) if shape is not None else bucket_shape(batch, align)
    n, e, l, b = shape["N"], shape["E"], shape["L"], shape["B"]
    if n < s["N"] or e < s["E"] or l < s["L"] or b < s["B"]:
        raise ValueError(f"bucket {shape} is smaller than the batch {s}")
    if (n > s["N"] or e > s["E"] or l > s["L"]) and b <= s["B"]:
        raise ValueError("padding needs one spare graph slot for the dummy graph (B_pad > B)")
    y = batch.y.reshape(s["B"], -1)
    dev = batch.x.device
    out = batch._like()
    out.num_graphs = b
    out.padded = True
    out.x = _pad_rows(batch.x, n)
    out.edge_attr = _pad_rows(batch.edge_attr, e)
    out.lg_edge_attr = _pad_rows(batch.lg_edge_attr, l)
    out.edge_index = _pad_index(batch.edge_index, e)
    out.lg_edge_index = _pad_index(batch.lg_edge_index, l)
    out.batch = _pad_rows(batch.batch, n, value=s["B"])                     # padded atoms -> the dummy graph
    out.global_x = _pad_rows(batch.global_x.reshape(s["B"], -1), b).reshape(-1, 1)
    out.sg_one_hot = _pad_rows(batch.sg_one_hot.reshape(s["B"], -1), b).reshape(-1, 1)
    out.y = _pad_rows(y, b, value=1.0).reshape(-1)                           # log(1) = 0: finite z-scores
    ti = batch.train_idx
    out.train_idx = _pad_rows(ti, b, value=-1) if isinstance(ti, Tensor) else ti
    if batch.lg_active_rows is not None:      # bucketed too (extra rows in the bound are merely isolated rows treated as active)
        out.lg_active_rows = min(e, round_up_bucket(batch.lg_active_rows, align))
    mask = torch.zeros(b, dtype=torch.float32, device=dev)
    mask[:s["B"]] = 1.0
    return out, mask
