"""``torch_geometric.nn.{TransformerConv, global_mean_pool}`` restated (oracle; test infrastructure).

Algorithm restated from the published PyG 2.7.0 ``TransformerConv``
(``torch_geometric/nn/conv/transformer_conv.py``, pinned by
``/root/reference/requirements.txt:9``; source absent from ``/root/reference``)
and anchored on the reference's call sites ``scripts/train.py:308,315,326,334``
(``TransformerConv(hidden, hidden // heads, heads=heads, edge_dim=hidden,
dropout=dropout, beta=True)`` called as ``conv(x, edge_index, edge_attr)``) and
``scripts/train.py:388,562`` (``global_mean_pool(node_state, data.batch)``).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .utils import scatter, softmax


class TransformerConv(nn.Module):
    """Graph transformer operator (Shi et al. 2021) with edge features and beta-gated skip.

    ``flow='source_to_target'``: ``edge_index[0]`` = source ``j``, ``edge_index[1]``
    = target ``i`` (aggregation index).  Sub-module registration order mirrors
    PyG: ``lin_key, lin_query, lin_value, lin_edge, lin_skip, lin_beta``.
    """

    def __init__(self, in_channels: int, out_channels: int, heads: int = 1, concat: bool = True,
                 beta: bool = False, dropout: float = 0.0, edge_dim: Optional[int] = None,
                 bias: bool = True, root_weight: bool = True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.heads = heads
        self.beta = beta and root_weight
        self.root_weight = root_weight
        self.concat = concat
        self.dropout = dropout
        self.edge_dim = edge_dim

        self.lin_key = nn.Linear(in_channels, heads * out_channels)
        self.lin_query = nn.Linear(in_channels, heads * out_channels)
        self.lin_value = nn.Linear(in_channels, heads * out_channels)
        if edge_dim is not None:
            self.lin_edge = nn.Linear(edge_dim, heads * out_channels, bias=False)
        else:
            self.lin_edge = self.register_parameter("lin_edge", None)
        if concat:
            self.lin_skip = nn.Linear(in_channels, heads * out_channels, bias=bias)
            if self.beta:
                self.lin_beta = nn.Linear(3 * heads * out_channels, 1, bias=False)
            else:
                self.lin_beta = self.register_parameter("lin_beta", None)
        else:
            self.lin_skip = nn.Linear(in_channels, out_channels, bias=bias)
            if self.beta:
                self.lin_beta = nn.Linear(3 * out_channels, 1, bias=False)
            else:
                self.lin_beta = self.register_parameter("lin_beta", None)

    def forward(self, x: Tensor, edge_index: Tensor, edge_attr: Optional[Tensor] = None) -> Tensor:
        H, C = self.heads, self.out_channels
        query = self.lin_query(x).view(-1, H, C)
        key = self.lin_key(x).view(-1, H, C)
        value = self.lin_value(x).view(-1, H, C)

        src, dst = edge_index[0], edge_index[1]
        n_dst = x.size(0)
        # --- message ---
        query_i = query.index_select(0, dst)
        key_j = key.index_select(0, src)
        value_j = value.index_select(0, src)
        e = None
        if self.lin_edge is not None:
            assert edge_attr is not None
            e = self.lin_edge(edge_attr).view(-1, H, C)
            key_j = key_j + e
        alpha = (query_i * key_j).sum(dim=-1) / math.sqrt(C)
        alpha = softmax(alpha, dst, None, n_dst)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        msg = value_j
        if e is not None:
            msg = msg + e
        msg = msg * alpha.view(-1, H, 1)
        # --- aggregate ('add') ---
        out = scatter(msg, dst, 0, dim_size=n_dst, reduce="sum")

        if self.concat:
            out = out.view(-1, H * C)
        else:
            out = out.mean(dim=1)

        if self.root_weight:
            x_r = self.lin_skip(x)
            if self.lin_beta is not None:
                beta = self.lin_beta(torch.cat([out, x_r, out - x_r], dim=-1)).sigmoid()
                out = beta * x_r + (1 - beta) * out
            else:
                out = out + x_r
        return out


def global_mean_pool(x: Tensor, batch: Optional[Tensor], size: Optional[int] = None) -> Tensor:
    if batch is None:
        return x.mean(dim=0, keepdim=True)
    n = int(batch.max()) + 1 if size is None else size
    return scatter(x, batch, 0, dim_size=n, reduce="mean")
