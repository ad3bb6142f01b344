"""CPU oracle: own-code restatement of the reference's ALIGNN model classes.

TEST INFRASTRUCTURE ONLY -- imported by ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs; never by the
product package.

PARITY UNPINNED by the reference's own tests (it has no numeric assertions
for this path); pinned here by (1) bit-identical agreement with the
reference's own classes imported from ``/root/reference/scripts/train.py``
wherever that tree is mounted (``tests/test_oracle_vs_reference.py``), (2) the
committed golden vectors under ``tests/golden/`` generated from those classes
by ``oracle/gen_golden.py``, and (3) the known-answer tests on the PyG shim.

What is restated (reference file:line):

* ``EdgeUpdateBlock``      -- ``scripts/train.py:303-317``
* ``NodeUpdateBlock``      -- ``scripts/train.py:320-336``
* ``AlignnRegressor``      -- ``scripts/train.py:339-401``
* ``HeteroAlignnRegressor``-- ``scripts/train.py:528-586``
* Gaussian NLL train loss  -- ``scripts/train.py:655-681``
* ensemble mixture moments -- ``scripts/train.py:876-894`` / ``scripts/predict.py:604-616``

The conv arithmetic comes from the PyG shim (``oracle/pyg_shim``), i.e. the
same leaf ops the reference classes use when run under the shim.
"""
from __future__ import annotations

import os
import sys
from typing import List, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor, nn

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pyg_shim")
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)

from torch_geometric.nn import TransformerConv, global_mean_pool  # noqa: E402  (the shim)

MIN_LOGVAR_FLOOR = -2.9  # scripts/train.py:39


def _two_layer(d_in: int, width: int) -> nn.Sequential:
    return nn.Sequential(nn.Linear(d_in, width), nn.ReLU(), nn.Linear(width, width))


def _check_heads(width: int, heads: int) -> None:
    if width % heads != 0:
        raise ValueError("hidden size must be divisible by number of heads")


class EdgeUpdateBlock(nn.Module):
    """Line-graph conv: bonds are nodes, angles are edges."""

    def __init__(self, hidden: int, heads: int, dropout: float):
        super().__init__()
        _check_heads(hidden, heads)
        self.conv = TransformerConv(hidden, hidden // heads, heads=heads, edge_dim=hidden,
                                    dropout=dropout, beta=True)
        self.norm = nn.LayerNorm(hidden)
        self.dropout = nn.Dropout(dropout)

    def forward(self, bond_state: Tensor, lg_index: Tensor, angle_emb: Tensor) -> Tensor:
        if min(bond_state.numel(), angle_emb.numel(), lg_index.numel()) == 0:
            return bond_state
        upd = self.norm(self.conv(bond_state, lg_index, angle_emb))
        return bond_state + self.dropout(F.relu(upd))


class NodeUpdateBlock(nn.Module):
    """Atom-graph conv with projected bond states as edge attributes."""

    def __init__(self, hidden_node: int, hidden_edge: int, heads: int, dropout: float):
        super().__init__()
        _check_heads(hidden_node, heads)
        self.edge_proj = nn.Linear(hidden_edge, hidden_edge)
        self.conv = TransformerConv(hidden_node, hidden_node // heads, heads=heads,
                                    edge_dim=hidden_edge, dropout=dropout, beta=True)
        self.norm = nn.LayerNorm(hidden_node)
        self.dropout = nn.Dropout(dropout)

    def forward(self, atom_state: Tensor, index: Tensor, bond_state: Tensor) -> Tensor:
        if min(bond_state.numel(), index.numel()) == 0:
            return atom_state
        upd = self.norm(self.conv(atom_state, index, self.edge_proj(bond_state)))
        return atom_state + self.dropout(F.relu(upd))


def _trunk(base: "AlignnRegressor", data) -> Tensor:
    """Encoders -> interleaved blocks -> mean pool -> concat globals -> feat_proj."""
    atoms = base.node_encoder(data.x)
    width = atoms.size(-1)
    dev = data.x.device
    if data.edge_attr.numel() > 0:
        bonds = base.edge_encoder(data.edge_attr)
    else:
        bonds = torch.zeros(data.edge_index.size(1), width, device=dev)
    if base.angle_encoder is not None and data.lg_edge_attr.numel() > 0:
        angles = base.angle_encoder(data.lg_edge_attr)
    else:
        angles = torch.zeros(data.lg_edge_index.size(1), width, device=dev)
    for eb, nb in zip(base.edge_blocks, base.node_blocks):
        bonds = eb(bonds, data.lg_edge_index, angles)
        atoms = nb(atoms, data.edge_index, bonds)
    pooled = global_mean_pool(atoms, data.batch)
    n_graphs = pooled.size(0)

    def _per_graph(t: Tensor) -> Tensor:
        if t.dim() == 1:
            t = t.unsqueeze(0)
        return t.reshape(n_graphs, -1)

    feats = torch.cat([pooled, _per_graph(data.global_x), _per_graph(data.sg_one_hot)], dim=1)
    return base.feat_proj(base.dropout(feats))


class AlignnRegressor(nn.Module):
    def __init__(self, node_dim: int, edge_dim: int, angle_dim: int, global_dim: int, target_dim: int,
                 hidden: int, layers: int, heads: int, dropout: float):
        super().__init__()
        if heads <= 0:
            raise ValueError("heads must be positive")
        if target_dim <= 0:
            raise ValueError("target_dim must be positive")
        _check_heads(hidden, heads)
        self.hidden = hidden
        self.heads = heads
        self.node_encoder = _two_layer(node_dim, hidden)
        self.edge_encoder = _two_layer(edge_dim, hidden)
        self.angle_encoder = _two_layer(angle_dim, hidden) if angle_dim > 0 else None
        self.edge_blocks = nn.ModuleList(EdgeUpdateBlock(hidden, heads, dropout) for _ in range(layers))
        self.node_blocks = nn.ModuleList(NodeUpdateBlock(hidden, hidden, heads, dropout) for _ in range(layers))
        self.dropout = nn.Dropout(dropout)
        self.feat_proj = nn.Sequential(nn.Linear(hidden + global_dim, hidden), nn.ReLU(), nn.Dropout(dropout))
        self.output_heads = nn.ModuleList(nn.Linear(hidden, 1) for _ in range(target_dim))

    def forward(self, data) -> Tensor:
        shared = _trunk(self, data)
        return torch.cat([h(shared) for h in self.output_heads], dim=1)


class HeteroAlignnRegressor(nn.Module):
    def __init__(self, base: AlignnRegressor, target_dim: int):
        super().__init__()
        self.base = base
        width = base.feat_proj[0].out_features
        self.mean_heads = nn.ModuleList(nn.Linear(width, 1) for _ in range(target_dim))
        self.logvar_heads = nn.ModuleList(nn.Linear(width, 1) for _ in range(target_dim))

    def embed(self, data) -> Tensor:
        return _trunk(self.base, data)

    def forward(self, data) -> Tuple[Tensor, Tensor]:
        shared = _trunk(self.base, data)
        mean = torch.cat([h(shared) for h in self.mean_heads], dim=1)
        logvar = torch.cat([h(shared) for h in self.logvar_heads], dim=1)
        return mean, logvar


def gaussian_nll_loss(mean: Tensor, logvar: Tensor, target_z: Tensor, log_sigma_l2: float = 0.1,
                      min_logvar_floor: float = MIN_LOGVAR_FLOOR) -> Tensor:
    """Training loss of ``train_epoch_hetero`` without sample weights (train.py:655-681)."""
    lv = torch.clamp(logvar, min=min_logvar_floor)
    diff = mean - target_z.to(mean.dtype)
    nll = 0.5 * (lv + diff.pow(2) / torch.exp(lv))
    loss = nll.mean(dim=1).mean()
    if log_sigma_l2 > 0.0:
        loss = loss + float(log_sigma_l2) * (0.5 * lv).pow(2).mean()
    return loss


def ensemble_moments(member_means: Sequence[Tensor], member_logvars: Sequence[Tensor],
                     min_logvar_floor: float = MIN_LOGVAR_FLOOR) -> Tuple[Tensor, Tensor, Tensor]:
    """Mixture mean / variance / std in z-space (train.py:876-894, predict.py:604-616)."""
    mu = torch.stack(list(member_means), dim=0)
    var = torch.stack([torch.exp(torch.clamp(lv, min=min_logvar_floor)) for lv in member_logvars], dim=0)
    mean_z = mu.mean(dim=0)
    var_z = var.mean(dim=0) + mu.pow(2).mean(dim=0) - mean_z.pow(2)
    std_z = torch.sqrt(torch.clamp(var_z, min=1e-12))
    return mean_z, var_z, std_z


def build_hetero(node_dim=206, edge_dim=36, angle_dim=11, global_dim=289, target_dim=2, hidden=256,
                 layers=4, heads=4, dropout=0.0, seed: int = 42) -> HeteroAlignnRegressor:
    """Default-arch member (train.py:1084-1086,1489-1505) with seeded default ``nn.Linear`` init."""
    torch.manual_seed(seed)
    base = AlignnRegressor(node_dim, edge_dim, angle_dim, global_dim, target_dim, hidden, layers, heads, dropout)
    return HeteroAlignnRegressor(base, target_dim)
