"""Bond features and line-graph construction on the device.

Replaces the two per-structure Python loops of the reference's featuriser, ``build_graph_from_structure``
(``scripts/fetch.py:385-396`` bonds -> ``edge_attr``; ``:417-447`` line graph -> ``lg_edge_index`` / ``lg_edge_attr``) for
one structure or for many concatenated ones, through ``alignn_bond_features`` / ``alignn_linegraph_count`` /
``alignn_linegraph_fill`` (``csrc/dataprep.cu``): float64 arithmetic in the reference's order, float32 / int64 results.
The neighbour list itself (``fetch.py:189-247``: pymatgen ``get_neighbors`` / CrystalNN) stays outside -- it is the input.

No CPU path: inputs must live on a CUDA device.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib, ops

_P = ops._p


def default_basis(rbf_n: int = 32, rbf_cutoff: float = 8.0, angle_n: int = 8, rbf_gamma: Optional[float] = None):
    """Centres and widths exactly as ``fetch.py:702-708`` builds them (``--rbf-n 32 --rbf-cutoff 8.0 --angle-n 8``)."""
    rbf_centers = torch.from_numpy(np.linspace(0.0, rbf_cutoff, rbf_n))        # np.linspace, as the reference
    if rbf_gamma is None:
        spacing = (rbf_cutoff - 0.0) / max(1, rbf_n - 1)
        rbf_gamma = float(1.0 / (spacing + 1e-8) ** 2)
    angle_centers = torch.from_numpy(np.linspace(0.0, math.pi, angle_n))
    angle_gamma = float((angle_n - 1) / (math.pi + 1e-8)) ** 2
    return rbf_centers, float(rbf_gamma), angle_centers, angle_gamma


def build_bond_and_line_graph(frac: Tensor, lattice: Tensor, en: Tensor, bond_src: Tensor, bond_dst: Tensor,
                              bond_image: Tensor, rbf_centers: Tensor, rbf_gamma: float, angle_centers: Tensor,
                              angle_gamma: float, atom_graph: Optional[Tensor] = None,
                              graph_bond_ptr: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """``frac [A,3] f64``, ``lattice [G,3,3] f64`` (rows = lattice vectors), ``en [A] f64``, bonds ``(src, dst, image)`` i-major
    (``int64 [E]``, ``int64 [E]``, ``int32 [E,3]``; global atom ids when several structures are concatenated, with
    ``atom_graph [A]`` naming each atom's structure and ``graph_bond_ptr [G+1]`` its first bond).  Returns ``edge_index``
    (global atom ids), ``edge_attr [E, n_rbf+4]``, ``lg_edge_index [2, L]`` (bond ids local to each structure),
    ``lg_edge_attr [L, n_ang+3]``, ``dirv [E,3] f64`` and ``angle_ptr [E+1]``.  One host sync (the angle count)."""
    for t in (frac, lattice, en, bond_src, bond_dst, bond_image):
        if not t.is_cuda:
            raise RuntimeError("build_bond_and_line_graph: inputs must be CUDA tensors (no CPU fallback path)")
    dev = frac.device
    f64 = lambda t: t.to(dev, torch.float64).contiguous()         # noqa: E731
    frac, lattice, en = f64(frac).reshape(-1, 3), f64(lattice).reshape(-1, 9), f64(en).reshape(-1)
    rbf_centers, angle_centers = f64(rbf_centers), f64(angle_centers)
    src, dst = bond_src.to(torch.int64).contiguous(), bond_dst.to(torch.int64).contiguous()
    img = bond_image.to(torch.int32).contiguous().reshape(-1, 3)
    n_atoms, n_bonds = int(frac.size(0)), int(src.numel())
    if atom_graph is not None:
        atom_graph = atom_graph.to(dev, torch.int64).contiguous()
    elif lattice.size(0) != 1:
        raise ValueError("several lattices need atom_graph")
    if graph_bond_ptr is not None:
        graph_bond_ptr = graph_bond_ptr.to(dev, torch.int64).contiguous()
    if n_bonds > 1 and bool((src[1:] < src[:-1]).any()):
        raise ValueError("bonds must be emitted source-major (i-major), as fetch.py:189-207 does")
    n_rbf, n_ang = int(rbf_centers.numel()), int(angle_centers.numel())
    lib = _lib.load()
    dirv = torch.empty(n_bonds, 3, dtype=torch.float64, device=dev)
    edge_attr = torch.empty(n_bonds, n_rbf + 4, dtype=torch.float32, device=dev)
    counts = torch.zeros(n_bonds, dtype=torch.int64, device=dev)
    out_ptr = torch.searchsorted(src, torch.arange(n_atoms + 1, device=dev, dtype=torch.int64)).contiguous()
    ag = _P(atom_graph) if atom_graph is not None else None
    with torch.cuda.device(dev), ops._Launch("bond_features", 2, (n_bonds,)):
        rc = lib.alignn_bond_features(_P(frac), _P(lattice), ag, _P(en), _P(src), _P(dst), _P(img), n_bonds, _P(rbf_centers),
                                      n_rbf, float(rbf_gamma), _P(dirv), _P(edge_attr), ops._stream())
        _lib.check(rc, "alignn_bond_features")
        rc = lib.alignn_linegraph_count(_P(src), _P(dst), _P(img), _P(out_ptr), n_bonds, _P(counts), ops._stream())
        _lib.check(rc, "alignn_linegraph_count")
    angle_ptr = torch.zeros(n_bonds + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=angle_ptr[1:])
    n_angles = int(angle_ptr[-1].item()) if n_bonds else 0
    lg_index = torch.empty(2, n_angles, dtype=torch.int64, device=dev)
    lg_attr = torch.empty(n_angles, n_ang + 3, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), ops._Launch("linegraph_fill", 1, (n_bonds, n_angles)):
        rc = lib.alignn_linegraph_fill(_P(frac), _P(lattice), ag, _P(graph_bond_ptr) if graph_bond_ptr is not None else None,
                                       _P(src), _P(dst), _P(img), _P(out_ptr), _P(dirv), _P(angle_ptr), n_bonds,
                                       _P(angle_centers), n_ang, float(angle_gamma), _P(lg_index), n_angles, _P(lg_attr),
                                       ops._stream())
    _lib.check(rc, "alignn_linegraph_fill")
    return {"edge_index": torch.stack([src, dst]), "edge_attr": edge_attr, "lg_edge_index": lg_index,
            "lg_edge_attr": lg_attr, "dirv": dirv, "angle_ptr": angle_ptr}
