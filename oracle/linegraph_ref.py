"""CPU restatement of the reference's bond / line-graph featuriser -- TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``/root/reference/scripts/fetch.py``:

* ``edge_geom``             -- ``_edge_geom``               (``fetch.py:250-263``)
* ``angle_between``         -- ``_angle_between_vectors``   (``fetch.py:266-273``)
* ``rbf_expand``            -- ``_rbf_expand``              (``fetch.py:311-316``)
* ``basis``                 -- centres / widths             (``fetch.py:702-708``)
* ``build_bond_and_line_graph`` -- the two loops of ``build_graph_from_structure``: bonds (``fetch.py:385-396``) and the
  line graph (``fetch.py:417-447``), and the tensor conversion of ``to_pyg_data`` (``fetch.py:629-633``: ``float`` lists ->
  ``torch.float`` / ``torch.long``).

Plain numpy / Python loops (float64, exactly the reference's operation order).  Pinned by ``tests/golden/linegraph_*.pt``,
which ``oracle/gen_golden_linegraph.py`` produced by running the reference's OWN ``build_graph_from_structure`` on
duck-typed structures (pymatgen is not installed; only its attribute surface is faked, none of the arithmetic).
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch


def basis(rbf_n: int = 32, rbf_cutoff: float = 8.0, angle_n: int = 8):
    """``(rbf_centers, rbf_gamma, angle_centers, angle_gamma)`` -- fetch.py:702-708 (defaults: ``--rbf-n 32``,
    ``--rbf-cutoff``, ``--angle-n 8``)."""
    rbf_centers = np.linspace(0.0, rbf_cutoff, rbf_n)
    spacing = (rbf_cutoff - 0.0) / max(1, rbf_n - 1)
    rbf_gamma = float(1.0 / (spacing + 1e-8) ** 2)
    angle_centers = np.linspace(0.0, math.pi, angle_n)
    angle_gamma = float((angle_n - 1) / (math.pi + 1e-8)) ** 2
    return rbf_centers, rbf_gamma, angle_centers, angle_gamma


def edge_geom(frac: np.ndarray, lattice: np.ndarray, i: int, j: int, jimage: Sequence[int]):
    """Distance and unit direction i -> j (cartesian).  ``lattice.get_cartesian_coords(d)`` is ``d @ matrix``."""
    fi = np.asarray(frac[i], dtype=float)
    fj = np.asarray(frac[j], dtype=float)
    dfrac = (fj + np.asarray(jimage, dtype=float)) - fi
    vec_cart = np.asarray(np.dot(dfrac, lattice), dtype=float)
    dist = float(np.linalg.norm(vec_cart))
    dirv = tuple((vec_cart / dist).tolist()) if dist > 0 else (0.0, 0.0, 0.0)
    return dist, dirv


def angle_between(u: np.ndarray, v: np.ndarray) -> float:
    nu = np.linalg.norm(u)
    nv = np.linalg.norm(v)
    if nu == 0 or nv == 0:
        return 0.0
    cos_t = float(np.clip(np.dot(u, v) / (nu * nv), -1.0, 1.0))
    return float(math.acos(cos_t))


def rbf_expand(r: float, centers: np.ndarray, gamma: float) -> List[float]:
    return np.exp(-gamma * (r - centers) ** 2).astype(float).tolist()


def build_bond_and_line_graph(frac: np.ndarray, lattice: np.ndarray, en: Sequence[float],
                              edges: Sequence[Tuple[int, int, Tuple[int, int, int]]],
                              rbf_centers: np.ndarray, rbf_gamma: float, angle_centers: np.ndarray,
                              angle_gamma: float) -> Dict[str, torch.Tensor]:
    """``edges``: directed bonds ``(i, j, jimage)`` in the reference's emission order (i-major, fetch.py:189-207)."""
    n_atoms = len(frac)
    neigh_map: Dict[int, list] = {i: [] for i in range(n_atoms)}
    for i, j, jimage in edges:
        neigh_map[i].append((j, tuple(jimage)))
    edge_index: List[List[int]] = [[], []]
    edge_attr: List[List[float]] = []
    bond_nodes_map = {}
    for idx, (i, j, jimage) in enumerate(edges):                                    # fetch.py:389-398
        dist, dirv = edge_geom(frac, lattice, i, j, jimage)
        delta_en = abs(en[i] - en[j])
        rbf = rbf_expand(dist, rbf_centers, rbf_gamma)
        edge_index[0].append(int(i))
        edge_index[1].append(int(j))
        edge_attr.append(rbf + [float(delta_en), float(dirv[0]), float(dirv[1]), float(dirv[2])])
        bond_nodes_map[(i, j, tuple(jimage))] = idx
    lg_edge_index: List[List[int]] = [[], []]
    lg_edge_attr: List[List[float]] = []
    for i, j, jimage in edges:                                                      # fetch.py:421-447
        jimage = tuple(jimage)
        for k, kimage in neigh_map.get(j, []):
            rev_im = (-int(jimage[0]), -int(jimage[1]), -int(jimage[2]))
            if k == i and kimage == rev_im:
                continue
            _, dir_ji = edge_geom(frac, lattice, j, i, rev_im)
            _, dir_jk = edge_geom(frac, lattice, j, k, kimage)
            angle = angle_between(np.array(dir_ji), np.array(dir_jk))
            ang_feat = np.exp(-angle_gamma * (angle - angle_centers) ** 2).astype(float).tolist()
            e1 = bond_nodes_map.get((i, j, jimage))
            e2 = bond_nodes_map.get((j, k, kimage))
            if e1 is None or e2 is None:
                continue
            lg_edge_index[0].append(e1)
            lg_edge_index[1].append(e2)
            lg_edge_attr.append(ang_feat + [float(angle), float(math.cos(angle)), float(math.sin(angle))])
    n_ang = len(angle_centers) + 3
    return {
        "edge_index": torch.tensor(edge_index, dtype=torch.long).reshape(2, -1),
        "edge_attr": torch.tensor(edge_attr, dtype=torch.float).reshape(-1, len(rbf_centers) + 4),
        "lg_edge_index": torch.tensor(lg_edge_index, dtype=torch.long).reshape(2, -1),
        "lg_edge_attr": torch.tensor(lg_edge_attr, dtype=torch.float).reshape(-1, n_ang),
    }


def random_crystal(n_atoms: int, seed: int, cutoff: float = 3.2, box: float = 4.0, shell: int = 1):
    """A small random periodic cell + its directed bond list ``(i, j, jimage)``, i-major like fetch.py:189-207
    (brute-force neighbour search over the 27 (``shell=1``) nearest images; self-image bonds and repeated ``(i, j)`` pairs
    with different images occur, as in real crystals, fetch.py:195)."""
    rng = np.random.default_rng(seed)
    lattice = np.eye(3) * box + rng.normal(scale=0.35, size=(3, 3))
    frac = rng.random((n_atoms, 3))
    en = (0.8 + 3.0 * rng.random(n_atoms)).tolist()
    edges = []
    rng_im = range(-shell, shell + 1)
    for i in range(n_atoms):
        for j in range(n_atoms):
            for a in rng_im:
                for b in rng_im:
                    for c in rng_im:
                        if i == j and a == b == c == 0:
                            continue
                        d = np.linalg.norm(np.dot((frac[j] + np.array([a, b, c], dtype=float)) - frac[i], lattice))
                        if d < cutoff:
                            edges.append((i, j, (a, b, c)))
    return frac, lattice, en, edges
