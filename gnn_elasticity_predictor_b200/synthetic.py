"""Synthetic crystal graphs in the reference's ``Data`` / PyG-``Batch`` layout.

The generator restates the *index conventions* of the reference's featuriser
(``/root/reference/scripts/fetch.py``) -- not its chemistry:

* bonds are emitted source-major, ``edge_index[0] = i`` (source), ``edge_index[1] = j``
  (neighbour), in the order of the ``for i ...: for nbr ...`` loop at ``fetch.py:389-396``;
* the line graph has one angle edge ``(i->j) -> (j->k)`` for every outgoing bond of ``j``
  except the exact reverse of ``(i->j)``, emitted e1-major (``fetch.py:421-444``);
* ``global_x`` is ``[59, 1]``, ``sg_one_hot`` is ``[230, 1]``, ``y`` is ``[2]`` (``fetch.py:614-651``).

Topology: ``A`` atoms on a ring, atom ``i`` bonded to ``(i +- d) mod A``
for ``d = 1..K/2``: K-regular, every bond has its reverse, so ``E = A*K`` and ``L = E*(K-1)``.
``dups=True`` repeats ~10 % of the bonds and adds self loops (periodic-image realism).

Collate follows PyG's default rules (what ``torch_geometric.loader.DataLoader`` does for the
reference, ``train.py:2037``): ``edge_index`` is offset by the running atom count; ``lg_edge_index`` is
offset by the running **atom** count too when ``lg_inc='pyg'`` (faithful to the reference, which
never overrides ``Data.__inc__``) or by the running **bond** count when ``lg_inc='bonds'``
(geometrically correct).  The model is index-agnostic; both are supported everywhere.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import torch
from torch import Tensor

NODE_DIM, EDGE_DIM, ANGLE_DIM, GLOBAL_SCALARS, SG_DIM, TARGET_DIM = 206, 36, 11, 59, 230, 2
GLOBAL_DIM = GLOBAL_SCALARS + SG_DIM
# log-transform statistics of artifacts/ensemble/scaler_state.pt (reference), used for target scale
LOG_MEANS = (4.3228, 3.5567)
LOG_STDS = (0.9051, 0.9405)


@dataclass
class CrystalGraph:
    """One structure in the reference's per-graph ``Data`` layout."""
    x: Tensor               # [A, node_dim] fp32
    edge_index: Tensor      # [2, E] int64, source-major
    edge_attr: Tensor       # [E, edge_dim]
    lg_edge_index: Tensor   # [2, L] int64 (bond ids local to the graph), e1-major
    lg_edge_attr: Tensor    # [L, angle_dim]
    global_x: Tensor        # [59, 1]
    sg_one_hot: Tensor      # [230, 1]
    y: Tensor               # [2]

    @property
    def num_nodes(self) -> int:
        return int(self.x.size(0))


class GraphBatch:
    """Duck-typed PyG ``Batch``: the attributes the model and trainer read."""

    _TENSORS = ("x", "edge_index", "edge_attr", "lg_edge_index", "lg_edge_attr", "global_x",
                "sg_one_hot", "batch", "y", "train_idx")

    def __init__(self, **kw):
        self.num_graphs = int(kw.pop("num_graphs"))
        self.lg_inc = kw.pop("lg_inc", "pyg")
        for k in self._TENSORS:
            setattr(self, k, kw.pop(k, None))
        if kw:
            raise TypeError(f"unexpected fields {sorted(kw)}")
        # host-side fact known at collate time: bond rows >= this bound have no line-graph neighbours (with PyG's
        # default collate the reference offsets lg_edge_index by atoms, so for B > 1 most bond rows are isolated)
        lg = self.lg_edge_index
        self.lg_active_rows = (int(lg.max()) + 1 if lg.numel() > 0 else 0) if isinstance(lg, Tensor) and not lg.is_cuda \
            else None
        self.padded = False                 # True for batches produced by batching.pad_batch (index -1 = padding)
        # second host-side fact: are the source rows (edge_index[0], lg_edge_index[0]) already non-decreasing?  True for
        # everything fetch.py emits (source-major loops, fetch.py:389-396, 421-444) and preserved by PyG's collate; the
        # graph plan then skips the source sort (csrc/plan.cu verifies the hint on the device)
        self.source_sorted = tuple(
            bool(isinstance(t, Tensor) and not t.is_cuda and t.dim() == 2
                 and (t.size(1) < 2 or bool((t[0, 1:] >= t[0, :-1]).all())))
            for t in (self.edge_index, self.lg_edge_index))

    def _like(self) -> "GraphBatch":
        out = GraphBatch.__new__(GraphBatch)
        out.num_graphs, out.lg_inc, out.lg_active_rows = self.num_graphs, self.lg_inc, self.lg_active_rows
        out.source_sorted = getattr(self, "source_sorted", (False, False))
        out.padded = getattr(self, "padded", False)
        return out

    def to(self, device, non_blocking: bool = False) -> "GraphBatch":
        out = self._like()
        for k in self._TENSORS:
            v = getattr(self, k)
            setattr(out, k, v.to(device, non_blocking=non_blocking) if isinstance(v, Tensor) else v)
        return out

    def pin_memory(self) -> "GraphBatch":
        out = self._like()
        for k in self._TENSORS:
            v = getattr(self, k)
            setattr(out, k, v.pin_memory() if isinstance(v, Tensor) else v)
        return out

    def tensors(self) -> Dict[str, Tensor]:
        return {k: getattr(self, k) for k in self._TENSORS if isinstance(getattr(self, k), Tensor)}

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors().values())

    @property
    def sizes(self) -> Dict[str, int]:
        return {"B": self.num_graphs, "N": int(self.x.size(0)), "E": int(self.edge_index.size(1)),
                "L": int(self.lg_edge_index.size(1))}


def ring_topology(atoms: int, k: int, dups: bool = False, gen: Optional[torch.Generator] = None):
    """``(edge_index[2,E], lg_edge_index[2,L])`` of one K-regular ring crystal (local ids)."""
    if k % 2 or k <= 0:
        raise ValueError("K must be a positive even number")
    if atoms <= k:
        raise ValueError("need atoms > K so that ring neighbours are distinct")
    half = k // 2
    offs = torch.tensor([d for d in range(1, half + 1)] + [-d for d in range(1, half + 1)], dtype=torch.long)
    src = torch.arange(atoms, dtype=torch.long).repeat_interleave(k)
    dst = (src + offs.repeat(atoms)) % atoms
    img = offs.repeat(atoms)                       # stands in for the periodic image of the bond
    if dups:
        if gen is None:
            gen = torch.Generator().manual_seed(0)
        # repeat ~10 % of bonds as a second periodic image, plus one self loop on ~10 % of atoms;
        # keep the source-major emission order of fetch.py (all bonds of atom i are contiguous)
        rep = torch.rand(src.numel(), generator=gen) < 0.10
        loops = torch.rand(atoms, generator=gen) < 0.10
        s_list, d_list, i_list = [], [], []
        for a in range(atoms):
            sl = slice(a * k, (a + 1) * k)
            s_list.append(src[sl]); d_list.append(dst[sl]); i_list.append(img[sl])
            r = rep[sl]
            if bool(r.any()):
                s_list.append(src[sl][r]); d_list.append(dst[sl][r]); i_list.append(img[sl][r] + 1000)
            if bool(loops[a]):
                s_list.append(torch.tensor([a])); d_list.append(torch.tensor([a])); i_list.append(torch.tensor([2000]))
        src, dst, img = torch.cat(s_list), torch.cat(d_list), torch.cat(i_list)
    edge_index = torch.stack([src, dst])
    e = src.numel()
    # line graph: bond e1=(i->j) feeds every bond e2=(j->k') leaving j, except e1's exact reverse
    # (same atoms, negated image).  Bonds leaving j are contiguous (source-major order).
    counts = torch.bincount(src, minlength=atoms)
    start = torch.cumsum(counts, 0) - counts
    deg_out_of_dst = counts[dst]                                     # candidates per e1
    e1 = torch.arange(e, dtype=torch.long).repeat_interleave(deg_out_of_dst)
    first = torch.cumsum(deg_out_of_dst, 0) - deg_out_of_dst
    within = torch.arange(e1.numel(), dtype=torch.long) - first.repeat_interleave(deg_out_of_dst)
    e2 = start[dst[e1]] + within
    is_reverse = (dst[e2] == src[e1]) & (img[e2] == -img[e1])
    keep = ~is_reverse
    lg_edge_index = torch.stack([e1[keep], e2[keep]])
    return edge_index, lg_edge_index


def make_crystal(atoms: int, k: int, gen: torch.Generator, dups: bool = False,
                 node_dim: int = NODE_DIM, edge_dim: int = EDGE_DIM, angle_dim: int = ANGLE_DIM,
                 global_scalars: int = GLOBAL_SCALARS, sg_dim: int = SG_DIM) -> CrystalGraph:
    ei, lg = ring_topology(atoms, k, dups=dups, gen=gen)
    e, l = ei.size(1), lg.size(1)
    x = torch.randn(atoms, node_dim, generator=gen)
    edge_attr = torch.rand(e, edge_dim, generator=gen)
    lg_attr = torch.rand(l, angle_dim, generator=gen)
    gx = torch.randn(global_scalars, 1, generator=gen)
    sg = torch.zeros(sg_dim, 1)
    sg[int(torch.randint(0, sg_dim, (1,), generator=gen))] = 1.0
    mu, sd = torch.tensor(LOG_MEANS), torch.tensor(LOG_STDS)
    y = torch.exp(mu + sd * torch.randn(2, generator=gen))
    return CrystalGraph(x, ei, edge_attr, lg, lg_attr, gx, sg, y)


def collate(graphs: Sequence[CrystalGraph], lg_inc: str = "pyg") -> GraphBatch:
    """Concatenate graphs with PyG's default ``Batch`` rules (see module docstring)."""
    if lg_inc not in ("pyg", "bonds"):
        raise ValueError("lg_inc must be 'pyg' or 'bonds'")
    xs, eis, eas, lgs, las, gxs, sgs, ys, bvec = [], [], [], [], [], [], [], [], []
    n_off = e_off = 0
    for g, d in enumerate(graphs):
        xs.append(d.x); eas.append(d.edge_attr); las.append(d.lg_edge_attr)
        gxs.append(d.global_x); sgs.append(d.sg_one_hot); ys.append(d.y)
        eis.append(d.edge_index + n_off)
        lgs.append(d.lg_edge_index + (n_off if lg_inc == "pyg" else e_off))
        bvec.append(torch.full((d.num_nodes,), g, dtype=torch.long))
        n_off += d.num_nodes
        e_off += int(d.edge_index.size(1))
    return GraphBatch(
        x=torch.cat(xs), edge_index=torch.cat(eis, dim=1), edge_attr=torch.cat(eas),
        lg_edge_index=torch.cat(lgs, dim=1), lg_edge_attr=torch.cat(las),
        global_x=torch.cat(gxs), sg_one_hot=torch.cat(sgs), batch=torch.cat(bvec), y=torch.cat(ys),
        train_idx=torch.arange(len(graphs), dtype=torch.long), num_graphs=len(graphs), lg_inc=lg_inc)


def synthetic_batch(n_graphs: int, atoms: int, k: int, seed: int = 0, lg_inc: str = "pyg", dups: bool = False,
                    node_dim: int = NODE_DIM, edge_dim: int = EDGE_DIM, angle_dim: int = ANGLE_DIM,
                    vectorized: Optional[bool] = None) -> GraphBatch:
    """Seeded batch of ``n_graphs`` identical-topology crystals with independent features.

    ``vectorized`` (default: on when ``dups`` is off) builds the topology once and draws all features
    in bulk -- same layout and index rules as ``collate(make_crystal(...))``, far faster for
    the 10^5..10^7-edge batches of the BASELINE configs.
    """
    gen = torch.Generator().manual_seed(int(seed))
    if vectorized is None:
        vectorized = not dups
    if not vectorized or dups:
        graphs = [make_crystal(atoms, k, gen, dups=dups, node_dim=node_dim, edge_dim=edge_dim,
                               angle_dim=angle_dim) for _ in range(n_graphs)]
        return collate(graphs, lg_inc=lg_inc)
    ei, lg = ring_topology(atoms, k)
    e, l = ei.size(1), lg.size(1)
    g_ids = torch.arange(n_graphs, dtype=torch.long)
    edge_index = (ei.unsqueeze(1) + (g_ids * atoms).view(1, -1, 1)).reshape(2, -1)
    lg_step = atoms if lg_inc == "pyg" else e
    lg_edge_index = (lg.unsqueeze(1) + (g_ids * lg_step).view(1, -1, 1)).reshape(2, -1)
    n = n_graphs * atoms
    x = torch.randn(n, node_dim, generator=gen)
    edge_attr = torch.rand(n_graphs * e, edge_dim, generator=gen)
    lg_attr = torch.rand(n_graphs * l, angle_dim, generator=gen)
    gx = torch.randn(n_graphs * GLOBAL_SCALARS, 1, generator=gen)
    sg = torch.zeros(n_graphs, SG_DIM)
    sg[g_ids, torch.randint(0, SG_DIM, (n_graphs,), generator=gen)] = 1.0
    mu, sd = torch.tensor(LOG_MEANS), torch.tensor(LOG_STDS)
    y = torch.exp(mu + sd * torch.randn(n_graphs, 2, generator=gen)).reshape(-1)
    return GraphBatch(x=x, edge_index=edge_index, edge_attr=edge_attr, lg_edge_index=lg_edge_index,
                      lg_edge_attr=lg_attr, global_x=gx, sg_one_hot=sg.reshape(-1, 1),
                      batch=g_ids.repeat_interleave(atoms), y=y, train_idx=g_ids.clone(),
                      num_graphs=n_graphs, lg_inc=lg_inc)


def zscore_targets(y: Tensor, num_graphs: int) -> Tensor:
    """``LogTransformer.transform_tensor`` with the shipped statistics (reference ``train.py:268-280``):
    ``z = (log(y) - mean) / std`` per target."""
    y = y.view(num_graphs, -1)
    mu = torch.tensor(LOG_MEANS, dtype=y.dtype, device=y.device)
    sd = torch.tensor(LOG_STDS, dtype=y.dtype, device=y.device)
    return (torch.log(y.clamp(min=1e-12)) - mu) / sd
