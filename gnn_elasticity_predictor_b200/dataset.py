"""Device-resident graph store + on-GPU collate.

What it replaces in the reference: ``PtGraphDataset.__getitem__`` (``scripts/train.py:130-172``) ``torch.load``s one file
per sample per epoch, reshapes and standardises it on the host (``:200-217``), and ``torch_geometric.loader.DataLoader``
(``train.py:2037``) collates the ``Data`` objects on the host and ships the batch over PCIe every step.  Here the whole
dataset is uploaded ONCE -- every field of ``fetch.py:614-651`` concatenated over graphs, indices local to each graph,
standardisation already applied -- and a batch is one C-ABI call (``alignn_collate``, ``csrc/dataprep.cu``) that writes
exactly what PyG's default ``Batch.from_data_list`` would, optionally straight into a shape bucket (``batching.py``) so
that ``engine.TrainStep`` replays one CUDA graph per bucket.  Per step the host sends ``B`` graph ids instead of ~88 MB.

Bit-exact against the host collate (``synthetic.collate``, itself pinned to the PyG rules in ``tests/test_synthetic.py``).
There is no CPU path: the store lives on a CUDA device.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from . import _lib, batching, ops
from .synthetic import CrystalGraph, GraphBatch

_P = ops._p


class DeviceGraphStore:
    """All graphs of a dataset in HBM.  ``collate(ids)`` -> ``GraphBatch`` on the device."""

    def __init__(self, graphs: Sequence[CrystalGraph], device, lg_inc: str = "pyg",
                 standardize: Optional[Dict[str, Tensor]] = None, scalar_dim: int = 6):
        if lg_inc not in ("pyg", "bonds"):
            raise ValueError("lg_inc must be 'pyg' or 'bonds'")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("DeviceGraphStore lives on a CUDA device (there is no CPU path)")
        if not graphs:
            raise ValueError("empty dataset")
        self.device, self.lg_inc = dev, lg_inc
        g0 = graphs[0]
        self.dims = dict(node=int(g0.x.size(1)), edge=int(g0.edge_attr.size(1)), angle=int(g0.lg_edge_attr.size(1)),
                         glob=int(g0.global_x.numel()), sg=int(g0.sg_one_hot.numel()), target=int(g0.y.numel()))
        n = np.array([g.x.size(0) for g in graphs], dtype=np.int64)
        e = np.array([g.edge_index.size(1) for g in graphs], dtype=np.int64)
        l = np.array([g.lg_edge_index.size(1) for g in graphs], dtype=np.int64)
        self.node_ptr_h, self.bond_ptr_h, self.angle_ptr_h = (np.concatenate([[0], np.cumsum(a)]) for a in (n, e, l))
        # host-side facts per graph, so that a batch's bounds are known without touching the device
        self._lg_max_h = np.array([int(g.lg_edge_index.max()) + 1 if g.lg_edge_index.numel() else 0 for g in graphs],
                                  dtype=np.int64)
        self._sorted_h = np.array([[_nondecreasing(g.edge_index[0]), _nondecreasing(g.lg_edge_index[0])] for g in graphs])
        ends = lambda t: (int(t[0, 0]), int(t[0, -1])) if t.size(1) else (None, None)      # noqa: E731
        self._src_ends_h = [(ends(g.edge_index), ends(g.lg_edge_index)) for g in graphs]
        x = torch.cat([g.x.reshape(-1, self.dims["node"]) for g in graphs]).float()
        gx = torch.stack([g.global_x.reshape(-1) for g in graphs]).float()
        if standardize:                                        # train.py:200-217, applied once instead of per sample per epoch
            x, gx = self._standardize(x.clone(), gx.clone(), standardize, scalar_dim)
        up = lambda t: t.contiguous().to(dev)                 # noqa: E731
        self.x = up(x)
        self.edge_attr = up(torch.cat([g.edge_attr.reshape(-1, self.dims["edge"]) for g in graphs]).float())
        self.lg_edge_attr = up(torch.cat([g.lg_edge_attr.reshape(-1, self.dims["angle"]) for g in graphs]).float())
        self.global_x = up(gx)
        self.sg_one_hot = up(torch.stack([g.sg_one_hot.reshape(-1) for g in graphs]).float())
        self.y = up(torch.stack([g.y.reshape(-1) for g in graphs]).float())
        self.edge_index = up(torch.cat([g.edge_index for g in graphs], dim=1).long())
        self.lg_edge_index = up(torch.cat([g.lg_edge_index for g in graphs], dim=1).long())
        self.node_ptr, self.bond_ptr, self.angle_ptr = (torch.from_numpy(a).to(dev) for a in
                                                        (self.node_ptr_h, self.bond_ptr_h, self.angle_ptr_h))
        self.n_graphs = len(graphs)
        s = self._struct = _lib.GraphStoreStruct()
        for name in ("x", "edge_attr", "lg_edge_attr", "global_x", "sg_one_hot", "y", "edge_index", "lg_edge_index",
                     "node_ptr", "bond_ptr", "angle_ptr"):
            setattr(s, name, getattr(self, name).data_ptr())
        s.n_graphs, s.n_nodes, s.n_bonds, s.n_angles = self.n_graphs, int(n.sum()), int(e.sum()), int(l.sum())
        s.node_dim, s.edge_dim, s.angle_dim = self.dims["node"], self.dims["edge"], self.dims["angle"]
        s.global_dim, s.sg_dim, s.target_dim = self.dims["glob"], self.dims["sg"], self.dims["target"]
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)

    @staticmethod
    def _standardize(x, gx, st, scalar_dim):
        sd = min(scalar_dim, x.size(1))
        if st.get("scalar_mean") is not None and st.get("scalar_std") is not None:
            x[:, :sd] = (x[:, :sd] - st["scalar_mean"].to(x.dtype)) / st["scalar_std"].to(x.dtype)
        if x.size(1) > sd and st.get("embed_mean") is not None and st.get("embed_std") is not None:
            x[:, sd:] = (x[:, sd:] - st["embed_mean"].to(x.dtype)) / st["embed_std"].to(x.dtype)
        if st.get("global_mean") is not None and st.get("global_std") is not None:
            gx = (gx - st["global_mean"].to(gx.dtype)) / st["global_std"].to(gx.dtype)
        return x, gx

    def __len__(self) -> int:
        return self.n_graphs

    def _batch_sorted(self, ids_h):
        """``GraphBatch.source_sorted`` of the collated selection, from per-graph host metadata: every graph's source row is
        sorted and consecutive graphs do not overlap after their increments (they do under ``lg_inc='pyg'``)."""
        if ids_h.size == 0:
            return (True, True)
        n_inc = np.concatenate([[0], np.cumsum(self.node_ptr_h[ids_h + 1] - self.node_ptr_h[ids_h])])[:-1]
        e_inc = np.concatenate([[0], np.cumsum(self.bond_ptr_h[ids_h + 1] - self.bond_ptr_h[ids_h])])[:-1]
        lg_inc = e_inc if self.lg_inc == "bonds" else n_inc
        g_ok = bool(self._sorted_h[ids_h, 0].all()) and _chain_sorted([self._src_ends_h[i][0] for i in ids_h], n_inc)
        lg_ok = bool(self._sorted_h[ids_h, 1].all()) and _chain_sorted([self._src_ends_h[i][1] for i in ids_h], lg_inc)
        return (g_ok, lg_ok)

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in (self.x, self.edge_attr, self.lg_edge_attr, self.global_x,
                                                          self.sg_one_hot, self.y, self.edge_index, self.lg_edge_index))

    def sizes_of(self, ids) -> Dict[str, int]:
        """Exact batch sizes of a selection, from the host copies of the offsets (no device access)."""
        ids = np.asarray(ids, dtype=np.int64).reshape(-1)
        if ids.size and (ids.min() < 0 or ids.max() >= self.n_graphs):
            raise IndexError(f"graph id outside [0, {self.n_graphs})")
        cnt = lambda p: p[ids + 1] - p[ids]                    # noqa: E731
        n, e, l = cnt(self.node_ptr_h), cnt(self.bond_ptr_h), cnt(self.angle_ptr_h)
        return {"B": int(ids.size), "N": int(n.sum()), "E": int(e.sum()), "L": int(l.sum()),
                "max_N": int(n.max()) if ids.size else 0, "max_E": int(e.max()) if ids.size else 0,
                "max_L": int(l.max()) if ids.size else 0}

    def collate(self, ids, pad_to_bucket: bool = False, align: int = 256, shape: Optional[Dict[str, int]] = None,
                ids_device: Optional[Tensor] = None, validate: bool = False, out: Optional[GraphBatch] = None) -> GraphBatch:
        """The batch PyG's ``Batch.from_data_list([dataset[i] for i in ids])`` would build, on the device.

        ``pad_to_bucket`` / ``shape``: write into a shape bucket (``batching.bucket_shape`` semantics; the returned batch has
        ``padded = True`` and ``mask`` attached as ``batch.loss_mask``).  ``ids_device``: the same ids already on the device
        (int64) -- otherwise they are uploaded here (``B`` x 8 bytes).  ``out``: an existing batch of exactly the target shape
        (e.g. ``TrainStep.static_inputs(...)[0]``, the input buffers of a captured CUDA graph) to write into."""
        ids_h = np.asarray(ids, dtype=np.int64).reshape(-1)
        s = self.sizes_of(ids_h)
        b = s["B"]
        if out is not None:                          # the destination fixes the shape (a bucket if it is larger)
            shape = dict(out.sizes)
        if shape is None and pad_to_bucket:
            shape = {"N": batching.round_up_bucket(s["N"] + 1, align), "E": batching.round_up_bucket(s["E"], align),
                     "L": batching.round_up_bucket(s["L"], align), "B": batching.round_up_bucket(b + 1, 8)}
        if shape is None:
            shape = {"N": s["N"], "E": s["E"], "L": s["L"], "B": b}
        elif shape["N"] < s["N"] or shape["E"] < s["E"] or shape["L"] < s["L"] or shape["B"] < b:
            raise ValueError(f"bucket {shape} is smaller than the selection {s}")
        padded = (shape["N"], shape["E"], shape["L"], shape["B"]) != (s["N"], s["E"], s["L"], b)
        if padded and shape["B"] <= b:
            raise ValueError("padding needs one spare graph slot for the dummy graph (B_pad > B)")
        dev, d = self.device, self.dims
        sel = ids_device if ids_device is not None else torch.from_numpy(ids_h).to(dev, non_blocking=True)
        f32, i64 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int64, device=dev)
        if out is not None:
            return self._collate_into(out, ids_h, s, shape, padded, sel, align, validate, check_facts=True)
        out = GraphBatch.__new__(GraphBatch)
        out.num_graphs, out.lg_inc, out.padded = shape["B"], self.lg_inc, padded
        out.x = torch.empty(shape["N"], d["node"], **f32)
        out.edge_attr = torch.empty(shape["E"], d["edge"], **f32)
        out.lg_edge_attr = torch.empty(shape["L"], d["angle"], **f32)
        out.global_x = torch.empty(shape["B"] * d["glob"], 1, **f32)
        out.sg_one_hot = torch.empty(shape["B"] * d["sg"], 1, **f32)
        out.y = torch.empty(shape["B"] * d["target"], **f32)
        out.edge_index = torch.empty(2, shape["E"], **i64)
        out.lg_edge_index = torch.empty(2, shape["L"], **i64)
        out.batch = torch.empty(shape["N"], **i64)
        out.train_idx = torch.empty(shape["B"], **i64)
        return self._collate_into(out, ids_h, s, shape, padded, sel, align, validate)

    def _collate_into(self, out: GraphBatch, ids_h, s, shape, padded: bool, sel: Tensor, align: int, validate: bool,
                      check_facts: bool = False) -> GraphBatch:
        dev, b = self.device, s["B"]
        # host-side facts of the batch (what GraphBatch.__init__ derives from host tensors), from per-graph host metadata
        if b:
            ptr = self.bond_ptr_h if self.lg_inc == "bonds" else self.node_ptr_h
            inc = np.concatenate([[0], np.cumsum(ptr[ids_h + 1] - ptr[ids_h])])[:-1]
            has = self._lg_max_h[ids_h] > 0
            active = int((inc[has] + self._lg_max_h[ids_h][has]).max()) if has.any() else 0
        else:
            active = 0
        active = min(shape["E"], batching.round_up_bucket(active, align)) if padded else active
        sorted_ = self._batch_sorted(ids_h)
        if check_facts and (getattr(out, "lg_active_rows", None) != active or getattr(out, "source_sorted", None) != sorted_
                            or bool(getattr(out, "padded", False)) != padded):
            # a captured CUDA graph was specialised on these facts: refuse rather than replay it on a batch it does not fit
            raise ValueError("collate(out=...): the selection's lg_active_rows / source_sorted / padding differ from the "
                             "destination batch's")
        f32, i64 = dict(dtype=torch.float32, device=dev), dict(dtype=torch.int64, device=dev)
        seg = torch.empty(3, b + 1, **i64)
        bo = _lib.BatchOutStruct()
        for name in ("x", "edge_attr", "lg_edge_attr", "global_x", "sg_one_hot", "y", "edge_index", "lg_edge_index",
                     "batch", "train_idx"):
            setattr(bo, name, getattr(out, name).data_ptr())
        bo.n_graphs, bo.n_nodes, bo.n_bonds, bo.n_angles = shape["B"], shape["N"], shape["E"], shape["L"]
        lib = _lib.load()
        n_launch = 10
        with torch.cuda.device(dev), ops._Launch("collate", n_launch, (b, s["N"], s["E"], s["L"])):
            totals = (ctypes.c_int64 * 3)(s["N"], s["E"], s["L"])
            maxima = (ctypes.c_int64 * 3)(s["max_N"], s["max_E"], s["max_L"])
            rc = lib.alignn_collate(ctypes.byref(self._struct), _P(sel), b, 1 if self.lg_inc == "bonds" else 0,
                                    ctypes.byref(bo), _P(seg), totals, maxima, _P(self.status), ops._stream())
        _lib.check(rc, "alignn_collate")
        if validate and int(self.status.item()) & 3:
            raise RuntimeError(f"alignn_collate status {int(self.status.item())}")
        out.lg_active_rows, out.source_sorted = active, sorted_
        out.seg_ptr = seg
        if padded:
            mask = torch.zeros(shape["B"], **f32)
            mask[:b] = 1.0
            out.loss_mask = mask
        return out


def _chain_sorted(ends, incs) -> bool:
    """Per-graph (first, last) source ids + per-graph increments -> is the concatenated source row non-decreasing?"""
    prev = None
    for (first, last), inc in zip(ends, incs):
        if first is None:
            continue
        if prev is not None and first + inc < prev:
            return False
        prev = last + inc
    return True


def _nondecreasing(t: Tensor) -> bool:
    return bool(t.numel() < 2 or bool((t[1:] >= t[:-1]).all()))
