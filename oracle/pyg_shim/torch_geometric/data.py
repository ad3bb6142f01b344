"""``torch_geometric.data.{Data, Batch, Dataset}`` restated (oracle; test infrastructure).

Only the collate rules the reference depends on are reproduced
(``/root/reference/scripts/fetch.py:614-651`` builds the ``Data`` objects;
``scripts/train.py:2037`` batches them with the default ``DataLoader``):

* ``Data.__cat_dim__(key)``: ``-1`` if ``'index'`` occurs in ``key`` else ``0``.
* ``Data.__inc__(key)``: ``num_nodes`` if ``'index'`` occurs in ``key`` else ``0``.
  The reference never overrides this, so ``lg_edge_index`` (bond ids) is offset
  by the running *atom* count -- the batching quirk of SURVEY.md section 8(a) A9 --
  and so are ``sample_index`` / ``train_idx``.
* 0-dim tensors are stacked; other tensors concatenated; non-tensors collected
  into lists; ``batch`` assigns each node its graph id; ``ptr`` holds node offsets.
"""
from __future__ import annotations

from typing import Any, Iterable, List

import torch
from torch import Tensor


class Data:
    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, **kwargs):
        self.__dict__["_store"] = {}
        for k, v in (("x", x), ("edge_index", edge_index), ("edge_attr", edge_attr), ("y", y)):
            if v is not None:
                self._store[k] = v
        for k, v in kwargs.items():
            self._store[k] = v

    # attribute plumbing -------------------------------------------------
    def __getattr__(self, key: str) -> Any:
        store = self.__dict__.get("_store", {})
        if key in store:
            return store[key]
        raise AttributeError(f"'{type(self).__name__}' object has no attribute '{key}'")

    def __setattr__(self, key: str, value: Any) -> None:
        self._store[key] = value

    def __contains__(self, key: str) -> bool:
        return key in self._store

    def keys(self) -> List[str]:
        return list(self._store.keys())

    def __getitem__(self, key: str) -> Any:
        return self._store[key]

    def __setitem__(self, key: str, value: Any) -> None:
        self._store[key] = value

    @property
    def num_nodes(self) -> int:
        if "num_nodes" in self._store:
            return int(self._store["num_nodes"])
        x = self._store.get("x")
        if isinstance(x, Tensor):
            return int(x.size(0))
        ei = self._store.get("edge_index")
        if isinstance(ei, Tensor) and ei.numel() > 0:
            return int(ei.max()) + 1
        return 0

    @property
    def num_edges(self) -> int:
        ei = self._store.get("edge_index")
        return int(ei.size(1)) if isinstance(ei, Tensor) else 0

    # collate rules --------------------------------------------------------
    def __cat_dim__(self, key: str, value: Any) -> int:
        return -1 if "index" in key else 0

    def __inc__(self, key: str, value: Any) -> int:
        return self.num_nodes if "index" in key else 0

    def to(self, device, non_blocking: bool = False):
        out = self.__class__.__new__(self.__class__)
        out.__dict__["_store"] = {
            k: (v.to(device, non_blocking=non_blocking) if isinstance(v, Tensor) else v)
            for k, v in self._store.items()
        }
        for k, v in self.__dict__.items():
            if k != "_store":
                out.__dict__[k] = v
        return out

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        out.__dict__["_store"] = {
            k: (v.clone() if isinstance(v, Tensor) else v) for k, v in self._store.items()
        }
        return out

    def __repr__(self) -> str:
        parts = []
        for k, v in self._store.items():
            parts.append(f"{k}={list(v.shape)}" if isinstance(v, Tensor) else f"{k}={type(v).__name__}")
        return f"{type(self).__name__}({', '.join(parts)})"


class Batch(Data):
    @classmethod
    def from_data_list(cls, data_list: Iterable[Data]) -> "Batch":
        data_list = list(data_list)
        if not data_list:
            raise ValueError("empty data list")
        keys = data_list[0].keys()
        out = cls.__new__(cls)
        out.__dict__["_store"] = {}
        incs = {k: 0 for k in keys}
        collected = {k: [] for k in keys}
        batch_vec = []
        ptr = [0]
        for g, d in enumerate(data_list):
            n = d.num_nodes
            for k in keys:
                v = d[k]
                if isinstance(v, Tensor):
                    if v.dim() == 0:
                        v = v.unsqueeze(0)
                    inc = incs[k]
                    if inc != 0:
                        v = v + inc
                    collected[k].append(v)
                    incs[k] += d.__inc__(k, d[k])
                else:
                    collected[k].append(v)
            batch_vec.append(torch.full((n,), g, dtype=torch.long))
            ptr.append(ptr[-1] + n)
        for k in keys:
            vals = collected[k]
            if isinstance(vals[0], Tensor):
                out._store[k] = torch.cat(vals, dim=data_list[0].__cat_dim__(k, vals[0]))
            else:
                out._store[k] = vals
        out._store["batch"] = torch.cat(batch_vec) if batch_vec else torch.zeros(0, dtype=torch.long)
        out._store["ptr"] = torch.tensor(ptr, dtype=torch.long)
        out.__dict__["_num_graphs"] = len(data_list)
        return out

    @property
    def num_graphs(self) -> int:
        return int(self.__dict__.get("_num_graphs", 0))

    @property
    def num_nodes(self) -> int:
        return int(self._store["x"].size(0))


class Dataset(torch.utils.data.Dataset):
    """Placeholder for ``torch_geometric.data.Dataset`` (imported but unused by the hot path)."""

    def __init__(self, root=None, transform=None, pre_transform=None, pre_filter=None):
        super().__init__()
        self.root = root
        self.transform = transform
