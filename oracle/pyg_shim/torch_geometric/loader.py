"""``torch_geometric.loader.DataLoader`` restated (oracle; test infrastructure).

A ``torch.utils.data.DataLoader`` whose ``collate_fn`` is
``Batch.from_data_list`` -- the behaviour the reference relies on at
``/root/reference/scripts/train.py:1392,1397,1600,1610,2037``.
"""
from __future__ import annotations

import torch

from .data import Batch


class DataLoader(torch.utils.data.DataLoader):
    def __init__(self, dataset, batch_size: int = 1, shuffle: bool = False, follow_batch=None,
                 exclude_keys=None, **kwargs):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size=batch_size, shuffle=shuffle,
                         collate_fn=lambda items: Batch.from_data_list(items), **kwargs)
