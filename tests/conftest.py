"""pytest configuration: `gpu` marker, import paths, shared helpers."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, name), weights_only=False)


class Bag:
    """Attribute bag standing in for a PyG Batch when feeding stored golden tensors to a model."""

    def __init__(self, tensors, num_graphs):
        for k, v in tensors.items():
            setattr(self, k, v)
        self.num_graphs = num_graphs

    def to(self, device):
        return Bag({k: (v.to(device) if torch.is_tensor(v) else v) for k, v in self.__dict__.items()
                    if k != "num_graphs"}, self.num_graphs)


def rel_err(a, b):
    """max |a-b| / max(|b|_max, tiny): the relative error used for every tolerance in this suite."""
    a, b = a.double(), b.double()
    denom = max(float(b.abs().max()), 1e-30)
    return float((a - b).abs().max()) / denom
