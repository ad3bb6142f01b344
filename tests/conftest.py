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
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    denom = max(float(b.abs().max()), 1e-30)
    return float((a - b).abs().max()) / denom


def grad_err(got, want, gmax):
    """gradient error relative to max(|want|_max, 1e-3 * gmax), gmax = largest |gradient| over ALL parameters.

    The floor matters for tensors whose true gradient is (near) zero by symmetry -- e.g. lin_key.bias: the
    segment softmax is invariant to a shift of all keys of a row, so d/d b_k == 0 exactly and every
    implementation (including the fp32 CPU oracle vs the fp64 oracle) only produces rounding noise there."""
    got, want = got.detach().double().cpu(), want.detach().double().cpu()
    return float((got - want).abs().max()) / max(float(want.abs().max()), 1e-3 * gmax, 1e-30)


def grads_gmax(grads):
    return max(float(v.detach().abs().max()) for v in grads.values())
