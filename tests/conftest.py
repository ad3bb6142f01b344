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


def per_tensor_report(got, want):
    """name -> (rel, share): rel = max|got - want| / max|want| of THAT tensor; share = max|want| / largest |gradient|."""
    gmax = max(grads_gmax(want), 1e-30)
    out = {}
    for k, w in want.items():
        w = w.detach().double().cpu()
        g = got[k].detach().double().cpu()
        wmax = float(w.abs().max())
        out[k] = (float((g - w).abs().max()) / max(wmax, 1e-30), wmax / gmax)
    return out


def check_per_tensor(got, want, tol, allow=None, label=None):
    """Every gradient tensor within ``tol`` of ITS OWN scale (max|err| / max|want| per tensor), except tensors matched by
    ``allow`` = {name suffix: (bound, reason)}: those are printed with their measured error, their bound and the reason,
    and still asserted against that bound.  A suffix bound that is a tuple ("abs", x) compares max|err| against
    x * (largest gradient of the model) -- for tensors whose TRUE gradient is exactly zero."""
    allow = allow or {}
    rep = per_tensor_report(got, want)
    gmax = max(grads_gmax(want), 1e-30)
    lines, bad = [], []
    for k, (rel, share) in sorted(rep.items(), key=lambda kv: -kv[1][0]):
        rule = next(((s, b) for s, b in allow.items() if k.endswith(s)), None)
        if rule is None:
            if rel > tol:
                bad.append(f"{k}: rel {rel:.3e} > {tol:g} (share of gmax {share:.2e})")
            continue
        bound, reason = rule[1]
        if isinstance(bound, tuple):
            err = float((got[k].detach().double().cpu() - want[k].detach().double().cpu()).abs().max()) / gmax
            ok, shown = err <= bound[1], f"abs err {err:.3e} of gmax (bound {bound[1]:g})"
        else:
            ok, shown = rel <= bound, f"rel {rel:.3e} (bound {bound:g}; default {tol:g})"
        lines.append(f"  allow-listed {k}: {shown}, share {share:.2e} -- {reason}")
        if not ok:
            bad.append(f"{k}: {shown} -- allow-listed bound exceeded")
    worst = sorted(rep.items(), key=lambda kv: -kv[1][0])[:8]
    text = "\n".join([f"[per-tensor gradient parity] {label or ''} tol {tol:g}"]
                     + [f"  {k}: rel {r:.3e} share {s:.2e}" for k, (r, s) in worst] + lines)
    print(text)
    out_dir = os.path.join(ROOT, "gpurun_out")
    if label and os.path.isdir(out_dir):
        with open(os.path.join(out_dir, f"parity_{label}.txt"), "w") as fh:
            fh.write(text + "\n" + "\n".join(f"{k}\t{r:.4e}\t{s:.3e}" for k, (r, s) in sorted(rep.items())) + "\n")
    assert not bad, "\n".join(bad)
    return rep
