"""Golden vectors for the ensemble post-processing, produced by the reference's OWN ``ensemble_collect``,
``conformal_calibration``, ``apply_conformal_intervals`` and ``LogTransformer`` (``/root/reference/scripts/train.py``,
imported unmodified under the PyG shim).  The "models" are stand-ins that return stored (mean, logvar) pairs; every number
in the stored outputs is computed by the reference's code.  Run from the repo root in the build container:

    python oracle/gen_golden_ensemble.py         # writes tests/golden/ensemble_post.pt

TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from oracle import ensemble_ref  # noqa: E402


class _Member(torch.nn.Module):
    def __init__(self, outs):
        super().__init__()
        self.outs, self.i = outs, 0

    def forward(self, batch):
        out = self.outs[self.i]
        self.i += 1
        return out


class _Batch:
    def __init__(self, y):
        self.y, self.num_graphs = y, y.size(0)

    def to(self, device):
        return self


def main():
    ref = oracle.load_reference_train_module()
    assert ref is not None, "needs /root/reference"
    g = torch.Generator().manual_seed(7)
    n_members, batches, n_targets = 5, (40, 24), 2
    means = [[torch.randn(b, n_targets, generator=g) for b in batches] for _ in range(n_members)]
    logvars = [[torch.randn(b, n_targets, generator=g) * 1.5 - 2.0 for b in batches] for _ in range(n_members)]   # some below the floor
    ys = [torch.exp(torch.tensor([4.3228, 3.5567]) + torch.tensor([0.9051, 0.9405]) * torch.randn(b, n_targets, generator=g))
          for b in batches]
    models = [_Member(list(zip(means[m], logvars[m]))) for m in range(n_members)]
    loader = [_Batch(y) for y in ys]
    mean_z, targets, std_z = ref.ensemble_collect(models, loader, torch.device("cpu"), hetero=True)
    tr = ref.LogTransformer()
    tr.load_state_dict({"means": np.array([4.3228, 3.5567]), "stds": np.array([0.9051, 0.9405])})
    blob = {"member_means": torch.stack([torch.cat(m) for m in means]), "member_logvars": torch.stack([torch.cat(v) for v in logvars]),
            "targets": targets, "mean_z": mean_z, "std_z": std_z,
            "log_means": torch.tensor(tr.means, dtype=torch.float32), "log_stds": torch.tensor(tr.stds, dtype=torch.float32)}
    for method in ("scaled", "absolute"):
        conf = ref.conformal_calibration(mean_z, std_z, targets, tr, alpha=0.1, method=method)
        mo, lo, hi = ref.apply_conformal_intervals(mean_z, std_z, conf, tr)
        mo_z, lo_z, hi_z = ref.apply_conformal_intervals(mean_z, std_z, conf, None)
        blob[method] = {"q": conf["q"], "method": conf["method"], "mean": mo, "lower": lo, "upper": hi, "lower_z": lo_z,
                        "upper_z": hi_z}
    # the restatement must reproduce the reference bit for bit
    blob["batch_sizes"] = list(batches)
    mz, sz = ensemble_ref.moments_batched(blob["member_means"], blob["member_logvars"], batches)
    assert torch.equal(mz, mean_z) and torch.equal(sz, std_z)
    tz = ensemble_ref.to_z(targets, blob["log_means"], blob["log_stds"])
    for method in ("scaled", "absolute"):
        conf = ensemble_ref.calibration(mz, sz, tz, 0.1, method)
        assert torch.equal(conf["q"], blob[method]["q"]), method
        mo, lo, hi = ensemble_ref.intervals(mz, sz, conf["q"], method, blob["log_means"], blob["log_stds"])
        assert torch.equal(mo, blob[method]["mean"]) and torch.equal(lo, blob[method]["lower"]) and torch.equal(hi, blob[method]["upper"])
    path = os.path.join(ROOT, "tests", "golden", "ensemble_post.pt")
    torch.save(blob, path)
    print(f"{path}: {os.path.getsize(path) / 1024:.0f} KB; restatement bit-identical to the reference")


if __name__ == "__main__":
    main()
