"""CPU restatement of the reference's ensemble post-processing -- TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Follows ``/root/reference/scripts/train.py``: the member-loop tail of ``ensemble_collect`` (``:876-894``, ``:899-903``),
``conformal_calibration`` (``:1029-1050``), ``apply_conformal_intervals`` (``:1053-1076``) and
``LogTransformer.transform_tensor / inverse_transform_tensor`` (``:268-296``).  Pinned by ``tests/golden/ensemble_post.pt``,
produced by the reference's OWN functions (``oracle/gen_golden_ensemble.py``)."""
from __future__ import annotations

import math
from typing import Optional

import torch

MIN_LOGVAR_FLOOR = -2.9          # train.py:39


def moments(member_means: torch.Tensor, member_logvars: Optional[torch.Tensor], floor: float = MIN_LOGVAR_FLOOR):
    stacked_means = member_means                                                   # [M, B, T]
    mean_z = stacked_means.mean(dim=0)
    if member_logvars is None:
        return mean_z, None, None
    stacked_vars = torch.exp(torch.clamp(member_logvars, min=floor))
    var_z = stacked_vars.mean(dim=0) + stacked_means.pow(2).mean(dim=0) - mean_z.pow(2)
    return mean_z, var_z, torch.sqrt(torch.clamp(var_z, min=1e-12))


def moments_batched(member_means, member_logvars, batch_sizes, floor: float = MIN_LOGVAR_FLOOR):
    """``ensemble_collect`` forms the moments batch by batch and concatenates (``train.py:899-903``); torch's reduction over
    the member axis rounds differently for different batch sizes, so bit-identity needs the same split."""
    mz, sz, o = [], [], 0
    for b in batch_sizes:
        m, _, s = moments(member_means[:, o:o + b], member_logvars[:, o:o + b], floor)
        mz.append(m); sz.append(s); o += b
    return torch.cat(mz), torch.cat(sz)


def to_z(targets: torch.Tensor, log_means: torch.Tensor, log_stds: torch.Tensor) -> torch.Tensor:
    return (torch.log(torch.clamp(targets, min=1e-12)) - log_means) / log_stds


def calibration(mean_z, std_z, targets_z, alpha: float, method: str):
    if method == "scaled" and std_z is not None:
        s = (targets_z - mean_z).abs() / torch.clamp(std_z, min=1e-12)
    else:
        s = (targets_z - mean_z).abs()
        method = "absolute"
    n = s.size(0)
    q_level = min(max(math.ceil((n + 1) * (1 - alpha)) / n, 0.0), 1.0)
    return {"q": torch.quantile(s, q_level, dim=0), "method": method, "alpha": alpha}


def intervals(mean_z, std_z, q, method: str, log_means=None, log_stds=None):
    if method == "scaled" and std_z is not None:
        lower_z, upper_z = mean_z - q * std_z, mean_z + q * std_z
    else:
        lower_z, upper_z = mean_z - q, mean_z + q
    if log_means is None:
        return mean_z, lower_z, upper_z
    inv = lambda z: torch.exp(z * log_stds + log_means)                              # noqa: E731
    return inv(mean_z), inv(lower_z), inv(upper_z)
