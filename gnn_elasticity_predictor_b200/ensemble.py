"""Ensemble member loops of the reference, kept on the device.

Restates the per-batch member loop + mixture moments of ``ensemble_collect`` (reference
``scripts/train.py:876-894``), ``predict.ensemble_predict`` (``scripts/predict.py:604-623``) and
``evaluate.collect_member_predictions`` (``scripts/evaluate.py:244-261``):

    mean_z = mean_m(mu_m);  var_z = mean_m(exp(clamp(logvar_m, floor))) + mean_m(mu_m^2) - mean_z^2
    std_z  = sqrt(clamp(var_z, 1e-12))

The graph plans (CSR/CSC sorts) are built once per batch and shared by all members.
"""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
from torch import Tensor

MIN_LOGVAR_FLOOR = -2.9  # reference scripts/train.py:39


def ensemble_moments(member_means: Tensor, member_logvars: Tensor,
                     min_logvar_floor: float = MIN_LOGVAR_FLOOR) -> Tuple[Tensor, Tensor, Tensor]:
    """``member_*``: ``[M, B, T]``.  Returns ``(mean_z, var_z, std_z)``, each ``[B, T]`` fp32."""
    mu = member_means.float()
    var = torch.exp(torch.clamp(member_logvars.float(), min=min_logvar_floor))
    mean_z = mu.mean(dim=0)
    var_z = var.mean(dim=0) + mu.pow(2).mean(dim=0) - mean_z.pow(2)
    return mean_z, var_z, torch.sqrt(torch.clamp(var_z, min=1e-12))


@torch.no_grad()
def ensemble_forward(models: Sequence[torch.nn.Module], batch, min_logvar_floor: float = MIN_LOGVAR_FLOOR):
    """All members on one batch (members share the batch's graph plans); returns the mixture moments."""
    if not models:
        raise ValueError("no ensemble members")
    base = models[0].base
    if getattr(batch, "_alignn_plans", None) is None:
        base.build_plans(batch)
    means, logvars = [], []
    for m in models:
        mean, logvar = m(batch)
        means.append(mean)
        logvars.append(logvar)
    return ensemble_moments(torch.stack(means), torch.stack(logvars), min_logvar_floor)


def lognormal_to_linear(mean_z: Tensor, std_z: Tensor, log_means: Tensor, log_stds: Tensor) -> Tuple[Tensor, Tensor]:
    """z-space moments -> linear-space mean and sigma via the log-normal moment formulas of
    ``predict.ensemble_predict`` (``scripts/predict.py:618-623``)."""
    log_mean = mean_z * log_stds + log_means
    log_std = std_z * log_stds
    mean_lin = torch.exp(log_mean)
    var_lin = (torch.exp(log_std.pow(2)) - 1.0) * torch.exp(2 * log_mean + log_std.pow(2))
    return mean_lin, torch.sqrt(torch.clamp(var_lin, min=0.0))
