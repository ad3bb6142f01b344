"""Ensemble member loops of the reference, kept on the device.

Restates the per-batch member loop + mixture moments of ``ensemble_collect`` (reference
``scripts/train.py:876-894``), ``predict.ensemble_predict`` (``scripts/predict.py:604-623``) and
``evaluate.collect_member_predictions`` (``scripts/evaluate.py:244-261``):

    mean_z = mean_m(mu_m);  var_z = mean_m(exp(clamp(logvar_m, floor))) + mean_m(mu_m^2) - mean_z^2
    std_z  = sqrt(clamp(var_z, 1e-12))

The graph plans (CSR/CSC sorts) are built once per batch and shared by all members.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Sequence, Tuple

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
    out = ensemble_post(torch.stack(means), torch.stack(logvars), min_logvar_floor)       # one kernel over the members
    return out["mean_z"], out["var_z"], out["std_z"]


def lognormal_to_linear(mean_z: Tensor, std_z: Tensor, log_means: Tensor, log_stds: Tensor) -> Tuple[Tensor, Tensor]:
    """z-space moments -> linear-space mean and sigma via the log-normal moment formulas of
    ``predict.ensemble_predict`` (``scripts/predict.py:618-623``)."""
    log_mean = mean_z * log_stds + log_means
    log_std = std_z * log_stds
    mean_lin = torch.exp(log_mean)
    var_lin = (torch.exp(log_std.pow(2)) - 1.0) * torch.exp(2 * log_mean + log_std.pow(2))
    return mean_lin, torch.sqrt(torch.clamp(var_lin, min=0.0))


def ensemble_post(member_means: Tensor, member_logvars: Optional[Tensor], min_logvar_floor: float = MIN_LOGVAR_FLOOR,
                  q: Optional[Tensor] = None, method: str = "scaled", log_means: Optional[Tensor] = None,
                  log_stds: Optional[Tensor] = None) -> Dict[str, Tensor]:
    """Mixture moments + conformal interval + inverse log-transform of the stacked member outputs ``[M, B, T]`` as ONE
    kernel (``alignn_ensemble_post``, ``csrc/post.cu``): the tail of ``ensemble_collect`` (reference ``train.py:876-894,
    903``), ``apply_conformal_intervals`` (``train.py:1053-1076``) and ``LogTransformer.inverse_transform_tensor``
    (``train.py:281-296``).  Returns ``mean_z, var_z, std_z`` and -- when ``q`` and / or the transformer statistics are
    given -- ``mean, lower, upper`` (original units when the statistics are given)."""
    from . import _lib, ops
    if not member_means.is_cuda:
        raise RuntimeError("ensemble_post: member outputs must be CUDA tensors (no CPU fallback path)")
    if member_means.dim() != 3:
        raise ValueError("member_means must be [members, graphs, targets]")
    if method not in ("scaled", "absolute"):
        raise ValueError("method must be 'scaled' or 'absolute'")
    dev = member_means.device
    f = lambda t: None if t is None else t.to(dev, torch.float32).contiguous()          # noqa: E731
    mu, lv, q, lm, ls = f(member_means), f(member_logvars), f(q), f(log_means), f(log_stds)
    if lv is None:
        method = "absolute"                         # reference: scaled needs std_z (train.py:1044-1048)
    m, b, t = mu.shape
    out = {k: torch.empty(b, t, dtype=torch.float32, device=dev) for k in ("mean_z", "var_z", "std_z")}
    want = q is not None or lm is not None
    if want:
        out.update({k: torch.empty(b, t, dtype=torch.float32, device=dev) for k in ("mean", "lower", "upper")})
    lib = _lib.load()
    p_ = ops._p
    with torch.cuda.device(dev), ops._Launch("ensemble_post", 1, (m, b, t)):
        rc = lib.alignn_ensemble_post(p_(mu), p_(lv), m, b, t, float(min_logvar_floor), p_(q), 1 if method == "scaled" else 0,
                                      p_(lm), p_(ls), p_(out["mean_z"]), p_(out["var_z"]), p_(out["std_z"]),
                                      p_(out.get("mean")), p_(out.get("lower")), p_(out.get("upper")), ops._stream())
    _lib.check(rc, "alignn_ensemble_post")
    return out


def conformal_calibration(mean_z: Tensor, std_z: Optional[Tensor], targets_z: Tensor, alpha: float,
                          method: str = "scaled") -> Dict[str, object]:
    """Split-conformal quantile of the calibration scores (reference ``train.py:1029-1050``; ``targets_z`` already in
    z-space, i.e. after ``LogTransformer.transform_tensor``).  Runs once per calibration set, wherever its inputs live."""
    if method == "scaled" and std_z is not None:
        s = (targets_z - mean_z).abs() / torch.clamp(std_z, min=1e-12)
    else:
        s = (targets_z - mean_z).abs()
        method = "absolute"
    n = s.size(0)
    q_level = min(max(math.ceil((n + 1) * (1 - alpha)) / n, 0.0), 1.0)
    return {"q": torch.quantile(s, q_level, dim=0), "method": method, "alpha": alpha}
