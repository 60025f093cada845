"""Batched objective functions over [n_sets x T] simulated series against one observed series.

The reference computes them one run at a time through HydroErr (python/src/hydrobricks/evaluation/metrics.py:40-62,
models/model.py:620-708); HydroErr is an unpinned third-party package absent from this image, so these follow the
published formulas and state their conventions — PARITY UNPINNED against HydroErr:
  NSE      = 1 - sum((s-o)^2) / sum((o-mean(o))^2)
  KGE-2012 = 1 - sqrt((r-1)^2 + (gamma-1)^2 + (beta-1)^2), gamma = CV_s/CV_o (std with ddof=1), beta = mean_s/mean_o
  KGE-2009 = the same with alpha = std_s/std_o in place of gamma; me = mean(s-o), mae = mean|s-o|, mse, rmse = sqrt(mse),
  nrmse_mean = rmse/mean(o), r_squared = Pearson r^2
Time steps where the observation is NaN are dropped (HydroErr's remove-NaN treatment); `warmup` leading steps are
skipped as SpotpySetup does (trainer.py:1091-1141). Works on torch tensors (GPU, zero-copy on the engine's outlet)
and on numpy arrays.
"""
from __future__ import annotations

import numpy as np


def _prep(sim, obs, warmup):
    import torch

    if isinstance(sim, np.ndarray):
        sim = torch.from_numpy(sim)
    obs = torch.as_tensor(np.asarray(obs, dtype=np.float64) if not hasattr(obs, "device") else obs,
                          dtype=torch.float64, device=sim.device)
    sim, obs = sim[:, warmup:], obs[warmup:]
    keep = ~torch.isnan(obs)
    return sim[:, keep], obs[keep]


def nse(sim, obs, warmup: int = 0):
    s, o = _prep(sim, obs, warmup)
    return 1 - ((s - o) ** 2).sum(dim=1) / ((o - o.mean()) ** 2).sum()


def kge_2012(sim, obs, warmup: int = 0):
    import torch

    s, o = _prep(sim, obs, warmup)
    sm, om = s.mean(dim=1), o.mean()
    sd, od = s - sm[:, None], o - om
    r = (sd * od).sum(dim=1) / torch.sqrt((sd ** 2).sum(dim=1) * (od ** 2).sum())
    n = o.numel()
    cv_s = torch.sqrt((sd ** 2).sum(dim=1) / (n - 1)) / sm
    cv_o = torch.sqrt((od ** 2).sum() / (n - 1)) / om
    return 1 - torch.sqrt((r - 1) ** 2 + (cv_s / cv_o - 1) ** 2 + (sm / om - 1) ** 2)


def kge_2009(sim, obs, warmup: int = 0):
    import torch

    s, o = _prep(sim, obs, warmup)
    sm, om = s.mean(dim=1), o.mean()
    sd, od = s - sm[:, None], o - om
    r = (sd * od).sum(dim=1) / torch.sqrt((sd ** 2).sum(dim=1) * (od ** 2).sum())
    alpha = torch.sqrt((sd ** 2).sum(dim=1)) / torch.sqrt((od ** 2).sum())
    return 1 - torch.sqrt((r - 1) ** 2 + (alpha - 1) ** 2 + (sm / om - 1) ** 2)


def r_squared(sim, obs, warmup: int = 0):
    import torch

    s, o = _prep(sim, obs, warmup)
    sd, od = s - s.mean(dim=1)[:, None], o - o.mean()
    r = (sd * od).sum(dim=1) / torch.sqrt((sd ** 2).sum(dim=1) * (od ** 2).sum())
    return r * r


def me(sim, obs, warmup: int = 0):
    s, o = _prep(sim, obs, warmup)
    return (s - o).mean(dim=1)


def mae(sim, obs, warmup: int = 0):
    s, o = _prep(sim, obs, warmup)
    return (s - o).abs().mean(dim=1)


def mse(sim, obs, warmup: int = 0):
    s, o = _prep(sim, obs, warmup)
    return ((s - o) ** 2).mean(dim=1)


def rmse(sim, obs, warmup: int = 0):
    return mse(sim, obs, warmup).sqrt()


def nrmse_mean(sim, obs, warmup: int = 0):
    s, o = _prep(sim, obs, warmup)
    return ((s - o) ** 2).mean(dim=1).sqrt() / o.mean()
