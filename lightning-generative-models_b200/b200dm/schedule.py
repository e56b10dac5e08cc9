"""Host-side noise schedules and the 13 fp32 coefficient tables of GaussianDiffusion.

Restates reference models/generative/diffusion/ddpm.py:491-529 (beta schedules) and :577-662
(buffer registration).  Runs once at construction on the host in float64 and is cast to float32
afterwards exactly like the reference's `register_buffer(name, val.to(torch.float32))`, because the
rounding order is part of parity (SURVEY.md §4 KATs).  Not a hot-path component.
"""
from __future__ import annotations

import math
from typing import Dict

import torch

_F64 = torch.float64


def _betas_from_alpha_bar(abar: torch.Tensor) -> torch.Tensor:
    abar = abar / abar[0]
    return torch.clip(1 - abar[1:] / abar[:-1], 0, 0.999)


def beta_schedule(name: str, T: int, **kw) -> torch.Tensor:
    if name == "linear":                                   # ddpm.py:491-498
        k = 1000 / T
        return torch.linspace(k * 1e-4, k * 2e-2, T, dtype=_F64)
    grid = torch.linspace(0, T, T + 1, dtype=_F64) / T
    if name == "cosine":                                   # ddpm.py:501-511
        s = kw.get("s", 0.008)
        return _betas_from_alpha_bar(torch.cos((grid + s) / (1 + s) * math.pi * 0.5) ** 2)
    if name == "sigmoid":                                  # ddpm.py:514-529
        lo, hi, tau = kw.get("start", -3), kw.get("end", 3), kw.get("tau", 1)
        s_lo, s_hi = torch.tensor(lo / tau).sigmoid(), torch.tensor(hi / tau).sigmoid()
        abar = (s_hi - ((grid * (hi - lo) + lo) / tau).sigmoid()) / (s_hi - s_lo)
        return _betas_from_alpha_bar(abar)
    raise ValueError(f"unknown beta schedule {name}")


def make_buffers(timesteps: int = 1000, schedule: str = "sigmoid", objective: str = "pred_v",
                 min_snr_loss_weight: bool = False, min_snr_gamma: float = 5,
                 schedule_fn_kwargs: dict | None = None) -> Dict[str, torch.Tensor]:
    b = beta_schedule(schedule, timesteps, **(schedule_fn_kwargs or {}))
    a = 1.0 - b
    abar = torch.cumprod(a, dim=0)
    abar_prev = torch.cat((torch.ones(1, dtype=_F64), abar[:-1]))
    post_var = b * (1.0 - abar_prev) / (1.0 - abar)
    snr = abar / (1 - abar)
    clipped = snr.clone()
    if min_snr_loss_weight:
        clipped.clamp_(max=min_snr_gamma)
    if objective == "pred_noise":
        lw = clipped / snr
    elif objective == "pred_x0":
        lw = clipped
    elif objective == "pred_v":
        lw = clipped / (snr + 1)
    else:
        raise ValueError(f"unknown objective {objective}")
    tables = {
        "betas": b,
        "alphas_cumprod": abar,
        "alphas_cumprod_prev": abar_prev,
        "sqrt_alphas_cumprod": abar.sqrt(),
        "sqrt_one_minus_alphas_cumprod": (1.0 - abar).sqrt(),
        "log_one_minus_alphas_cumprod": (1.0 - abar).log(),
        "sqrt_recip_alphas_cumprod": (1.0 / abar).sqrt(),
        "sqrt_recipm1_alphas_cumprod": (1.0 / abar - 1).sqrt(),
        "posterior_variance": post_var,
        "posterior_log_variance_clipped": post_var.clamp(min=1e-20).log(),
        "posterior_mean_coef1": b * abar_prev.sqrt() / (1.0 - abar),
        "posterior_mean_coef2": (1.0 - abar_prev) * a.sqrt() / (1.0 - abar),
        "loss_weight": lw,
    }
    return {k: v.to(torch.float32) for k, v in tables.items()}


def ddim_time_pairs(num_timesteps: int, sampling_timesteps: int):
    """ddpm.py:792-798 — [(T-1, ...), ..., (t_last, -1)]."""
    times = torch.linspace(-1, num_timesteps - 1, steps=sampling_timesteps + 1)
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))
