"""Helpers of mdt/models/edm_diffusion/utils.py that the hot path uses (host side, plain torch)."""
from __future__ import annotations

import math

import torch


def append_dims(x: torch.Tensor, target_dims: int) -> torch.Tensor:
    """utils.py:146-151 -- trailing singleton dims until ``x.ndim == target_dims``."""
    missing = target_dims - x.ndim
    if missing < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x.reshape(x.shape + (1,) * missing)


def rand_log_logistic(shape, loc=0.0, scale=1.0, min_value=0.0, max_value=float("inf"), device="cpu", dtype=torch.float32):
    """Truncated log-logistic sigma density used for training (utils.py:159-166): fp64 uniform -> logit.
    Kept in torch so the RNG stream matches the reference's."""
    lo = torch.as_tensor(min_value, device=device, dtype=torch.float64)
    hi = torch.as_tensor(max_value, device=device, dtype=torch.float64)
    cdf_lo = torch.sigmoid((lo.log() - loc) / scale)
    cdf_hi = torch.sigmoid((hi.log() - loc) / scale)
    u = torch.rand(shape, device=device, dtype=torch.float64) * (cdf_hi - cdf_lo) + cdf_lo
    return (torch.logit(u) * scale + loc).exp().to(dtype)


def rand_log_normal(shape, loc=0.0, scale=1.0, device="cpu", dtype=torch.float32):
    """utils.py:154-156"""
    return (torch.randn(shape, device=device, dtype=dtype) * scale + loc).exp()


def rand_log_uniform(shape, min_value, max_value, device="cpu", dtype=torch.float32):
    """utils.py:169-173"""
    lo, hi = math.log(min_value), math.log(max_value)
    return (torch.rand(shape, device=device, dtype=dtype) * (hi - lo) + lo).exp()


def rand_uniform(shape, min_value, max_value, device="cpu", dtype=torch.float32):
    """utils.py:201-203"""
    return torch.rand(shape, device=device, dtype=dtype) * (max_value - min_value) + min_value
