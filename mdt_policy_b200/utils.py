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


def rand_v_diffusion(shape, sigma_data=1.0, min_value=0.0, max_value=float("inf"), device="cpu", dtype=torch.float32):
    """utils.py:176-181 -- truncated v-diffusion density: uniform in the arctan CDF, mapped back with tan"""
    cdf_lo = math.atan(min_value / sigma_data) * 2 / math.pi
    cdf_hi = math.atan(max_value / sigma_data) * 2 / math.pi
    u = torch.rand(shape, device=device, dtype=dtype) * (cdf_hi - cdf_lo) + cdf_lo
    return torch.tan(u * math.pi / 2) * sigma_data


def rand_split_log_normal(shape, loc, scale_1, scale_2, device="cpu", dtype=torch.float32):
    """utils.py:184-191 -- half-normal magnitudes placed left (scale_1) or right (scale_2) of loc in log space"""
    mag = torch.randn(shape, device=device, dtype=dtype).abs()
    side = torch.rand(shape, device=device, dtype=dtype)
    left = side < scale_1 / (scale_1 + scale_2)
    return torch.where(left, loc - mag * scale_1, loc + mag * scale_2).exp()


def rand_discrete(shape, values, device="cpu", dtype=torch.float32):
    """utils.py:194-198 -- uniform draw from a 1-D tensor of candidate sigmas"""
    idx = torch.randint(0, len(values), shape, device=device)
    return values.index_select(0, idx).to(dtype)
