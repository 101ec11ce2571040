"""Deterministic synthetic weights and inputs (numpy PCG64, independent of torch's RNG and of the reference).

Used by the golden-vector generator, the parity tests and bench.py so that the reference (in the authoring
container), the CPU oracle and the CUDA path all see bit-identical parameters and inputs without shipping
90 MB of weights: every tensor is regenerated from (seed, tensor name).
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def _rng(seed: int, name: str) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))


def synthetic_tensor(name: str, shape, seed: int, profile: str = "trained") -> torch.Tensor:
    """profile "init": the reference init (normal(0, 0.02) weights, zero biases, unit LN; mdtv_transformer.py:197-206).
    profile "trained": O(1) activations everywhere -- weights ~ N(0, 1/fan_in), biases ~ N(0, 0.1^2),
    LN weights 1 + N(0, 0.1^2) -- so every bias / gate / shift path carries signal."""
    g = _rng(seed, name)
    shape = tuple(shape)
    leaf = name.rsplit(".", 1)[-1]
    is_ln = ".ln" in name or name.endswith("ln.weight") or ".ln_" in name
    if profile == "init":
        if leaf == "bias":
            a = np.zeros(shape)
        elif is_ln:
            a = np.ones(shape)
        else:
            a = g.standard_normal(shape) * 0.02
    elif profile == "trained":
        if leaf == "bias":
            a = g.standard_normal(shape) * 0.1
        elif is_ln:
            a = 1.0 + g.standard_normal(shape) * 0.1
        elif len(shape) == 2:
            a = g.standard_normal(shape) / np.sqrt(shape[1])
        else:
            a = g.standard_normal(shape) * 0.02
    else:
        raise ValueError(profile)
    return torch.from_numpy(a.astype(np.float32))


def synthetic_state_dict(named_shapes, seed: int, profile: str = "trained") -> dict:
    """named_shapes: iterable of (name, shape) -- e.g. [(n, p.shape) for n, p in module.named_parameters()]."""
    return {n: synthetic_tensor(n, s, seed, profile) for n, s in named_shapes}


def synthetic_inputs(B: int, seed: int, n_state_tokens=3, obs_dim=384, goal_dim=512, T=10, A=7, sigma_max=80.0):
    """BASELINE.md section 2 inputs: state_images ~ N(0,1) (B,3,384), goal ~ N(0,1) (B,1,512), x_T = randn * sigma_max."""
    g = _rng(seed, "inputs")
    f = lambda *s: torch.from_numpy(g.standard_normal(s).astype(np.float32))
    return {"state_images": f(B, n_state_tokens, obs_dim), "goal": f(B, 1, goal_dim), "x_T": f(B, T, A) * sigma_max,
            "noise": f(B, T, A), "actions": torch.from_numpy(g.uniform(-1, 1, (B, T, A)).astype(np.float32))}
