"""Drop-in for mdt.models.edm_diffusion.score_wrappers.GCDenoiser (score_wrappers.py:18-100).

Same constructor (``inner_model`` config-or-module, ``sigma_data``), same methods and the same
``state_dict`` keys / ``named_parameters()`` order (``inner_model.*``).  ``forward`` fuses the Karras
preconditioner (c_in on the way in, c_out/c_skip on the way out) into the CUDA kernels; ``sample`` is the
additive fast path that runs a whole sampler loop in one CUDA graph.
"""
from __future__ import annotations

import importlib

import torch
from torch import nn

from .utils import append_dims


def _instantiate(cfg):
    """hydra.utils.instantiate when hydra is present (the reference's call, score_wrappers.py:28), otherwise a
    minimal ``_target_`` resolver so the module also works without Hydra installed."""
    if isinstance(cfg, nn.Module):
        return cfg
    try:
        import hydra  # type: ignore
        return hydra.utils.instantiate(cfg)
    except ImportError:
        cfg = dict(cfg)
        target = cfg.pop("_target_")
        cfg.pop("_recursive_", None)
        mod, cls = target.rsplit(".", 1)
        return getattr(importlib.import_module(mod), cls)(**cfg)


class GCDenoiser(nn.Module):
    """A Karras et al. (EDM) preconditioner around the score network."""

    def __init__(self, inner_model, sigma_data=1.):
        super().__init__()
        self.inner_model = _instantiate(inner_model)
        self.sigma_data = sigma_data
        # the fused precondition / sampler-update kernels need sigma_data inside the library handle
        self.inner_model.sigma_data = float(sigma_data)

    def get_scalings(self, sigma):
        """score_wrappers.py:31-43"""
        sd2 = self.sigma_data ** 2
        denom = sigma ** 2 + sd2
        return sd2 / denom, sigma * self.sigma_data / denom ** 0.5, 1 / denom ** 0.5

    def forward(self, state, action, goal, sigma, **kwargs):
        """score_wrappers.py:65-80: inner(state, action*c_in, goal, sigma) * c_out + action * c_skip.  Without autograd the
        three scalings are applied inside the CUDA kernels; with autograd (training path) they are plain tensor ops around
        the differentiable score network."""
        if self.inner_model.wants_grad(action, goal):
            c_skip, c_out, c_in = [append_dims(x, action.ndim) for x in self.get_scalings(sigma)]
            return self.inner_model(state, action * c_in, goal, sigma, **kwargs) * c_out + action * c_skip
        return self.inner_model(state, action, goal, sigma, _precondition=True, **kwargs)

    def loss(self, state, action, goal, noise, sigma, **kwargs):
        """score_wrappers.py:45-63.  With autograd enabled the score network runs through the training path
        (mdt_policy_b200/training.py: fp32 CUDA kernels with hand-written backward) and `loss.backward()` fills the
        parameter gradients; under torch.no_grad() the inference kernels evaluate the same value."""
        c_skip, c_out, c_in = [append_dims(x, action.ndim) for x in self.get_scalings(sigma)]
        noised_input = action + noise * append_dims(sigma, action.ndim)
        model_output = self.inner_model(state, noised_input * c_in, goal, sigma, **kwargs)
        target = (action - c_skip * noised_input) / c_out
        return (model_output - target).pow(2).flatten(1).mean(), model_output

    def forward_context_only(self, state, action, goal, sigma, **kwargs):
        """score_wrappers.py:82-97: encoder only (used by the CLA loss)."""
        return self.inner_model.forward_enc_only(state, action, goal, sigma, **kwargs)

    def sample(self, state, x_t, goal, sigmas, sampler: str = "ddim", **kwargs):
        """Fast path: the whole sampler loop (encode once + N steps) as one CUDA graph."""
        return self.inner_model.sample(state, x_t, goal, sigmas, sampler=sampler, **kwargs)

    def get_params(self):
        return self.inner_model.parameters()
