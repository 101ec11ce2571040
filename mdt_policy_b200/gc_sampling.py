"""Noise schedules and samplers with the call contract of mdt/models/edm_diffusion/gc_sampling.py.

``sample_ddim / sample_euler / sample_heun / sample_dpmpp_2m`` keep the reference signature
``(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None, ...)``.
When ``model`` is this package's ``GCDenoiser`` and nothing in the call needs per-step host interaction
(no callback, no scaler, no churn) the whole loop runs as ONE CUDA graph (``GCDenoiser.sample`` ->
``mdtb200_sample``): encoder + cross-attention K/V once, the AdaLN table for all steps once, then N fused
steps whose sampler update is folded into the output-head kernel.  Otherwise the generic loop below calls
``model(...)`` once per evaluation exactly like the reference (each call = one ``mdtb200_denoise``).

Samplers of the reference that are not restated here (lms, dpm_2, dpmpp_2s, ...) take any callable with the
``model(state, action, goal, sigma)`` contract, so the reference's own functions run unchanged on top of this
package's ``GCDenoiser``.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import utils

# --------------------------------------------------------------------------------------------------- schedules


def append_zero(action):
    return torch.cat([action, action.new_zeros([1])])


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7., device='cpu'):
    """gc_sampling.py:26-32 -- Karras et al. (2022) rho-schedule."""
    ramp = torch.linspace(0, 1, n)
    lo, hi = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    return append_zero((hi + ramp * (lo - hi)) ** rho).to(device)


def get_sigmas_exponential(n, sigma_min, sigma_max, device='cpu'):
    """gc_sampling.py:35-38 -- log-linear schedule (the default: conf/model/mdtv_agent.yaml noise_scheduler)."""
    return append_zero(torch.linspace(math.log(sigma_max), math.log(sigma_min), n, device=device).exp())


def get_sigmas_linear(n, sigma_min, sigma_max, device='cpu'):
    """gc_sampling.py:41-44"""
    return append_zero(torch.linspace(sigma_max, sigma_min, n, device=device))


def cosine_beta_schedule(n, s=0.008, device='cpu'):
    """gc_sampling.py:47-58"""
    steps = n + 1
    grid = np.linspace(0, steps, steps)
    acp = np.cos(((grid / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    acp = acp / acp[0]
    betas = np.clip(1 - (acp[1:] / acp[:-1]), a_min=0, a_max=0.999)
    return append_zero(torch.tensor(np.flip(betas).copy(), device=device, dtype=torch.float32))


def get_sigmas_ve(n, sigma_min=0.02, sigma_max=100, device='cpu'):
    """gc_sampling.py:61-68"""
    t = torch.linspace(0, n + 1, n, device=device)
    return append_zero(torch.sqrt((sigma_max ** 2) * ((sigma_min ** 2 / sigma_max ** 2) ** (t / (n - 1)))))


def get_sigmas_vp(n, beta_d=19.9, beta_min=0.1, eps_s=1e-3, device='cpu'):
    """gc_sampling.py:84-88"""
    t = torch.linspace(1, eps_s, n, device=device)
    return append_zero(torch.sqrt(torch.exp(beta_d * t ** 2 / 2 + beta_min * t) - 1))


def get_iddpm_sigmas(n, sigma_min=0.02, sigma_max=100, M=1000, j_0=0, C_1=0.001, C_2=0.008, device='cpu'):
    """gc_sampling.py:71-81 (fp64 recursion, fp32 result)"""
    idx = torch.arange(n, dtype=torch.float64, device=device)
    u = torch.zeros(M + 1, dtype=torch.float64, device=device)

    def alpha_bar(j):
        return (0.5 * np.pi * j / M / (C_2 + 1)).sin() ** 2

    for j in torch.arange(M, j_0, -1, device=device):
        u[j - 1] = ((u[j] ** 2 + 1) / (alpha_bar(j - 1) / alpha_bar(j)).clip(min=C_1) - 1).sqrt()
    kept = u[torch.logical_and(u >= sigma_min, u <= sigma_max)]
    sig = kept[((len(kept) - 1) / (n - 1) * idx).round().to(torch.int64)]
    return append_zero(sig).to(torch.float32)


# --------------------------------------------------------------------------------------------------- helpers

def to_d(action, sigma, denoised):
    """Karras ODE derivative (gc_sampling.py:91-93)."""
    return (action - denoised) / utils.append_dims(sigma, action.ndim)


def get_ancestral_step(sigma_from, sigma_to, eta=1.):
    """gc_sampling.py:102-109"""
    if not eta:
        return sigma_to, 0.
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    sigma_down = (sigma_to ** 2 - sigma_up ** 2) ** 0.5
    return sigma_down, sigma_up


def _fused_ok(model, scaler, extra_args, callback) -> bool:
    return (hasattr(model, "sample") and hasattr(model, "inner_model") and scaler is None and callback is None
            and not extra_args)


def _on_cuda(*tensors) -> bool:
    return all(t.is_cuda for t in tensors if isinstance(t, torch.Tensor))


# --------------------------------------------------------------------------------------------------- samplers
# Every sampler below first tries the fused CUDA path (GCDenoiser.sample -> mdtb200_sample: the whole loop is one CUDA graph).
# Anything the graph does not cover (callbacks, scalers, churn, extra_args, foreign models)
# runs through ONE generic driver, `_integrate`, parameterised by a per-sampler update rule; the rules are the formulas of
# mdt/models/edm_diffusion/gc_sampling.py (cited per rule), the signatures and callback payloads are the reference's.

def _log_step(x, denoised, sigma, sigma_next):
    """exponential-integrator step in t = -log sigma: x <- (s'/s) x - expm1(-(t' - t)) D  (DPM-Solver-1 == DDIM, :948-950)"""
    t, t_next = -sigma.log(), -sigma_next.log()
    return ((-t_next).exp() / (-t).exp()) * x - torch.expm1(t - t_next) * denoised


def _rule_ddim(ctx, x, denoised, sigma, sigma_next):                       # :922-951
    return _log_step(x, denoised, sigma, sigma_next)


def _rule_euler(ctx, x, denoised, sigma, sigma_next):                      # :201-207
    return x + to_d(x, sigma, denoised) * (sigma_next - sigma)


def _rule_heun(ctx, x, denoised, sigma, sigma_next):                       # :296-309
    d, dt = to_d(x, sigma, denoised), sigma_next - sigma
    if sigma_next == 0:
        return x + d * dt
    x_mid = x + d * dt
    d_mid = to_d(x_mid, sigma_next, ctx["eval"](x_mid, sigma_next))
    return x + (d + d_mid) / 2 * dt


def _rule_euler_ancestral(ctx, x, denoised, sigma, sigma_next):            # :240-251
    sigma_down, sigma_up = get_ancestral_step(sigma, sigma_next, eta=ctx["eta"])
    x = x + to_d(x, sigma, denoised) * (sigma_down - sigma)
    if sigma_down > 0:
        x = x + torch.randn_like(x) * sigma_up
    return x


def _rule_dpmpp_2m(ctx, x, denoised, sigma, sigma_next):                   # :716-732
    prev, ctx["prev"] = ctx.get("prev"), (denoised, sigma)
    if prev is None or sigma_next == 0:
        return _log_step(x, denoised, sigma, sigma_next)
    old_denoised, sigma_prev = prev
    r = (sigma_prev.log() - sigma.log()) / (sigma.log() - sigma_next.log())
    return _log_step(x, (1 + 1 / (2 * r)) * denoised - (1 / (2 * r)) * old_denoised, sigma, sigma_next)


def _rule_dpm_2(ctx, x, denoised, sigma, sigma_next):                      # :340-372: midpoint rule, midpoint taken in log sigma
    d = to_d(x, sigma, denoised)
    if sigma_next == 0:
        return x + d * (sigma_next - sigma)
    sigma_mid = sigma.log().lerp(sigma_next.log(), 0.5).exp()
    x_mid = x + d * (sigma_mid - sigma)
    return x + to_d(x_mid, sigma_mid, ctx["eval"](x_mid, sigma_mid)) * (sigma_next - sigma)


def _rule_dpm_2_ancestral(ctx, x, denoised, sigma, sigma_next):            # :389-408
    sigma_down, sigma_up = get_ancestral_step(sigma, sigma_next, eta=ctx["eta"])
    if sigma_down == 0:
        return x + to_d(x, sigma, denoised) * (sigma_down - sigma)
    x = _rule_dpm_2(ctx, x, denoised, sigma, sigma_down)
    return x + torch.randn_like(x) * sigma_up


def _two_stage_log_step(ctx, x, denoised, sigma, sigma_to):
    """DPM-Solver++(2S) step from sigma to sigma_to (:983-990 / :906-913): half step in t = -log sigma, re-evaluation there, full step
    with the midpoint prediction"""
    t, t_to = -sigma.log(), -sigma_to.log()
    h = t_to - t
    s = t + 0.5 * h
    x_2 = ((-s).exp() / (-t).exp()) * x - torch.expm1(-h * 0.5) * denoised
    denoised_2 = ctx["eval"](x_2, (-s).exp())
    return ((-t_to).exp() / (-t).exp()) * x - torch.expm1(-h) * denoised_2


def _rule_dpmpp_2s(ctx, x, denoised, sigma, sigma_next):                   # :975-992
    if sigma_next == 0:
        return x + to_d(x, sigma, denoised) * (sigma_next - sigma)
    return _two_stage_log_step(ctx, x, denoised, sigma, sigma_next)


def _rule_dpmpp_2s_ancestral(ctx, x, denoised, sigma, sigma_next):         # :895-917 (the noise term is evaluated on every step)
    sigma_down, sigma_up = get_ancestral_step(sigma, sigma_next, eta=ctx["eta"])
    if sigma_down == 0:
        x = x + to_d(x, sigma, denoised) * (sigma_down - sigma)
    else:
        x = _two_stage_log_step(ctx, x, denoised, sigma, sigma_down)
    return x + ctx["noise_sampler"](sigma, sigma_next) * ctx["s_noise"] * sigma_up


def linear_multistep_coeff(order, t, i, j):
    """gc_sampling.py:413-427: integral over [t_i, t_{i+1}] of the j-th Lagrange basis polynomial through the last `order` nodes"""
    if order - 1 > i:
        raise ValueError(f'Order {order} too high for step {i}')

    def basis(tau):
        prod = 1.
        for k in range(order):
            if k != j:
                prod *= (tau - t[i - k]) / (t[i - j] - t[i - k])
        return prod
    from scipy import integrate
    return integrate.quad(basis, t[i], t[i + 1], epsrel=1e-4)[0]


def _rule_lms(ctx, x, denoised, sigma, sigma_next):                        # :452-465: Adams-Bashforth in sigma over the derivative history
    ds, i, order = ctx.setdefault("ds", []), ctx["i"], ctx["order"]
    ds.append(to_d(x, sigma, denoised))
    if len(ds) > order:
        ds.pop(0)
    cur = min(i + 1, order)
    coeffs = [linear_multistep_coeff(cur, ctx["sigmas_np"], i, j) for j in range(cur)]
    return x + sum(c * d for c, d in zip(coeffs, reversed(ds)))


def _integrate(rule, model, state, x, goal, sigmas, scaler=None, extra_args=None, callback=None, cb_key="x", churn=None, eta=1., **more):
    extra_args = {} if extra_args is None else extra_args
    ones = x.new_ones([x.shape[0]])
    ctx = {"eta": eta, "eval": lambda xx, sig: model(state, xx, goal, sig * ones, **extra_args)}
    ctx.update(more)
    n = len(sigmas) - 1
    for i in range(n):
        ctx["i"] = i
        sigma_hat = sigmas[i]
        if churn is not None:               # Karras "churn": the reference draws eps every step, also when gamma == 0 (:195)
            s_churn, s_tmin, s_tmax, s_noise = churn
            gamma = min(s_churn / n, 2 ** 0.5 - 1) if s_tmin <= sigmas[i] <= s_tmax else 0.
            eps = torch.randn_like(x) * s_noise
            sigma_hat = sigmas[i] * (gamma + 1)
            if gamma > 0:
                x = x + eps * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
        denoised = ctx["eval"](x, sigma_hat)
        if callback is not None:
            callback({cb_key: x, 'i': i, 'sigma': sigmas[i], 'sigma_hat': sigma_hat, 'denoised': denoised})
        x = rule(ctx, x, denoised, sigma_hat, sigmas[i + 1])
        if scaler is not None:
            x = scaler.clip_output(x)
    return x


def _fused(model, name, state, action, goal, sigmas, scaler, extra_args, callback):
    if _fused_ok(model, scaler, extra_args or {}, callback) and _on_cuda(action, goal):
        return model.sample(state, action, goal, sigmas, sampler=name)
    return None


@torch.no_grad()
def sample_ddim(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None, eta=1.):
    """DPM-Solver-1 / DDIM (gc_sampling.py:922-951)."""
    out = _fused(model, "ddim", state, action, goal, sigmas, scaler, extra_args, callback)
    return out if out is not None else _integrate(_rule_ddim, model, state, action, goal, sigmas, None, extra_args, callback, "action")


@torch.no_grad()
def sample_euler(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None,
                 s_churn=0., s_tmin=0., s_tmax=float('inf'), s_noise=1.):
    """Karras Algorithm 2 without the 2nd-order correction (gc_sampling.py:164-210)."""
    out = _fused(model, "euler", state, action, goal, sigmas, scaler, extra_args, callback) if s_churn == 0 else None
    return out if out is not None else _integrate(_rule_euler, model, state, action, goal, sigmas, scaler, extra_args, callback,
                                                  churn=(s_churn, s_tmin, s_tmax, s_noise))


@torch.no_grad()
def sample_heun(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None,
                s_churn=0., s_tmin=0., s_tmax=float('inf'), s_noise=1.):
    """Karras Algorithm 2 (Heun), gc_sampling.py:256-311: Euler predictor + trapezoidal corrector, plain Euler on
    the last step (sigma_next == 0)."""
    graph_ok = s_churn == 0 and float(sigmas[-1]) == 0.0 and bool((sigmas[:-1] > 0).all())
    out = _fused(model, "heun", state, action, goal, sigmas, scaler, extra_args, callback) if graph_ok else None
    return out if out is not None else _integrate(_rule_heun, model, state, action, goal, sigmas, scaler, extra_args, callback,
                                                  churn=(s_churn, s_tmin, s_tmax, s_noise))


@torch.no_grad()
def sample_euler_ancestral(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None,
                           disable=None, eta=1.):
    """gc_sampling.py:213-253 (stochastic: the noise of every step comes from torch's generator, as in the reference).
    Fused path: the draws are made up front in the reference's order (one randn_like per step whose sigma_down > 0) and the whole
    loop runs as one CUDA graph (mdtb200_sample_ancestral); the generic driver covers callbacks / scalers / foreign models."""
    if _fused_ok(model, scaler, extra_args or {}, callback) and _on_cuda(action, goal):
        sig = sigmas.detach().float().cpu()
        n = sig.numel() - 1
        noise = torch.zeros((n,) + tuple(action.shape), dtype=torch.float32, device=action.device)
        for i in range(n):
            if get_ancestral_step(sig[i], sig[i + 1], eta=eta)[0] > 0:
                noise[i] = torch.randn_like(action)
        return model.sample(state, action, goal, sigmas, sampler="euler_ancestral", noise=noise, eta=eta)
    return _integrate(_rule_euler_ancestral, model, state, action, goal, sigmas, scaler, extra_args, callback, eta=eta)


@torch.no_grad()
def sample_dpmpp_2m(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None):
    """DPM-Solver++(2M), gc_sampling.py:699-733."""
    out = _fused(model, "dpmpp_2m", state, action, goal, sigmas, scaler, extra_args, callback)
    return out if out is not None else _integrate(_rule_dpmpp_2m, model, state, action, goal, sigmas, None, extra_args, callback, "action")


# ---- the remaining samplers of the reference module (two score evaluations per step or a derivative history): generic driver only ----

def default_noise_sampler(x):
    """gc_sampling.py:97-99"""
    return lambda sigma, sigma_next: torch.randn_like(x)


@torch.no_grad()
def sample_dpm_2(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None,
                 s_churn=0., s_tmin=0., s_tmax=float('inf'), s_noise=1.):
    """DPM-Solver-2 / Karras Algorithm 2 with the midpoint in log sigma (gc_sampling.py:315-372)."""
    return _integrate(_rule_dpm_2, model, state, action, goal, sigmas, scaler, extra_args, callback, "action",
                      churn=(s_churn, s_tmin, s_tmax, s_noise))


@torch.no_grad()
def sample_dpm_2_ancestral(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None, eta=1.):
    """Ancestral sampling with DPM-Solver-2 steps (gc_sampling.py:375-410)."""
    return _integrate(_rule_dpm_2_ancestral, model, state, action, goal, sigmas, scaler, extra_args, callback, eta=eta)


@torch.no_grad()
def sample_lms(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None, order=4):
    """Linear multistep sampler (gc_sampling.py:430-466)."""
    return _integrate(_rule_lms, model, state, action, goal, sigmas, scaler, extra_args, callback, order=order,
                      sigmas_np=sigmas.detach().cpu().numpy())


@torch.no_grad()
def sample_dpmpp_2_with_lms(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None):
    """gc_sampling.py:797-831: the same update as sample_dpmpp_2m (the scaler is accepted and ignored there as well)."""
    out = _fused(model, "dpmpp_2m", state, action, goal, sigmas, None, extra_args, callback)
    return out if out is not None else _integrate(_rule_dpmpp_2m, model, state, action, goal, sigmas, None, extra_args, callback, "action")


@torch.no_grad()
def sample_dpmpp_2s(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None, eta=1.):
    """DPM-Solver++(2S) (gc_sampling.py:956-994)."""
    return _integrate(_rule_dpmpp_2s, model, state, action, goal, sigmas, scaler, extra_args, callback, "action")


@torch.no_grad()
def sample_dpmpp_2s_ancestral(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None,
                              eta=1., s_noise=1., noise_sampler=None):
    """Ancestral sampling with DPM-Solver++(2S) steps (gc_sampling.py:874-920)."""
    return _integrate(_rule_dpmpp_2s_ancestral, model, state, action, goal, sigmas, scaler, extra_args, callback, "action", eta=eta,
                      s_noise=s_noise, noise_sampler=default_noise_sampler(action) if noise_sampler is None else noise_sampler)


SAMPLERS = {
    "ddim": sample_ddim, "euler": sample_euler, "heun": sample_heun, "dpmpp_2m": sample_dpmpp_2m,
    "euler_ancestral": sample_euler_ancestral, "dpm": sample_dpm_2, "ancestral": sample_dpm_2_ancestral, "lms": sample_lms,
    "dpmpp_2s": sample_dpmpp_2s, "dpmpp_2s_ancestral": sample_dpmpp_2s_ancestral, "dpmpp_2_with_lms": sample_dpmpp_2_with_lms,
}
