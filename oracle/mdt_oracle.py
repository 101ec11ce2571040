"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the MDT denoising hot path.

A functional, from-scratch restatement of the reference algorithm (score network +
EDM preconditioner + samplers) on top of plain torch CPU ops.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this file, and only as the checker / the CPU arm -- the product package
``mdt_policy_b200`` never imports it and has no CPU fallback.

Why torch and not numpy/C: the reference *is* torch (ATen ``linear``/``layer_norm``/SDPA/
``gelu``), so a torch restatement (a) reproduces its fp32 rounding to ~1e-6 and (b) times
the same MKL/oneDNN kernels the reference would run on the host cores, which is what the
``cpu_baseline`` is supposed to represent (kind = "port").  Running the same functions on
a float64 state dict gives the high-precision truth used to separate "GPU error" from
"fp32 reference noise".

Pinning: the reference has no tests/golden vectors for this path, so
the oracle is pinned against outputs of the reference itself, imported in the authoring
container by ``oracle/ref_shim.py``; the vectors are committed under ``tests/golden/``
together with ``tests/golden/make_golden.py`` (the generating script), and
``tests/test_oracle_vs_golden.py`` re-checks them on every run.

Every function cites the reference lines it follows (paths relative to /root/reference).
Tensors: ``P`` is a flat ``{name: tensor}`` dict with the reference's state-dict keys
(prefix ``inner_model.``).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch
import torch.nn.functional as F


@dataclass
class OracleCfg:
    """Subset of conf/model/model/mdtv_transformer.yaml the arithmetic depends on."""
    embed_dim: int = 384
    n_heads: int = 8
    n_enc_layers: int = 4
    n_dec_layers: int = 4
    action_dim: int = 7
    action_seq_len: int = 10
    goal_seq_len: int = 1
    sigma_data: float = 0.5
    variant: str = "mdtv"      # "mdtv" (mdtv_transformer.py) | "mdt" (mdt_transformer.py)
    prefix: str = "inner_model."


# --------------------------------------------------------------------------- blocks

def _ln(x, w, b=None):
    # transformer_blocks.py:29-38 (bias-free LayerNorm, eps 1e-5); ln3 is nn.LayerNorm (:205)
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def _attention(P, pre, n_heads, x, kv_src, mask_mode):
    """transformer_blocks.py:119-158.  q/k/v have biases, c_proj has none (bias=False).

    mask_mode: "full" (encoder), "causal" -> mask[i, j] = (j <= i), top-left aligned also
    for the non-square cross-attention (Block passes `causal` to cross_att, :201-204).
    """
    B, Tq, C = x.shape
    Tk = kv_src.shape[1]
    hd = C // n_heads
    q = F.linear(x, P[pre + "query.weight"], P[pre + "query.bias"]).view(B, Tq, n_heads, hd).transpose(1, 2)
    k = F.linear(kv_src, P[pre + "key.weight"], P[pre + "key.bias"]).view(B, Tk, n_heads, hd).transpose(1, 2)
    v = F.linear(kv_src, P[pre + "value.weight"], P[pre + "value.bias"]).view(B, Tk, n_heads, hd).transpose(1, 2)
    mask = None
    if mask_mode == "causal":
        i = torch.arange(Tq, device=x.device).view(Tq, 1)
        j = torch.arange(Tk, device=x.device).view(1, Tk)
        mask = j <= i
    y = F.scaled_dot_product_attention(q, k, v, attn_mask=mask)  # scale 1/sqrt(hd)
    y = y.transpose(1, 2).reshape(B, Tq, C)
    return F.linear(y, P[pre + "c_proj.weight"], P.get(pre + "c_proj.bias"))


def _mlp(P, pre, x):
    # transformer_blocks.py:175-180: c_fc -> exact (erf) GELU -> c_proj, no biases
    h = F.gelu(F.linear(x, P[pre + "c_fc.weight"], P.get(pre + "c_fc.bias")))
    return F.linear(h, P[pre + "c_proj.weight"], P.get(pre + "c_proj.bias"))


def _encoder_block(P, pre, nh, x):
    # Block.forward, transformer_blocks.py:209-214 (no cross-attention in the encoder)
    x = x + _attention(P, pre + "attn.", nh, _ln(x, P[pre + "ln_1.weight"]), _ln(x, P[pre + "ln_1.weight"]), "full")
    x = x + _mlp(P, pre + "mlp.", _ln(x, P[pre + "ln_2.weight"]))
    return x


def _conditioned_block(P, pre, nh, x, c, ctx):
    # ConditionedBlock.forward, transformer_blocks.py:292-309; AdaLNZero :245-260;
    # modulate(x, shift, scale) = shift + x*scale (:262-263) -- note: no "1 +".
    mod = F.linear(F.silu(c), P[pre + "adaLN_zero.modulation.1.weight"], P[pre + "adaLN_zero.modulation.1.bias"])
    sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=-1)
    a = sh1 + _ln(x, P[pre + "ln_1.weight"]) * sc1
    x = x + g1 * _attention(P, pre + "attn.", nh, a, a, "causal")
    q_in = _ln(x, P[pre + "ln3.weight"], P[pre + "ln3.bias"])
    x = x + _attention(P, pre + "cross_att.", nh, q_in, ctx, "causal")
    a = sh2 + _ln(x, P[pre + "ln_2.weight"]) * sc2
    x = x + g2 * _mlp(P, pre + "mlp.", a)
    return x


# --------------------------------------------------------------------------- score network

def _goal_mlp(P, pre, g):
    # nn.Sequential(Linear, GELU, Linear)  mdtv_transformer.py:86-101
    return F.linear(F.gelu(F.linear(g, P[pre + "0.weight"], P[pre + "0.bias"])), P[pre + "2.weight"], P[pre + "2.bias"])


def encode(P, cfg: OracleCfg, state: dict, goal, context_only: bool = False):
    """forward_enc_only, mdtv_transformer.py:213-222 (eval mode, uncond=False).

    context_only=True is the GCDenoiser.forward_context_only entry.  It only matters for the
    MDT variant, whose forward() goes through enc_only_forward (always goal_emb,
    mdt_transformer.py:211-215) while forward_enc_only is modality-aware (:257-262).

    state: {'state_images': (B, n_tok, obs_dim), 'modality': 'lang'|'vis'} for MDT-V;
           {'static': (B,1,obs), 'gripper': (B,1,obs), 'modality': ...} for MDT
           (mdt_transformer.py:211-229, :309-324).
    """
    p = cfg.prefix
    if goal.dim() == 2:                                  # preprocess_goals :246-250
        goal = goal.unsqueeze(1)
    n_state = state["state_images"].shape[1] if cfg.variant == "mdtv" else 1
    if goal.shape[1] == n_state and cfg.goal_seq_len == 1:
        goal = goal[:, :1, :]
    lang = state.get("modality") == "lang" and (p + "lang_emb.0.weight") in P
    if cfg.variant == "mdt" and not context_only:
        lang = False
    g = _goal_mlp(P, p + ("lang_emb." if lang else "goal_emb."), goal)       # :268-273
    if cfg.variant == "mdtv":
        s = F.linear(state["state_images"], P[p + "tok_emb.weight"], P[p + "tok_emb.bias"])   # :260-266
        x = torch.cat([g, s], dim=1)                     # concatenate_inputs :284-300; pos_emb NOT applied
    else:
        st = F.linear(state["static"].float(), P[p + "tok_emb.weight"], P[p + "tok_emb.bias"])
        gr = F.linear(state["gripper"].float(), P[p + "incam_embed.weight"], P[p + "incam_embed.bias"])
        s = torch.stack((st, gr), dim=2).reshape(st.shape[0], 2, cfg.embed_dim)
        pe = P[p + "pos_emb"]                            # mdt_transformer.py:318-324
        g = g + pe[:, : cfg.goal_seq_len, :]
        s = s + pe[:, cfg.goal_seq_len: cfg.goal_seq_len + 1, :]
        x = torch.cat([g, s], dim=1)
    for l in range(cfg.n_enc_layers):
        x = _encoder_block(P, f"{p}encoder.blocks.{l}.", cfg.n_heads, x)
    return _ln(x, P[p + "encoder.ln.weight"])            # transformer_blocks.py:379


def sigma_embedding(P, cfg: OracleCfg, sigma):
    """process_sigma_embeddings :238-244 + SinusoidalPosEmb :13-25 + sigma_emb :169-174."""
    p = cfg.prefix
    e = sigma.log() / 4
    half = cfg.embed_dim // 2
    f = torch.exp(torch.arange(half, dtype=e.dtype, device=e.device) * -(math.log(10000) / (half - 1)))
    ang = e[:, None] * f[None, :]
    pe = torch.cat((ang.sin(), ang.cos()), dim=-1)
    h = F.mish(F.linear(pe, P[p + "sigma_emb.1.weight"], P[p + "sigma_emb.1.bias"]))
    return F.linear(h, P[p + "sigma_emb.3.weight"], P[p + "sigma_emb.3.bias"]).unsqueeze(1)   # (B,1,d)


def decode(P, cfg: OracleCfg, ctx, actions, sigma):
    """forward_dec_only, mdtv_transformer.py:224-236 (embed dropout p=0)."""
    p = cfg.prefix
    c = sigma_embedding(P, cfg, sigma)
    x = F.linear(actions, P[p + "action_emb.weight"], P[p + "action_emb.bias"])
    # (MDT's dec_only_forward, mdt_transformer.py:231-242, is identical: no pos_emb on actions)
    for l in range(cfg.n_dec_layers):
        x = _conditioned_block(P, f"{p}decoder.blocks.{l}.", cfg.n_heads, x, c, ctx)
    x = _ln(x, P[p + "decoder.ln.weight"])
    return F.linear(x, P[p + "action_pred.weight"], P[p + "action_pred.bias"])


def inner_forward(P, cfg, state, actions, goal, sigma):
    # MDTVTransformer.forward :208-211 -- encoder re-run on every call, as the reference does
    return decode(P, cfg, encode(P, cfg, state, goal), actions, sigma)


# --------------------------------------------------------------------------- EDM wrapper

def get_scalings(sigma, sigma_data):
    # score_wrappers.py:31-43
    c_skip = sigma_data ** 2 / (sigma ** 2 + sigma_data ** 2)
    c_out = sigma * sigma_data / (sigma ** 2 + sigma_data ** 2) ** 0.5
    c_in = 1 / (sigma ** 2 + sigma_data ** 2) ** 0.5
    return c_skip, c_out, c_in


def denoiser_forward(P, cfg, state, action, goal, sigma):
    # GCDenoiser.forward, score_wrappers.py:65-80
    c_skip, c_out, c_in = [s[:, None, None] for s in get_scalings(sigma, cfg.sigma_data)]
    return inner_forward(P, cfg, state, action * c_in, goal, sigma) * c_out + action * c_skip


def denoiser_loss(P, cfg, state, action, goal, noise, sigma):
    # GCDenoiser.loss, score_wrappers.py:45-63 (eval-mode arithmetic: no dropout / goal masking)
    c_skip, c_out, c_in = [s[:, None, None] for s in get_scalings(sigma, cfg.sigma_data)]
    noised = action + noise * sigma[:, None, None]
    out = inner_forward(P, cfg, state, noised * c_in, goal, sigma)
    target = (action - c_skip * noised) / c_out
    return (out - target).pow(2).flatten(1).mean(), out


def forward_context_only(P, cfg, state, action, goal, sigma):
    # score_wrappers.py:82-97
    return encode(P, cfg, state, goal, context_only=True)


# --------------------------------------------------------------------------- schedules

def _append_zero(s):
    return torch.cat([s, s.new_zeros([1])])


def get_sigmas_exponential(n, sigma_min, sigma_max):
    # gc_sampling.py:35-38
    return _append_zero(torch.linspace(math.log(sigma_max), math.log(sigma_min), n).exp())


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0):
    # gc_sampling.py:26-32
    ramp = torch.linspace(0, 1, n)
    lo, hi = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    return _append_zero((hi + ramp * (lo - hi)) ** rho)


def get_sigmas_linear(n, sigma_min, sigma_max):
    # gc_sampling.py:41-44
    return _append_zero(torch.linspace(sigma_max, sigma_min, n))


# --------------------------------------------------------------------------- samplers
# All take model(x, sigma_vec) -> denoised, i.e. the (state, goal) arguments are bound.

def sample_ddim(model, x, sigmas):
    # gc_sampling.py:922-951 ; last step: log(0) = -inf -> x = 0*x + 1*D
    ones = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        d = model(x, sigmas[i] * ones)
        t, t_next = -sigmas[i].log(), -sigmas[i + 1].log()
        h = t_next - t
        x = ((-t_next).exp() / (-t).exp()) * x - torch.expm1(-h) * d
    return x


def sample_euler(model, x, sigmas):
    # gc_sampling.py:164-210 with s_churn = 0 (gamma = 0, no noise injection)
    ones = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        den = model(x, sigmas[i] * ones)
        d = (x - den) / sigmas[i]                          # to_d :91-93
        x = x + d * (sigmas[i + 1] - sigmas[i])
    return x


def sample_heun(model, x, sigmas):
    # gc_sampling.py:256-311 with s_churn = 0
    ones = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        den = model(x, sigmas[i] * ones)
        d = (x - den) / sigmas[i]
        dt = sigmas[i + 1] - sigmas[i]
        if sigmas[i + 1] == 0:
            x = x + d * dt
        else:
            x2 = x + d * dt
            den2 = model(x2, sigmas[i + 1] * ones)
            d2 = (x2 - den2) / sigmas[i + 1]
            x = x + (d + d2) / 2 * dt
    return x


def sample_dpmpp_2m(model, x, sigmas):
    # gc_sampling.py:699-733
    ones = x.new_ones([x.shape[0]])
    old = None
    for i in range(len(sigmas) - 1):
        den = model(x, sigmas[i] * ones)
        t, t_next = -sigmas[i].log(), -sigmas[i + 1].log()
        h = t_next - t
        if old is None or sigmas[i + 1] == 0:
            x = ((-t_next).exp() / (-t).exp()) * x - torch.expm1(-h) * den
        else:
            h_last = t - (-sigmas[i - 1].log())
            r = h_last / h
            dd = (1 + 1 / (2 * r)) * den - (1 / (2 * r)) * old
            x = ((-t_next).exp() / (-t).exp()) * x - torch.expm1(-h) * dd
        old = den
    return x


SAMPLERS = {"ddim": sample_ddim, "euler": sample_euler, "heun": sample_heun, "dpmpp_2m": sample_dpmpp_2m}


def sample(P, cfg, state, x_T, goal, sigmas, sampler="ddim"):
    """MDTVAgent.sample_loop (mdtv_agent.py:593-658) for the fused samplers: N sequential
    full GCDenoiser.forward evaluations (encoder recomputed each time, like the reference)."""
    def model(x, s):
        return denoiser_forward(P, cfg, state, x, goal, s)
    with torch.no_grad():
        return SAMPLERS[sampler](model, x_T, sigmas)


def evals_per_call(sampler, n_steps):
    """Number of score-network evaluations one sampling call performs."""
    return 2 * n_steps - 1 if sampler == "heun" else n_steps
