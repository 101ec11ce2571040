"""Shared test plumbing: golden fixtures, oracle parameter dicts, product-model construction."""
from __future__ import annotations

import json
import os

import numpy as np
import torch

from mdt_policy_b200.synthetic import synthetic_state_dict, synthetic_inputs
from oracle import mdt_oracle as orc

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    meta = json.loads(str(z["meta"]))
    return meta, {k: torch.from_numpy(z[k]) for k in z.files if k != "meta"}


def manifest(key):
    with open(os.path.join(GOLDEN, "param_manifest.json")) as f:
        return [(n, tuple(s)) for n, s in json.load(f)[key]]


def mdtv_shapes(n_enc, n_dec, d=384, goal_dim=512, obs_dim=384):
    """(name, shape) list of GCDenoiser(MDTVTransformer) for arbitrary layer counts, derived from the 4+4 manifest."""
    out = []
    for n, s in manifest("mdtv_4_4"):
        if ".blocks." in n:
            stack, rest = n.split(".blocks.")
            idx, tail = rest.split(".", 1)
            if int(idx) != 0:
                continue
            for l in range(n_enc if stack.endswith("encoder") else n_dec):
                out.append((f"{stack}.blocks.{l}.{tail}", s))
        else:
            out.append((n, s))
    # restore reference ordering: encoder blocks, encoder.ln, decoder blocks, decoder.ln are already grouped by stack
    return out


def oracle_params(shapes, seed, profile, dtype=torch.float32):
    return {k: v.to(dtype) for k, v in synthetic_state_dict(shapes, seed, profile).items()}


def mdtv_inner_cfg(n_enc=4, n_dec=4, **over):
    cfg = dict(
        _target_="mdt_policy_b200.networks.MDTVTransformer", action_dim=7, obs_dim=384, goal_dim=512, proprio_dim=8,
        goal_conditioned=True, embed_dim=384, n_dec_layers=n_dec, n_enc_layers=n_enc, n_obs_token=3, goal_seq_len=1,
        obs_seq_len=1, action_seq_len=10, embed_pdrob=0, goal_drop=0, attn_pdrop=0.3, resid_pdrop=0.1, mlp_pdrop=0.05,
        n_heads=8, device="cuda", linear_output=True, use_rot_embed=False, use_abs_pos_emb=True, bias=False,
        use_ada_conditioning=True, use_noise_encoder=False, use_modality_encoder=True, use_mlp_goal=True)
    cfg.update(over)
    return cfg


def mdt_inner_cfg(**over):
    cfg = dict(
        _target_="mdt_policy_b200.networks.MDTTransformer", action_dim=7, obs_dim=512, goal_dim=512, proprio_dim=8,
        goal_conditioned=True, embed_dim=512, n_dec_layers=6, n_enc_layers=4, goal_seq_len=1, obs_seq_len=1,
        action_seq_len=10, embed_pdrob=0, goal_drop=0, attn_pdrop=0.3, resid_pdrop=0.1, mlp_pdrop=0.05, n_heads=8,
        device="cuda", linear_output=True, use_rot_embed=False, use_abs_pos_emb=True, bias=False,
        use_ada_conditioning=True, use_noise_encoder=False, use_modality_encoder=True, use_mlp_goal=True)
    cfg.update(over)
    return cfg


def build_product(inner_cfg, seed, profile, device="cuda"):
    """Product GCDenoiser with synthetic weights loaded through load_state_dict (the checkpoint path)."""
    from mdt_policy_b200 import GCDenoiser
    model = GCDenoiser(inner_cfg, sigma_data=0.5)
    sd = synthetic_state_dict([(n, p.shape) for n, p in model.named_parameters()], seed, profile)
    model.load_state_dict(sd, strict=True)
    return model.to(device).eval()


def perceiver_shapes(depth, n_lat, d=384, inner=512, n_time=1):
    """(name, shape) list of the reference PerceiverResampler (perceiver_resampler.py:86-117), in registration order."""
    s = [("latents", (n_lat, d)), ("time_pos_emb", (n_time, 1, d))]
    for l in range(depth):
        p = f"layers.{l}."
        s += [(p + "0.norm_media.weight", (d,)), (p + "0.norm_media.bias", (d,)), (p + "0.norm_latents.weight", (d,)),
              (p + "0.norm_latents.bias", (d,)), (p + "0.to_q.weight", (inner, d)), (p + "0.to_k.weight", (inner, d)),
              (p + "0.to_v.weight", (inner, d)), (p + "0.to_out.weight", (d, inner)), (p + "1.0.weight", (d,)), (p + "1.0.bias", (d,)),
              (p + "1.1.weight", (4 * d, d)), (p + "1.3.weight", (d, 4 * d))]
    return s + [("norm.weight", (d,)), ("norm.bias", (d,))]


def perceiver_state(named_shapes, seed, profile="trained"):
    """synthetic PerceiverResampler weights.  latents and time embedding ~ N(0, 1) (their reference initialisation).
    profile "trained": LayerNorms ~ 1 + N(0, 0.1^2) / N(0, 0.1^2), projections ~ N(0, 1/fan_in) -- O(1) activations everywhere, which
    drives the softmax of layers >= 2 into saturation (scores sigma ~ 27): a deliberately ill-conditioned stress case.
    profile "init": LayerNorms 1 / 0, projections ~ N(0, 0.02^2) -- the regime of a freshly initialised model."""
    from mdt_policy_b200.synthetic import synthetic_tensor
    out = {}
    for n, shp in named_shapes:
        if n in ("latents", "time_pos_emb"):
            out[n] = synthetic_tensor(n, shp, seed, "init") * 50.0
        elif "norm" in n or ".1.0." in n:
            out[n] = synthetic_tensor("p." + n.replace("norm", "ln_norm").replace(".1.0.", ".1.ln_0."), shp, seed, profile)
        else:
            out[n] = synthetic_tensor(n, shp, seed, profile)
    return out
