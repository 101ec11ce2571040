"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference PerceiverResampler (the module right before the denoising path:
mdt/models/networks/transformers/perceiver_resampler.py:11-163, called at mdt/models/mdtv_agent.py:392-403 on the (B, 1, 392, 384)
Voltron token sequence).  Plain torch ops, functional over a parameter dict with the reference's state-dict keys.  Pinned against
outputs of the reference itself: tests/golden/perceiver.npz (tests/golden/make_golden_perceiver.py).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _ln(x, P, pre):
    return F.layer_norm(x, (x.shape[-1],), P[pre + "weight"], P[pre + "bias"], 1e-5)


def attention_layer(P, pre, features, latents, heads, dim_head):
    """PerceiverAttentionLayer.forward (perceiver_resampler.py:32-83): the latents attend to [features ; latents]."""
    B, nf, _ = features.shape
    nq = latents.shape[1]
    x = _ln(features, P, pre + "norm_media.")
    lat = _ln(latents, P, pre + "norm_latents.")
    q = F.linear(lat, P[pre + "to_q.weight"]).view(B, nq, heads, dim_head).transpose(1, 2)
    kv_in = torch.cat((x, lat), dim=-2)
    k = F.linear(kv_in, P[pre + "to_k.weight"]).view(B, nf + nq, heads, dim_head).transpose(1, 2)
    v = F.linear(kv_in, P[pre + "to_v.weight"]).view(B, nf + nq, heads, dim_head).transpose(1, 2)
    sim = (q * dim_head ** -0.5) @ k.transpose(-1, -2)
    sim = sim - sim.amax(dim=-1, keepdim=True)
    out = sim.softmax(dim=-1) @ v
    out = out.transpose(1, 2).reshape(B, nq, heads * dim_head)
    return F.linear(out, P[pre + "to_out.weight"])


def perceiver_forward(P, x_f, depth, heads=8, dim_head=64, mask=None, prefix=""):
    """PerceiverResampler.forward (perceiver_resampler.py:126-163): x_f (B, T, n, d) -> (B, num_latents, d)."""
    B, T, _, d = x_f.shape
    tpe = P[prefix + "time_pos_emb"][:T].unsqueeze(0).expand(B, -1, -1, -1)
    if mask is not None:
        tpe = tpe * mask.unsqueeze(-1).unsqueeze(-1)
    feats = (x_f + tpe).reshape(B, -1, d)
    x = P[prefix + "latents"].unsqueeze(0).expand(B, -1, -1)
    for l in range(depth):
        pre = f"{prefix}layers.{l}."
        x = x + attention_layer(P, pre + "0.", feats, x, heads, dim_head)
        h = _ln(x, P, pre + "1.0.")                                            # feed_forward_layer: LN, Linear, GELU, Linear
        x = x + F.linear(F.gelu(F.linear(h, P[pre + "1.1.weight"])), P[pre + "1.3.weight"])
    return _ln(x, P, prefix + "norm.")
