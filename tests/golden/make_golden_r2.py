"""Round-2 pins, generated from the UNMODIFIED reference (authoring container only):  python tests/golden/make_golden_r2.py

  r2_pins.npz
    density_*            the training sigma densities (mdt/models/edm_diffusion/utils.py:154-203) under torch.manual_seed(7)
    attn_self / attn_cross   the reference Attention module (transformer_blocks.py:66-158), causal, 10x10 and 10x4 (top-left mask)
    dec1_raw             MDTVTransformer.forward_dec_only (1 decoder layer) on a random context: one ConditionedBlock end to end
    smp_*                reference samplers with torch-drawn noise (churned Euler / Heun, Euler ancestral) under torch.manual_seed(123)
"""
from __future__ import annotations

import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from mdt_policy_b200.synthetic import synthetic_state_dict, synthetic_inputs, synthetic_tensor  # noqa: E402
from tests.golden.make_golden import build, save  # noqa: E402


@torch.no_grad()
def main():
    _, gcs = ref_shim.load_reference()
    import importlib
    rutils = importlib.import_module("mdt.models.edm_diffusion.utils")
    blocks = importlib.import_module("mdt.models.networks.transformers.transformer_blocks")
    out = {}
    # ---- sigma densities
    torch.manual_seed(7)
    out["density_loglogistic"] = rutils.rand_log_logistic((512,), loc=math.log(0.5), scale=0.5, min_value=0.001, max_value=80.0)
    out["density_lognormal"] = rutils.rand_log_normal((512,), loc=-1.2, scale=1.2)
    out["density_loguniform"] = rutils.rand_log_uniform((512,), 0.001, 80.0)
    out["density_uniform"] = rutils.rand_uniform((512,), 0.001, 80.0)
    # ---- Attention module, causal: self (10x10) and cross (10x4, top-left aligned mask)
    att = blocks.Attention(384, 8, 0.0, 0.0, block_size=20, causal=True, bias=False).eval()
    att.load_state_dict(synthetic_state_dict([(n, p.shape) for n, p in att.named_parameters()], 31, "trained"))
    x = synthetic_tensor("attn.x", (2, 10, 384), 31, "init") * 50.0      # N(0, 1)
    ctx = synthetic_tensor("attn.ctx", (2, 4, 384), 31, "init") * 50.0
    out["attn_self"] = att(x)
    out["attn_cross"] = att(x, context=ctx)
    # ---- one ConditionedBlock end to end: forward_dec_only with a single decoder layer on a random context
    model = build(ref_shim.mdtv_inner_cfg(n_enc_layers=1, n_dec_layers=1), seed=32, profile="trained")
    inp = synthetic_inputs(5, seed=42)
    context = synthetic_tensor("dec1.ctx", (5, 4, 384), 32, "init") * 50.0
    sigma = torch.tensor([80.0, 7.0, 0.5, 0.05, 0.001])
    out["dec1_raw"] = model.inner_model.forward_dec_only(context, inp["noise"] * sigma[:, None, None], sigma)
    # ---- samplers that draw noise from torch's generator
    model = build(ref_shim.mdtv_inner_cfg(n_enc_layers=1, n_dec_layers=1), seed=33, profile="trained")
    inp = synthetic_inputs(3, seed=43)
    state = {"state_images": inp["state_images"], "modality": "lang"}
    sig = gcs.get_sigmas_exponential(6, 0.01, 80.0)
    torch.manual_seed(123)
    out["smp_euler_churn"] = gcs.sample_euler(model, state, inp["x_T"], inp["goal"], sig, disable=True, s_churn=2.0)
    torch.manual_seed(123)
    out["smp_heun_churn"] = gcs.sample_heun(model, state, inp["x_T"], inp["goal"], sig, disable=True, s_churn=2.0)
    torch.manual_seed(123)
    out["smp_euler_ancestral"] = gcs.sample_euler_ancestral(model, state, inp["x_T"], inp["goal"], sig, disable=True)
    torch.manual_seed(123)
    out["smp_euler_nochurn_rngstate"] = gcs.sample_euler(model, state, inp["x_T"], inp["goal"], sig, disable=True)
    out["rng_after_euler"] = torch.rand(4)      # the reference draws eps every step even without churn: pins the RNG stream
    save("r2_pins", meta=dict(case="r2_pins", attn_seed=31, dec1_seed=32, dec1_input_seed=42, smp_seed=33, smp_input_seed=43), **out)


if __name__ == "__main__":
    main()
