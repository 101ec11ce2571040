"""Generates tests/golden/*.npz from the UNMODIFIED reference (run in the authoring container only):

    python tests/golden/make_golden.py

The reference (/root/reference, read-only) is imported through oracle/ref_shim.py; parameters and inputs
come from mdt_policy_b200.synthetic (numpy PCG64 keyed by (seed, tensor name)), so the fixtures only need
to store the reference's OUTPUTS.  Every case is reproducible from (config, seed, profile) recorded in the
npz.  The reference has no tests or golden vectors of its own for this path.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from mdt_policy_b200.synthetic import synthetic_state_dict, synthetic_inputs  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def mdt_inner_cfg(**over):
    """conf/model/model/mdt_transformer.yaml resolved."""
    cfg = dict(
        _target_="mdt.models.networks.mdt_transformer.MDTTransformer", action_dim=7, obs_dim=512, goal_dim=512,
        proprio_dim=8, goal_conditioned=True, embed_dim=512, n_dec_layers=6, n_enc_layers=4, goal_seq_len=1,
        obs_seq_len=1, action_seq_len=10, embed_pdrob=0, goal_drop=0, attn_pdrop=0.3, resid_pdrop=0.1,
        mlp_pdrop=0.05, n_heads=8, device="cpu", linear_output=True, use_rot_embed=False, use_abs_pos_emb=True,
        bias=False, use_ada_conditioning=True, use_noise_encoder=False, use_modality_encoder=True, use_mlp_goal=True)
    cfg.update(over)
    return cfg


def build(inner_cfg, seed, profile, dtype=torch.float32):
    model = ref_shim.build_reference_denoiser(inner_cfg, sigma_data=0.5)
    sd = synthetic_state_dict([(n, p.shape) for n, p in model.named_parameters()], seed, profile)
    model.load_state_dict(sd, strict=True)
    return model.to(dtype).eval()


def save(name, **arrays):
    meta = arrays.pop("meta")
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=json.dumps(meta),
                        **{k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in arrays.items()})
    print("wrote", name, {k: tuple(np.asarray(v).shape) for k, v in arrays.items()})


@torch.no_grad()
def main():
    torch.set_num_threads(os.cpu_count())
    _, gcs = ref_shim.load_reference()

    # ---- parameter manifests (names, shapes, order): the boundary contract for checkpoints / EMA zip
    manifest = {}
    for key, cfg in (("mdtv_4_4", ref_shim.mdtv_inner_cfg()), ("mdt_4_6", mdt_inner_cfg())):
        m = ref_shim.build_reference_denoiser(cfg)
        manifest[key] = [[n, list(p.shape)] for n, p in m.named_parameters()]
    with open(os.path.join(OUT, "param_manifest.json"), "w") as f:
        json.dump(manifest, f)
    print("wrote param_manifest.json")

    # ---- BASELINE config 1: single denoise forward, B=4, 2 enc + 2 dec, sigma in {80, 0.5, 0.001, mixed}
    for profile in ("trained", "init"):
        cfg = ref_shim.mdtv_inner_cfg(n_enc_layers=2, n_dec_layers=2)
        model = build(cfg, seed=11, profile=profile)
        inp = synthetic_inputs(4, seed=21)
        outs = {}
        for modality in ("lang", "vis"):
            state = {"state_images": inp["state_images"], "modality": modality}
            for tag, sig in (("80", torch.full((4,), 80.0)), ("0p5", torch.full((4,), 0.5)), ("0p001", torch.full((4,), 0.001)),
                             ("mixed", torch.tensor([80.0, 3.1, 0.5, 0.001]))):
                x = inp["noise"] * sig[:, None, None]
                outs[f"fwd_{modality}_{tag}"] = model(state, x, inp["goal"], sig)
                if modality == "lang" and tag == "mixed":
                    outs["raw_lang_mixed"] = model.inner_model(state, x, inp["goal"], sig)
                    loss, mo = model.loss(state, inp["actions"], inp["goal"], inp["noise"], sig)
                    outs["loss_lang_mixed"] = loss
                    outs["loss_out_lang_mixed"] = mo
            outs[f"ctx_{modality}"] = model.forward_context_only(state, inp["noise"], inp["goal"], torch.ones(4))
        save(f"config1_{profile}", meta=dict(case="config1", enc=2, dec=2, B=4, weight_seed=11, input_seed=21, profile=profile), **outs)

    # ---- BASELINE config 2: 10-step sampling, B=256, MDT-V 4+4 (fp32 reference and fp64 "truth")
    for profile in ("trained", "init"):
        cfg = ref_shim.mdtv_inner_cfg()
        inp = synthetic_inputs(256, seed=22)
        outs = {}
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            model = build(cfg, seed=12, profile=profile, dtype=dtype)
            state = {"state_images": inp["state_images"].to(dtype), "modality": "lang"}
            goal, xT = inp["goal"].to(dtype), inp["x_T"].to(dtype)
            for smin in (0.001, 1.0):
                sig = gcs.get_sigmas_exponential(10, smin, 80.0).to(dtype)
                outs[f"ddim_smin{smin}_{tag}"] = gcs.sample_ddim(model, state, xT, goal, sig, disable=True)
            if dtype == torch.float32:
                sig = gcs.get_sigmas_exponential(10, 0.001, 80.0)
                n = 32   # the other fused samplers on the first 32 samples
                st32 = {"state_images": state["state_images"][:n], "modality": "lang"}
                outs["euler_b32"] = gcs.sample_euler(model, st32, xT[:n], goal[:n], sig, disable=True)
                outs["heun_b32"] = gcs.sample_heun(model, st32, xT[:n], goal[:n], sig, disable=True)
                outs["dpmpp_2m_b32"] = gcs.sample_dpmpp_2m(model, st32, xT[:n], goal[:n], sig, disable=True)
                sigk = gcs.get_sigmas_karras(5, 0.001, 80.0)
                outs["ddim_karras5_b32"] = gcs.sample_ddim(model, st32, xT[:n], goal[:n], sigk, disable=True)
                outs["ctx_b32"] = model.forward_context_only(st32, xT[:n], goal[:n], torch.ones(n))
        save(f"config2_{profile}", meta=dict(case="config2", enc=4, dec=4, B=256, weight_seed=12, input_seed=22, profile=profile,
                                             sigma_max=80.0, steps=10), **outs)

    # ---- BASELINE-literal "6 layers" variant (6 enc + 6 dec), B=32
    cfg = ref_shim.mdtv_inner_cfg(n_enc_layers=6, n_dec_layers=6)
    model = build(cfg, seed=13, profile="trained")
    inp = synthetic_inputs(32, seed=23)
    state = {"state_images": inp["state_images"], "modality": "vis"}
    sig = gcs.get_sigmas_exponential(10, 0.001, 80.0)
    save("config2_6x6_trained", meta=dict(case="config2_6x6", enc=6, dec=6, B=32, weight_seed=13, input_seed=23, profile="trained"),
         ddim=gcs.sample_ddim(model, state, inp["x_T"], inp["goal"], sig, disable=True))

    # ---- MDT (ResNet) variant, d=512, 4 enc + 6 dec, B=8
    model = build(mdt_inner_cfg(), seed=14, profile="trained")
    inp = synthetic_inputs(8, seed=24, n_state_tokens=2, obs_dim=512)
    state = {"static": inp["state_images"][:, :1], "gripper": inp["state_images"][:, 1:], "modality": "lang"}
    sigv = torch.tensor([80.0, 20.0, 5.0, 1.0, 0.5, 0.1, 0.01, 0.001])
    x = inp["noise"] * sigv[:, None, None]
    sig = gcs.get_sigmas_exponential(10, 0.001, 80.0)
    save("mdt_trained", meta=dict(case="mdt", enc=4, dec=6, B=8, weight_seed=14, input_seed=24, profile="trained"),
         fwd=model(state, x, inp["goal"], sigv),
         ctx=model.forward_context_only(state, x, inp["goal"], sigv),
         ddim=gcs.sample_ddim(model, state, inp["x_T"], inp["goal"], sig, disable=True))

    # ---- training: gradient fingerprints of the reference's GCDenoiser.loss (dropout 0, 2+2 layers, B=16)
    cfg = ref_shim.mdtv_inner_cfg(n_enc_layers=2, n_dec_layers=2, attn_pdrop=0, resid_pdrop=0, mlp_pdrop=0, embed_pdrob=0, goal_drop=0)
    model = build(cfg, seed=15, profile="trained").train()
    inp = synthetic_inputs(16, seed=25)
    sigma = torch.exp(torch.linspace(3.0, -4.0, 16))
    with torch.enable_grad():
        loss, _ = model.loss({"state_images": inp["state_images"], "modality": "lang"}, inp["actions"], inp["goal"], inp["noise"], sigma)
        loss.backward()
    norms, probes = [], []
    for n, p in model.named_parameters():
        g = p.grad.flatten() if p.grad is not None else torch.zeros(p.numel())
        norms.append(float(g.norm()))
        probes.append(g[torch.linspace(0, g.numel() - 1, 8).long()].clone())
    save("train_grads", meta=dict(case="train_grads", enc=2, dec=2, B=16, weight_seed=15, input_seed=25, profile="trained"),
         loss=loss.detach(), norms=torch.tensor(norms), probes=torch.stack(probes))

    # ---- schedules
    sch = {}
    for n in (1, 3, 5, 10, 20):
        sch[f"exponential_{n}"] = gcs.get_sigmas_exponential(n, 0.001, 80.0)
        sch[f"karras_{n}"] = gcs.get_sigmas_karras(n, 0.001, 80.0)
        sch[f"linear_{n}"] = gcs.get_sigmas_linear(n, 0.001, 80.0)
        if n > 1:
            sch[f"ve_{n}"] = gcs.get_sigmas_ve(n, 0.001, 80.0)
        sch[f"vp_{n}"] = gcs.get_sigmas_vp(n)
        sch[f"cosine_beta_{n}"] = gcs.cosine_beta_schedule(n)
    sch["iddpm_10"] = gcs.get_iddpm_sigmas(10, 0.001, 80.0)
    save("schedules", meta=dict(case="schedules", sigma_min=0.001, sigma_max=80.0), **sch)


if __name__ == "__main__":
    main()
