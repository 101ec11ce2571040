"""TEST INFRASTRUCTURE ONLY -- never imported by the product package.

Imports the *unmodified* reference (``/root/reference``, read-only) in the authoring
container so that golden vectors can be generated from it and the oracle restatement
(``oracle/mdt_oracle.py``) can be pinned against it.

The reference's hot-path modules need five packages at *import* time that are not
installed here (hydra, omegaconf, matplotlib, torchsde, torchdiffeq -- see
``mdt/models/edm_diffusion/score_wrappers.py:3``, ``gc_sampling.py:7-10``,
``networks/transformers/transformer_blocks.py:1,17``).  None of them is used by the
arithmetic of the path, so they are replaced by empty stub modules plus a 6-line
``hydra.utils.instantiate`` that resolves ``_target_``.

``/root/reference`` does not exist on the GPU box: nothing under ``tests -m gpu``,
``bench.py`` or ``__graft_entry__.smoke()`` may call :func:`load_reference`.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MDT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mdt", "models", "edm_diffusion"))


def _stub(name: str, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def _instantiate(cfg, *args, **kwargs):
    cfg = dict(cfg)
    target = cfg.pop("_target_")
    cfg.pop("_recursive_", None)
    module_name, cls_name = target.rsplit(".", 1)
    cls = getattr(importlib.import_module(module_name), cls_name)
    return cls(*args, **cfg, **kwargs)


def load_reference():
    """Returns (GCDenoiser class, gc_sampling module) of the reference."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    if "omegaconf" not in sys.modules:
        _stub("omegaconf", DictConfig=dict, OmegaConf=object)
    if "matplotlib" not in sys.modules:
        mp = _stub("matplotlib")
        mp.pyplot = _stub("matplotlib.pyplot", cla=lambda *a, **k: None)
    if "torchsde" not in sys.modules:
        _stub("torchsde", BrownianTree=object)
    if "torchdiffeq" not in sys.modules:
        _stub("torchdiffeq", odeint=None)
    if "hydra" not in sys.modules:
        h = _stub("hydra")
        h.utils = _stub("hydra.utils", instantiate=_instantiate)
    from mdt.models.edm_diffusion.score_wrappers import GCDenoiser  # type: ignore
    from mdt.models.edm_diffusion import gc_sampling  # type: ignore
    return GCDenoiser, gc_sampling


def mdtv_inner_cfg(n_enc_layers=4, n_dec_layers=4, embed_dim=384, n_heads=8, **over):
    """conf/model/model/mdtv_transformer.yaml resolved with conf/config_d.yaml constants."""
    cfg = dict(
        _target_="mdt.models.networks.mdtv_transformer.MDTVTransformer",
        action_dim=7, obs_dim=384, goal_dim=512, proprio_dim=8, goal_conditioned=True,
        embed_dim=embed_dim, n_dec_layers=n_dec_layers, n_enc_layers=n_enc_layers, n_obs_token=3,
        goal_seq_len=1, obs_seq_len=1, action_seq_len=10, embed_pdrob=0, goal_drop=0,
        attn_pdrop=0.3, resid_pdrop=0.1, mlp_pdrop=0.05, n_heads=n_heads, device="cpu",
        linear_output=True, use_rot_embed=False, use_abs_pos_emb=True, bias=False,
        use_ada_conditioning=True, use_noise_encoder=False, use_modality_encoder=True,
        use_mlp_goal=True,
    )
    cfg.update(over)
    return cfg


def build_reference_denoiser(inner_cfg=None, sigma_data=0.5):
    import contextlib
    import io
    GCDenoiser, _ = load_reference()
    with contextlib.redirect_stdout(io.StringIO()):  # the ctor prints dims
        model = GCDenoiser(inner_cfg or mdtv_inner_cfg(), sigma_data=sigma_data)
    return model.eval()
