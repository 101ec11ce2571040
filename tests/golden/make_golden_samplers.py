"""Goldens for the remaining samplers of the reference module, generated from the UNMODIFIED reference (authoring
container only):  python tests/golden/make_golden_samplers.py  ->  samplers.npz

Each entry is the output of the reference's own function on a 1+1-layer MDT-V with synthetic weights (seed 33), inputs of seed 43,
a 6-step exponential schedule 80 -> 0.01, under torch.manual_seed(123) (the stochastic ones draw from torch's global generator)."""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from mdt_policy_b200.synthetic import synthetic_inputs  # noqa: E402
from tests.golden.make_golden import build, save  # noqa: E402


@torch.no_grad()
def main():
    _, gcs = ref_shim.load_reference()
    model = build(ref_shim.mdtv_inner_cfg(n_enc_layers=1, n_dec_layers=1), seed=33, profile="trained")
    inp = synthetic_inputs(3, seed=43)
    state = {"state_images": inp["state_images"], "modality": "lang"}
    sig = gcs.get_sigmas_exponential(6, 0.01, 80.0)
    out = {}
    for key, fn, kw in (("dpm_2", gcs.sample_dpm_2, {}), ("dpm_2_churn", gcs.sample_dpm_2, dict(s_churn=2.0)),
                        ("dpm_2_ancestral", gcs.sample_dpm_2_ancestral, {}), ("dpm_2_ancestral_eta05", gcs.sample_dpm_2_ancestral, dict(eta=0.5)),
                        ("lms", gcs.sample_lms, {}), ("lms_order2", gcs.sample_lms, dict(order=2)),
                        ("dpmpp_2_with_lms", gcs.sample_dpmpp_2_with_lms, {}), ("dpmpp_2s", gcs.sample_dpmpp_2s, {}),
                        ("dpmpp_2s_ancestral", gcs.sample_dpmpp_2s_ancestral, {}),
                        ("dpmpp_2s_ancestral_eta0", gcs.sample_dpmpp_2s_ancestral, dict(eta=0.0, s_noise=0.7))):
        torch.manual_seed(123)
        out[key] = fn(model, state, inp["x_T"], inp["goal"], sig, disable=True, **kw)
        out["rng_after_" + key] = torch.rand(2)
        assert torch.isfinite(out[key]).all(), key
    # the three sigma densities of mdt/models/edm_diffusion/utils.py that r2_pins.npz does not cover
    import importlib
    rutils = importlib.import_module("mdt.models.edm_diffusion.utils")
    torch.manual_seed(9)
    out["density_v_diffusion"] = rutils.rand_v_diffusion((512,), sigma_data=0.5, min_value=0.001, max_value=80.0)
    out["density_split_log_normal"] = rutils.rand_split_log_normal((512,), loc=-1.2, scale_1=0.8, scale_2=1.6)
    out["density_discrete"] = rutils.rand_discrete((512,), gcs.get_sigmas_exponential(50, 0.001, 80.0))
    save("samplers", meta=dict(case="samplers", smp_seed=33, smp_input_seed=43), **out)


if __name__ == "__main__":
    main()
