"""The remaining samplers of the reference module (dpm_2, dpm_2_ancestral, lms, dpmpp_2_with_lms, dpmpp_2s,
dpmpp_2s_ancestral), restated on the generic driver of mdt_policy_b200.gc_sampling, against outputs of the reference's own functions
(tests/golden/make_golden_samplers.py).  CPU: the driver with the oracle as the model callable, incl. the RNG stream; GPU: the same
functions with the product model (mdtb200_denoise per evaluation)."""
import pytest
import torch

from oracle import mdt_oracle as orc
from tests import helpers as H
from mdt_policy_b200 import gc_sampling as gcs
from mdt_policy_b200.synthetic import synthetic_inputs

CASES = [("dpm_2", gcs.sample_dpm_2, {}), ("dpm_2_churn", gcs.sample_dpm_2, dict(s_churn=2.0)),
         ("dpm_2_ancestral", gcs.sample_dpm_2_ancestral, {}), ("dpm_2_ancestral_eta05", gcs.sample_dpm_2_ancestral, dict(eta=0.5)),
         ("lms", gcs.sample_lms, {}), ("lms_order2", gcs.sample_lms, dict(order=2)),
         ("dpmpp_2_with_lms", gcs.sample_dpmpp_2_with_lms, {}), ("dpmpp_2s", gcs.sample_dpmpp_2s, {}),
         ("dpmpp_2s_ancestral", gcs.sample_dpmpp_2s_ancestral, {}),
         ("dpmpp_2s_ancestral_eta0", gcs.sample_dpmpp_2s_ancestral, dict(eta=0.0, s_noise=0.7))]


@pytest.mark.parametrize("key,fn,kw", CASES, ids=[c[0] for c in CASES])
def test_restated_sampler_matches_reference_output_and_rng_stream(key, fn, kw):
    meta, gold = H.load_golden("samplers")
    P = H.oracle_params(H.mdtv_shapes(1, 1), meta["smp_seed"], "trained")
    cfg = orc.OracleCfg(n_enc_layers=1, n_dec_layers=1)
    inp = synthetic_inputs(3, seed=meta["smp_input_seed"])
    state = {"state_images": inp["state_images"], "modality": "lang"}

    def model(state, action, goal, sigma):
        return orc.denoiser_forward(P, cfg, state, action, goal, sigma)

    sig = gcs.get_sigmas_exponential(6, 0.01, 80.0)
    with torch.no_grad():
        torch.manual_seed(123)
        got = fn(model, state, inp["x_T"], inp["goal"], sig, **kw)
    scale = max(1.0, float(gold[key].abs().max()))
    assert (got - gold[key]).abs().max() < 2e-5 * scale
    assert torch.equal(torch.rand(2), gold["rng_after_" + key])          # consumed exactly the reference's draws


def test_remaining_sigma_densities_match_reference_streams():
    """rand_v_diffusion / rand_split_log_normal / rand_discrete (utils.py:176-198) under the reference's seed; 'v-diffusion' through
    DenoiseAgent.make_sample_density as mdtv_agent.py:577-580 builds it"""
    import math
    from mdt_policy_b200 import utils as U, DenoiseAgent
    _, gold = H.load_golden("samplers")
    torch.manual_seed(9)
    assert torch.equal(U.rand_v_diffusion((512,), sigma_data=0.5, min_value=0.001, max_value=80.0), gold["density_v_diffusion"])
    assert torch.equal(U.rand_split_log_normal((512,), loc=-1.2, scale_1=0.8, scale_2=1.6), gold["density_split_log_normal"])
    assert torch.equal(U.rand_discrete((512,), gcs.get_sigmas_exponential(50, 0.001, 80.0)), gold["density_discrete"])
    agent = DenoiseAgent(model=None, device="cpu", sigma_sample_density_type="v-diffusion")
    torch.manual_seed(9)
    assert torch.equal(agent.make_sample_density()(shape=(512,), device="cpu"), gold["density_v_diffusion"])


def test_agent_dispatch_covers_the_reference_sampler_names():
    """mdtv_agent.py:619-656: every sampler_type the reference dispatches that works there is accepted (three are not: dpm_adaptive /
    dpm_fast raise NameError in the reference itself, dpmpp_2m_sde needs torchsde)"""
    from mdt_policy_b200 import DenoiseAgent
    calls = []

    class Dummy:
        training = False

        def __call__(self, state, action, goal, sigma):
            calls.append(float(sigma[0]))
            return action * 0.5

    agent = DenoiseAgent(Dummy(), device="cpu")
    sig = gcs.get_sigmas_exponential(4, 0.01, 80.0)
    x = torch.randn(2, 10, 7)
    for name in ("lms", "heun", "euler", "ancestral", "euler_ancestral", "dpm", "dpmpp_2s_ancestral", "dpmpp_2m", "ddim", "dpmpp_2s",
                 "dpmpp_2_with_lms"):
        out = agent.sample_loop(sig, x, {"state_images": torch.zeros(2, 3, 384)}, torch.zeros(2, 1, 512), None, name)
        assert out.shape == x.shape and torch.isfinite(out).all(), name
    with pytest.raises(ValueError):
        agent.sample_loop(sig, x, {}, torch.zeros(2, 1, 512), None, "dpm_fast")


@pytest.mark.gpu
@pytest.mark.parametrize("key,fn,kw", [c for c in CASES if "ancestral" not in c[0] and "churn" not in c[0]],
                         ids=[c[0] for c in CASES if "ancestral" not in c[0] and "churn" not in c[0]])
def test_cuda_model_under_the_deterministic_samplers_vs_reference(key, fn, kw):
    meta, gold = H.load_golden("samplers")
    model = H.build_product(H.mdtv_inner_cfg(1, 1), meta["smp_seed"], "trained")
    inp = {k: v.cuda() for k, v in synthetic_inputs(3, seed=meta["smp_input_seed"]).items()}
    state = {"state_images": inp["state_images"], "modality": "lang"}
    sig = gcs.get_sigmas_exponential(6, 0.01, 80.0).cuda()
    got = fn(model, state, inp["x_T"], inp["goal"], sig, **kw).cpu()
    assert (got - gold[key]).abs().max() < 1e-4 * max(1.0, float(gold[key].abs().max()))
