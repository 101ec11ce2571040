"""Round-2 parity pins against outputs of the reference itself (tests/golden/r2_pins.npz, made by make_golden_r2.py):
the training sigma densities (a17), the reference samplers that draw noise from torch's generator (RNG-stream parity of
the generic sampler driver), the Attention module with the non-square causal mask and one ConditionedBlock end to end."""
import math
import os

import pytest
import torch

from oracle import mdt_oracle as orc
from tests import helpers as H
from mdt_policy_b200 import gc_sampling as gcs, utils
from mdt_policy_b200.agent import DenoiseAgent
from mdt_policy_b200.synthetic import synthetic_inputs, synthetic_state_dict, synthetic_tensor


def test_sigma_densities_bit_equal_reference():
    _, gold = H.load_golden("r2_pins")
    torch.manual_seed(7)
    assert torch.equal(utils.rand_log_logistic((512,), loc=math.log(0.5), scale=0.5, min_value=0.001, max_value=80.0), gold["density_loglogistic"])
    assert torch.equal(utils.rand_log_normal((512,), loc=-1.2, scale=1.2), gold["density_lognormal"])
    assert torch.equal(utils.rand_log_uniform((512,), 0.001, 80.0), gold["density_loguniform"])
    assert torch.equal(utils.rand_uniform((512,), 0.001, 80.0), gold["density_uniform"])
    # the agent's factory (mdtv_agent.py:552-591) hands out the same density with the shipped knobs
    agent = DenoiseAgent(model=None, device="cpu")
    torch.manual_seed(7)
    assert torch.equal(agent.make_sample_density()(shape=(512,), device="cpu"), gold["density_loglogistic"])


def test_generic_sampler_driver_matches_reference_rng_stream():
    """churned Euler / Heun and Euler-ancestral draw their noise from torch's global generator: same seed -> same actions
    as the reference's own loops, and the generator is left in the same state (the reference draws eps every step)."""
    meta, gold = H.load_golden("r2_pins")
    P = H.oracle_params(H.mdtv_shapes(1, 1), meta["smp_seed"], "trained")
    cfg = orc.OracleCfg(n_enc_layers=1, n_dec_layers=1)
    inp = synthetic_inputs(3, seed=meta["smp_input_seed"])
    state = {"state_images": inp["state_images"], "modality": "lang"}

    def model(state, action, goal, sigma):
        return orc.denoiser_forward(P, cfg, state, action, goal, sigma)

    sig = gcs.get_sigmas_exponential(6, 0.01, 80.0)
    with torch.no_grad():
        for key, fn, kw in (("smp_euler_churn", gcs.sample_euler, dict(s_churn=2.0)), ("smp_heun_churn", gcs.sample_heun, dict(s_churn=2.0)),
                            ("smp_euler_ancestral", gcs.sample_euler_ancestral, {})):
            torch.manual_seed(123)
            got = fn(model, state, inp["x_T"], inp["goal"], sig, **kw)
            scale = max(1.0, float(gold[key].abs().max()))
            assert (got - gold[key]).abs().max() < 2e-5 * scale, key
        torch.manual_seed(123)
        got = gcs.sample_euler(model, state, inp["x_T"], inp["goal"], sig)
        assert (got - gold["smp_euler_nochurn_rngstate"]).abs().max() < 2e-5
        assert torch.equal(torch.rand(4), gold["rng_after_euler"])


def test_deepcopy_and_pickle_of_a_used_model_drop_the_engine():
    import copy, pickle
    m = H.build_product(H.mdtv_inner_cfg(1, 1), 3, "init", device="cpu")
    m.inner_model.__dict__["_engines"] = {"cuda:0": object()}        # what a CUDA call leaves behind (ctypes handles)
    c = copy.deepcopy(m)
    assert "_engines" not in c.inner_model.__dict__ and c.inner_model.__dict__["_weights_dirty"]
    assert len(pickle.dumps(m)) > 1000
    assert [n for n, _ in c.named_parameters()] == [n for n, _ in m.named_parameters()]


# ------------------------------------------------------------------------------------------------------------------ GPU

@pytest.mark.gpu
def test_attention_module_vs_reference_nonsquare_causal_mask():
    """training-path attention block (projection GEMMs + attention kernel) against the reference Attention module: causal
    self-attention 10x10 and cross-attention 10x4 with the top-left aligned mask."""
    from mdt_policy_b200 import training
    from mdt_policy_b200.networks import _Attention
    meta, gold = H.load_golden("r2_pins")
    att = _Attention(384, False, 0.0, 0.0)
    att.load_state_dict(synthetic_state_dict([(n, p.shape) for n, p in att.named_parameters()], meta["attn_seed"], "trained"))
    att = att.cuda()
    x = (synthetic_tensor("attn.x", (2, 10, 384), meta["attn_seed"], "init") * 50.0).cuda()
    ctx = (synthetic_tensor("attn.ctx", (2, 4, 384), meta["attn_seed"], "init") * 50.0).cuda()
    with torch.no_grad():
        got_self = training._attention_block(att, 8, x, x, True, False).cpu()
        got_cross = training._attention_block(att, 8, x, ctx, True, False).cpu()
    assert (got_self - gold["attn_self"]).abs().max() < 1e-4 * max(1.0, float(gold["attn_self"].abs().max()))
    assert (got_cross - gold["attn_cross"]).abs().max() < 1e-4 * max(1.0, float(gold["attn_cross"].abs().max()))


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["bf16x3", "fp32"])
@pytest.mark.parametrize("cross_fused", ["1", "0"])
def test_single_conditioned_block_vs_reference(precision, cross_fused, monkeypatch):
    """forward_dec_only with one decoder layer on a random context = sigma embedding + one ConditionedBlock (AdaLN, causal
    self-attention, 10x4 cross-attention, MLP) + head; with the algebraic cross-attention kernel and with the 5-kernel chain."""
    monkeypatch.setenv("MDTB200_CROSS_FUSED", cross_fused)
    meta, gold = H.load_golden("r2_pins")
    model = H.build_product(H.mdtv_inner_cfg(1, 1, precision=precision), meta["dec1_seed"], "trained")
    inp = synthetic_inputs(5, seed=meta["dec1_input_seed"])
    context = (synthetic_tensor("dec1.ctx", (5, 4, 384), meta["dec1_seed"], "init") * 50.0).cuda()
    sigma = torch.tensor([80.0, 7.0, 0.5, 0.05, 0.001]).cuda()
    with torch.no_grad():
        got = model.inner_model.forward_dec_only(context, (inp["noise"] * sigma.cpu()[:, None, None]).cuda(), sigma).cpu()
    ref = gold["dec1_raw"]
    tol = 2e-5 if precision == "fp32" else 1e-4
    assert (got - ref).abs().max() < tol * max(1.0, float(ref.abs().max())), float((got - ref).abs().max())


@pytest.mark.gpu
def test_invalidate_weights_after_inplace_data_write():
    """in-place writes through .data bump no version counter (PyTorch semantics): invalidate_weights() makes the engine re-pack;
    re-assigning .data is detected automatically."""
    model = H.build_product(H.mdtv_inner_cfg(1, 1), 3, "trained")
    inp = {k: v.cuda() for k, v in synthetic_inputs(4, seed=5).items()}
    state = {"state_images": inp["state_images"], "modality": "lang"}
    sig = torch.full((4,), 1.5).cuda()
    with torch.no_grad():
        a = model(state, inp["x_T"], inp["goal"], sig).clone()
        w = model.inner_model.action_pred.weight
        w.data.mul_(0.5)
        model.inner_model.invalidate_weights()
        b = model(state, inp["x_T"], inp["goal"], sig).clone()
        assert (a - b).abs().max() > 1e-3
        w.data = w.data * 2.0                       # new storage, same Parameter object and version
        c = model(state, inp["x_T"], inp["goal"], sig)
        assert (a - c).abs().max() < 1e-5


@pytest.mark.gpu
def test_persistent_fused_decoder_opt_in_matches_default_path(monkeypatch):
    """MDTB200_FUSED=1: the one-launch persistent decoder (fused_decoder.cuh) gives the same actions as the kernel graph."""
    inp = {k: v.cuda() for k, v in synthetic_inputs(37, seed=8).items()}
    state = {"state_images": inp["state_images"], "modality": "lang"}
    sig = gcs.get_sigmas_exponential(4, 0.001, 80.0, "cuda")
    outs = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("MDTB200_FUSED", flag)
        model = H.build_product(H.mdtv_inner_cfg(2, 2), 9, "trained")
        for name, fn in (("ddim", gcs.sample_ddim), ("heun", gcs.sample_heun), ("dpmpp_2m", gcs.sample_dpmpp_2m)):
            outs[(flag, name)] = fn(model, state, inp["x_T"], inp["goal"], sig, disable=True).cpu()
        outs[(flag, "launches")] = model.inner_model.launch_count()
    for name in ("ddim", "heun", "dpmpp_2m"):
        ref = outs[("0", name)]
        assert (ref - outs[("1", name)]).abs().max() < 5e-5 * max(1.0, float(ref.abs().max())), name   # Heun divides by small sigmas
    assert outs[("1", "launches")] < outs[("0", "launches")]
