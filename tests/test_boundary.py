"""CPU-side checks of the drop-in boundary: parameter names/order vs the reference manifest, the C-ABI library
loads and exports every symbol include/mdtb200.h declares, no CPU fallback, schedules bit-equal to the reference."""
import ctypes
import os
import re

import pytest
import torch

from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_param_names_shapes_order_match_reference_mdtv():
    from mdt_policy_b200 import GCDenoiser
    model = GCDenoiser(H.mdtv_inner_cfg(), sigma_data=0.5)
    mine = [(n, tuple(p.shape)) for n, p in model.named_parameters()]
    assert mine == H.manifest("mdtv_4_4")
    assert list(model.state_dict().keys()) == [n for n, _ in H.manifest("mdtv_4_4")]


def test_param_names_shapes_order_match_reference_mdt():
    from mdt_policy_b200 import GCDenoiser
    model = GCDenoiser(H.mdt_inner_cfg(), sigma_data=0.5)
    mine = [(n, tuple(p.shape)) for n, p in model.named_parameters()]
    assert mine == H.manifest("mdt_4_6")


def test_reference_init_statistics():
    """mdtv_transformer.py:197-206: normal(0, 0.02) weights, zero biases, unit LayerNorm."""
    from mdt_policy_b200 import GCDenoiser
    torch.manual_seed(0)
    m = GCDenoiser(H.mdtv_inner_cfg(), sigma_data=0.5).inner_model
    assert abs(float(m.decoder.blocks[0].mlp.c_fc.weight.detach().std()) - 0.02) < 1e-3
    assert float(m.decoder.blocks[0].attn.key.bias.abs().max()) == 0.0
    assert torch.equal(m.decoder.blocks[0].ln3.weight, torch.ones(384))
    assert abs(float(m.pos_emb.detach().std()) - 0.02) < 2e-3


def test_library_exports_every_declared_symbol():
    from mdt_policy_b200 import _lib
    lib = _lib.load()     # builds with nvcc if the .so is absent (cross-compiles without a GPU)
    header = open(os.path.join(ROOT, "include", "mdtb200.h")).read()
    declared = set(re.findall(r"\b(mdtb200_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mdtb200_abi_version() == _lib.ABI_VERSION
    # the enums the Python side mirrors
    for key, table in (("VARIANT", _lib.VARIANT), ("PREC", _lib.PRECISION), ("SAMPLER", _lib.SAMPLER)):
        for name, val in table.items():
            m = re.search(rf"MDTB200_{key}_{name.upper()}\s*=\s*(\d+)", header)
            assert m and int(m.group(1)) == val, (key, name)
    assert ctypes.sizeof(_lib.MdtConfig) == 14 * 4


def test_create_rejects_bad_config_without_gpu_work():
    from mdt_policy_b200 import _lib
    lib = _lib.load()
    cfg = _lib.MdtConfig(abi_version=99)
    h = ctypes.c_void_p()
    assert lib.mdtb200_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"abi_version" in lib.mdtb200_last_error(None)
    cfg = _lib.MdtConfig(abi_version=1, variant=0, embed_dim=100, n_heads=8, n_enc_layers=1, n_dec_layers=1, action_dim=7,
                         action_seq_len=10, goal_dim=512, obs_dim=384, n_state_tokens=3, precision=0, max_batch=4, sigma_data=0.5)
    assert lib.mdtb200_create(ctypes.byref(cfg), ctypes.byref(h)) == -5
    assert b"embed_dim" in lib.mdtb200_last_error(None)


def test_no_cpu_fallback():
    from mdt_policy_b200 import GCDenoiser
    model = GCDenoiser(H.mdtv_inner_cfg(n_enc=1, n_dec=1), sigma_data=0.5).eval()
    state = {"state_images": torch.zeros(2, 3, 384), "modality": "lang"}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(state, torch.zeros(2, 10, 7), torch.zeros(2, 1, 512), torch.ones(2))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.forward_context_only(state, torch.zeros(2, 10, 7), torch.zeros(2, 1, 512), torch.ones(2))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mdt_policy_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|oracle[./]|ref_shim|/root/reference", src, re.M), \
                    f"{f} reaches into oracle/ or the reference tree"


def test_unsupported_configs_fail_loudly():
    from mdt_policy_b200 import GCDenoiser
    with pytest.raises(NotImplementedError):
        GCDenoiser(H.mdtv_inner_cfg(use_ada_conditioning=False))
    with pytest.raises(NotImplementedError):
        GCDenoiser(H.mdtv_inner_cfg(use_rot_embed=True))
    model = GCDenoiser(H.mdtv_inner_cfg(n_enc=1, n_dec=1)).train()
    args = ({"state_images": torch.zeros(1, 3, 384)}, torch.zeros(1, 10, 7), torch.zeros(1, 1, 512), torch.ones(1))
    with pytest.raises(RuntimeError, match="no CPU fallback"):       # autograd on -> training path, CUDA only
        model(*args)
    with torch.no_grad(), pytest.raises(NotImplementedError, match="train-mode"):   # the inference kernels have no dropout
        model(*args)


def test_schedules_bit_equal_reference():
    from mdt_policy_b200 import gc_sampling as gcs
    _, gold = H.load_golden("schedules")
    for n in (1, 3, 5, 10, 20):
        assert torch.equal(gcs.get_sigmas_exponential(n, 0.001, 80.0), gold[f"exponential_{n}"])
        assert torch.equal(gcs.get_sigmas_karras(n, 0.001, 80.0), gold[f"karras_{n}"])
        assert torch.equal(gcs.get_sigmas_linear(n, 0.001, 80.0), gold[f"linear_{n}"])
        assert torch.equal(gcs.get_sigmas_vp(n), gold[f"vp_{n}"])
        assert torch.equal(gcs.cosine_beta_schedule(n), gold[f"cosine_beta_{n}"])
        if n > 1:
            assert torch.equal(gcs.get_sigmas_ve(n, 0.001, 80.0), gold[f"ve_{n}"])
    assert torch.equal(gcs.get_iddpm_sigmas(10, 0.001, 80.0), gold["iddpm_10"])


def test_generic_sampler_loops_match_oracle_on_cpu_callable():
    """The generic (non-fused) sampler loops are host logic: drive them with a CPU callable (the oracle's denoiser)
    and compare with the oracle's own samplers."""
    from mdt_policy_b200 import gc_sampling as gcs
    from oracle import mdt_oracle as orc
    from mdt_policy_b200.synthetic import synthetic_inputs
    P = H.oracle_params(H.mdtv_shapes(1, 1), 5, "trained")
    cfg = orc.OracleCfg(n_enc_layers=1, n_dec_layers=1)
    inp = synthetic_inputs(3, seed=6)
    state = {"state_images": inp["state_images"], "modality": "lang"}

    def model(state, action, goal, sigma):
        return orc.denoiser_forward(P, cfg, state, action, goal, sigma)

    sig = gcs.get_sigmas_exponential(6, 0.01, 80.0)
    with torch.no_grad():
        for name in ("ddim", "euler", "heun", "dpmpp_2m"):
            got = gcs.SAMPLERS[name](model, state, inp["x_T"], inp["goal"], sig)
            want = orc.sample(P, cfg, state, inp["x_T"], inp["goal"], sig, name)
            assert (got - want).abs().max() < 1e-5, name


def test_train_goal_preprocessing_follows_reference_rules():
    """preprocess_goals in train mode (mdt_transformer.py:293-305): 2-D goals get a token axis, one-goal-per-state-step inputs keep the
    first token, concatenated vision goals (2 * obs_dim) are truncated, goal_drop masks element-wise only while training"""
    from mdt_policy_b200 import GCDenoiser, training
    net = GCDenoiser(H.mdt_inner_cfg(n_enc_layers=1, n_dec_layers=1, goal_drop=0.5), sigma_data=0.5).inner_model
    states = {"static": torch.zeros(4, 1, 512), "gripper": torch.zeros(4, 1, 512)}
    net.eval()
    g = torch.randn(4, 512)
    assert training._prep_goal_train(net, states, g).shape == (4, 1, 512)
    assert torch.equal(training._prep_goal_train(net, states, g)[:, 0], g)                     # eval: no mask
    g2 = torch.randn(4, 1, 1024)
    assert torch.equal(training._prep_goal_train(net, states, g2), g2[:, :, :512])           # 2 * obs_dim -> obs_dim
    net.train()
    torch.manual_seed(0)
    m = training._prep_goal_train(net, states, torch.ones(64, 512))
    frac = float((m == 0).float().mean())
    assert 0.4 < frac < 0.6 and set(m.unique().tolist()) == {0.0, 1.0}
