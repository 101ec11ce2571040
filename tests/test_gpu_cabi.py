"""The C ABI called directly with ctypes, the way a non-Python host would bind it: error codes and
messages, call-order rules, and the host-buffer sampling entry point."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import mdt_oracle as orc
from tests import helpers as H
from mdt_policy_b200 import _lib
from mdt_policy_b200.synthetic import synthetic_inputs, synthetic_state_dict

pytestmark = pytest.mark.gpu


def _cfg(**over):
    base = dict(abi_version=1, variant=0, embed_dim=384, n_heads=8, n_enc_layers=1, n_dec_layers=1, action_dim=7, action_seq_len=10,
                goal_dim=512, obs_dim=384, n_state_tokens=3, precision=0, max_batch=8, sigma_data=0.5)
    base.update(over)
    return _lib.MdtConfig(**base)


def _ptr(t):
    return C.c_void_p(t.data_ptr())


def test_call_order_and_error_codes():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.mdtb200_create(C.byref(_cfg()), C.byref(h)) == 0
    x = torch.zeros(2, 10, 7, device="cuda"); sig = torch.ones(2, device="cuda"); out = torch.empty_like(x)
    # not committed
    assert lib.mdtb200_denoise(h, _ptr(x), _ptr(sig), 2, 1, _ptr(out), None) == -2
    assert b"not committed" in lib.mdtb200_last_error(h)
    # commit with missing weights
    assert lib.mdtb200_commit_weights(h, None) == -2
    assert b"was not bound" in lib.mdtb200_last_error(h)
    shapes = H.mdtv_shapes(1, 1)
    sd = {k: v.cuda() for k, v in synthetic_state_dict(shapes, 71, "trained").items()}
    for k, v in sd.items():
        assert lib.mdtb200_bind_weight(h, k.encode(), _ptr(v), v.numel()) == 0
    # wrong size for one tensor
    bad = torch.zeros(5, device="cuda")
    lib.mdtb200_bind_weight(h, b"inner_model.action_pred.bias", _ptr(bad), 5)
    assert lib.mdtb200_commit_weights(h, None) == -1
    assert b"expected 7 elements" in lib.mdtb200_last_error(h)
    for k, v in sd.items():
        lib.mdtb200_bind_weight(h, k.encode(), _ptr(v), v.numel())
    assert lib.mdtb200_commit_weights(h, None) == 0
    # denoise before any context
    assert lib.mdtb200_denoise(h, _ptr(x), _ptr(sig), 2, 1, _ptr(out), None) == -2
    assert b"encode first" in lib.mdtb200_last_error(h)
    # batch out of range / null pointers / bad sampler
    assert lib.mdtb200_encode(h, _ptr(x), _ptr(x), 1, 9, None, None) == -1
    assert lib.mdtb200_encode(h, None, _ptr(x), 1, 2, None, None) == -1
    assert lib.mdtb200_sample(h, 7, _ptr(sig), 1, _ptr(x), _ptr(x), 1, 2, _ptr(x), None) == -1
    assert lib.mdtb200_sample(h, 0, _ptr(sig), 0, _ptr(x), _ptr(x), 1, 2, _ptr(x), None) == -1
    lib.mdtb200_destroy(h)
    lib.mdtb200_destroy(None)   # no-op


def test_sample_host_buffers_vs_oracle():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.mdtb200_create(C.byref(_cfg(n_enc_layers=2, n_dec_layers=2)), C.byref(h)) == 0
    shapes = H.mdtv_shapes(2, 2)
    sd = synthetic_state_dict(shapes, 72, "trained")
    dev = {k: v.cuda() for k, v in sd.items()}
    for k, v in dev.items():
        assert lib.mdtb200_bind_weight(h, k.encode(), _ptr(v), v.numel()) == 0
    assert lib.mdtb200_commit_weights(h, None) == 0
    B = 5
    inp = synthetic_inputs(B, seed=73)
    sig = orc.get_sigmas_exponential(6, 0.01, 80.0)
    goal = np.ascontiguousarray(inp["goal"][:, 0].numpy()); state = np.ascontiguousarray(inp["state_images"].numpy())
    for sampler, name in ((0, "ddim"), (1, "euler"), (2, "heun"), (3, "dpmpp_2m")):
        x = np.ascontiguousarray(inp["x_T"].numpy().copy())
        s = np.ascontiguousarray(sig.numpy())
        rc = lib.mdtb200_sample_host(h, sampler, s.ctypes.data_as(C.c_void_p), 6, goal.ctypes.data_as(C.c_void_p),
                                     state.ctypes.data_as(C.c_void_p), 1, B, x.ctypes.data_as(C.c_void_p), None)
        assert rc == 0, lib.mdtb200_last_error(h)
        want = orc.sample(sd, orc.OracleCfg(n_enc_layers=2, n_dec_layers=2), {"state_images": inp["state_images"], "modality": "lang"},
                          inp["x_T"], inp["goal"], sig, name)
        assert np.abs(x - want.numpy()).max() < 2e-5, name
    assert lib.mdtb200_launch_count(h) > 0
    lib.mdtb200_destroy(h)
