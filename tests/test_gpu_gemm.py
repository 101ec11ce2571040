"""The tcgen05/TMEM/TMA GEMM kernel on its own (through the test-only C-ABI entry mdtb200_debug_gemm) against an
fp64 torch matmul: every epilogue, both tile widths, ragged M, K = 384 / 1536 / 512 / 2048."""
import ctypes as C

import pytest
import torch

from mdt_policy_b200 import _lib

pytestmark = pytest.mark.gpu


def _handle(precision):
    lib = _lib.load()
    h = C.c_void_p()
    cfg = _lib.MdtConfig(abi_version=1, variant=0, embed_dim=384, n_heads=8, n_enc_layers=1, n_dec_layers=1, action_dim=7,
                         action_seq_len=10, goal_dim=512, obs_dim=384, n_state_tokens=3, precision=_lib.PRECISION[precision],
                         max_batch=4, sigma_data=0.5)
    rc = lib.mdtb200_create(C.byref(cfg), C.byref(h))
    assert rc == 0, lib.mdtb200_last_error(None)
    return lib, h


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


SHAPES = [  # (M, N, K, epi, rows_per_group)
    (2560, 1536, 384, 1, 1),     # MLP c_fc + GELU           (BN=128)
    (2560, 384, 1536, 5, 10),    # MLP c_proj + gate + res   (BN=64)
    (2560, 1152, 384, 0, 1),     # fused QKV
    (2560, 384, 384, 4, 1),      # attention c_proj + res
    (1024, 3072, 384, 0, 1),     # cross-attention K|V of 4 layers
    (37, 384, 384, 0, 1),        # ragged M, single partial tile
    (130, 1536, 384, 1, 1),      # M just past one tile
    (300, 512, 2048, 5, 10),     # MDT (d=512) c_proj
    (1280, 2048, 512, 1, 1),     # MDT c_fc
]


@pytest.mark.parametrize("precision,tol", [("bf16x3", 2e-5), ("bf16", 2e-2)])
@pytest.mark.parametrize("M,N,K,epi,rpg", SHAPES)
def test_tc_gemm_vs_fp64(precision, tol, M, N, K, epi, rpg):
    lib, h = _handle(precision)
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + epi)
    A = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K ** 0.5
    bias = torch.randn(N, generator=g) * 0.1
    R = torch.randn(M, N, generator=g)
    gate = torch.randn((M + rpg - 1) // rpg, N, generator=g)
    want = A.double() @ W.double().T + bias.double()
    if epi == 1:
        want = torch.nn.functional.gelu(want)
    elif epi == 4:
        want = R.double() + want
    elif epi == 5:
        want = R.double() + gate.double().repeat_interleave(rpg, 0)[:M] * want
    dA, dW, db, dR, dg = (t.cuda() for t in (A, W, bias, R, gate))
    out = torch.full((M, N), float("nan"), device="cuda")
    rc = lib.mdtb200_debug_gemm(h, _p(dA), _p(dW), _p(db), _p(dR) if epi >= 4 else None, _p(dg) if epi == 5 else None,
                                M, N, K, epi, rpg, _p(out), None)
    assert rc == 0, lib.mdtb200_last_error(h)
    err = (out.cpu().double() - want).abs().max().item()
    scale = want.abs().max().item()
    assert err < tol * scale, (err, scale)
    lib.mdtb200_destroy(h)
