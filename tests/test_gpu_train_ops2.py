"""Fused training primitives (mdt_policy_b200/train_ops.py -> libmdtb200.so) against plain PyTorch fp32/fp64 references of the same ops.
Tolerances: bf16x3 GEMMs 2e-5 relative to the output scale (operand split error 2^-17); everything else fp32 rounding."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from mdt_policy_b200 import train_ops
    return train_ops


def _join(t16):
    K = t16.shape[1] // 2
    return t16[:, :K].float() + t16[:, K:].float()


def _rel(got, ref):
    return float((got.double() - ref.double()).abs().max() / ref.double().abs().max().clamp_min(1e-30))


def test_split_is_hi_plus_lo_and_colsum():
    O = _ops()
    torch.manual_seed(0)
    x = torch.randn(333, 384, device="cuda") * 3
    x16, cs = O.split(x, want_colsum=True)
    assert _rel(_join(x16), x) < 2 ** -16
    assert _rel(cs, x.double().sum(0)) < 1e-6
    h = torch.randn_like(x)
    for act, fn in ((O.ACT_GELU, torch.nn.functional.gelu), (O.ACT_MISH, torch.nn.functional.mish), (O.ACT_SILU, torch.nn.functional.silu)):
        hh = h.double().requires_grad_(True)
        (g,) = torch.autograd.grad(fn(hh).sum(), hh)
        d16 = O.split(x, h=h, act=act)
        assert _rel(_join(d16), x.double() * g) < 2e-5


@pytest.mark.parametrize("M,N,K", [(5120, 384, 384), (1536, 1152, 384), (512, 2304, 384), (5120, 384, 1536), (200, 128, 64), (5120, 1536, 384)])
def test_gemm16_three_roles_match_fp64(M, N, K):
    O = _ops()
    torch.manual_seed(1)
    x, w, dy = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") / K ** 0.5, torch.randn(M, N, device="cuda")
    b = torch.randn(N, device="cuda")
    x16, w16, dy16 = O.split(x), O.split(w), O.split(dy)
    y = O.gemm16(0, x16, w16, M, N, K, bias=b)
    assert _rel(y, x.double() @ w.double().t() + b.double()) < 2e-5
    dx = O.gemm16(1, dy16, w16, M, N, K)
    assert _rel(dx, dy.double() @ w.double()) < 2e-5
    ref_dw = dy.double().t() @ x.double()
    for splits in (1, None, 4 if M >= 1024 else 2):
        dw = O.gemm16(2, dy16, x16, M, N, K, splits=splits)
        assert _rel(dw, ref_dw) < 2e-5, splits
    # determinism of the split-K reduction
    a, b2 = O.gemm16(2, dy16, x16, M, N, K, splits=3), O.gemm16(2, dy16, x16, M, N, K, splits=3)
    assert torch.equal(a, b2)


def test_gemm16_gelu16_epilogue():
    O = _ops()
    torch.manual_seed(2)
    M, N, K = 640, 1536, 384
    x, w, b = torch.randn(M, K, device="cuda"), torch.randn(N, K, device="cuda") / K ** 0.5, torch.randn(N, device="cuda")
    h, g16 = O.gemm16(0, O.split(x), O.split(w), M, N, K, bias=b, epi=O.EPI_GELU16)
    ref = x.double() @ w.double().t() + b.double()
    assert _rel(h, ref) < 2e-5
    assert _rel(_join(g16), torch.nn.functional.gelu(ref)) < 3e-5


def test_weight_bank_groups():
    O = _ops()
    torch.manual_seed(3)
    ws = [torch.randn(384, 384, device="cuda") for _ in range(3)]
    bs = [torch.randn(384, device="cuda") for _ in range(3)]
    w2 = [torch.randn(1536, 384, device="cuda")]
    bank = O.WeightBank([("qkv", ws, bs), ("fc", w2, [None])], torch.device("cuda"))
    bank.refresh()
    assert _rel(_join(bank.w16["qkv"]), torch.cat(ws)) < 2 ** -16
    assert torch.equal(bank.bias["qkv"], torch.cat(bs))
    assert _rel(_join(bank.w16["fc"]), w2[0]) < 2 ** -16 and bank.bias["fc"] is None
    ws[1].mul_(2.0)
    bank.refresh()
    assert _rel(_join(bank.w16["qkv"]), torch.cat(ws)) < 2 ** -16 and bank.valid()


@pytest.mark.parametrize("mod", [False, True])
def test_ln_fwd16_and_bwd2(mod):
    O = _ops()
    torch.manual_seed(4)
    B, T, d = 37, 10, 384
    x = torch.randn(B * T, d, device="cuda")
    w, b = torch.randn(d, device="cuda"), torch.randn(d, device="cuda")
    modt = torch.randn(B, 6 * d, device="cuda")
    sh, sc = (modt[:, :d], modt[:, d:2 * d]) if mod else (None, None)
    dy, dres = torch.randn(B * T, d, device="cuda"), torch.randn(B * T, d, device="cuda")

    def ref(xx, ww, bb, shh, scc):
        n = torch.nn.functional.layer_norm(xx, (d,), ww, bb, 1e-5)
        if shh is None:
            return n
        return shh.repeat_interleave(T, 0) + n * scc.repeat_interleave(T, 0)

    xx, ww, bb = x.double().requires_grad_(True), w.double().requires_grad_(True), b.double().requires_grad_(True)
    shh = sh.double().clone().requires_grad_(True) if mod else None
    scc = sc.double().clone().requires_grad_(True) if mod else None
    yr = ref(xx, ww, bb, shh, scc)
    y16, y = O.ln_fwd16(x, w, b, sh, sc, 6 * d, T, want32=True)
    assert _rel(y, yr) < 1e-5 and _rel(_join(y16), yr) < 2e-5
    grads = torch.autograd.grad(yr, [xx, ww, bb] + ([shh, scc] if mod else []), dy.double())
    dmod = torch.zeros(B, 6 * d, device="cuda")
    dx, dw, db = O.ln_bwd2(x, dy, w, b, sc, 6 * d, dres, dmod[:, :d] if mod else None, dmod[:, d:2 * d] if mod else None, 6 * d, T)
    assert _rel(dx, grads[0] + dres.double()) < 1e-5
    assert _rel(dw, grads[1]) < 1e-5 and _rel(db, grads[2]) < 1e-5
    if mod:
        assert _rel(dmod[:, :d], grads[3]) < 1e-5 and _rel(dmod[:, d:2 * d], grads[4]) < 1e-5


@pytest.mark.parametrize("gated,p", [(True, 0.0), (False, 0.0), (True, 0.25)])
def test_res_drop_fwd_bwd(gated, p):
    O = _ops()
    torch.manual_seed(5)
    B, T, d = 29, 10, 384
    x, f, dout = (torch.randn(B * T, d, device="cuda") for _ in range(3))
    modt = torch.randn(B, 6 * d, device="cuda")
    gate = modt[:, 2 * d:3 * d] if gated else None
    out = O.res_drop_fwd(x, f, gate, 6 * d, T, p, 99)
    gfull = gate.repeat_interleave(T, 0) if gated else torch.ones_like(x)
    mask = (out - x) / (gfull * f)            # = dropout mask / (1 - p)
    if p == 0:
        assert _rel(out, x + gfull * f) < 1e-6
    else:
        keep = mask.abs() > 0.5
        assert abs(float(keep.float().mean()) - (1 - p)) < 0.02
        assert torch.allclose(mask[keep], torch.full_like(mask[keep], 1 / (1 - p)), rtol=2e-2)      # (out - x) / (g f) cancels
        mask = keep.float() / (1 - p)
    m = mask if p > 0 else torch.ones_like(x)
    dmod = torch.zeros(B, 6 * d, device="cuda")
    df16, bsum = O.res_drop_bwd(dout, f, gate, 6 * d, dmod[:, 2 * d:3 * d] if gated else None, 6 * d, T, p, 99, want_bias=True)
    ref_df = gfull * m * dout
    assert _rel(_join(df16), ref_df) < 2e-5
    assert _rel(bsum, ref_df.double().sum(0)) < 1e-5
    if gated:
        assert _rel(dmod[:, 2 * d:3 * d], (dout * m * f).double().view(B, T, d).sum(1)) < 1e-5


def test_narrow_linear_ops():
    O = _ops()
    torch.manual_seed(6)
    M, d, J = 5120, 384, 7
    x, W, b = torch.randn(M, d, device="cuda"), torch.randn(J, d, device="cuda"), torch.randn(J, device="cuda")
    assert _rel(O.narrow_fwd(x, W, b), x.double() @ W.double().t() + b.double()) < 1e-5
    dy = torch.randn(M, J, device="cuda")
    assert _rel(O.narrow_wgrad(x, dy, False), dy.double().t() @ x.double()) < 1e-5          # d action_pred.weight (J, d)
    a = torch.randn(M, J, device="cuda")
    g = torch.randn(M, d, device="cuda")
    assert _rel(O.narrow_wgrad(g, a, True), g.double().t() @ a.double()) < 1e-5            # d action_emb.weight (d, J)


@pytest.mark.parametrize("M", [5120, 200])
def test_dgrad_with_gelu_backward_epilogue(M):
    """dh16 = split((dy . W) * GELU'(h)) and its column sums, produced by ONE GEMM (mode 1, epilogue 7)"""
    O = _ops()
    torch.manual_seed(7)
    N, K = 384, 1536                     # c_proj: dy (M, d), W (d, 4d) -> dg (M, 4d)
    dy, w, h = torch.randn(M, N, device="cuda"), torch.randn(N, K, device="cuda") / N ** 0.5, torch.randn(M, K, device="cuda")
    dh16, cs = O.dgrad_gelu_bwd16(O.split(dy), O.split(w), h, M, N, K, want_colsum=True)
    hh = h.double().requires_grad_(True)
    (g,) = torch.autograd.grad(torch.nn.functional.gelu(hh).sum(), hh)
    ref = (dy.double() @ w.double()) * g
    assert _rel(_join(dh16), ref) < 3e-5
    assert _rel(cs, ref.sum(0)) < 1e-5


@pytest.mark.parametrize("p", [0.0, 0.3])
def test_attention_backward_emitting_operands_matches_generic_kernel(p):
    """attention_bwd2_kernel (specialised, writes dqkv16 + bias partials) vs attention_bwd_kernel + split on the same inputs / dropout seed"""
    O = _ops()
    torch.manual_seed(8)
    B, H, hd, T = 33, 8, 48, 10
    D = H * hd
    qkv, dy = torch.randn(B * T, 3 * D, device="cuda"), torch.randn(B * T, D, device="cuda")
    ref = torch.empty_like(qkv)
    O.attn_bwd(qkv, 3 * D, qkv[:, D:], qkv[:, 2 * D:], 3 * D, dy, ref, 3 * D, ref[:, D:], ref[:, 2 * D:], 3 * D, B, H, hd, T, T, True, p, 1234)
    d16, bias = O.attn_bwd16_self(qkv, dy, B, H, hd, T, True, p, 1234, True)
    assert _rel(_join(d16), ref) < 2e-5
    assert _rel(bias, ref.double().sum(0)) < 1e-5
    # cross shape: 10 queries x 4 context tokens, strided k/v views of a wider buffer
    Tk = 4
    q = torch.randn(B * T, D, device="cuda")
    kv_all = torch.randn(B, Tk, 8 * D, device="cuda")
    kv = kv_all[..., 2 * D:4 * D]
    rq, rkv = torch.empty_like(q), torch.empty(B, Tk, 2 * D, device="cuda")
    O.attn_bwd(q, D, kv, kv[..., D:], kv.stride(1), dy, rq, D, rkv, rkv[..., D:], 2 * D, B, H, hd, T, Tk, True, p, 77)
    dq16, dkv, bq = O.attn_bwd16_cross(q, kv, dy, B, H, hd, T, Tk, True, p, 77, True)
    assert _rel(_join(dq16), rq) < 2e-5 and _rel(dkv, rkv) < 1e-5 and _rel(bq, rq.double().sum(0)) < 1e-5
