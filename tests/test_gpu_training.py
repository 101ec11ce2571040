"""Training path (BASELINE config 3 shape: GCDenoiser.loss + gradients): every primitive op against torch autograd in
fp64, then the full loss / all 157 parameter gradients against autograd over the CPU oracle, and against gradient
fingerprints produced by the reference itself (tests/golden/train_grads.npz)."""
import ctypes as C

import pytest
import torch

from oracle import mdt_oracle as orc
from tests import helpers as H
from mdt_policy_b200 import _lib, training as T
from mdt_policy_b200.synthetic import synthetic_inputs

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / (b.double().abs().max() + 1e-30))


@pytest.mark.parametrize("tc", [True, False])
@pytest.mark.parametrize("M,N,K", [(5120, 384, 384), (100, 1536, 384), (37, 384, 1536), (640, 7, 384), (640, 384, 7), (30, 2304, 384),
                                   (5120, 1536, 384), (1000, 384, 1536)])
def test_linear_fwd_bwd(M, N, K, tc, monkeypatch):
    """forward / input gradient / weight gradient in both GEMM back ends: tcgen05 bf16x3 (2^-17 operand rounding) and exact fp32"""
    monkeypatch.setattr(T, "USE_TC", tc)
    tol = 4e-5 if tc else 5e-6
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).double().requires_grad_()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).double().requires_grad_()
    b = torch.randn(N, generator=g).double().requires_grad_()
    dy = torch.randn(M, N, generator=g).double()
    y = torch.nn.functional.linear(x, w, b)
    y.backward(dy)
    xc, wc, bc = (t.detach().float().cuda().requires_grad_() for t in (x, w, b))
    yc = T.Linear.apply(xc, wc, bc)
    yc.backward(dy.float().cuda())
    assert _rel(yc, y) < tol and _rel(xc.grad, x.grad) < tol and _rel(wc.grad, w.grad) < tol and _rel(bc.grad, b.grad) < 5e-6


@pytest.mark.parametrize("kind,fn", [(1, torch.nn.functional.gelu), (2, torch.nn.functional.mish), (3, torch.nn.functional.silu)])
def test_activations(kind, fn):
    x = (torch.randn(4096, generator=torch.Generator().manual_seed(kind)) * 3).double().requires_grad_()
    y = fn(x)
    y.backward(torch.ones_like(y) * 0.7)
    xc = x.detach().float().cuda().requires_grad_()
    yc = T.Act.apply(xc, kind)
    yc.backward(torch.full_like(yc, 0.7))
    assert _rel(yc, y) < 2e-6 and _rel(xc.grad, x.grad) < 5e-6


@pytest.mark.parametrize("mod,bias", [(True, False), (False, True), (False, False)])
def test_layernorm_modulate(mod, bias):
    g = torch.Generator().manual_seed(5)
    B, Tn, d = 6, 10, 384
    x = (torch.randn(B, Tn, d, generator=g) * 2 + 0.5).double().requires_grad_()
    w = (1 + 0.1 * torch.randn(d, generator=g)).double().requires_grad_()
    b = (0.1 * torch.randn(d, generator=g)).double().requires_grad_() if bias else None
    sh = torch.randn(B, d, generator=g).double().requires_grad_() if mod else None
    sc = torch.randn(B, d, generator=g).double().requires_grad_() if mod else None
    dy = torch.randn(B, Tn, d, generator=g).double()
    n = torch.nn.functional.layer_norm(x, (d,), w, b, 1e-5)
    y = sh[:, None] + n * sc[:, None] if mod else n
    y.backward(dy)
    f = lambda t: t.detach().float().cuda().requires_grad_() if t is not None else None
    xc, wc, bc, shc, scc = f(x), f(w), f(b), f(sh), f(sc)
    yc = T.LayerNormMod.apply(xc, wc, bc, shc, scc)
    yc.backward(dy.float().cuda())
    assert _rel(yc, y) < 5e-6 and _rel(xc.grad, x.grad) < 2e-5 and _rel(wc.grad, w.grad) < 2e-5
    if bias:
        assert _rel(bc.grad, b.grad) < 2e-5
    if mod:
        assert _rel(shc.grad, sh.grad) < 2e-5 and _rel(scc.grad, sc.grad) < 2e-5


@pytest.mark.parametrize("Tq,Tk,causal,hd", [(10, 10, True, 48), (10, 4, True, 48), (4, 4, False, 48), (10, 3, True, 64)])
def test_attention_fwd_bwd(Tq, Tk, causal, hd):
    g = torch.Generator().manual_seed(Tq * 7 + Tk)
    B, Hh = 5, 8
    D = Hh * hd
    q, k, v = (torch.randn(B, t, D, generator=g).double().requires_grad_() for t in (Tq, Tk, Tk))
    dy = torch.randn(B, Tq, D, generator=g).double()
    sp = lambda t, n: t.view(B, n, Hh, hd).transpose(1, 2)
    mask = (torch.arange(Tk)[None, :] <= torch.arange(Tq)[:, None]) if causal else None
    y = torch.nn.functional.scaled_dot_product_attention(sp(q, Tq), sp(k, Tk), sp(v, Tk), attn_mask=mask).transpose(1, 2).reshape(B, Tq, D)
    y.backward(dy)
    qc, kc, vc = (t.detach().float().cuda().requires_grad_() for t in (q, k, v))
    yc = T.Attention.apply(qc, kc, vc, Hh, causal, 0.0, 0)
    yc.backward(dy.float().cuda())
    assert _rel(yc, y) < 5e-6 and _rel(qc.grad, q.grad) < 2e-5 and _rel(kc.grad, k.grad) < 2e-5 and _rel(vc.grad, v.grad) < 2e-5


@pytest.mark.parametrize("gated", [True, False])
def test_gate_residual(gated):
    g = torch.Generator().manual_seed(9)
    B, Tn, d = 4, 10, 384
    x, f = (torch.randn(B, Tn, d, generator=g).double().requires_grad_() for _ in range(2))
    gate = torch.randn(B, d, generator=g).double().requires_grad_() if gated else None
    dy = torch.randn(B, Tn, d, generator=g).double()
    y = x + (gate[:, None] * f if gated else f)
    y.backward(dy)
    xc, fc = (t.detach().float().cuda().requires_grad_() for t in (x, f))
    gc = gate.detach().float().cuda().requires_grad_() if gated else None
    yc = T.GateResidual.apply(xc, fc, gc)
    yc.backward(dy.float().cuda())
    assert _rel(yc, y) < 2e-6 and _rel(xc.grad, x.grad) < 2e-6 and _rel(fc.grad, f.grad) < 2e-6
    if gated:
        assert _rel(gc.grad, gate.grad) < 5e-6


def _no_dropout(cfg):
    cfg.update(attn_pdrop=0, resid_pdrop=0, mlp_pdrop=0, embed_pdrob=0, goal_drop=0)
    return cfg


@pytest.mark.parametrize("variant", ["mdtv", "mdt"])
def test_loss_and_all_gradients_vs_oracle_autograd(variant):
    """GCDenoiser.loss in train mode (dropout 0): value and every parameter gradient vs autograd over the fp32 CPU oracle
    (rel-tol 1e-3 of the gradient's max, abs 1e-6 -- SURVEY 8d config 3)."""
    if variant == "mdtv":
        cfg, ocfg, Bn = _no_dropout(H.mdtv_inner_cfg(2, 2)), orc.OracleCfg(n_enc_layers=2, n_dec_layers=2), 16
        inp = synthetic_inputs(Bn, seed=92)
        state = {"state_images": inp["state_images"], "modality": "lang"}
    else:
        cfg, ocfg, Bn = _no_dropout(H.mdt_inner_cfg(n_enc_layers=1, n_dec_layers=2)), orc.OracleCfg(embed_dim=512, n_enc_layers=1, n_dec_layers=2, variant="mdt"), 8
        inp = synthetic_inputs(Bn, seed=93, n_state_tokens=2, obs_dim=512)
        state = {"static": inp["state_images"][:, :1], "gripper": inp["state_images"][:, 1:], "modality": "lang"}
    model = H.build_product(cfg, 91, "trained").train()
    P = {k: v.clone().requires_grad_() for k, v in H.oracle_params([(n, p.shape) for n, p in model.named_parameters()], 91, "trained").items()}
    sigma = torch.exp(torch.linspace(3.0, -4.0, Bn))
    loss_o, out_o = orc.denoiser_loss(P, ocfg, state, inp["actions"], inp["goal"], inp["noise"], sigma)
    loss_o.backward()
    dstate = {k: (v.cuda() if isinstance(v, torch.Tensor) else v) for k, v in state.items()}
    loss, out = model.loss(dstate, inp["actions"].cuda(), inp["goal"].cuda(), inp["noise"].cuda(), sigma.cuda())
    loss.backward()
    assert abs(float(loss) - float(loss_o)) < 1e-5 * max(1.0, abs(float(loss_o)))
    assert (out.detach().cpu() - out_o.detach()).abs().max() < 2e-5 * max(1.0, float(out_o.abs().max()))
    worst = ("", 0.0)
    for name, p in model.named_parameters():
        go = P[name].grad
        if go is None:                       # pos_emb / proprio_emb (and lang_emb under the MDT forward) are unused
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        assert p.grad is not None, name
        err = float((p.grad.cpu() - go).abs().max())
        tol = 1e-3 * float(go.abs().max()) + 1e-6
        if err / tol > worst[1]:
            worst = (name, err / tol)
        assert err < tol, (name, err, float(go.abs().max()))
    print("worst gradient error / tolerance:", worst)


def test_input_gradients_vs_oracle_autograd():
    """d loss / d (noised actions via the action tensor, goal, state tokens): the fused training graph also carries gradients to its
    inputs (guidance-style uses), checked against autograd over the oracle"""
    cfg, ocfg, Bn = _no_dropout(H.mdtv_inner_cfg(2, 2)), orc.OracleCfg(n_enc_layers=2, n_dec_layers=2), 8
    inp = synthetic_inputs(Bn, seed=94)
    model = H.build_product(cfg, 91, "trained").train()
    P = H.oracle_params([(n, p.shape) for n, p in model.named_parameters()], 91, "trained")
    sigma = torch.exp(torch.linspace(2.0, -3.0, Bn))
    xin = (inp["actions"] + inp["noise"] * sigma[:, None, None])
    a_o, g_o, s_o = (t.clone().requires_grad_() for t in (xin, inp["goal"], inp["state_images"]))
    out_o = orc.denoiser_forward(P, ocfg, {"state_images": s_o, "modality": "lang"}, a_o, g_o, sigma)
    out_o.square().sum().backward()
    a_c, g_c, s_c = (t.clone().cuda().requires_grad_() for t in (xin, inp["goal"], inp["state_images"]))
    out = model({"state_images": s_c, "modality": "lang"}, a_c, g_c, sigma.cuda())
    out.square().sum().backward()
    for name, got, want in (("actions", a_c.grad, a_o.grad), ("goal", g_c.grad, g_o.grad), ("state", s_c.grad, s_o.grad)):
        assert got is not None, name
        assert float((got.cpu() - want).abs().max()) < 1e-3 * float(want.abs().max()) + 1e-6, name


def test_gradients_vs_reference_fingerprints():
    """The same check against the reference itself: fingerprints (L2 norm + 8 probe entries per parameter) of the gradients
    of the reference's GCDenoiser.loss, generated by tests/golden/make_golden.py."""
    meta, gold = H.load_golden("train_grads")
    model = H.build_product(_no_dropout(H.mdtv_inner_cfg(meta["enc"], meta["dec"])), meta["weight_seed"], "trained").train()
    Bn = meta["B"]
    inp = {k: v.cuda() for k, v in synthetic_inputs(Bn, seed=meta["input_seed"]).items()}
    sigma = torch.exp(torch.linspace(3.0, -4.0, Bn)).cuda()
    loss, _ = model.loss({"state_images": inp["state_images"], "modality": "lang"}, inp["actions"], inp["goal"], inp["noise"], sigma)
    loss.backward()
    assert abs(float(loss) - float(gold["loss"])) < 1e-5 * max(1.0, float(gold["loss"]))
    names = [n for n, _ in model.named_parameters()]
    for i, name in enumerate(names):
        p = dict(model.named_parameters())[name]
        gn, probe = float(gold["norms"][i]), gold["probes"][i]
        if gn == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0
            continue
        flat = p.grad.flatten().cpu()
        idx = torch.linspace(0, flat.numel() - 1, 8).long()
        assert abs(float(flat.norm()) - gn) < 1e-3 * gn + 1e-7, name
        assert (flat[idx] - probe).abs().max() < 1e-3 * float(flat.abs().max()) + 1e-6, name


def test_optimizer_step_reduces_loss_and_resyncs_inference_weights():
    """A few AdamW steps on one synthetic batch through the CUDA training path lower the loss, and the inference engine
    picks the updated weights up afterwards (version-counter resync)."""
    model = H.build_product(_no_dropout(H.mdtv_inner_cfg(1, 1)), 95, "trained").train()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, betas=(0.9, 0.9), weight_decay=0.05)
    inp = {k: v.cuda() for k, v in synthetic_inputs(32, seed=96).items()}
    sigma = torch.exp(torch.linspace(2.0, -3.0, 32)).cuda()
    state = {"state_images": inp["state_images"], "modality": "lang"}
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        loss, _ = model.loss(state, inp["actions"], inp["goal"], inp["noise"], sigma)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0]
    model.eval()
    with torch.no_grad():
        l_inf, _ = model.loss(state, inp["actions"], inp["goal"], inp["noise"], sigma)      # inference kernels, new weights
        model.train()
    with torch.enable_grad():
        l_tr, _ = model.loss(state, inp["actions"], inp["goal"], inp["noise"], sigma)
    assert abs(float(l_inf) - float(l_tr)) < 1e-3 * max(1.0, float(l_tr))


def test_dropout_ops_mask_consistency():
    """Element dropout: keep fraction ~ 1-p, kept values scaled by 1/(1-p), backward uses the SAME mask.  Attention dropout:
    with V = one-hot the output reveals the (dropped, rescaled) probabilities; forward and backward agree with a torch
    computation that applies that very mask."""
    x = torch.randn(1 << 18, device="cuda", requires_grad=True)
    y = T.Dropout.apply(x, 0.3, 12345)
    keep = (y != 0)
    assert abs(float(keep.float().mean()) - 0.7) < 0.01
    assert torch.allclose(y[keep], x.detach()[keep] / 0.7, rtol=1e-6)
    y.backward(torch.ones_like(y))
    assert torch.equal(x.grad != 0, keep) and torch.allclose(x.grad[keep], torch.full_like(x.grad[keep], 1 / 0.7))
    assert not torch.equal(T.Dropout.apply(x.detach(), 0.3, 999) != 0, keep)        # another seed, another mask

    B, Hh, Tq, Tk, hd = 3, 8, 10, 10, 16
    D = Hh * hd
    g = torch.Generator().manual_seed(3)
    q, k = (torch.randn(B, t, D, generator=g).cuda() for t in (Tq, Tk))
    v_onehot = torch.zeros(B, Tk, Hh, hd); v_onehot[:, torch.arange(Tk), :, torch.arange(Tk)] = 1.0     # v[b, j, h, c] = (c == j)
    p_used = T.Attention.apply(q, k, v_onehot.view(B, Tk, D).cuda(), Hh, True, 0.3, 777).view(B, Tq, Hh, hd)[..., :Tk]   # (B,Tq,H,Tk)
    sp = lambda t, n: t.view(B, n, Hh, hd).transpose(1, 2)
    s = (sp(q, Tq) @ sp(k, Tk).transpose(-1, -2)) / hd ** 0.5
    mask = torch.arange(Tk, device="cuda")[None, :] <= torch.arange(Tq, device="cuda")[:, None]
    P = torch.softmax(s.masked_fill(~mask, float("-inf")), dim=-1)                  # (B,H,Tq,Tk)
    pu = p_used.permute(0, 2, 1, 3)
    m = torch.where(pu != 0, torch.full_like(pu, 1 / 0.7), torch.zeros_like(pu))
    live = P > 1e-6
    assert torch.allclose(pu[live & (pu != 0)], (P * m)[live & (pu != 0)], rtol=2e-5, atol=1e-7)
    assert abs(float((pu[live] != 0).float().mean()) - 0.7) < 0.05
    # gradients with the same mask
    v = torch.randn(B, Tk, D, generator=g).cuda()
    qa, ka, va = (t.clone().requires_grad_() for t in (q, k, v))
    ya = T.Attention.apply(qa, ka, va, Hh, True, 0.3, 777)
    dy = torch.randn(B, Tq, D, generator=g).cuda()
    ya.backward(dy)
    qb, kb, vb = (t.clone().double().requires_grad_() for t in (q, k, v))
    sb = (sp(qb, Tq) @ sp(kb, Tk).transpose(-1, -2)) / hd ** 0.5
    Pb = torch.softmax(sb.masked_fill(~mask, float("-inf")), dim=-1) * m.double()
    yb = (Pb @ sp(vb, Tk)).transpose(1, 2).reshape(B, Tq, D)
    yb.backward(dy.double())
    assert _rel(ya, yb) < 5e-6 and _rel(qa.grad, qb.grad) < 2e-5 and _rel(ka.grad, kb.grad) < 2e-5 and _rel(va.grad, vb.grad) < 2e-5


def test_training_with_shipped_dropout_probabilities():
    """The shipped config (attn 0.3 / resid 0.1 / mlp 0.05) trains: finite loss and gradients, reproducible under
    torch.manual_seed, different under another seed, and the mean loss over seeds is close to the dropout-free loss scale."""
    model = H.build_product(H.mdtv_inner_cfg(2, 2), 97, "trained").train()
    inp = {k: v.cuda() for k, v in synthetic_inputs(32, seed=98).items()}
    sigma = torch.exp(torch.linspace(2.0, -3.0, 32)).cuda()
    state = {"state_images": inp["state_images"], "modality": "lang"}

    def run(seed):
        torch.manual_seed(seed)
        model.zero_grad(set_to_none=True)
        loss, _ = model.loss(state, inp["actions"], inp["goal"], inp["noise"], sigma)
        loss.backward()
        return float(loss), model.inner_model.decoder.blocks[0].mlp.c_fc.weight.grad.clone()

    l1, g1 = run(1)
    l1b, g1b = run(1)
    l2, g2 = run(2)
    assert l1 == l1b and torch.equal(g1, g1b)
    assert l1 != l2 and not torch.equal(g1, g2)
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    model.eval()
    with torch.no_grad():
        l_eval = float(model.loss(state, inp["actions"], inp["goal"], inp["noise"], sigma)[0])
    model.train()
    mean_l = sum(run(s)[0] for s in range(3, 9)) / 6
    assert 0.3 * l_eval < mean_l < 3.0 * l_eval


def test_back_to_back_steps_without_host_sync_match_synchronised_steps():
    """Regression (round 2): the tensor-core GEMM requested its W tiles ahead of the programmatic-dependent-launch wait, which is
    only valid for static inference weights -- in the training ops W is written by the split kernel launched right before, so
    un-synchronised training steps could read it early (timing-dependent NaNs).  Dropout off exposes it fastest (fewer kernels)."""
    import math
    from mdt_policy_b200 import utils as U

    def run(sync):
        model = H.build_product(H.mdtv_inner_cfg(4, 4, attn_pdrop=0.0, resid_pdrop=0.0, mlp_pdrop=0.0, max_batch=512), 12, "trained").train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.9), weight_decay=0.05)
        inp = {k: v.cuda() for k, v in synthetic_inputs(512, seed=31).items()}
        torch.manual_seed(0)
        sig = U.rand_log_logistic((512,), loc=math.log(0.5), scale=0.5, min_value=0.001, max_value=80.0).cuda()
        state = {"state_images": inp["state_images"], "modality": "lang"}
        losses = []
        for _ in range(8):
            opt.zero_grad(set_to_none=True)
            loss, _ = model.loss(state, inp["actions"], inp["goal"], inp["noise"], sig)
            loss.backward()
            opt.step()
            losses.append(loss.detach())
            if sync:
                torch.cuda.synchronize()
        torch.cuda.synchronize()
        return torch.stack(losses).cpu()

    a, b = run(True), run(False)
    assert torch.isfinite(b).all(), b
    assert torch.equal(a, b), (a, b)
