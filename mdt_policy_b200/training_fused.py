"""Training path, fused composition: one autograd node per residual branch of the score network.

Same semantics as training.py (train-mode forward of mdtv_transformer.py:208-236 / transformer_blocks.py Block :209-214 and
ConditionedBlock :292-309, hash-based dropout masks), but instead of one autograd Function per primitive each residual branch

    x + gate * dropout(c_proj(attention(q(LN(x)), k, v)))          x + gate * dropout(c_proj(GELU(c_fc(LN(x)))))

is ONE Function whose forward and backward chain the kernels of train_ops.py by hand: LayerNorm / attention / the residual
backward / the activation backward emit the split-bf16 GEMM operands directly, every GEMM (forward, dgrad, wgrad) reads one
row-major split per tensor (no transposed copies), q/k/v -- and the cross-attention K/V and AdaLN modulations of ALL layers --
are single GEMMs over row-stacked weight operands that a WeightBank refreshes with one launch per step, the residual-stream
gradient is added inside the LayerNorm backward, and bias / LayerNorm / AdaLN gradients come from per-CTA partial sums.
~330 kernel launches per step instead of ~760 (profiles/r02_train_*.md).
"""
from __future__ import annotations

import math

import torch
from torch.autograd import Function

from . import train_ops as O
from . import training as T


def eligible(net) -> bool:
    d = net.embed_dim
    if d % 128 or d > 512 or (d // net.n_heads) > 64 or (d // net.n_heads) % 4:
        return False
    for m in list(net.encoder.blocks) + list(net.decoder.blocks):
        for lin in (m.attn.query, m.attn.c_proj, m.mlp.c_fc, m.mlp.c_proj):
            if lin.weight.shape[0] % 64 or lin.weight.shape[1] % 64:
                return False
    return True


def _bias_list(mods):
    bs = [m.bias for m in mods]
    if any(b is None for b in bs):
        if not all(b is None for b in bs):
            raise RuntimeError("a weight group mixes biased and bias-free projections")
        return [None]
    return bs


def _bank(net):
    bank = net.__dict__.get("_train_bank")
    dev = net.action_emb.weight.device
    if bank is not None and bank.valid() and bank.device == dev:
        return bank
    groups = []

    def add(name, mods):
        groups.append((name, [m.weight for m in mods], _bias_list(mods)))

    for i, b in enumerate(net.encoder.blocks):
        add(f"enc{i}.qkv", [b.attn.query, b.attn.key, b.attn.value]); add(f"enc{i}.o", [b.attn.c_proj])
        add(f"enc{i}.fc", [b.mlp.c_fc]); add(f"enc{i}.proj", [b.mlp.c_proj])
    for i, b in enumerate(net.decoder.blocks):
        add(f"dec{i}.qkv", [b.attn.query, b.attn.key, b.attn.value]); add(f"dec{i}.o", [b.attn.c_proj])
        add(f"dec{i}.xq", [b.cross_att.query]); add(f"dec{i}.xo", [b.cross_att.c_proj])
        add(f"dec{i}.fc", [b.mlp.c_fc]); add(f"dec{i}.proj", [b.mlp.c_proj])
    add("dec.xkv", [m for b in net.decoder.blocks for m in (b.cross_att.key, b.cross_att.value)])
    add("dec.mod", [b.adaLN_zero.modulation[1] for b in net.decoder.blocks])
    # the embedding MLPs / token embeddings (goal, language goal, state tokens, sigma): one group per Linear whose dims allow the tensor cores
    seen = set()
    for name, mod in _embedding_linears(net):
        w = mod.weight
        if id(w) not in seen and w.shape[0] % 64 == 0 and w.shape[1] % 64 == 0:
            seen.add(id(w))
            add(name, [mod])
    bank = O.WeightBank(groups, dev)
    net.__dict__["_train_bank"] = bank
    return bank


def _embedding_linears(net):
    out = [("emb.goal0", net.goal_emb[0]), ("emb.goal2", net.goal_emb[2]), ("emb.lang0", net.lang_emb[0]), ("emb.lang2", net.lang_emb[2]),
           ("emb.tok", net.tok_emb), ("emb.sig1", net.sigma_emb[1]), ("emb.sig3", net.sigma_emb[3])]
    if getattr(net, "incam_embed", None) is not None:
        out.append(("emb.incam", net.incam_embed))
    return out


def _emb_lin(net, bank, mod, x):
    """Linear of an embedding path: through the weight bank (one split of x, pre-split weight) when the module is in it"""
    name = net.__dict__.setdefault("_emb_names", {}).get(id(mod.weight))
    if name is None:
        name = next((n for n, m in _embedding_linears(net) if m.weight is mod.weight and n in bank.w16), "")
        net.__dict__["_emb_names"][id(mod.weight)] = name
    return linear_group(x, bank, name, [mod]) if name else T._lin(mod, x)


def _rows(ws):
    out, r = [], 0
    for w in ws:
        out.append((r, r + w.shape[0]))
        r += w.shape[0]
    return out


class LinearG(Function):
    """y = x . cat(W_i)^T + cat(b_i): one GEMM over a row-stacked weight group of the bank (the params are passed for the autograd
    edges only).  dW comes back as row slices of one (sum N_i, K) weight-gradient GEMM."""

    @staticmethod
    def forward(ctx, x, bank, name, n_w, *params):
        ws = params[:n_w]
        K = ws[0].shape[1]
        N = sum(w.shape[0] for w in ws)
        x2 = T._c(x).reshape(-1, K)
        M = x2.shape[0]
        x16 = O.split(x2)
        y = O.gemm16(0, x16, bank.w16[name], M, N, K, bias=bank.bias[name])
        ctx.save_for_backward(x16)
        ctx.cfg = (bank, name, n_w, M, N, K, x.shape, _rows(ws), bank.bias[name] is not None)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        (x16,) = ctx.saved_tensors
        bank, name, n_w, M, N, K, xshape, rows, has_bias = ctx.cfg
        dy2 = T._c(dy).reshape(M, N)
        if has_bias:
            dy16, db = O.split(dy2, want_colsum=True)
        else:
            dy16, db = O.split(dy2), None
        dw = O.wgrad_begin(dy16, x16, M, N, K)
        dx = O.gemm16(1, dy16, bank.w16[name], M, N, K).view(xshape) if ctx.needs_input_grad[0] else None
        O.wgrad_join(dy16.device)
        gw = [dw[a:b] for a, b in rows]
        gb = [db[a:b] for a, b in rows] if has_bias else []
        return (dx, None, None, None, *gw, *gb)


def linear_group(x, bank, name, mods):
    ps = [m.weight for m in mods] + [m.bias for m in mods if m.bias is not None]
    return LinearG.apply(x, bank, name, len(mods), *ps)


class AttnBranch(Function):
    """out = x + gate * dropout(c_proj(attention(q, k, v))), q (and k, v for self-attention) projected from LN(x)(+modulate)."""

    @staticmethod
    def forward(ctx, x, kv, shift, scale, gate, cfg, ln_w, ln_b, *params):
        bank, n_in, n_o, H, causal, p_attn, p_res = cfg
        B, Tq, d = x.shape
        M, hd = B * Tq, d // H
        x2 = T._c(x).view(M, d)
        mstride = shift.stride(0) if shift is not None else 0
        gstride = gate.stride(0) if gate is not None else 0
        a16 = O.ln_fwd16(x2, ln_w, ln_b, shift, scale, mstride, Tq)
        seed_a = T._new_seed() if p_attn > 0 else 0
        seed_r = T._new_seed() if p_res > 0 else 0
        if kv is None:
            qkv = O.gemm16(0, a16, bank.w16[n_in], M, 3 * d, d, bias=bank.bias[n_in])
            q, k, v, ldq, ldkv, Tk = qkv, qkv[:, d:], qkv[:, 2 * d:], 3 * d, 3 * d, Tq
        else:
            qkv = O.gemm16(0, a16, bank.w16[n_in], M, d, d, bias=bank.bias[n_in])
            if kv.stride(-1) != 1 or kv.stride(0) != kv.shape[1] * kv.stride(1):
                raise RuntimeError("cross-attention K|V must be a last-dim slice of a contiguous (B, Tk, *) tensor")
            q, k, v, ldq, ldkv, Tk = qkv, kv, kv[..., d:], d, kv.stride(1), kv.shape[1]
        y16 = O.attn_fwd16(q, ldq, k, v, ldkv, B, H, hd, Tq, Tk, causal, p_attn, seed_a)
        f = O.gemm16(0, y16, bank.w16[n_o], M, d, d, bias=bank.bias[n_o])
        out = O.res_drop_fwd(x2, f, gate, gstride, Tq, p_res, seed_r)
        ctx.save_for_backward(x2, a16, qkv, kv, y16, f if gate is not None else None, scale, gate, ln_w, ln_b)
        ctx.cfg = (cfg, B, Tq, Tk, d, seed_a, seed_r, shift is not None, len(params))
        return out.view(B, Tq, d)

    @staticmethod
    def backward(ctx, dout):
        x2, a16, qkv, kv, y16, f, scale, gate, ln_w, ln_b = ctx.saved_tensors
        (bank, n_in, n_o, H, causal, p_attn, p_res), B, Tq, Tk, d, seed_a, seed_r, has_mod, n_params = ctx.cfg
        M, hd = B * Tq, d // H
        dev = x2.device
        dout2 = T._c(dout).view(M, d)
        dmod = torch.empty(B, 3 * d, dtype=torch.float32, device=dev) if has_mod else None
        dshift, dscale, dgate = (dmod[:, :d], dmod[:, d:2 * d], dmod[:, 2 * d:]) if has_mod else (None, None, None)
        has_bo, has_bi = bank.bias[n_o] is not None, bank.bias[n_in] is not None
        r = O.res_drop_bwd(dout2, f, gate, gate.stride(0) if gate is not None else 0, dgate, 3 * d, Tq, p_res, seed_r, want_bias=has_bo)
        df16, dbo = r if has_bo else (r, None)
        dwo = O.wgrad_begin(df16, y16, M, d, d)
        dy = O.gemm16(1, df16, bank.w16[n_o], M, d, d)
        fast = (hd, Tq, Tk, bool(causal)) in O.ATTN_BWD16_SHAPES      # specialised kernel: emits the operand + bias partials itself
        if kv is None:
            Nin, dkv = 3 * d, None
            if fast:
                dq16, dbi = O.attn_bwd16_self(qkv, dy, B, H, hd, Tq, causal, p_attn, seed_a, has_bi)
            else:
                dqkv = torch.empty(M, 3 * d, dtype=torch.float32, device=dev)
                O.attn_bwd(qkv, 3 * d, qkv[:, d:], qkv[:, 2 * d:], 3 * d, dy, dqkv, 3 * d, dqkv[:, d:], dqkv[:, 2 * d:], 3 * d, B, H, hd, Tq, Tk,
                           causal, p_attn, seed_a)
        else:
            Nin = d
            if fast:
                dq16, dkv, dbi = O.attn_bwd16_cross(qkv, kv, dy, B, H, hd, Tq, Tk, causal, p_attn, seed_a, has_bi)
            else:
                dqkv = torch.empty(M, d, dtype=torch.float32, device=dev)
                dkv = torch.empty(B, Tk, 2 * d, dtype=torch.float32, device=dev)
                O.attn_bwd(qkv, d, kv, kv[..., d:], kv.stride(1), dy, dqkv, d, dkv, dkv[..., d:], 2 * d, B, H, hd, Tq, Tk, causal, p_attn, seed_a)
        if not fast:
            if has_bi:
                dq16, dbi = O.split(dqkv, want_colsum=True)
            else:
                dq16, dbi = O.split(dqkv), None
        dwi = O.wgrad_begin(dq16, a16, M, Nin, d)
        da = O.gemm16(1, dq16, bank.w16[n_in], M, Nin, d)
        dx, dlw, dlb = O.ln_bwd2(x2, da, ln_w, ln_b, scale, scale.stride(0) if scale is not None else 0, dout2, dshift, dscale, 3 * d, Tq)
        O.wgrad_join(dev)
        n_w = Nin // d
        gw = [dwi[i * d:(i + 1) * d] for i in range(n_w)] + [dwo]
        gb = ([dbi[i * d:(i + 1) * d] for i in range(n_w)] if has_bi else []) + ([dbo] if has_bo else [])
        grads = gw + gb
        if len(grads) != n_params:
            raise RuntimeError("internal: gradient list does not match the parameters passed to AttnBranch")
        return (dx.view(B, Tq, d), dkv, dshift, dscale, dgate, None, dlw, dlb, *grads)


class MLPBranch(Function):
    """out = x + gate * dropout(c_proj(GELU(c_fc(LN(x)(+modulate)))))"""

    @staticmethod
    def forward(ctx, x, shift, scale, gate, cfg, ln_w, ln_b, *params):
        bank, n_fc, n_proj, p_drop = cfg
        B, Tq, d = x.shape
        M = B * Tq
        F = bank.w16[n_fc].shape[0]
        x2 = T._c(x).view(M, d)
        a16 = O.ln_fwd16(x2, ln_w, ln_b, shift, scale, shift.stride(0) if shift is not None else 0, Tq)
        h, g16 = O.gemm16(0, a16, bank.w16[n_fc], M, F, d, bias=bank.bias[n_fc], epi=O.EPI_GELU16)
        f = O.gemm16(0, g16, bank.w16[n_proj], M, d, F, bias=bank.bias[n_proj])
        seed = T._new_seed() if p_drop > 0 else 0
        out = O.res_drop_fwd(x2, f, gate, gate.stride(0) if gate is not None else 0, Tq, p_drop, seed)
        ctx.save_for_backward(x2, a16, h, g16, f if gate is not None else None, scale, gate, ln_w, ln_b)
        ctx.cfg = (cfg, B, Tq, d, F, seed, shift is not None, len(params))
        return out.view(B, Tq, d)

    @staticmethod
    def backward(ctx, dout):
        x2, a16, h, g16, f, scale, gate, ln_w, ln_b = ctx.saved_tensors
        (bank, n_fc, n_proj, p_drop), B, Tq, d, F, seed, has_mod, n_params = ctx.cfg
        M = B * Tq
        dout2 = T._c(dout).view(M, d)
        dmod = torch.empty(B, 3 * d, dtype=torch.float32, device=x2.device) if has_mod else None
        dshift, dscale, dgate = (dmod[:, :d], dmod[:, d:2 * d], dmod[:, 2 * d:]) if has_mod else (None, None, None)
        has_bp, has_bf = bank.bias[n_proj] is not None, bank.bias[n_fc] is not None
        r = O.res_drop_bwd(dout2, f, gate, gate.stride(0) if gate is not None else 0, dgate, 3 * d, Tq, p_drop, seed, want_bias=has_bp)
        df16, dbp = r if has_bp else (r, None)
        dwp = O.wgrad_begin(df16, g16, M, d, F)                      # (d, F)
        # dL/dh = (df . Wproj) * GELU'(h): the activation backward, the operand split and the c_fc bias-gradient partials all happen in the
        # dgrad GEMM's epilogue (no fp32 (M, F) gradient is ever written)
        if has_bf:
            dh16, dbf = O.dgrad_gelu_bwd16(df16, bank.w16[n_proj], h, M, d, F, want_colsum=True)
        else:
            dh16, dbf = O.dgrad_gelu_bwd16(df16, bank.w16[n_proj], h, M, d, F), None
        dwf = O.wgrad_begin(dh16, a16, M, F, d)
        da = O.gemm16(1, dh16, bank.w16[n_fc], M, F, d)
        dx, dlw, dlb = O.ln_bwd2(x2, da, ln_w, ln_b, scale, scale.stride(0) if scale is not None else 0, dout2, dshift, dscale, 3 * d, Tq)
        O.wgrad_join(x2.device)
        grads = [dwf, dwp] + ([dbf] if has_bf else []) + ([dbp] if has_bp else [])
        if len(grads) != n_params:
            raise RuntimeError("internal: gradient list does not match the parameters passed to MLPBranch")
        return (dx.view(B, Tq, d), dshift, dscale, dgate, None, dlw, dlb, *grads)


class FinalNorm(Function):
    """plain LayerNorm (encoder.ln / decoder.ln) with the partial-sum backward"""

    @staticmethod
    def forward(ctx, x, w, b):
        B, Tq, d = x.shape
        x2 = T._c(x).view(B * Tq, d)
        ctx.save_for_backward(x2, w, b)
        ctx.shape = (B, Tq, d)
        return O.ln_fwd(x2, w, b).view(B, Tq, d)

    @staticmethod
    def backward(ctx, dy):
        x2, w, b = ctx.saved_tensors
        B, Tq, d = ctx.shape
        dx, dw, db = O.ln_bwd2(x2, T._c(dy).view(B * Tq, d), w, b, None, 0, None, None, None, 0, Tq)
        return dx.view(B, Tq, d), dw, db


class NarrowIn(Function):
    """y = x W^T + b for a few input features (action_emb: 7 -> d)"""

    @staticmethod
    def forward(ctx, x, w, b):
        K, N = w.shape[1], w.shape[0]
        x2 = T._c(x).reshape(-1, K)
        y = torch.empty(x2.shape[0], N, dtype=torch.float32, device=x.device)
        T._gemm(0, x2, w, b, y, x2.shape[0], N, K)
        ctx.save_for_backward(x2, w)
        ctx.cfg = (x.shape, b is not None)
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        xshape, has_bias = ctx.cfg
        M, N, K = x2.shape[0], w.shape[0], w.shape[1]
        dy2 = T._c(dy).reshape(M, N)
        dx = None
        if ctx.needs_input_grad[0]:          # only when the noised actions themselves carry a gradient (guidance-style uses)
            dx = torch.empty(M, K, dtype=torch.float32, device=dy.device)
            T._gemm(1, dy2, w, None, dx, M, N, K)
            dx = dx.view(xshape)
        dw = O.narrow_wgrad(dy2, x2, True)
        db = T._colsum(dy2) if has_bias else None
        return dx, dw, db


class NarrowOut(Function):
    """y = x W^T + b for a few output features (action_pred: d -> 7)"""

    @staticmethod
    def forward(ctx, x, w, b):
        K = w.shape[1]
        x2 = T._c(x).reshape(-1, K)
        y = O.narrow_fwd(x2, w, b)
        ctx.save_for_backward(x2, w)
        ctx.cfg = (x.shape, b is not None)
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        xshape, has_bias = ctx.cfg
        M, J = x2.shape[0], w.shape[0]
        dy2 = T._c(dy).reshape(M, J)
        dx = torch.empty_like(x2)
        T._gemm(1, dy2, w, None, dx, M, J, w.shape[1])
        dw = O.narrow_wgrad(x2, dy2, False)
        db = T._colsum(dy2) if has_bias else None
        return dx.view(xshape), dw, db


def _attn_params(att):
    ws = [att.query.weight, att.key.weight, att.value.weight, att.c_proj.weight]
    return ws + [m.bias for m in (att.query, att.key, att.value, att.c_proj) if m.bias is not None]


def _xattn_params(att):
    return [att.query.weight, att.c_proj.weight] + [m.bias for m in (att.query, att.c_proj) if m.bias is not None]


def _mlp_params(m):
    return [m.c_fc.weight, m.c_proj.weight] + [x.bias for x in (m.c_fc, m.c_proj) if x.bias is not None]


def encode_train(net, states, goals):
    """forward_enc_only with gradients (mdtv_transformer.py:213-222 / mdt_transformer.py:211-229)."""
    bank = _bank(net)
    bank.refresh()
    net.__dict__["_train_bank_fresh"] = True
    goals = T._prep_goal_train(net, states, goals)
    train = net.training
    lang = net.use_modality_encoder and states.get("modality") == "lang" and net._variant == "mdtv"
    gm = net.lang_emb if lang else net.goal_emb
    g = _emb_lin(net, bank, gm[2], T.Act.apply(_emb_lin(net, bank, gm[0], goals[:, :1, :].float()), T.ACT_GELU))
    if net._variant == "mdtv":
        s = _emb_lin(net, bank, net.tok_emb, states["state_images"].float())
    else:
        st = _emb_lin(net, bank, net.tok_emb, states["static"].float())
        gr = _emb_lin(net, bank, net.incam_embed, states["gripper"].float())
        s = torch.cat((st, gr), dim=1)
        g = T._drop(g + net.pos_emb[:, : net.goal_seq_len, :], net.drop.p if train else 0.0)
        s = T._drop(s + net.pos_emb[:, net.goal_seq_len: net.goal_seq_len + 1, :], net.drop.p if train else 0.0)
    x = torch.cat([g, s], dim=1).contiguous()
    H = net.n_heads
    for i, blk in enumerate(net.encoder.blocks):
        a = blk.attn
        cfg = (bank, f"enc{i}.qkv", f"enc{i}.o", H, False, a.attn_dropout.p if train else 0.0, a.resid_dropout.p if train else 0.0)
        x = AttnBranch.apply(x, None, None, None, None, cfg, blk.ln_1.weight, blk.ln_1.bias, *_attn_params(a))
        cfg = (bank, f"enc{i}.fc", f"enc{i}.proj", blk.mlp.dropout.p if train else 0.0)
        x = MLPBranch.apply(x, None, None, None, cfg, blk.ln_2.weight, blk.ln_2.bias, *_mlp_params(blk.mlp))
    return FinalNorm.apply(x, net.encoder.ln.weight, net.encoder.ln.bias)


def decode_train(net, ctx, actions, sigma):
    """forward_dec_only with gradients (mdtv_transformer.py:224-236, ConditionedBlock :292-309)."""
    bank = _bank(net)
    if not net.__dict__.pop("_train_bank_fresh", False):
        bank.refresh()
    d = net.embed_dim
    half = d // 2
    e = sigma.float().log() / 4
    f = torch.exp(torch.arange(half, device=sigma.device, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    ang = e[:, None] * f[None, :]
    pe = torch.cat((ang.sin(), ang.cos()), dim=-1)                       # no parameters, no gradient: input preparation
    c = _emb_lin(net, bank, net.sigma_emb[3], T.Act.apply(_emb_lin(net, bank, net.sigma_emb[1], pe), T.ACT_MISH))            # (B, d)
    sc_c = T.Act.apply(c, T.ACT_SILU)
    train = net.training
    blocks = list(net.decoder.blocks)
    mod_all = linear_group(sc_c, bank, "dec.mod", [b.adaLN_zero.modulation[1] for b in blocks])                       # (B, L * 6d)
    kv_all = linear_group(ctx, bank, "dec.xkv", [m for b in blocks for m in (b.cross_att.key, b.cross_att.value)])   # (B, Tc, L * 2d)
    mods, kvs = mod_all.split(6 * d, dim=-1), kv_all.split(2 * d, dim=-1)
    x = T._drop(NarrowIn.apply(actions, net.action_emb.weight, net.action_emb.bias), net.drop.p if train else 0.0)
    H = net.n_heads
    for i, blk in enumerate(blocks):
        sh1, s1, g1, sh2, s2, g2 = mods[i].split(d, dim=-1)
        a = blk.attn
        cfg = (bank, f"dec{i}.qkv", f"dec{i}.o", H, True, a.attn_dropout.p if train else 0.0, a.resid_dropout.p if train else 0.0)
        x = AttnBranch.apply(x, None, sh1, s1, g1, cfg, blk.ln_1.weight, blk.ln_1.bias, *_attn_params(a))
        a = blk.cross_att
        cfg = (bank, f"dec{i}.xq", f"dec{i}.xo", H, True, a.attn_dropout.p if train else 0.0, a.resid_dropout.p if train else 0.0)
        x = AttnBranch.apply(x, kvs[i], None, None, None, cfg, blk.ln3.weight, blk.ln3.bias, *_xattn_params(a))
        cfg = (bank, f"dec{i}.fc", f"dec{i}.proj", blk.mlp.dropout.p if train else 0.0)
        x = MLPBranch.apply(x, sh2, s2, g2, cfg, blk.ln_2.weight, blk.ln_2.bias, *_mlp_params(blk.mlp))
    x = FinalNorm.apply(x, net.decoder.ln.weight, net.decoder.ln.bias)
    return NarrowOut.apply(x, net.action_pred.weight, net.action_pred.bias)


def forward_train(net, states, actions, goals, sigma):
    ctx = encode_train(net, states, goals)
    net.latent_encoder_emb = ctx
    return decode_train(net, ctx, T._c(actions.float()), sigma)
