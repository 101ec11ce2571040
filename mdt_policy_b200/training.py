"""Training path: GCDenoiser.loss with gradients (mdt/models/edm_diffusion/score_wrappers.py:45-63 ->
mdtv_transformer.py:208-236 -> transformer_blocks.py Block / ConditionedBlock / Attention / MLP / LayerNorm).

The graph is held by torch.autograd (plumbing: saved tensors, gradient accumulation over the residual stream and the
shared sigma embedding), but every node is one of the autograd Functions below whose forward AND backward are the
exact-fp32 CUDA kernels of libmdtb200.so (C ABI "training primitives": mdtb200_op_*).  Semantics are the reference's
train-mode forward: dropout on the attention probabilities (attn_pdrop), after both c_proj projections (resid_pdrop,
mlp_pdrop) and on the action embedding (embed_pdrob), goal masking (goal_drop).  Dropout masks come from a counter-based
hash of (seed, element index) -- one base seed per forward pass is drawn from torch's CPU generator and every dropout site derives its
own from it, so torch.manual_seed makes a step reproducible -- and cannot equal the reference's Philox stream: with p = 0 gradients match the reference to fp32
rounding, with p > 0 the check is statistical (tests/test_gpu_training.py).
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch
from torch.autograd import Function

from . import _lib

ACT_GELU, ACT_MISH, ACT_SILU = 1, 2, 3

# Linear layers (forward, input gradient, weight gradient) run on the tcgen05 tensor cores with split-bf16 ("bf16x3") operands
# whenever the shapes allow it (reduce and output-column dims multiples of 64); MDTB200_TRAIN_TC=0 forces the exact-fp32
# CUDA-core GEMMs everywhere (the 7-wide action embedding / output head always use them).
USE_TC = os.environ.get("MDTB200_TRAIN_TC", "1") != "0"
# One autograd node per residual branch with operand-emitting kernels (training_fused.py) whenever the architecture allows it
# (embed_dim multiple of 128 and <= 512, projection dims multiples of 64); MDTB200_TRAIN_FUSED=0 keeps the per-primitive graph below.
USE_FUSED = USE_TC and os.environ.get("MDTB200_TRAIN_FUSED", "1") != "0"


def _fused(net):
    if not USE_FUSED or not net.action_emb.weight.is_cuda:     # CPU tensors: the primitives below raise "no CPU fallback"
        return None
    ok = net.__dict__.get("_train_fused_ok")
    from . import training_fused
    if ok is None:
        ok = net.__dict__["_train_fused_ok"] = training_fused.eligible(net)
    return training_fused if ok else None


def _gemm(mode, A, B, bias, Cout, M, N, K):
    lib = _lib.load()
    ok = USE_TC and K % 64 == 0 and ((mode == 0 and N % 64 == 0) or (mode == 1 and N % 64 == 0) or mode == 2) and M * N * K >= (1 << 22)
    if ok:
        scratch = torch.empty(lib.mdtb200_op_gemm_tc_scratch(mode, M, N, K), dtype=torch.bfloat16, device=A.device)
        _chk(lib.mdtb200_op_gemm_tc(mode, _p(A), _p(B), _p(bias), _p(Cout), M, N, K, _p(scratch), _stream(A)), "op_gemm_tc")
    else:
        _chk(lib.mdtb200_op_gemm(mode, _p(A), _p(B), _p(bias), _p(Cout), M, N, K, 0, _stream(A)), "op_gemm")


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream(t):
    # raw handle of torch's current stream on the tensor's device (torch.cuda.current_stream() builds a Stream object: ~7 us per call,
    # 300+ calls per training step)
    return _lib.current_stream_ptr(t.device.index)


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {_lib.load().mdtb200_last_error(None).decode()}")


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _require_cuda(t):
    if not t.is_cuda:
        raise RuntimeError("mdt_policy_b200 training path runs only on CUDA tensors; there is no CPU fallback")


def _colsum(src2d):
    M, Cn = src2d.shape
    out = torch.empty(Cn, dtype=torch.float32, device=src2d.device)
    scratch = torch.empty(((M + 63) // 64) * Cn, dtype=torch.float32, device=src2d.device)
    _chk(_lib.load().mdtb200_op_colsum(_p(src2d), _p(out), _p(scratch), M, Cn, 0, _stream(src2d)), "op_colsum")
    return out


def _group_sum(src2d, G, T):
    Cn = src2d.shape[1]
    out = torch.empty(G, Cn, dtype=torch.float32, device=src2d.device)
    _chk(_lib.load().mdtb200_op_group_sum(_p(src2d), _p(out), G, T, Cn, 0, _stream(src2d)), "op_group_sum")
    return out


class Linear(Function):
    """y = x W^T + b  (nn.Linear);  dx = dy W, dW = dy^T x, db = colsum(dy)"""

    @staticmethod
    def forward(ctx, x, w, b):
        _require_cuda(x)
        K, N = w.shape[1], w.shape[0]
        x2 = _c(x).reshape(-1, K)
        y = torch.empty(x2.shape[0], N, dtype=torch.float32, device=x.device)
        _gemm(0, x2, w, b, y, x2.shape[0], N, K)
        ctx.save_for_backward(x2, w)
        ctx.has_bias = b is not None
        ctx.xshape = x.shape
        return y.view(*x.shape[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, w = ctx.saved_tensors
        lib = _lib.load()
        N, K = w.shape
        M = x2.shape[0]
        dy2 = _c(dy).reshape(M, N)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, K, dtype=torch.float32, device=dy.device)
            _gemm(1, dy2, w, None, dx, M, N, K)
            dx = dx.view(ctx.xshape)
        if ctx.needs_input_grad[1]:
            dw = torch.empty(N, K, dtype=torch.float32, device=dy.device)
            _gemm(2, dy2, x2, None, dw, M, N, K)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = _colsum(dy2)
        return dx, dw, db


class Act(Function):
    @staticmethod
    def forward(ctx, x, kind):
        _require_cuda(x)
        x = _c(x)
        y = torch.empty_like(x)
        _chk(_lib.load().mdtb200_op_act(_p(x), None, _p(y), x.numel(), kind, _stream(x)), "op_act fwd")
        ctx.save_for_backward(x)
        ctx.kind = kind
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = _c(dy)
        dx = torch.empty_like(x)
        _chk(_lib.load().mdtb200_op_act(_p(x), _p(dy), _p(dx), x.numel(), ctx.kind, _stream(x)), "op_act bwd")
        return dx, None


class LayerNormMod(Function):
    """y = shift + LN(x; w, b) * scale (shift/scale (B, d) per sample, or None for a plain LayerNorm)."""

    @staticmethod
    def forward(ctx, x, w, b, shift, scale):
        _require_cuda(x)
        B, T, d = x.shape
        x = _c(x)
        shift = _c(shift) if shift is not None else None
        scale = _c(scale) if scale is not None else None
        y = torch.empty_like(x)
        _chk(_lib.load().mdtb200_op_ln_fwd(_p(x), _p(w), _p(b), _p(shift), _p(scale), d, T, B * T, d, _p(y), _stream(x)), "op_ln_fwd")
        ctx.save_for_backward(x, w, b, scale)
        ctx.mod = shift is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, b, scale = ctx.saved_tensors
        B, T, d = x.shape
        M = B * T
        dy = _c(dy)
        dx = torch.empty_like(x)
        t_dw = torch.empty(M, d, dtype=torch.float32, device=x.device)
        t_db = torch.empty(M, d, dtype=torch.float32, device=x.device) if b is not None else None
        t_dsc = torch.empty(M, d, dtype=torch.float32, device=x.device) if ctx.mod else None
        _chk(_lib.load().mdtb200_op_ln_bwd(_p(x), _p(dy), _p(w), _p(b), _p(scale), d, T, M, d, _p(dx), _p(t_dw), _p(t_db), _p(t_dsc),
                                           _stream(x)), "op_ln_bwd")
        dw = _colsum(t_dw)
        db = _colsum(t_db) if b is not None else None
        dshift = _group_sum(dy.view(M, d), B, T) if ctx.mod else None
        dscale = _group_sum(t_dsc, B, T) if ctx.mod else None
        return dx, dw, db, dshift, dscale


_seed_pool = [0, 0]       # [base seed of the current forward pass, sites served from it]


def begin_forward_seeds():
    """one draw from torch's CPU generator per forward pass (torch.manual_seed -> reproducible); the dropout sites of that pass take
    base + k * golden-ratio increments (the mask hash mixes seed and element index, so consecutive seeds are independent streams)"""
    _seed_pool[0] = int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())
    _seed_pool[1] = 0


def _new_seed() -> int:
    """63-bit seed of the next dropout site: derived from the forward pass' base seed when one is active (forward_train), else one
    draw from torch's CPU generator per call"""
    if _seed_pool[0]:
        _seed_pool[1] += 1
        return (_seed_pool[0] + _seed_pool[1] * 0x9E3779B97F4A7C15) & (2 ** 62 - 1)
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def end_forward_seeds():
    _seed_pool[0] = 0


class Dropout(Function):
    """inverted dropout; the mask is a pure function of (seed, element index), regenerated in backward"""

    @staticmethod
    def forward(ctx, x, p, seed):
        _require_cuda(x)
        x = _c(x)
        y = torch.empty_like(x)
        _chk(_lib.load().mdtb200_op_dropout(_p(x), _p(y), x.numel(), float(p), seed, _stream(x)), "op_dropout")
        ctx.cfg = (float(p), seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        p, seed = ctx.cfg
        dy = _c(dy)
        dx = torch.empty_like(dy)
        _chk(_lib.load().mdtb200_op_dropout(_p(dy), _p(dx), dy.numel(), p, seed, _stream(dy)), "op_dropout bwd")
        return dx, None, None


def _drop(x, p):
    return Dropout.apply(x, p, _new_seed()) if p > 0 else x


class Attention(Function):
    """softmax(q k^T / sqrt(hd) + mask) v over heads, optional dropout on the probabilities; q (B,Tq,D), k/v (B,Tk,D)."""

    @staticmethod
    def forward(ctx, q, k, v, n_heads, causal, p_drop, seed):
        _require_cuda(q)
        q, k, v = _c(q), _c(k), _c(v)
        B, Tq, D = q.shape
        Tk = k.shape[1]
        y = torch.empty_like(q)
        _chk(_lib.load().mdtb200_op_attn_fwd(_p(q), D, _p(k), _p(v), D, _p(y), D, B, n_heads, D // n_heads, Tq, Tk, int(causal),
                                             float(p_drop), seed, _stream(q)), "op_attn_fwd")
        ctx.save_for_backward(q, k, v)
        ctx.cfg = (n_heads, int(causal), float(p_drop), seed)
        return y

    @staticmethod
    def backward(ctx, dy):
        q, k, v = ctx.saved_tensors
        H, causal, p_drop, seed = ctx.cfg
        B, Tq, D = q.shape
        Tk = k.shape[1]
        dy = _c(dy)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        _chk(_lib.load().mdtb200_op_attn_bwd(_p(q), D, _p(k), _p(v), D, _p(dy), D, _p(dq), D, _p(dk), _p(dv), D, B, H, D // H, Tq, Tk,
                                             causal, p_drop, seed, _stream(q)), "op_attn_bwd")
        return dq, dk, dv, None, None, None, None


class GateResidual(Function):
    """out = x + gate * f  (gate (B, d) broadcast over tokens, or None: out = x + f)."""

    @staticmethod
    def forward(ctx, x, f, gate):
        _require_cuda(x)
        x, f = _c(x), _c(f)
        gate = _c(gate) if gate is not None else None
        B, T, d = x.shape
        out = torch.empty_like(x)
        _chk(_lib.load().mdtb200_op_gate_res(_p(x), _p(f), _p(gate), _p(out), B * T, d, T, _stream(x)), "op_gate_res")
        ctx.save_for_backward(f, gate)
        return out

    @staticmethod
    def backward(ctx, dout):
        f, gate = ctx.saved_tensors
        B, T, d = f.shape
        dout = _c(dout)
        df = torch.empty_like(f)
        prod = torch.empty_like(f) if gate is not None else None
        _chk(_lib.load().mdtb200_op_gate_res_bwd(_p(dout), _p(f), _p(gate), _p(df), _p(prod), B * T, d, T, _stream(f)), "op_gate_res_bwd")
        dgate = _group_sum(prod.view(B * T, d), B, T) if gate is not None else None
        return dout, df, dgate


# ------------------------------------------------------------------------------------------------------------------
# train-mode forward of the score network, composed of the Functions above (same structure as the reference modules)

def _lin(mod, x):
    return Linear.apply(x, mod.weight, mod.bias)


def _attention_block(att, n_heads, x, kv_src, causal, train):
    # transformer_blocks.py:119-158: separate query / key / value projections with bias, SDPA with dropout_p = attn_pdrop in
    # train mode, c_proj (no bias) followed by resid_dropout
    q, k, v = _lin(att.query, x), _lin(att.key, kv_src), _lin(att.value, kv_src)
    pa = att.attn_dropout.p if train else 0.0
    y = Attention.apply(q, k, v, n_heads, causal, pa, _new_seed() if pa > 0 else 0)
    return _drop(_lin(att.c_proj, y), att.resid_dropout.p if train else 0.0)


def _mlp(m, x, train):
    # transformer_blocks.py:175-180
    return _drop(_lin(m.c_proj, Act.apply(_lin(m.c_fc, x), ACT_GELU)), m.dropout.p if train else 0.0)


def _check_no_dropout(net):
    """kept for callers that need the dropout-free semantics (gradient parity tests)"""
    bad = [n for n, m in net.named_modules() if isinstance(m, torch.nn.Dropout) and m.p > 0 and n != "proprio_drop"]
    if net.training and (bad or net.cond_mask_prob > 0):
        raise RuntimeError(f"dropout is active ({bad[:3]}...)")


def _mask_goal(net, goals):
    # mask_cond, mdtv_transformer.py:302-309 (input preparation: element-wise Bernoulli mask on the goal embedding)
    if net.training and net.cond_mask_prob > 0:
        return goals * (1.0 - torch.bernoulli(torch.full_like(goals, net.cond_mask_prob)))
    return goals


def _prep_goal_train(net, states, goals):
    """preprocess_goals in train mode (mdtv_transformer.py:246-258 / mdt_transformer.py:293-305): (B, d) -> (B, 1, d), first goal token
    when one goal per state step is given, vision-goal truncation to obs_dim, then the Bernoulli goal mask."""
    if goals.dim() == 2:
        goals = goals[:, None, :]
    if goals.shape[1] == net._states_length(states) and net.goal_seq_len == 1:
        goals = goals[:, :1, :]
    if goals.shape[-1] == 2 * net.obs_dim:
        goals = goals[:, :, : net.obs_dim]
    return _mask_goal(net, goals)


def encode_train(net, states, goals):
    """forward_enc_only with gradients (mdtv_transformer.py:213-222 / mdt_transformer.py:211-229)."""
    fused = _fused(net)
    if fused is not None:
        return fused.encode_train(net, states, goals)
    goals = _prep_goal_train(net, states, goals)
    train = net.training
    lang = net.use_modality_encoder and states.get("modality") == "lang" and net._variant == "mdtv"
    gm = net.lang_emb if lang else net.goal_emb
    g = _lin(gm[2], Act.apply(_lin(gm[0], goals[:, :1, :].float()), ACT_GELU))
    if net._variant == "mdtv":
        s = _lin(net.tok_emb, states["state_images"].float())
    else:
        st = _lin(net.tok_emb, states["static"].float())
        gr = _lin(net.incam_embed, states["gripper"].float())
        s = torch.cat((st, gr), dim=1)
        # apply_position_embeddings, mdt_transformer.py:317-323: embedding dropout (embed_pdrob) after the position embedding
        g = _drop(g + net.pos_emb[:, : net.goal_seq_len, :], net.drop.p if train else 0.0)
        s = _drop(s + net.pos_emb[:, net.goal_seq_len: net.goal_seq_len + 1, :], net.drop.p if train else 0.0)
    x = torch.cat([g, s], dim=1).contiguous()
    for blk in net.encoder.blocks:          # Block.forward, transformer_blocks.py:209-214
        a = LayerNormMod.apply(x, blk.ln_1.weight, blk.ln_1.bias, None, None)
        x = GateResidual.apply(x, _attention_block(blk.attn, net.n_heads, a, a, False, train), None)
        a = LayerNormMod.apply(x, blk.ln_2.weight, blk.ln_2.bias, None, None)
        x = GateResidual.apply(x, _mlp(blk.mlp, a, train), None)
    return LayerNormMod.apply(x, net.encoder.ln.weight, net.encoder.ln.bias, None, None)


def decode_train(net, ctx, actions, sigma):
    """forward_dec_only with gradients (mdtv_transformer.py:224-236, ConditionedBlock :292-309)."""
    fused = _fused(net)
    if fused is not None:
        return fused.decode_train(net, ctx, actions, sigma)
    d = net.embed_dim
    half = d // 2
    e = sigma.float().log() / 4
    f = torch.exp(torch.arange(half, device=sigma.device, dtype=torch.float32) * -(math.log(10000) / (half - 1)))
    ang = e[:, None] * f[None, :]
    pe = torch.cat((ang.sin(), ang.cos()), dim=-1)                       # no parameters, no gradient: input preparation
    c = _lin(net.sigma_emb[3], Act.apply(_lin(net.sigma_emb[1], pe), ACT_MISH))            # (B, d)
    sc_c = Act.apply(c, ACT_SILU)
    train = net.training
    x = _drop(_lin(net.action_emb, actions), net.drop.p if train else 0.0)
    for blk in net.decoder.blocks:
        mod = _lin(blk.adaLN_zero.modulation[1], sc_c)                  # (B, 6d)
        sh1, s1, g1, sh2, s2, g2 = mod.chunk(6, dim=-1)
        a = LayerNormMod.apply(x, blk.ln_1.weight, blk.ln_1.bias, sh1, s1)
        x = GateResidual.apply(x, _attention_block(blk.attn, net.n_heads, a, a, True, train), g1)
        a = LayerNormMod.apply(x, blk.ln3.weight, blk.ln3.bias, None, None)
        x = GateResidual.apply(x, _attention_block(blk.cross_att, net.n_heads, a, ctx, True, train), None)
        a = LayerNormMod.apply(x, blk.ln_2.weight, blk.ln_2.bias, sh2, s2)
        x = GateResidual.apply(x, _mlp(blk.mlp, a, train), g2)
    x = LayerNormMod.apply(x, net.decoder.ln.weight, net.decoder.ln.bias, None, None)
    return _lin(net.action_pred, x)


def forward_train(net, states, actions, goals, sigma):
    if net.training:
        begin_forward_seeds()
    try:
        return _forward_train(net, states, actions, goals, sigma)
    finally:
        end_forward_seeds()


def _forward_train(net, states, actions, goals, sigma):
    fused = _fused(net)
    if fused is not None:
        return fused.forward_train(net, states, actions, goals, sigma)
    ctx = encode_train(net, states, goals)
    net.latent_encoder_emb = ctx
    return decode_train(net, ctx, _c(actions.float()), sigma)
