"""Host-side mirror of the reference score networks.

``MDTVTransformer`` / ``MDTTransformer`` keep the reference's constructor signature, parameter names and
``named_parameters()`` order (mdt/models/networks/mdtv_transformer.py:35-206, mdt_transformer.py:38-190,
networks/transformers/transformer_blocks.py) so that reference checkpoints, the EMA list-zip
(mdt/models/mdtv_agent.py:145-162, mdt/evaluation/utils.py:93-100) and Hydra configs keep working with only
the ``_target_`` strings changed.  The sub-modules below are *parameter containers*: all arithmetic runs in
``libmdtb200.so`` (mdt_policy_b200/csrc) through the C ABI in include/mdtb200.h.  There is no PyTorch/CPU
implementation of the forward pass in this package -- calling it on CPU tensors raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn as nn

from . import _lib

__all__ = ["MDTVTransformer", "MDTTransformer"]


# ------------------------------------------------------------------------------------------------------------------
# parameter containers (same attribute names / registration order as transformer_blocks.py)

class _Placeholder(nn.Module):
    """Keeps nn.Sequential indices aligned with the reference (activation / sinusoidal-embedding slots)."""

    def __init__(self, what: str):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


class _LayerNorm(nn.Module):          # transformer_blocks.py:29-38
    def __init__(self, ndim: int, bias: bool):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(ndim))
        self.bias = nn.Parameter(torch.zeros(ndim)) if bias else None


class _Attention(nn.Module):          # transformer_blocks.py:66-95: key, query, value (bias), c_proj (bias=bias)
    def __init__(self, n_embd: int, bias: bool, attn_pdrop: float, resid_pdrop: float):
        super().__init__()
        self.key = nn.Linear(n_embd, n_embd)
        self.query = nn.Linear(n_embd, n_embd)
        self.value = nn.Linear(n_embd, n_embd)
        self.c_proj = nn.Linear(n_embd, n_embd, bias=bias)
        self.attn_dropout = nn.Dropout(attn_pdrop)
        self.resid_dropout = nn.Dropout(resid_pdrop)


class _MLP(nn.Module):                # transformer_blocks.py:161-173
    def __init__(self, n_embd: int, bias: bool, dropout: float):
        super().__init__()
        self.c_fc = nn.Linear(n_embd, 4 * n_embd, bias=bias)
        self.gelu = _Placeholder("GELU(erf)")
        self.c_proj = nn.Linear(4 * n_embd, n_embd, bias=bias)
        self.dropout = nn.Dropout(dropout)


class _Block(nn.Module):              # transformer_blocks.py:183-207
    def __init__(self, n_embd, bias, attn_pdrop, resid_pdrop, mlp_pdrop, cross: bool):
        super().__init__()
        self.ln_1 = _LayerNorm(n_embd, bias)
        self.attn = _Attention(n_embd, bias, attn_pdrop, resid_pdrop)
        if cross:
            self.cross_att = _Attention(n_embd, bias, attn_pdrop, resid_pdrop)
            self.ln3 = nn.LayerNorm(n_embd)
        self.ln_2 = _LayerNorm(n_embd, bias)
        self.mlp = _MLP(n_embd, bias, mlp_pdrop)


class _AdaLNZero(nn.Module):          # transformer_blocks.py:245-260
    def __init__(self, hidden: int):
        super().__init__()
        self.modulation = nn.Sequential(_Placeholder("SiLU"), nn.Linear(hidden, 6 * hidden, bias=True))


class _ConditionedBlock(_Block):      # transformer_blocks.py:266-290
    def __init__(self, n_embd, bias, attn_pdrop, resid_pdrop, mlp_pdrop):
        super().__init__(n_embd, bias, attn_pdrop, resid_pdrop, mlp_pdrop, cross=True)
        self.adaLN_zero = _AdaLNZero(n_embd)


class _Stack(nn.Module):              # TransformerEncoder :344-380 / TransformerFiLMDecoder :509-569
    def __init__(self, blocks, n_embd, bias):
        super().__init__()
        self.blocks = nn.Sequential(*blocks)
        self.ln = _LayerNorm(n_embd, bias)


def _goal_mlp(goal_dim, d):
    return nn.Sequential(nn.Linear(goal_dim, 2 * d), _Placeholder("GELU(erf)"), nn.Linear(2 * d, d))


def _cond_mlp(in_dim, d, first=None):
    mods = ([first] if first is not None else []) + [nn.Linear(in_dim, 2 * d), _Placeholder("Mish"), nn.Linear(2 * d, d)]
    return nn.Sequential(*mods)


# ------------------------------------------------------------------------------------------------------------------
# engine: one libmdtb200 handle per (module, device)

def _ptr(t: torch.Tensor) -> C.c_void_p:
    return C.c_void_p(t.data_ptr())


def _f32c(t: torch.Tensor, what: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t if t.is_contiguous() else t.contiguous()


class _Engine:
    def __init__(self, owner: "_ScoreNetBase", device: torch.device, max_batch: int):
        self.lib = _lib.load()
        self.device = device
        self.max_batch = int(max_batch)
        self.handle = C.c_void_p()
        self.weights_key = None
        cfg = _lib.MdtConfig(
            abi_version=_lib.ABI_VERSION, variant=_lib.VARIANT[owner._variant], embed_dim=owner.embed_dim,
            n_heads=owner.n_heads, n_enc_layers=owner.n_enc_layers, n_dec_layers=owner.n_dec_layers,
            action_dim=owner.action_dim, action_seq_len=owner.action_seq_len, goal_dim=owner.goal_dim,
            obs_dim=owner.obs_dim, n_state_tokens=owner.n_state_tokens,
            precision=_lib.PRECISION[owner.precision], max_batch=self.max_batch, sigma_data=float(owner.sigma_data))
        with torch.cuda.device(device):
            rc = self.lib.mdtb200_create(C.byref(cfg), C.byref(self.handle))
        if rc != 0:
            raise RuntimeError(f"mdtb200_create failed ({rc}): {self.lib.mdtb200_last_error(None).decode()}")

    def close(self):
        if self.handle:
            self.lib.mdtb200_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def check(self, rc: int, what: str):
        if rc < 0:
            raise RuntimeError(f"{what} failed ({rc}): {self.lib.mdtb200_last_error(self.handle).decode()}")
        return rc

    @property
    def stream(self) -> C.c_void_p:
        return _lib.current_stream_ptr(self.device.index)      # raw handle: no Stream object per call

    def sync_weights(self, owner: nn.Module, prefix: str = "inner_model."):
        # The module tree is static, so walk it once and keep (name, owning _parameters dict, leaf) triples; every call
        # then re-reads the live Parameter objects through the dicts (catches re-assigned Parameters as well as in-place
        # updates: .to(), load_state_dict, EMA swap, optimizer steps bump data_ptr / _version).
        slots = self.__dict__.get("_slots")
        if slots is None:
            slots = []
            for mname, mod in owner.named_modules():
                for leaf in mod._parameters:
                    if mod._parameters[leaf] is not None:
                        slots.append(((mname + "." if mname else "") + leaf, mod._parameters, leaf))
            seen, uniq = set(), []
            for full, d, leaf in slots:              # shared modules (lang_emb is goal_emb without a modality encoder)
                if id(d[leaf]) not in seen:
                    seen.add(id(d[leaf])); uniq.append((full, d, leaf))
            slots = self._slots = uniq
        # cheap per-call check first (the e2e path pays for every microsecond here): identity and version counters of the live
        # Parameter objects; storage moves (.to(), .cuda()) mark the module dirty through _ScoreNetBase._apply
        # (id, version counter and storage address: `p.data = t` keeps id and version but moves the address).  What this cannot see
        # is an in-place write THROUGH `.data` / `.detach()` views (`p.data.copy_(ema)`): those bump no counter by design -- callers
        # doing that must call `model.invalidate_weights()`.
        quick = 0
        for _, d, leaf in slots:
            p = d[leaf]
            quick += id(p) + p._version + p.data_ptr()
        if quick == self.__dict__.get("_quick_key") and not owner.__dict__.get("_weights_dirty", True):
            return
        params = [(full, d[leaf]) for full, d, leaf in slots]
        key = tuple([(p.data_ptr(), p._version) for _, p in params])
        if key == self.weights_key and not owner.__dict__.get("_weights_force", False):
            self._quick_key = quick
            owner.__dict__["_weights_dirty"] = False
            return
        for name, p in params:
            if p.device != self.device:
                raise RuntimeError(f"parameter {name} lives on {p.device}, engine on {self.device}")
            t = p.detach()
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise RuntimeError(f"parameter {name} must be contiguous fp32 (got {t.dtype})")
            self.check(self.lib.mdtb200_bind_weight(self.handle, (prefix + name).encode(), _ptr(t), t.numel()), "bind_weight")
        self.check(self.lib.mdtb200_commit_weights(self.handle, self.stream), "commit_weights")
        self.weights_key = key
        self._quick_key = quick
        owner.__dict__["_weights_dirty"] = False
        owner.__dict__["_weights_force"] = False

    def launch_count(self) -> int:
        return int(self.lib.mdtb200_launch_count(self.handle))

    def gemm_time_us(self, M: int, N: int, K: int, epi: int, iters: int = 200) -> float:
        """average duration of the tensor-core GEMM kernel for one shape (bench.py's per-kernel roofline)"""
        out = C.c_float()
        with torch.cuda.device(self.device):
            self.check(self.lib.mdtb200_debug_gemm_time(self.handle, M, N, K, epi, iters, C.byref(out), self.stream), "debug_gemm_time")
        return float(out.value)

    def debug_buffer(self, name: str, numel: int) -> torch.Tensor:
        out = torch.empty(numel, dtype=torch.float32, device=self.device)
        n = self.check(self.lib.mdtb200_debug_copy(self.handle, name.encode(), _ptr(out), numel, self.stream), "debug_copy")
        return out[:n]


# ------------------------------------------------------------------------------------------------------------------

class _ScoreNetBase(nn.Module):
    _variant = "mdtv"

    # set by subclasses: embed_dim, n_heads, n_enc_layers, n_dec_layers, action_dim, action_seq_len, goal_dim,
    # obs_dim, n_state_tokens, goal_seq_len
    precision = "bf16x3"
    sigma_data = 0.5            # overwritten by GCDenoiser (needed by the fused precondition/sampler kernels)
    max_batch = 256

    def _init_weights(self, module):
        # mdtv_transformer.py:197-206
        if isinstance(module, (nn.Linear, nn.Embedding)):
            torch.nn.init.normal_(module.weight, mean=0.0, std=0.02)
            if isinstance(module, nn.Linear) and module.bias is not None:
                torch.nn.init.zeros_(module.bias)
        elif isinstance(module, nn.LayerNorm):
            torch.nn.init.zeros_(module.bias)
            torch.nn.init.ones_(module.weight)
        elif isinstance(module, _ScoreNetBase):
            torch.nn.init.normal_(module.pos_emb, mean=0.0, std=0.02)

    # -- engine plumbing -------------------------------------------------------------------------------------------
    def _engine(self, device: torch.device, batch: int) -> _Engine:
        if device.type != "cuda":
            raise RuntimeError(
                "mdt_policy_b200 runs the score network only through its sm_100a CUDA library; "
                f"got tensors on '{device}'. There is no CPU fallback.")
        engines = self.__dict__.setdefault("_engines", {})
        dev = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        eng = engines.get(dev)
        if eng is None or batch > eng.max_batch:
            if eng is not None:
                eng.close()
            eng = _Engine(self, dev, max(int(self.max_batch), int(batch)))
            engines[dev] = eng
        eng.sync_weights(self)
        return eng

    def invalidate_weights(self):
        """Forces the next call to re-pack the weights into the CUDA library.  Needed only after in-place writes through
        ``param.data`` / ``param.detach()`` (e.g. ``p.data.copy_(ema_p)``), which PyTorch does not version; ``load_state_dict``,
        optimizer steps, ``.to()`` and ``param.data = tensor`` are detected automatically."""
        self.__dict__["_weights_dirty"] = True
        self.__dict__["_weights_force"] = True

    def __getstate__(self):
        # engines hold ctypes handles of the CUDA library: never copied / pickled -- a copy builds its own handle lazily
        state = self.__dict__.copy()
        state.pop("_engines", None)
        state["_weights_dirty"] = True
        return state

    def get_block_size(self):
        return self.block_size

    def get_params(self):
        return self.parameters()

    def launch_count(self) -> int:
        return sum(e.launch_count() for e in self.__dict__.get("_engines", {}).values())

    # -- input canonicalisation ------------------------------------------------------------------------------------
    def _prep_goal(self, goals: torch.Tensor, states_length: int, uncond: bool) -> torch.Tensor:
        # preprocess_goals, mdtv_transformer.py:246-258 (eval semantics; goal_drop masking is training-only)
        if goals.dim() == 2:
            goals = goals[:, None, :]
        if goals.shape[1] == states_length and self.goal_seq_len == 1:
            goals = goals[:, :1, :]
        if goals.shape[-1] == 2 * self.obs_dim:
            goals = goals[:, :, : self.obs_dim]
        if goals.shape[1] != 1:
            raise NotImplementedError("goal_seq_len > 1 is not supported by the CUDA path (shipped configs use 1)")
        if goals.shape[-1] != self.goal_dim:
            raise ValueError(f"goal feature dim {goals.shape[-1]} != goal_dim {self.goal_dim}")
        if uncond:
            goals = torch.zeros_like(goals)
        return _f32c(goals[:, 0, :], "goal")

    def _prep_state(self, states: dict) -> torch.Tensor:
        raise NotImplementedError

    def _modality(self, states: dict, context_only: bool) -> int:
        lang = self.use_modality_encoder and "modality" in states and states["modality"] == "lang"
        return _lib.MODALITY_LANG if lang else _lib.MODALITY_VIS

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .float() may move parameter storage without touching version counters: force a full weight check
        self.__dict__["_weights_dirty"] = True
        return super()._apply(fn, *args, **kwargs)

    def _check_mode(self):
        if not self.training:
            return
        if self.cond_mask_prob > 0 or any(isinstance(m, nn.Dropout) and m.p > 0 for m in self.modules()):
            raise NotImplementedError(
                "train-mode forward under torch.no_grad(): the inference kernels implement eval semantics only (no dropout / goal "
                "masking); call .eval(), or enable autograd to run the training path (mdt_policy_b200/training.py)")

    # -- reference API ---------------------------------------------------------------------------------------------
    def _encode(self, states, goals, uncond, context_only, want_ctx=True):
        self._check_mode()
        state = self._prep_state(states)
        goal = self._prep_goal(goals, self._states_length(states), bool(uncond))
        B = state.shape[0]
        eng = self._engine(state.device, B)
        ctx = torch.empty(B, 1 + self.n_state_tokens, self.embed_dim, dtype=torch.float32, device=state.device) if want_ctx else None
        with torch.cuda.device(eng.device):
            eng.check(eng.lib.mdtb200_encode(eng.handle, _ptr(goal), _ptr(state), self._modality(states, context_only), B,
                                             _ptr(ctx) if want_ctx else None, eng.stream), "mdtb200_encode")
        if want_ctx:
            self.latent_encoder_emb = ctx
        return eng, ctx

    def _decode(self, eng, actions, sigma, precondition: bool):
        actions = _f32c(actions, "actions")
        B = actions.shape[0]
        if actions.shape[1:] != (self.action_seq_len, self.action_dim):
            raise ValueError(f"actions must be (B, {self.action_seq_len}, {self.action_dim}), got {tuple(actions.shape)}")
        sigma = _f32c(sigma.reshape(-1), "sigma").to(actions.device)
        if sigma.numel() == 1 and B > 1:
            sigma = sigma.expand(B).contiguous()
        if sigma.numel() != B:
            raise ValueError(f"sigma must have {B} entries, got {sigma.numel()}")
        out = torch.empty_like(actions)
        with torch.cuda.device(eng.device):
            eng.check(eng.lib.mdtb200_denoise(eng.handle, _ptr(actions), _ptr(sigma), B, int(precondition), _ptr(out), eng.stream),
                      "mdtb200_denoise")
        return out

    def wants_grad(self, *tensors) -> bool:
        """True when autograd must see this call: gradients enabled and a parameter (or an input) requires them."""
        if not torch.is_grad_enabled():
            return False
        return any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors) or any(p.requires_grad for p in self.parameters())

    def forward(self, states, actions, goals, sigma, uncond: Optional[bool] = False, _precondition: bool = False):
        if self.wants_grad(actions, goals):
            # training path: exact-fp32 CUDA kernels with hand-written backward, composed under torch.autograd (training.py)
            from . import training
            if _precondition:
                raise RuntimeError("internal: the fused preconditioner has no backward; GCDenoiser applies the scalings itself")
            if uncond:
                goals = torch.zeros_like(goals)
            return training.forward_train(self, states, actions, goals, sigma)
        if actions.shape[0] == 0:       # empty batch: nothing to launch (the reference's tensor ops return an empty result as well)
            return actions.new_zeros(actions.shape, dtype=torch.float32)
        eng, _ = self._encode(states, goals, uncond, context_only=False)
        return self._decode(eng, actions, sigma, _precondition)

    def forward_enc_only(self, states, actions=None, goals=None, sigma=None, uncond: Optional[bool] = False):
        if self.wants_grad(goals) and self._variant == "mdtv":
            from . import training
            return training.encode_train(self, states, torch.zeros_like(goals) if uncond else goals)
        _, ctx = self._encode(states, goals, uncond, context_only=True)
        return ctx

    def forward_dec_only(self, context, actions, sigma):
        self._check_mode()
        context = _f32c(context, "context")
        B = context.shape[0]
        eng = self._engine(context.device, B)
        with torch.cuda.device(eng.device):
            eng.check(eng.lib.mdtb200_set_context(eng.handle, _ptr(context), B, eng.stream), "mdtb200_set_context")
        return self._decode(eng, actions, sigma, False)

    # fused sampling (MDTVAgent.sample_loop for ddim / euler / heun / dpmpp_2m)
    def sample(self, states, x_t, goals, sigmas, sampler: str = "ddim", uncond: bool = False, noise=None, eta: float = 1.0):
        """noise / eta: euler_ancestral only -- noise (n_steps, B, T, A) holds the standard-normal draw of every step (zeros where the
        reference draws nothing), produced by the caller so that the RNG stream is the reference's (gc_sampling.sample_euler_ancestral)."""
        self._check_mode()
        if x_t.shape[0] == 0:
            return x_t.new_zeros(x_t.shape, dtype=torch.float32)
        state = self._prep_state(states)
        goal = self._prep_goal(goals, self._states_length(states), bool(uncond))
        B = state.shape[0]
        eng = self._engine(state.device, B)
        x = _f32c(x_t, "x_t").clone()
        sig = _f32c(sigmas, "sigmas").to(state.device)
        n_steps = sig.numel() - 1
        if sampler == "euler_ancestral":
            if noise is None or tuple(noise.shape) != (n_steps,) + tuple(x.shape):
                raise ValueError("euler_ancestral needs noise of shape (n_steps, B, T, A)")
            noise = _f32c(noise, "noise")
            with torch.cuda.device(eng.device):
                eng.check(eng.lib.mdtb200_sample_ancestral(eng.handle, _ptr(sig), n_steps, _ptr(goal), _ptr(state), self._modality(states, False), B,
                                                           _ptr(x), _ptr(noise), float(eta), eng.stream), "mdtb200_sample_ancestral")
            return x
        with torch.cuda.device(eng.device):
            eng.check(eng.lib.mdtb200_sample(eng.handle, _lib.SAMPLER[sampler], _ptr(sig), n_steps, _ptr(goal), _ptr(state),
                                             self._modality(states, False), B, _ptr(x), eng.stream), "mdtb200_sample")
        return x


    def sample_host(self, state_host, x_inout_host, goal_host, sigmas_host, sampler: str, modality_lang: bool, device):
        """Whole sampling call on HOST buffers through mdtb200_sample_host: the library stages state/goal/x_T/sigmas to the
        device, replays the graph and copies the actions back into `x_inout_host` (synchronous on return)."""
        self._check_mode()
        B = state_host.shape[0]
        eng = self._engine(device, B)
        n_steps = sigmas_host.numel() - 1
        with torch.cuda.device(eng.device):
            eng.check(eng.lib.mdtb200_sample_host(eng.handle, _lib.SAMPLER[sampler], _ptr(sigmas_host), n_steps, _ptr(goal_host),
                                                  _ptr(state_host), _lib.MODALITY_LANG if modality_lang else _lib.MODALITY_VIS, B,
                                                  _ptr(x_inout_host), eng.stream), "mdtb200_sample_host")
        return x_inout_host


class MDTVTransformer(_ScoreNetBase):
    """Drop-in for mdt.models.networks.mdtv_transformer.MDTVTransformer (same ctor arguments)."""

    _variant = "mdtv"

    def __init__(self, obs_dim: int, goal_dim: int, device: str, n_obs_token: int, goal_conditioned: bool,
                 action_dim: int, proprio_dim: int, embed_dim: int, embed_pdrob: float, attn_pdrop: float,
                 resid_pdrop: float, mlp_pdrop: float, n_dec_layers: int, n_enc_layers: int, n_heads: int,
                 goal_seq_len: int, obs_seq_len: int, action_seq_len: int, goal_drop: float = 0.1, bias=False,
                 use_mlp_goal: bool = False, use_abs_pos_emb: bool = True, use_rot_embed: bool = False,
                 rotary_xpos: bool = False, linear_output: bool = True, use_ada_conditioning: bool = False,
                 use_noise_encoder: bool = False, use_modality_encoder: bool = False,
                 precision: str = "bf16x3", max_batch: int = 256):
        super().__init__()
        _require_shipped_config(goal_conditioned, use_mlp_goal, use_rot_embed, linear_output, use_ada_conditioning,
                                use_noise_encoder, goal_seq_len, precision)
        self.device = device
        self.goal_conditioned = goal_conditioned
        self.obs_dim, self.goal_dim, self.embed_dim, self.action_dim = obs_dim, goal_dim, embed_dim, action_dim
        self.n_obs_token = n_obs_token
        self.n_heads, self.n_enc_layers, self.n_dec_layers = n_heads, n_enc_layers, n_dec_layers
        self.use_ada_conditioning = use_ada_conditioning
        self.use_modality_encoder = use_modality_encoder
        self.action_seq_len, self.goal_seq_len, self.obs_seq_len = action_seq_len, goal_seq_len, obs_seq_len
        self.n_state_tokens = obs_seq_len * n_obs_token
        self.block_size = goal_seq_len + action_seq_len + obs_seq_len * n_obs_token + 2
        seq_size = goal_seq_len + obs_seq_len * n_obs_token + action_seq_len
        self.cond_mask_prob = goal_drop
        self.use_rot_embed, self.use_abs_pos_emb = use_rot_embed, use_abs_pos_emb
        self.precision, self.max_batch = precision, max_batch
        self.latent_encoder_emb = None

        # registration order == mdtv_transformer.py:78-179 (pos_emb, a direct Parameter, lists first)
        self.tok_emb = nn.Linear(obs_dim, embed_dim)
        self.goal_emb = _goal_mlp(goal_dim, embed_dim)
        self.lang_emb = _goal_mlp(goal_dim, embed_dim) if use_modality_encoder else self.goal_emb
        self.pos_emb = nn.Parameter(torch.zeros(1, seq_size, embed_dim))   # never applied in MDT-V
        self.drop = nn.Dropout(embed_pdrob)
        self.proprio_drop = nn.Dropout(0.5)   # dead branch under the shipped configs (mdtv_transformer.py:106)
        self.encoder = _Stack([_Block(embed_dim, bias, attn_pdrop, resid_pdrop, mlp_pdrop, cross=False)
                               for _ in range(n_enc_layers)], embed_dim, bias)
        self.decoder = _Stack([_ConditionedBlock(embed_dim, bias, attn_pdrop, resid_pdrop, mlp_pdrop)
                               for _ in range(n_dec_layers)], embed_dim, bias)
        self.proprio_emb = _cond_mlp(proprio_dim, embed_dim)
        self.sigma_emb = _cond_mlp(embed_dim, embed_dim, first=_Placeholder("SinusoidalPosEmb"))
        self.action_emb = nn.Linear(action_dim, embed_dim)
        self.action_pred = nn.Linear(embed_dim, action_dim)
        self.apply(self._init_weights)

    def _states_length(self, states):
        return states["state_images"].size(1)

    def _prep_state(self, states):
        if "state_obs" in states:
            raise NotImplementedError("proprioceptive 'state_obs' tokens are not used by the shipped MDT-V config")
        s = states["state_images"]
        if s.dim() != 3 or s.shape[1] != self.n_state_tokens or s.shape[2] != self.obs_dim:
            raise ValueError(f"state_images must be (B, {self.n_state_tokens}, {self.obs_dim}), got {tuple(s.shape)}")
        return _f32c(s, "state_images")


class MDTTransformer(_ScoreNetBase):
    """Drop-in for mdt.models.networks.mdt_transformer.MDTTransformer (ResNet variant, d=512, 2 state tokens)."""

    _variant = "mdt"

    def __init__(self, obs_dim: int, goal_dim: int, device: str, goal_conditioned: bool, action_dim: int,
                 embed_dim: int, embed_pdrob: float, attn_pdrop: float, resid_pdrop: float, mlp_pdrop: float,
                 n_dec_layers: int, n_enc_layers: int, n_heads: int, goal_seq_len: int, obs_seq_len: int,
                 action_seq_len: int, proprio_dim: Optional[int] = None, goal_drop: float = 0.1, bias=False,
                 use_abs_pos_emb: bool = True, use_rot_embed: bool = False, rotary_xpos: bool = False,
                 linear_output: bool = True, use_ada_conditioning: bool = False, use_noise_encoder: bool = False,
                 latent_is_decoder: bool = False, use_modality_encoder: bool = False, use_mlp_goal: bool = False,
                 precision: str = "bf16x3", max_batch: int = 256):
        super().__init__()
        _require_shipped_config(goal_conditioned, use_mlp_goal, use_rot_embed, linear_output, use_ada_conditioning,
                                use_noise_encoder, goal_seq_len, precision)
        if not use_abs_pos_emb:
            raise NotImplementedError("MDTTransformer without use_abs_pos_emb is not supported")
        self.device = device
        self.goal_conditioned = goal_conditioned
        self.obs_dim, self.goal_dim, self.embed_dim, self.action_dim = obs_dim, goal_dim, embed_dim, action_dim
        self.n_heads, self.n_enc_layers, self.n_dec_layers = n_heads, n_enc_layers, n_dec_layers
        self.use_ada_conditioning = use_ada_conditioning
        self.use_modality_encoder = use_modality_encoder
        self.proprio_dim = proprio_dim
        self.latent_is_decoder = latent_is_decoder
        self.action_seq_len, self.goal_seq_len, self.obs_seq_len = action_seq_len, goal_seq_len, obs_seq_len
        self.n_state_tokens = 2
        self.block_size = goal_seq_len + action_seq_len + obs_seq_len + 1
        seq_size = goal_seq_len + action_seq_len
        self.cond_mask_prob = goal_drop
        self.use_rot_embed, self.use_abs_pos_emb = use_rot_embed, use_abs_pos_emb
        self.precision, self.max_batch = precision, max_batch
        self.latent_encoder_emb = None

        # registration order == mdt_transformer.py:82-190
        self.tok_emb = nn.Linear(obs_dim, embed_dim)
        self.incam_embed = nn.Linear(obs_dim, embed_dim)
        self.pos_emb = nn.Parameter(torch.zeros(1, seq_size, embed_dim))
        self.drop = nn.Dropout(embed_pdrob)
        self.goal_emb = _goal_mlp(goal_dim, embed_dim)
        self.lang_emb = _goal_mlp(goal_dim, embed_dim) if use_modality_encoder else self.goal_emb
        self.encoder = _Stack([_Block(embed_dim, bias, attn_pdrop, resid_pdrop, mlp_pdrop, cross=False)
                               for _ in range(n_enc_layers)], embed_dim, bias)
        self.decoder = _Stack([_ConditionedBlock(embed_dim, bias, attn_pdrop, resid_pdrop, mlp_pdrop)
                               for _ in range(n_dec_layers)], embed_dim, bias)
        self.sigma_emb = _cond_mlp(embed_dim, embed_dim, first=_Placeholder("SinusoidalPosEmb"))
        self.action_emb = nn.Linear(action_dim, embed_dim)
        self.action_pred = nn.Linear(embed_dim, action_dim)
        if proprio_dim is not None:      # registered last in the reference (mdt_transformer.py:179-184); unused
            self.proprio_emb = _cond_mlp(proprio_dim, embed_dim)
        self.apply(self._init_weights)

    def _states_length(self, states):
        return 1   # enc_only_forward passes t = 1 (mdt_transformer.py:213)

    def _modality(self, states, context_only):
        # forward() -> enc_only_forward always uses goal_emb (mdt_transformer.py:215); only the
        # forward_enc_only entry (GCDenoiser.forward_context_only) is modality-aware (:262, :280-285)
        return super()._modality(states, context_only) if context_only else _lib.MODALITY_VIS

    def _prep_state(self, states):
        st, gr = states["static"].to(torch.float32), states["gripper"].to(torch.float32)
        if st.shape[1:] != (1, self.obs_dim) or gr.shape[1:] != (1, self.obs_dim):
            raise ValueError(f"static/gripper must be (B, 1, {self.obs_dim})")
        return torch.cat((st, gr), dim=1).contiguous()


def _require_shipped_config(goal_conditioned, use_mlp_goal, use_rot_embed, linear_output, use_ada_conditioning,
                            use_noise_encoder, goal_seq_len, precision):
    bad = []
    if not goal_conditioned: bad.append("goal_conditioned=False")
    if not use_mlp_goal: bad.append("use_mlp_goal=False")
    if use_rot_embed: bad.append("use_rot_embed=True")
    if not linear_output: bad.append("linear_output=False")
    if not use_ada_conditioning: bad.append("use_ada_conditioning=False")
    if use_noise_encoder: bad.append("use_noise_encoder=True")
    if goal_seq_len != 1: bad.append(f"goal_seq_len={goal_seq_len}")
    if bad:
        raise NotImplementedError(
            "the CUDA score network implements the shipped configs (conf/model/model/mdt*_transformer.yaml); "
            "unsupported option(s): " + ", ".join(bad))
    if precision not in _lib.PRECISION:
        raise ValueError(f"precision must be one of {sorted(_lib.PRECISION)}, got {precision!r}")
