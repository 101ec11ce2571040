"""Host-side mirror of the reference ``PerceiverResampler`` (mdt/models/networks/transformers/perceiver_resampler.py:86-163).

Same constructor, parameter names and ``named_parameters()`` order as the reference (``latents``, ``time_pos_emb``,
``layers.{l}.0.{norm_media,norm_latents,to_q,to_k,to_v,to_out}``, ``layers.{l}.1.{0,1,3}``, ``norm``), so MDTVAgent checkpoints load
unchanged; the modules are parameter containers only -- ``forward`` runs in ``libmdtb200.so`` (csrc/perceiver.cuh: queries projected
into feature space, tcgen05 GEMMs with head-grouped weights, streaming score / weighted-sum kernels).  CUDA tensors only.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .networks import _Placeholder

__all__ = ["PerceiverResampler"]


class _PerceiverAttentionLayer(nn.Module):        # perceiver_resampler.py:14-30
    def __init__(self, dim: int, dim_head: int = 64, heads: int = 8):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        inner = dim_head * heads
        self.norm_media = nn.LayerNorm(dim)
        self.norm_latents = nn.LayerNorm(dim)
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(dim, inner, bias=False)
        self.to_v = nn.Linear(dim, inner, bias=False)
        self.to_out = nn.Linear(inner, dim, bias=False)


def _feed_forward_layer(dim: int, mult: int = 4):  # transformers/utils.py:16-28 (activation 'gelu')
    return nn.Sequential(nn.LayerNorm(dim), nn.Linear(dim, int(dim * mult), bias=False), _Placeholder("GELU(erf)"),
                         nn.Linear(int(dim * mult), dim, bias=False))


class PerceiverResampler(nn.Module):
    def __init__(self, dim: int, depth: int, dim_head: int = 64, heads: int = 8, num_latents: int = 64, num_time_embeds: int = 4,
                 ff_mult: int = 4, activation: str = "gelu", trainable: bool = True, max_batch: int = 256):
        super().__init__()
        if activation != "gelu":
            raise NotImplementedError("the CUDA perceiver implements the shipped 'gelu' feed-forward only")
        self.dim, self.depth, self.dim_head, self.heads = dim, depth, dim_head, heads
        self.num_queries, self.num_time_embeds, self.ff_mult, self.max_batch = num_latents, num_time_embeds, ff_mult, max_batch
        self.latents = nn.Parameter(torch.randn(num_latents, dim))
        self.time_pos_emb = nn.Parameter(torch.randn(num_time_embeds, 1, dim))
        self.layers = nn.ModuleList([nn.ModuleList([_PerceiverAttentionLayer(dim, dim_head, heads), _feed_forward_layer(dim, ff_mult)])
                                     for _ in range(depth)])
        self.norm = nn.LayerNorm(dim)
        for p in self.parameters():
            p.requires_grad = trainable

    # -- engine plumbing (one library handle per device; weights re-packed when they change) ---------------------------------
    def __getstate__(self):
        state = self.__dict__.copy()
        state.pop("_handles", None)
        return state

    def invalidate_weights(self):
        """after in-place writes through ``param.data`` (not versioned by PyTorch) -- see networks._ScoreNetBase.invalidate_weights"""
        for h in self.__dict__.get("_handles", {}).values():
            h["key"] = None

    def _handle(self, device: torch.device, batch: int, features: int):
        if device.type != "cuda":
            raise RuntimeError(f"mdt_policy_b200.PerceiverResampler runs only through its sm_100a CUDA library; got tensors on '{device}'. "
                               "There is no CPU fallback.")
        lib = _lib.load()
        handles = self.__dict__.setdefault("_handles", {})
        dev = torch.device("cuda", device.index if device.index is not None else torch.cuda.current_device())
        h = handles.get(dev)
        if h is None or batch > h["max_batch"] or features > h["max_features"]:
            if h is not None:
                lib.mdtb200_perceiver_destroy(h["ptr"])
            cfg = _lib.MdtPerceiverConfig(abi_version=_lib.ABI_VERSION, dim=self.dim, depth=self.depth, heads=self.heads, dim_head=self.dim_head,
                                          num_latents=self.num_queries, num_time_embeds=self.num_time_embeds, ff_mult=self.ff_mult,
                                          max_batch=max(int(self.max_batch), int(batch)), max_features=int(features))
            ptr = C.c_void_p()
            with torch.cuda.device(dev):
                rc = lib.mdtb200_perceiver_create(C.byref(cfg), C.byref(ptr))
            if rc != 0:
                raise RuntimeError(f"mdtb200_perceiver_create failed ({rc}): {lib.mdtb200_perceiver_last_error(None).decode()}")
            h = handles[dev] = {"ptr": ptr, "max_batch": cfg.max_batch, "max_features": cfg.max_features, "key": None}
        params = list(self.named_parameters())
        key = tuple((p.data_ptr(), p._version) for _, p in params)
        if key != h["key"]:
            stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
            for name, p in params:
                t = p.detach()
                if t.device != dev or t.dtype != torch.float32 or not t.is_contiguous():
                    raise RuntimeError(f"parameter {name} must be contiguous fp32 on {dev}")
                rc = lib.mdtb200_perceiver_bind_weight(h["ptr"], name.encode(), C.c_void_p(t.data_ptr()), t.numel())
                if rc != 0:
                    raise RuntimeError(f"perceiver bind_weight({name}) failed: {lib.mdtb200_perceiver_last_error(h['ptr']).decode()}")
            with torch.cuda.device(dev):
                rc = lib.mdtb200_perceiver_commit_weights(h["ptr"], stream)
            if rc != 0:
                raise RuntimeError(f"perceiver commit_weights failed ({rc}): {lib.mdtb200_perceiver_last_error(h['ptr']).decode()}")
            h["key"] = key
        return lib, h, dev

    def launch_count(self) -> int:
        lib = _lib.load()
        return sum(int(lib.mdtb200_perceiver_launch_count(h["ptr"])) for h in self.__dict__.get("_handles", {}).values())

    def __del__(self):
        try:
            lib = _lib.load()
            for h in self.__dict__.get("_handles", {}).values():
                lib.mdtb200_perceiver_destroy(h["ptr"])
        except Exception:  # noqa: BLE001
            pass

    @torch.no_grad()
    def forward(self, x_f: torch.Tensor, mask: torch.Tensor = None) -> torch.Tensor:
        """x_f (batch, n_frames, n_features, dim), mask (batch, n_frames) -> (batch, num_latents, dim)   (:126-163)"""
        if x_f.ndim != 4 or x_f.shape[-1] != self.dim:
            raise ValueError(f"x_f must be (B, T, n, {self.dim}), got {tuple(x_f.shape)}")
        B, T, n, _ = x_f.shape
        if T > self.num_time_embeds:
            raise ValueError(f"{T} frames but only {self.num_time_embeds} time embeddings")
        lib, h, dev = self._handle(x_f.device, B, T * n)
        x = x_f.float().contiguous()
        m = None if mask is None else mask.to(device=dev, dtype=torch.float32).contiguous()
        out = torch.empty(B, self.num_queries, self.dim, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            rc = lib.mdtb200_perceiver_forward(h["ptr"], C.c_void_p(x.data_ptr()), None if m is None else C.c_void_p(m.data_ptr()), B, T, n,
                                               C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
        if rc != 0:
            raise RuntimeError(f"mdtb200_perceiver_forward failed ({rc}): {lib.mdtb200_perceiver_last_error(h['ptr']).decode()}")
        return out
