"""Fused AdamW + EMA optimizer step on the CUDA library (``mdtb200_op_adamw_ema``).

The reference trains with ``torch.optim.AdamW`` (mdt/models/mdtv_agent.py:164-199) and keeps an exponential moving average of
the weights in a Lightning callback that walks the state dict tensor by tensor (mdt/callbacks/ema.py:106-126).  Here both are
ONE kernel launch per parameter group: a multi-tensor table {param, grad, exp_avg, exp_avg_sq, ema} is updated in place, with
``torch.optim.AdamW``'s exact update order and the callback's ``ema -= (1 - decay) * (ema - w)``.

    opt = FusedAdamWEMA(model.parameters(), lr=1e-4, betas=(0.9, 0.9), weight_decay=0.05, ema_decay=0.999)
    loss.backward(); opt.step()
    with opt.swap_ema():          # evaluate with the averaged weights (EMACallback.replace_model_weights / restore_original_weights)
        ...

``ema_decay=None`` disables the average; ``ema_schedule=(inv_gamma, power, min, max)`` reproduces ``EMACallback.get_decay``.
CUDA fp32 parameters only (there is no CPU fallback in this package).
"""
from __future__ import annotations

import contextlib
import ctypes as C
import math
import struct

import torch

from . import _lib

_CHUNK = 4096


class FusedAdamWEMA(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, ema_decay=None, ema_schedule=None,
                 ema_start_step=0):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        if ema_decay is not None and not (0 <= ema_decay <= 1):
            raise ValueError("EMA decay value must be between 0 and 1")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.ema_decay, self.ema_schedule, self.ema_start_step = ema_decay, ema_schedule, ema_start_step
        self._lib = _lib.load()
        self._maps = {}          # (group index, tensor sizes with grads) -> device block map

    # mdt/callbacks/ema.py:84-92
    def _decay(self, optimization_step):
        if self.ema_schedule is None:
            return self.ema_decay
        inv_gamma, power, lo, hi = self.ema_schedule
        step = max(0, optimization_step - self.ema_start_step - 1)
        return max(min(1 - (1 + step / inv_gamma) ** -power, hi), lo)

    @property
    def has_ema(self):
        return self.ema_decay is not None or self.ema_schedule is not None

    def ema_parameters(self):
        """EMA tensors in parameter order (parameters that never received a gradient report their live value)."""
        return [self.state[p]["ema"] if "ema" in self.state.get(p, {}) else p.detach() for g in self.param_groups for p in g["params"]]

    @contextlib.contextmanager
    def swap_ema(self):
        """Temporarily loads the averaged weights into the parameters (validation with EMA weights)."""
        params = [p for g in self.param_groups for p in g["params"] if "ema" in self.state.get(p, {})]
        backup = [p.detach().clone() for p in params]
        with torch.no_grad():
            for p in params:
                p.copy_(self.state[p]["ema"])
        try:
            yield
        finally:
            with torch.no_grad():
                for p, b in zip(params, backup):
                    p.copy_(b)

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            dev = params[0].device
            for p in params:
                if p.device != dev or p.dtype != torch.float32 or not p.is_cuda or not p.is_contiguous():
                    raise RuntimeError("FusedAdamWEMA needs contiguous fp32 CUDA parameters on one device (no CPU fallback)")
                if p.grad.is_sparse or p.grad.dtype != torch.float32:
                    raise RuntimeError("FusedAdamWEMA needs dense fp32 gradients")
                st = self.state[p]
                if not st:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.contiguous_format)
                    if self.has_ema:
                        st["ema"] = p.detach().clone()
            group["step"] = group.get("step", 0) + 1          # optimizer steps taken (EMA schedule / start step)
            beta1, beta2 = group["betas"]
            use_ema = self.has_ema and group["step"] >= self.ema_start_step
            decay = self._decay(group["step"]) if use_ema else 0.0
            # pointer table (gradients may be re-allocated by zero_grad(set_to_none=True): rebuilt every step, 56 bytes per tensor);
            # bias corrections use the per-parameter step count like torch.optim.AdamW (a parameter without gradient does not step)
            rows, keep = [], []
            for p in params:
                st = self.state[p]
                st["step"] = st.get("step", 0) + 1
                t = st["step"]
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                keep.append(g)
                hyper = struct.unpack("q", struct.pack("ff", group["lr"] / (1 - beta1 ** t), math.sqrt(1 - beta2 ** t)))[0]
                rows.append((p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                             st["ema"].data_ptr() if use_ema else 0, p.numel(), hyper))
            table = torch.tensor(rows, dtype=torch.int64).to(dev, non_blocking=True)
            key = (gi, tuple(p.numel() for p in params))
            blocks = self._maps.get(key)
            if blocks is None or blocks.device != dev:
                bm = [(i, c) for i, p in enumerate(params) for c in range((p.numel() + _CHUNK - 1) // _CHUNK)]
                blocks = torch.tensor(bm, dtype=torch.int32).to(dev)
                self._maps = {key: blocks}
            with torch.cuda.device(dev):
                rc = self._lib.mdtb200_op_adamw_ema(
                    C.c_void_p(table.data_ptr()), C.c_void_p(blocks.data_ptr()), blocks.shape[0], group["lr"], beta1, beta2, group["eps"],
                    group["weight_decay"], float(decay), int(use_ema),
                    C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
            if rc != 0:
                raise RuntimeError(f"mdtb200_op_adamw_ema failed ({rc}): {self._lib.mdtb200_last_error(None).decode()}")
        return loss
