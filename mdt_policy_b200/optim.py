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
``capturable=True`` keeps the step count on the device and the pointer table in pinned memory, so ``step()`` can be recorded into a
CUDA graph (``GraphedTrainStep`` below captures loss forward + backward + this step as ONE graph: the ~1500 launches of a training
step are issued by the GPU front end instead of by Python).  CUDA fp32 parameters only (there is no CPU fallback in this package).
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
                 ema_start_step=0, capturable=False):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        if ema_decay is not None and not (0 <= ema_decay <= 1):
            raise ValueError("EMA decay value must be between 0 and 1")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.ema_decay, self.ema_schedule, self.ema_start_step = ema_decay, ema_schedule, ema_start_step
        if capturable and (ema_schedule is not None or ema_start_step):
            raise ValueError("capturable=True supports a constant ema_decay only (host-side schedules are frozen inside a CUDA graph)")
        self.capturable = capturable
        self._cap = {}           # group index -> {"step": device int32, "pinned": host table, "table": device table}
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
            step_ptr = None
            if self.capturable:
                # device-side step count (incremented by a capturable op) + pinned pointer table copied asynchronously: everything
                # below is legal inside torch.cuda.graph(); on replay the same table is copied again and the count keeps advancing
                cap = self._cap.get(gi)
                if cap is None or cap["pinned"].shape[0] != len(rows):
                    cap = self._cap[gi] = {"step": torch.zeros(1, dtype=torch.int32, device=dev),
                                           "pinned": torch.empty(len(rows), 7, dtype=torch.int64).pin_memory(),
                                           "table": torch.empty(len(rows), 7, dtype=torch.int64, device=dev)}
                    cap["step"].fill_(max(self.state[p]["step"] for p in params) - 1)
                cap["pinned"].copy_(torch.tensor(rows, dtype=torch.int64))
                cap["table"].copy_(cap["pinned"], non_blocking=True)
                cap["step"].add_(1)
                cap["grads"] = keep
                table, step_ptr = cap["table"], C.c_void_p(cap["step"].data_ptr())
            else:
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
                    group["weight_decay"], float(decay), int(use_ema), step_ptr,
                    _lib.current_stream_ptr(dev.index))
            if rc != 0:
                raise RuntimeError(f"mdtb200_op_adamw_ema failed ({rc}): {self._lib.mdtb200_last_error(None).decode()}")
            # the kernel writes the parameters through raw pointers: tell autograd / the inference engine's staleness check
            # (networks._Engine.sync_weights keys on Parameter._version) that they changed
            torch.autograd.graph.increment_version(params)
        return loss


class GraphedTrainStep:
    """One training step -- ``GCDenoiser.loss`` forward, backward and the fused AdamW + EMA update -- captured into a CUDA graph.

        opt = FusedAdamWEMA(model.parameters(), ..., capturable=True)
        step = GraphedTrainStep(model, opt, state_images, goal, actions, noise, sigma)     # warm-up + capture on example tensors
        loss = step(state_images, goal, actions, noise, sigma)                             # copy-in, one graph replay

    The eager step issues ~1500 small kernels from Python (two thirds of its wall time is host sequencing); the replayed graph
    issues them from the GPU front end.  Dropout masks stay fresh on every replay through a device-side RNG epoch counter that the
    graph increments (``mdtb200_op_set_seed_epoch``); the optimizer's step count lives on the device as well.  Shapes are fixed at
    capture time (as for any CUDA graph); parameters without a gradient at capture time stay frozen.  As with any torch.cuda.graph
    capture of a backward pass, no autograd graph of an earlier eager step may still be alive (e.g. a kept `loss` tensor): its
    AccumulateGrad nodes are bound to the default stream and the capture fails with cudaErrorStreamCaptureImplicit.
    """

    def __init__(self, model, optimizer: FusedAdamWEMA, state_images, goal, actions, noise, sigma, modality="lang", warmup=3,
                 process_group=None, data_parallel=None):
        """data_parallel (default: torch.distributed is initialised with world_size > 1): gradients are averaged over the process
        group between two graphs -- [loss + backward + flatten] | one NCCL all-reduce of the flat gradient buffer | [unflatten + fused
        AdamW/EMA] -- which is DistributedDataParallel's arithmetic with a single bucket; parameters are broadcast from rank 0 first."""
        if not optimizer.capturable:
            raise ValueError("GraphedTrainStep needs FusedAdamWEMA(capturable=True)")
        import torch.distributed as dist
        dev = actions.device
        self.model, self.opt, self.modality = model, optimizer, modality
        self.pg = process_group
        self.dp = (dist.is_available() and dist.is_initialized() and dist.get_world_size(process_group) > 1) if data_parallel is None else bool(data_parallel)
        self.comm = True                  # set False to time the step without the all-reduce (exposed communication = difference)
        self._bump = [p for group in optimizer.param_groups for p in group["params"]][:1]
        self.static = [t.detach().clone() for t in (state_images, goal, actions, noise, sigma)]
        lib = _lib.load()
        self.epoch = torch.zeros(1, dtype=torch.int64, device=dev)
        if lib.mdtb200_op_set_seed_epoch(C.c_void_p(self.epoch.data_ptr())) != 0:
            raise RuntimeError("mdtb200_op_set_seed_epoch failed: " + lib.mdtb200_last_error(None).decode())
        if self.dp:
            for p in model.parameters():
                dist.broadcast(p.data, 0, group=process_group)
        # The score network stashes its encoder output WITH its autograd graph (`latent_encoder_emb`, as the reference does for the
        # CLA loss).  After eager steps on the default stream that graph keeps the encoder parameters' AccumulateGrad nodes alive and
        # bound to the legacy stream, which a capturing stream may not synchronise with (cudaErrorStreamCaptureImplicit): drop it.
        inner = getattr(model, "inner_model", None)
        if inner is not None and getattr(inner, "latent_encoder_emb", None) is not None:
            inner.latent_encoder_emb = None
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                if self.dp:
                    self._allreduce_eager()
                self.opt.step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import train_ops
        train_ops.presize_side_workspace(dev)      # the captured step runs its weight-gradient GEMMs on a side stream
        self.graph = torch.cuda.CUDAGraph()
        optimizer.zero_grad(set_to_none=True)
        if not self.dp:
            with torch.cuda.graph(self.graph):
                self.loss = self._fwd_bwd()
                self.opt.step()
            return
        with torch.cuda.graph(self.graph):
            self.loss = self._fwd_bwd()
            self.grads = [p.grad for group in optimizer.param_groups for p in group["params"] if p.grad is not None]
            self.flat = torch.cat([g.reshape(-1) for g in self.grads])
        self.graph2 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph2, pool=self.graph.pool()):
            torch._foreach_copy_(self.grads, [v.view_as(g) for v, g in zip(self.flat.split([g.numel() for g in self.grads]), self.grads)])
            self.opt.step()

    def _fwd_bwd(self):
        s, g, a, n, sig = self.static
        self.opt.zero_grad(set_to_none=True)
        self.epoch.add_(1)
        loss, _ = self.model.loss({"state_images": s, "modality": self.modality}, a, g, n, sig)
        loss.backward()
        return loss.detach()

    def _allreduce_eager(self):
        import torch.distributed as dist
        grads = [p.grad for group in self.opt.param_groups for p in group["params"] if p.grad is not None]
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.pg)
        torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)])

    def __call__(self, state_images, goal, actions, noise, sigma):
        for dst, src in zip(self.static, (state_images, goal, actions, noise, sigma)):
            dst.copy_(src, non_blocking=True)
        self.graph.replay()
        if self.dp:
            if self.comm:
                import torch.distributed as dist
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.pg)
            self.graph2.replay()
        torch.autograd.graph.increment_version(self._bump)      # a replay updates the parameters without touching their version counters
        return self.loss

    def close(self):
        """unregisters the RNG epoch counter (eager steps go back to host-drawn seeds only)"""
        _lib.load().mdtb200_op_set_seed_epoch(None)
