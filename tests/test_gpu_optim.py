"""Fused AdamW + EMA kernel (mdtb200_op_adamw_ema / mdt_policy_b200.optim.FusedAdamWEMA) against torch.optim.AdamW and the
reference EMA callback's formula (mdt/callbacks/ema.py:117-126); a 2-GPU DistributedDataParallel step of the training path
(BASELINE config 5: NCCL gradient all-reduce is the path's only collective)."""
import os

import pytest
import torch

from tests import helpers as H
from mdt_policy_b200.optim import FusedAdamWEMA, GraphedTrainStep
from mdt_policy_b200.synthetic import synthetic_inputs

pytestmark = pytest.mark.gpu


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    shapes = [(384, 384), (1536, 384), (7,), (384,), (3, 5, 11), (1,), (2304, 384), (13,)]
    return [torch.randn(s, generator=g).cuda().requires_grad_() for s in shapes]


def test_fused_adamw_matches_torch_and_ema_formula():
    ref_p, fus_p = _params(0), _params(0)
    ref = torch.optim.AdamW(ref_p, lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)
    fus = FusedAdamWEMA(fus_p, lr=3e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05, ema_decay=0.99)
    ema = [p.detach().clone() for p in ref_p]
    g = torch.Generator().manual_seed(1)
    for step in range(6):
        for a, b in zip(ref_p, fus_p):
            grad = torch.randn(a.shape, generator=g).cuda() * (10.0 if step == 3 else 1.0)
            a.grad = grad.clone()
            b.grad = grad.clone()
        if step == 4:                        # a parameter without a gradient is left untouched (and keeps its state)
            ref_p[2].grad = None
            fus_p[2].grad = None
        ref.step(); fus.step()
        for e, p in zip(ema, ref_p):         # mdt/callbacks/ema.py:117-126
            if p.grad is not None:
                diff = e - p.detach()
                diff.mul_(1.0 - 0.99)
                e.sub_(diff)
        for a, b in zip(ref_p, fus_p):
            assert (a - b).abs().max() <= 2e-6 * max(1.0, float(a.abs().max())), (step, tuple(a.shape))
    for e, f in zip(ema, fus.ema_parameters()):
        assert (e - f).abs().max() <= 2e-6 * max(1.0, float(e.abs().max()))
    before = [p.detach().clone() for p in fus_p]
    with fus.swap_ema():
        for p, e in zip(fus_p, ema):
            assert (p - e).abs().max() <= 2e-6 * max(1.0, float(e.abs().max()))
    for p, b in zip(fus_p, before):
        assert torch.equal(p, b)
    sd = fus.state_dict()                    # standard Optimizer state dict round trip
    fus2 = FusedAdamWEMA(_params(0), lr=3e-3, betas=(0.9, 0.95), weight_decay=0.05, ema_decay=0.99)
    fus2.load_state_dict(sd)
    assert len(fus2.state) == len(fus.state)


def test_ema_schedule_matches_callback_get_decay():
    opt = FusedAdamWEMA(_params(0), ema_schedule=(1.0, 2 / 3, 0.0, 0.9999))
    for t in (1, 2, 10, 1000):
        step = max(0, t - 0 - 1)
        assert abs(opt._decay(t) - max(min(1 - (1 + step / 1.0) ** -(2 / 3), 0.9999), 0.0)) < 1e-12


def test_training_step_with_fused_optimizer_reduces_loss():
    model = H.build_product(H.mdtv_inner_cfg(2, 2, attn_pdrop=0.0, resid_pdrop=0.0, mlp_pdrop=0.0), 15, "trained").train()
    opt = FusedAdamWEMA(model.parameters(), lr=2e-4, betas=(0.9, 0.9), weight_decay=0.05, ema_decay=0.9)
    inp = {k: v.cuda() for k, v in synthetic_inputs(16, seed=25).items()}
    state = {"state_images": inp["state_images"], "modality": "lang"}
    sigma = torch.exp(torch.linspace(3.0, -4.0, 16)).cuda()
    losses = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        loss, _ = model.loss(state, inp["actions"], inp["goal"], inp["noise"], sigma)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses
    # the inference engine re-packs the updated weights; EMA weights differ from the live ones
    with opt.swap_ema(), torch.no_grad():
        model.eval()
        out_ema = model(state, inp["x_T"], inp["goal"], sigma).clone()
    with torch.no_grad():
        out_live = model(state, inp["x_T"], inp["goal"], sigma)
    assert torch.isfinite(out_ema).all() and (out_ema - out_live).abs().max() > 0


def _ddp_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world)
    try:
        model = H.build_product(H.mdtv_inner_cfg(1, 1, attn_pdrop=0.0, resid_pdrop=0.0, mlp_pdrop=0.0), 15, "trained", device=f"cuda:{rank}").train()

        class LossModule(torch.nn.Module):
            def __init__(self, m):
                super().__init__()
                self.m = m

            def forward(self, s, g, a, n, sig):
                return self.m.loss({"state_images": s, "modality": "lang"}, a, g, n, sig)[0]

        net = torch.nn.parallel.DistributedDataParallel(LossModule(model), device_ids=[rank], find_unused_parameters=True)
        inp = {k: v.cuda(rank) for k, v in synthetic_inputs(8, seed=40 + rank).items()}
        sig = torch.exp(torch.linspace(2.0, -3.0, 8)).cuda(rank)
        args = (inp["state_images"], inp["goal"], inp["actions"], inp["noise"], sig)
        with net.no_sync():                                  # local gradients (no all-reduce)
            net(*args).backward()
        local = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
        model.zero_grad(set_to_none=True)
        net(*args).backward()                                # bucketed NCCL all-reduce
        worst = 0.0
        for n, p in model.named_parameters():
            if p.grad is None:
                continue
            mean = local[n].clone()
            dist.all_reduce(mean)
            mean /= world
            worst = max(worst, float((p.grad - mean).abs().max() / (mean.abs().max() + 1e-12)))
        opt = FusedAdamWEMA(model.parameters(), lr=1e-4, betas=(0.9, 0.9), weight_decay=0.05, ema_decay=0.999)
        opt.step()
        chk = torch.stack([p.detach().double().sum() for p in model.parameters()]).sum().reshape(1)
        both = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        ret[rank] = (worst, float(both[0]), float(both[1]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_ddp_two_gpu_gradients_are_rank_means_and_weights_stay_in_sync():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ddp_worker, args=(2, 29533, ret), nprocs=2, join=True)
    for rank in (0, 1):
        worst, c0, c1 = ret[rank]
        assert worst < 1e-4, worst            # DDP gradient == mean of the ranks' local gradients
        assert c0 == c1                        # identical weights on both ranks after the fused optimizer step


def test_inference_engine_sees_fused_optimizer_and_graph_replay_updates():
    """The fused optimizer kernel and a CUDA-graph replay write the parameters through raw pointers; the inference engine (weight arena
    packed at first use) must still notice: eval loss (inference kernels) == train-path loss (dropout 0) after every kind of update."""
    from mdt_policy_b200.optim import GraphedTrainStep
    model = H.build_product(H.mdtv_inner_cfg(1, 1, attn_pdrop=0.0, resid_pdrop=0.0, mlp_pdrop=0.0), 95, "trained")
    inp = {k: v.cuda() for k, v in synthetic_inputs(32, seed=96).items()}
    sigma = torch.exp(torch.linspace(2.0, -3.0, 32)).cuda()
    state = {"state_images": inp["state_images"], "modality": "lang"}
    args = (inp["state_images"], inp["goal"], inp["actions"], inp["noise"], sigma)

    def both():
        model.eval()
        with torch.no_grad():
            l_inf = float(model.loss(state, inp["actions"], inp["goal"], inp["noise"], sigma)[0])     # inference engine
        model.train()
        l_tr = float(model.loss(state, inp["actions"], inp["goal"], inp["noise"], sigma)[0])          # training kernels, live parameters
        return l_inf, l_tr

    l0, t0 = both()                                   # packs the weight arena of the inference engine
    assert abs(l0 - t0) < 1e-3 * max(1.0, t0)
    opt = FusedAdamWEMA(model.parameters(), lr=1e-3, betas=(0.9, 0.9), weight_decay=0.05, ema_decay=0.99, capturable=True)
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        model.loss(state, inp["actions"], inp["goal"], inp["noise"], sigma)[0].backward()
        opt.step()
    l1, t1 = both()
    assert t1 < t0 and abs(l1 - t1) < 1e-3 * max(1.0, t1), (l0, l1, t1)       # eager fused step: engine re-synced
    step = GraphedTrainStep(model, opt, *args)
    for _ in range(3):
        step(*args)
    l2, t2 = both()
    step.close()
    assert t2 < t1 and abs(l2 - t2) < 1e-3 * max(1.0, t2), (l1, l2, t2)       # graph replays: engine re-synced


def _graphed_dp_worker(rank, world, port, ret):
    """data-parallel GraphedTrainStep: [fwd + bwd] graph | one NCCL all-reduce | [fused AdamW/EMA] graph, vs an eager loop that averages
    the two ranks' gradients by hand (dropout off so both are deterministic)"""
    import torch.distributed as dist
    from mdt_policy_b200.optim import GraphedTrainStep
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        cfg = H.mdtv_inner_cfg(2, 2, attn_pdrop=0.0, resid_pdrop=0.0, mlp_pdrop=0.0)
        inp = {k: v.cuda(rank) for k, v in synthetic_inputs(16, seed=60 + rank).items()}
        sig = torch.exp(torch.linspace(2.0, -3.0, 16)).cuda(rank)
        args = (inp["state_images"], inp["goal"], inp["actions"], inp["noise"], sig)
        # eager reference: 3 warm-up steps + the 3 replays below, gradients averaged over the ranks by hand
        model = H.build_product(cfg, 15, "trained", device=f"cuda:{rank}").train()
        opt = FusedAdamWEMA(model.parameters(), lr=2e-4, betas=(0.9, 0.9), weight_decay=0.05, ema_decay=0.99)
        for _ in range(3 + 3):
            opt.zero_grad(set_to_none=True)
            loss, _ = model.loss({"state_images": args[0], "modality": "lang"}, args[2], args[1], args[3], args[4])
            loss.backward()
            for p in model.parameters():
                if p.grad is not None:
                    dist.all_reduce(p.grad)
                    p.grad /= world
            opt.step()
        want = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        model2 = H.build_product(cfg, 15, "trained", device=f"cuda:{rank}").train()
        opt2 = FusedAdamWEMA(model2.parameters(), lr=2e-4, betas=(0.9, 0.9), weight_decay=0.05, ema_decay=0.99, capturable=True)
        step = GraphedTrainStep(model2, opt2, *args)
        assert step.dp
        for _ in range(3):
            step(*args)
        torch.cuda.synchronize()
        got = torch.cat([p.detach().reshape(-1) for p in model2.parameters()])
        chk = got.double().sum().reshape(1)
        both = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        ret[rank] = (float((got - want).abs().max()), float(want.abs().max()), float(both[0]), float(both[1]))
        step.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_graphed_data_parallel_step_two_gpus_matches_hand_averaged_eager_steps():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_graphed_dp_worker, args=(2, 29535, ret), nprocs=2, join=True)
    for rank in (0, 1):
        err, scale, c0, c1 = ret[rank]
        assert err < 2e-5 * max(1.0, scale), (err, scale)     # same weights as the eager loop (different summation order in the flat all-reduce)
        assert c0 == c1                                        # and bit-identical across the ranks


def test_graphed_train_step_matches_eager_steps_and_redraws_dropout():
    """CUDA-graph replay of (loss fwd + bwd + fused AdamW/EMA): without dropout the replayed steps reproduce the eager steps'
    losses; with dropout every replay draws new masks (device-side RNG epoch) and the loss still goes down."""
    import copy
    inp = {k: v.cuda() for k, v in synthetic_inputs(32, seed=25).items()}
    sigma = torch.exp(torch.linspace(3.0, -4.0, 32)).cuda()
    args = (inp["state_images"], inp["goal"], inp["actions"], inp["noise"], sigma)

    def make(drop):
        p = dict(attn_pdrop=0.3, resid_pdrop=0.1, mlp_pdrop=0.05) if drop else dict(attn_pdrop=0.0, resid_pdrop=0.0, mlp_pdrop=0.0)
        return H.build_product(H.mdtv_inner_cfg(2, 2, **p), 15, "trained").train()

    # eager reference: 3 warm-up steps (the capture helper runs them too) + 4 steps
    model = make(False)
    opt = FusedAdamWEMA(model.parameters(), lr=2e-4, betas=(0.9, 0.9), weight_decay=0.05, ema_decay=0.99)
    eager = []
    for _ in range(3 + 1 + 4):               # warm-up, the captured (not executed) step does not count, 4 replays ... see below
        opt.zero_grad(set_to_none=True)
        loss, _ = model.loss({"state_images": args[0], "modality": "lang"}, args[2], args[1], args[3], args[4])
        loss.backward(); opt.step()
        eager.append(float(loss))
    model2 = make(False)
    opt2 = FusedAdamWEMA(model2.parameters(), lr=2e-4, betas=(0.9, 0.9), weight_decay=0.05, ema_decay=0.99, capturable=True)
    step = GraphedTrainStep(model2, opt2, *args)
    graphed = [float(step(*args)) for _ in range(4)]
    step.close()
    # capture does not execute: the replays are steps 4..7 of the eager sequence (after the 3 warm-up steps)
    for g, e in zip(graphed, eager[3:7]):
        assert abs(g - e) <= 2e-5 * max(1.0, abs(e)), (graphed, eager)
    for a, b in zip(model.parameters(), model2.parameters()):
        pass
    # dropout on: consecutive replays on the same batch must differ (fresh masks) and stay finite
    model3 = make(True)
    opt3 = FusedAdamWEMA(model3.parameters(), lr=0.0, betas=(0.9, 0.9), weight_decay=0.0, capturable=True)      # lr 0: weights frozen
    step3 = GraphedTrainStep(model3, opt3, *args)
    l = [float(step3(*args)) for _ in range(4)]
    step3.close()
    assert all(torch.isfinite(torch.tensor(l))) and len({round(v, 6) for v in l}) == 4, l
