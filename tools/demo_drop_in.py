"""End-to-end tour of the drop-in surface on one B200 (synthetic weights / inputs; prints one line per stage):

  1. the score model exactly as Hydra would build it from conf/model/model/mdtv_transformer.yaml with the two `_target_` strings changed,
  2. the PerceiverResampler producing the 3 state tokens from (B, 1, 392, 384) Voltron-like features,
  3. DenoiseAgent.denoise_actions (10-step DDIM, one CUDA graph) and the other fused samplers,
  4. a vectorised rollout of 256 environments (BatchedRollout: per-env action-chunk caches, one batched sampling call per step),
  5. training: loss.backward() + FusedAdamWEMA eagerly, then the same step replayed as a CUDA graph (GraphedTrainStep).

    python tools/demo_drop_in.py
"""
import math, os, sys, time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mdt_policy_b200 import GCDenoiser, DenoiseAgent, PerceiverResampler, BatchedRollout, utils as U      # noqa: E402
from mdt_policy_b200.optim import FusedAdamWEMA, GraphedTrainStep                                         # noqa: E402
from mdt_policy_b200.rollout import SyntheticVecEnv                                                        # noqa: E402
from mdt_policy_b200.synthetic import synthetic_state_dict, synthetic_inputs                               # noqa: E402

dev = torch.device("cuda")
B = 256


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3, out


# 1. the object hydra.utils.instantiate(cfg.model) would return (mdtv_transformer.yaml, `_recursive_: false`)
inner = dict(_target_="mdt_policy_b200.networks.MDTVTransformer", action_dim=7, obs_dim=384, goal_dim=512, proprio_dim=8,
             goal_conditioned=True, embed_dim=384, n_dec_layers=4, n_enc_layers=4, n_obs_token=3, goal_seq_len=1, obs_seq_len=1,
             action_seq_len=10, embed_pdrob=0, goal_drop=0, attn_pdrop=0.3, resid_pdrop=0.1, mlp_pdrop=0.05, n_heads=8, device="cuda",
             linear_output=True, use_rot_embed=False, use_abs_pos_emb=True, bias=False, use_ada_conditioning=True,
             use_noise_encoder=False, use_modality_encoder=True, use_mlp_goal=True)
model = GCDenoiser(inner, sigma_data=0.5)
model.load_state_dict(synthetic_state_dict([(n, p.shape) for n, p in model.named_parameters()], 12, "trained"))      # a checkpoint's state_dict
model = model.to(dev).eval()
print(f"1. GCDenoiser(MDTVTransformer): {sum(p.numel() for p in model.parameters()) / 1e6:.1f} M parameters, reference names / order")

# 2. state tokens from Voltron-like features
perc = PerceiverResampler(dim=384, depth=6, dim_head=64, heads=8, num_latents=3, num_time_embeds=1, max_batch=B).to(dev)
feats = torch.randn(B, 1, 392, 384, device=dev)
ms, tokens = timed(lambda: perc(feats))
state_images = tokens                                                  # (B, 3, 384): perceptual_emb['state_images'], mdtv_agent.py:392-403
print(f"2. PerceiverResampler (B, 1, 392, 384) -> {tuple(tokens.shape)}: {ms:.2f} ms")

# 3. sampling
goal = torch.randn(B, 1, 512, device=dev)
agent = DenoiseAgent(model, device=dev, num_sampling_steps=10, sampler_type="ddim", noise_scheduler="exponential", sigma_min=0.001, sigma_max=80.0)
state = {"state_images": state_images, "modality": "lang"}
for sampler in ("ddim", "euler", "heun", "dpmpp_2m", "euler_ancestral"):
    agent.sampler_type = sampler                                      # plain attributes, as mdt_evaluate.py:248-256 assigns them
    ms, actions = timed(lambda: agent.denoise_actions(None, state, goal, inference=True))
    print(f"3. denoise_actions[{sampler:15s}] -> {tuple(actions.shape)}: {ms:.2f} ms per 10-step call ({10 * (2 if sampler == 'heun' else 1) / ms * 1e3:.0f} evaluations/s)")
agent.sampler_type = "ddim"

# 4. 256 environments, action chunks of 10, episodes of random length
env = SyntheticVecEnv(B, device=dev, seed=1)
roll = BatchedRollout(agent, B)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(60):
    s_img, g = env.observe()
    actions = roll.step(s_img, g)
    roll.reset(env.step(actions))
torch.cuda.synchronize()
print(f"4. BatchedRollout: 60 simulator steps x {B} envs in {(time.perf_counter() - t0) * 1e3:.0f} ms, {roll.sampling_calls} sampling calls "
      f"({roll.samples_planned} chunks planned, {int(env.episodes.sum())} episodes finished)")

# 5. training
model.train()
inp = {k: v.to(dev) for k, v in synthetic_inputs(512, seed=31).items()}
sig = U.rand_log_logistic((512,), loc=math.log(0.5), scale=0.5, min_value=0.001, max_value=80.0, device="cpu").to(dev)
opt = FusedAdamWEMA(model.parameters(), lr=1e-4, betas=(0.9, 0.9), weight_decay=0.05, ema_decay=0.999, capturable=True)


def eager_step():
    opt.zero_grad(set_to_none=True)
    loss, _ = model.loss({"state_images": inp["state_images"], "modality": "lang"}, inp["actions"], inp["goal"], inp["noise"], sig)
    loss.backward()
    opt.step()
    return loss.detach()          # do not keep the autograd graph alive: its AccumulateGrad nodes are bound to this (default) stream


ms, loss = timed(eager_step, 10)
print(f"5a. eager step (loss.backward() + FusedAdamWEMA.step()), batch 512: {ms:.2f} ms, loss {float(loss):.4f}")
args = (inp["state_images"], inp["goal"], inp["actions"], inp["noise"], sig)
gstep = GraphedTrainStep(model, opt, *args)
ms, loss = timed(lambda: gstep(*args), 20)
print(f"5b. GraphedTrainStep replay: {ms:.2f} ms per step = {5120 / ms:.0f} k action-tokens/s, loss {float(loss):.4f}")
gstep.close()
model.eval()
with torch.no_grad():
    a = agent.denoise_actions(None, state, goal, inference=True)      # the inference engine re-commits the updated weights by itself
print(f"6. sampling with the trained weights: finite = {bool(torch.isfinite(a).all())}")
