"""Action chunking (MDTVAgent.step / reset, mdtv_agent.py:680-746) and the vectorised rollout driver against a restated B=1
reference loop (rollout_long_horizon.py:235-269): every environment must receive exactly the actions a private B=1 agent,
reset at the same moments, would have produced."""
import pytest
import torch

from mdt_policy_b200.agent import DenoiseAgent
from mdt_policy_b200.rollout import BatchedRollout, SyntheticVecEnv


class _ToyDenoiser(torch.nn.Module):
    """cheap deterministic stand-in for GCDenoiser with the same call contract (row-wise, so batching cannot change a row)"""

    def forward(self, state, action, goal, sigma):
        s = state["state_images"].mean(dim=(1, 2))[:, None, None]
        g = goal.reshape(goal.shape[0], -1)[:, :7][:, None, :]
        c = 1.0 / (1.0 + sigma.reshape(-1, 1, 1) ** 2)
        return action * c + (1 - c) * torch.tanh(g + s)


def _noise(ids, chunk):
    out = []
    for e, c in zip(ids.tolist(), chunk.tolist()):
        out.append(torch.randn((10, 7), generator=torch.Generator().manual_seed(1000 * e + c)) * 80.0)
    return torch.stack(out)


def _reference_b1_loop(make_agent, env_seed, n_envs, n_steps):
    """rollout_long_horizon.py:235-269 restated for one environment at a time: model.reset() at episode start, model.step()
    every simulator step (the embeddings are replayed from the same synthetic stream the batched run sees)."""
    env = SyntheticVecEnv(n_envs, seed=env_seed)
    agents = [make_agent() for _ in range(n_envs)]
    chunks = [0] * n_envs
    acts = []
    for _ in range(n_steps):
        emb, goal = env.observe()
        row = []
        for e, ag in enumerate(agents):
            if ag.rollout_step_counter % ag.multistep == 0:        # what MDTVAgent.step -> forward -> denoise_actions does, with OUR x_T
                x_T = _noise(torch.tensor([e]), torch.tensor([chunks[e]])).to(ag.device)
                ag.pred_action_seq = ag.denoise_actions(None, {"state_images": emb[e:e + 1].to(ag.device), "modality": "lang"},
                                                        goal[e:e + 1].to(ag.device), inference=True, x_T=x_T)
                chunks[e] += 1
            a = ag.pred_action_seq[0, ag.rollout_step_counter]
            ag.rollout_step_counter = (ag.rollout_step_counter + 1) % ag.multistep
            row.append(a.cpu())
        acts.append(torch.stack(row))
        for e in env.step(None).tolist():
            agents[e].reset()
    return torch.stack(acts)


def _batched_loop(agent, env_seed, n_envs, n_steps, bucket):
    env = SyntheticVecEnv(n_envs, seed=env_seed)
    drv = BatchedRollout(agent, n_envs, bucket=bucket, noise_fn=_noise)
    acts = []
    for _ in range(n_steps):
        emb, goal = env.observe()
        acts.append(drv.step(emb, goal).cpu())
        drv.reset(env.step(None))
    return torch.stack(acts), drv


def test_step_reset_chunking_semantics_b1():
    agent = DenoiseAgent(_ToyDenoiser(), device="cpu", sampler_type="euler", num_sampling_steps=4, multistep=4)
    emb = {"state_images": torch.randn(1, 3, 384)}
    goal = torch.randn(1, 1, 512)
    torch.manual_seed(0)
    a0 = agent.step(emb, goal)
    seq = agent.pred_action_seq.clone()
    assert a0.shape == (7,) and torch.equal(a0, seq[0, 0])
    for k in range(1, 4):                                  # cached actions: no new sampling call
        assert torch.equal(agent.step(emb, goal), seq[0, k]) and torch.equal(agent.pred_action_seq, seq)
    assert agent.rollout_step_counter == 0
    a4 = agent.step(emb, goal)                             # chunk used up -> re-plan (new noise)
    assert not torch.equal(agent.pred_action_seq, seq) and torch.equal(a4, agent.pred_action_seq[0, 0])
    agent.step(emb, goal)
    agent.reset()
    assert agent.rollout_step_counter == 0 and agent.pred_action_seq is None


@pytest.mark.parametrize("multistep", [10, 3])
def test_batched_rollout_equals_per_env_reference_loop_cpu(multistep):
    make = lambda: DenoiseAgent(_ToyDenoiser(), device="cpu", sampler_type="ddim", num_sampling_steps=5, multistep=multistep)
    ref = _reference_b1_loop(make, env_seed=3, n_envs=9, n_steps=60)
    got, drv = _batched_loop(make(), env_seed=3, n_envs=9, n_steps=60, bucket=4)
    assert torch.equal(ref, got)
    assert drv.sampling_calls < 60 and drv.samples_planned == int(drv.chunks_planned.sum())


@pytest.mark.gpu
def test_batched_rollout_equals_per_env_reference_loop_cuda():
    from tests import helpers as H
    model = H.build_product(H.mdtv_inner_cfg(1, 1), 7, "trained")
    make = lambda: DenoiseAgent(model, device="cuda", sampler_type="ddim", num_sampling_steps=3, multistep=5)
    ref = _reference_b1_loop(make, env_seed=5, n_envs=6, n_steps=23)
    got, drv = _batched_loop(make(), env_seed=5, n_envs=6, n_steps=23, bucket=4)
    assert (ref - got).abs().max() < 2e-5 * max(1.0, float(ref.abs().max()))
    assert drv.sampling_calls <= 23
