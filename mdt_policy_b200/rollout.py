"""Vectorised rollout driver: ``MDTVAgent.step`` (mdt/models/mdtv_agent.py:721-746) for MANY environments at once.

The reference evaluates one environment per process (mdt/rollout/rollout_long_horizon.py:235-269): ``model.reset()``, then
``model.step(obs, goal)`` every simulator step, which re-plans a 10-action chunk every ``multistep`` calls with a B=1 sampling
call.  The CUDA path is built for B=256 calls, so this driver keeps ONE chunk buffer and ONE step counter PER ENVIRONMENT,
collects the environments whose chunk is used up (they drift apart because episodes end at different times), runs one batched
sampling call for exactly those and hands every environment its next action.  Per-environment semantics are identical to the
reference's B=1 loop (tests/test_rollout.py runs both side by side).

The encoders (CLIP text, Voltron / perceiver) are out of scope: the driver is fed their outputs -- ``state_images``
(N, n_state_tokens, obs_dim) and ``latent_goal`` (N, 1, goal_dim) -- for the environments it asks for (``due()``).
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

__all__ = ["BatchedRollout", "SyntheticVecEnv"]


class BatchedRollout:
    def __init__(self, agent, n_envs: int, multistep: Optional[int] = None, bucket: int = 32,
                 noise_fn: Optional[Callable[[torch.Tensor, torch.Tensor], torch.Tensor]] = None):
        """agent: DenoiseAgent.  bucket: sampling batches are padded up to a multiple of this (each distinct batch size is one
        cached CUDA graph in the library).  noise_fn(env_ids, chunk_index) -> x_T (n, T, A): deterministic initial noise
        (tests / reproducible evaluation); default: torch.randn * sigma_max like MDTVAgent.denoise_actions."""
        self.agent = agent
        self.n_envs = int(n_envs)
        self.multistep = int(multistep if multistep is not None else agent.multistep)
        if not (1 <= self.multistep <= agent.act_window_size):
            raise ValueError("multistep must be in [1, act_window_size]")
        self.bucket = max(1, int(bucket))
        self.noise_fn = noise_fn
        dev = agent.device
        self.counter = torch.zeros(self.n_envs, dtype=torch.long)                 # rollout_step_counter of every env (host)
        self.chunks_planned = torch.zeros(self.n_envs, dtype=torch.long)
        self.pred = torch.zeros(self.n_envs, agent.act_window_size, agent.action_dim, device=dev)
        self.sampling_calls = 0
        self.samples_planned = 0

    def reset(self, env_ids=None):
        """model.reset() for the given environments (all when None): their next step() re-plans."""
        ids = torch.arange(self.n_envs) if env_ids is None else torch.as_tensor(env_ids, dtype=torch.long)
        self.counter[ids] = 0

    def due(self) -> torch.Tensor:
        """environments that need a new action chunk at the next step() (their embeddings must be passed to it)"""
        return torch.nonzero(self.counter % self.multistep == 0).flatten()

    @torch.no_grad()
    def step(self, state_images: torch.Tensor, latent_goal: torch.Tensor, modality: str = "lang") -> torch.Tensor:
        """One simulator step for all environments.  state_images / latent_goal: embeddings of the environments in due(), in
        that order (or of all environments: then the due rows are selected).  Returns the actions (n_envs, action_dim)."""
        due = self.due()
        if due.numel():
            if state_images.shape[0] == self.n_envs and due.numel() != self.n_envs:
                idx = due.to(state_images.device)
                state_images, latent_goal = state_images[idx], latent_goal[idx]
            if state_images.shape[0] != due.numel():
                raise ValueError(f"expected embeddings for {due.numel()} due environments, got {state_images.shape[0]}")
            self.pred[due.to(self.pred.device)] = self._plan(due, state_images, latent_goal, modality)
            self.chunks_planned[due] += 1
        dev = self.pred.device
        actions = self.pred[torch.arange(self.n_envs, device=dev), self.counter.to(dev)]
        self.counter += 1
        self.counter[self.counter == self.multistep] = 0
        return actions

    def _plan(self, due, state_images, latent_goal, modality):
        ag = self.agent
        n = due.numel()
        dev = ag.device
        if latent_goal.dim() == 2:
            latent_goal = latent_goal[:, None, :]
        if self.noise_fn is not None:
            x_T = self.noise_fn(due, self.chunks_planned[due]).to(dev)
        else:
            x_T = torch.randn((n, ag.act_window_size, ag.action_dim), device=dev) * ag.sigma_max
        pad = (-n) % self.bucket
        st, gl = state_images.to(dev), latent_goal.to(dev)
        if pad:           # pad with copies of the first row: the batch size stays on a small set of values (graph cache)
            st = torch.cat((st, st[:1].expand(pad, *st.shape[1:])), 0)
            gl = torch.cat((gl, gl[:1].expand(pad, *gl.shape[1:])), 0)
            x_T = torch.cat((x_T, x_T[:1].expand(pad, *x_T.shape[1:])), 0)
        seq = ag.denoise_actions(torch.zeros_like(gl), {"state_images": st, "modality": modality}, gl, inference=True, x_T=x_T)
        self.sampling_calls += 1
        self.samples_planned += n
        return seq[:n]


class SyntheticVecEnv:
    """Stand-in for a vector of simulators + encoders: every step yields fresh synthetic embeddings (what the Voltron / CLIP
    encoders would produce for the new observation); episodes end after a per-environment random number of steps, so the
    environments' chunk phases drift apart exactly like in a real long-horizon evaluation."""

    def __init__(self, n_envs, n_state_tokens=3, obs_dim=384, goal_dim=512, min_len=7, max_len=45, seed=0, device="cpu"):
        self.n, self.shape_s, self.shape_g = n_envs, (n_state_tokens, obs_dim), (1, goal_dim)
        self.g = torch.Generator().manual_seed(seed)
        self.min_len, self.max_len, self.device = min_len, max_len, device
        self.remaining = self._lengths(n_envs)
        self.goal = torch.randn((n_envs,) + self.shape_g, generator=self.g)
        self.episodes = torch.zeros(n_envs, dtype=torch.long)

    def _lengths(self, k):
        return torch.randint(self.min_len, self.max_len + 1, (k,), generator=self.g)

    def observe(self):
        """embeddings of the current observation of every environment"""
        return torch.randn((self.n,) + self.shape_s, generator=self.g).to(self.device), self.goal.to(self.device)

    def step(self, actions):
        """advances every environment; returns the ids whose episode just ended (they were reset to a new task / goal)"""
        self.remaining -= 1
        done = torch.nonzero(self.remaining <= 0).flatten()
        if done.numel():
            self.remaining[done] = self._lengths(done.numel())
            self.goal[done] = torch.randn((done.numel(),) + self.shape_g, generator=self.g)
            self.episodes[done] += 1
        return done
