"""The hot-path methods of the reference agents (mdt/models/mdtv_agent.py:508-678, twin in mdt_agent.py),
as a plain object around ``GCDenoiser`` -- no Lightning, no encoders.

``DenoiseAgent`` is what ``MDTVAgent.denoise_actions / sample_loop / get_noise_schedule / diffusion_loss`` do
once the perceptual and goal embeddings exist; the sampler knobs stay plain attributes that evaluation code
overwrites by assignment (mdt/evaluation/mdt_evaluate.py:248-256).
"""
from __future__ import annotations

import math
from functools import partial

import torch

from . import gc_sampling as gcs
from . import utils


class DenoiseAgent:
    def __init__(self, model, device="cuda", act_window_size=10, action_dim=7, num_sampling_steps=10,
                 sampler_type="ddim", noise_scheduler="exponential", sigma_data=0.5, sigma_min=0.001, sigma_max=80.0,
                 sigma_sample_density_type="loglogistic", multistep=10):
        self.model = model
        self.multistep = multistep
        self.reset()
        self.device = torch.device(device)
        self.act_window_size = act_window_size
        self.action_dim = action_dim
        self.num_sampling_steps = num_sampling_steps
        self.sampler_type = sampler_type
        self.noise_scheduler = noise_scheduler
        self.sigma_data, self.sigma_min, self.sigma_max = sigma_data, sigma_min, sigma_max
        self.sigma_sample_density_type = sigma_sample_density_type

    # mdtv_agent.py:680-686
    def reset(self):
        """Call at the beginning of a new rollout."""
        self.plan = None
        self.latent_goal = None
        self.rollout_step_counter = 0
        self.pred_action_seq = None

    # mdtv_agent.py:688-719 -- the part after the (out-of-scope) language / Voltron encoders: the caller passes their outputs
    def forward(self, perceptual_emb, latent_goal):
        if isinstance(perceptual_emb, dict) and "modality" not in perceptual_emb:
            perceptual_emb = dict(perceptual_emb, modality="lang")
        return self.denoise_actions(torch.zeros_like(latent_goal), perceptual_emb, latent_goal, inference=True)

    __call__ = forward

    # mdtv_agent.py:721-746 -- action chunking: a new action sequence every `multistep` calls, cached actions in between
    def step(self, perceptual_emb, latent_goal):
        if self.rollout_step_counter % self.multistep == 0:
            self.pred_action_seq = self(perceptual_emb, latent_goal)
        current_action = self.pred_action_seq[0, self.rollout_step_counter]
        if current_action.dim() == 2:
            current_action = current_action[:, None, :]
        self.rollout_step_counter += 1
        if self.rollout_step_counter == self.multistep:
            self.rollout_step_counter = 0
        return current_action

    # mdtv_agent.py:660-678 (memoised: the schedule only depends on the sampler knobs, which stay plain attributes)
    def get_noise_schedule(self, n_sampling_steps, noise_schedule_type):
        key = (n_sampling_steps, noise_schedule_type, self.sigma_min, self.sigma_max, str(self.device))
        cache = self.__dict__.setdefault("_schedule_cache", {})
        if key not in cache:
            cache[key] = self._build_noise_schedule(n_sampling_steps, noise_schedule_type)
        return cache[key]

    def _build_noise_schedule(self, n_sampling_steps, noise_schedule_type):
        if noise_schedule_type == 'karras':
            return gcs.get_sigmas_karras(n_sampling_steps, self.sigma_min, self.sigma_max, 7, self.device)
        if noise_schedule_type == 'exponential':
            return gcs.get_sigmas_exponential(n_sampling_steps, self.sigma_min, self.sigma_max, self.device)
        if noise_schedule_type == 'vp':
            return gcs.get_sigmas_vp(n_sampling_steps, device=self.device)
        if noise_schedule_type == 'linear':
            return gcs.get_sigmas_linear(n_sampling_steps, self.sigma_min, self.sigma_max, device=self.device)
        if noise_schedule_type == 'cosine_beta':
            return gcs.cosine_beta_schedule(n_sampling_steps, device=self.device)
        if noise_schedule_type == 've':
            return gcs.get_sigmas_ve(n_sampling_steps, self.sigma_min, self.sigma_max, device=self.device)
        if noise_schedule_type == 'iddpm':
            return gcs.get_iddpm_sigmas(n_sampling_steps, self.sigma_min, self.sigma_max, device=self.device)
        raise ValueError('Unknown noise schedule type')

    # mdtv_agent.py:593-658 (the samplers this package restates; the others accept the same model callable)
    def sample_loop(self, sigmas, x_t, state, goal, latent_plan, sampler_type, extra_args={}):
        s_churn = extra_args.get('s_churn', 0)
        s_min = extra_args.get('s_min', 0)
        if sampler_type == 'ddim':
            return gcs.sample_ddim(self.model, state, x_t, goal, sigmas, disable=True)
        if sampler_type == 'heun':
            return gcs.sample_heun(self.model, state, x_t, goal, sigmas, s_churn=s_churn, s_tmin=s_min, disable=True)
        if sampler_type == 'euler':
            return gcs.sample_euler(self.model, state, x_t, goal, sigmas, disable=True)
        if sampler_type == 'euler_ancestral':
            return gcs.sample_euler_ancestral(self.model, state, x_t, goal, sigmas, disable=True)
        if sampler_type == 'dpmpp_2m':
            return gcs.sample_dpmpp_2m(self.model, state, x_t, goal, sigmas, disable=True)
        # the remaining working samplers of the reference (generic per-step driver; 'dpm_adaptive' / 'dpm_fast' are broken in the
        # reference itself and 'dpmpp_2m_sde' needs torchsde: SURVEY 8 a5)
        if sampler_type == 'lms':
            return gcs.sample_lms(self.model, state, x_t, goal, sigmas, disable=True)
        if sampler_type == 'ancestral':
            return gcs.sample_dpm_2_ancestral(self.model, state, x_t, goal, sigmas, disable=True)
        if sampler_type == 'dpm':
            return gcs.sample_dpm_2(self.model, state, x_t, goal, sigmas, disable=True)
        if sampler_type == 'dpmpp_2s_ancestral':
            return gcs.sample_dpmpp_2s_ancestral(self.model, state, x_t, goal, sigmas, disable=True)
        if sampler_type == 'dpmpp_2s':
            return gcs.sample_dpmpp_2s(self.model, state, x_t, goal, sigmas, disable=True)
        if sampler_type == 'dpmpp_2_with_lms':
            return gcs.sample_dpmpp_2_with_lms(self.model, state, x_t, goal, sigmas, disable=True)
        raise ValueError('desired sampler type not found!')

    # mdtv_agent.py:523-550
    def denoise_actions(self, latent_plan, perceptual_emb, latent_goal, inference=False, extra_args={}, x_T=None):
        sampling_steps = self.num_sampling_steps if inference else 10
        if self.model.training:          # reference: self.model.eval() on every call (a ~200-module walk, 0.25 ms of host time)
            self.model.eval()
        ref = perceptual_emb['state_images'] if isinstance(perceptual_emb, dict) and 'state_images' in perceptual_emb else None
        if ref is not None and latent_goal.dim() < ref.dim():
            latent_goal = latent_goal.unsqueeze(1)
        sigmas = self.get_noise_schedule(sampling_steps, self.noise_scheduler)
        if x_T is None:
            x_T = torch.randn((len(latent_goal), self.act_window_size, self.action_dim), device=self.device) * self.sigma_max
        return self.sample_loop(sigmas, x_T, perceptual_emb, latent_goal, latent_plan, self.sampler_type, extra_args)

    def denoise_actions_host(self, state_images_host, latent_goal_host, x_T_host, modality="lang", out_host=None):
        """End-to-end call on HOST tensors (pinned or not): H2D copies of the inputs, the sampling graph, D2H copy
        of the actions, one synchronisation.  This is the call bench.py's ``e2e`` number times."""
        dev = self.device
        inner = getattr(self.model, "inner_model", None)
        fused = self.sampler_type in ("ddim", "euler", "heun", "dpmpp_2m") and getattr(inner, "_variant", None) == "mdtv"     # (ancestral: device path)
        host_ok = all((not t.is_cuda) and t.dtype == torch.float32 and t.is_contiguous()
                      for t in (state_images_host, latent_goal_host, x_T_host))     # raw float* through the C ABI: no silent casts
        if fused and host_ok:
            # straight through the C ABI's host-buffer entry point: no intermediate CUDA tensors on the Python side
            if self.model.training:
                self.model.eval()
            key = ("host", self.num_sampling_steps, self.noise_scheduler, self.sigma_min, self.sigma_max)
            cache = self.__dict__.setdefault("_schedule_cache", {})
            if key not in cache:
                dev_saved, self.device = self.device, torch.device("cpu")
                try:
                    cache[key] = self._build_noise_schedule(self.num_sampling_steps, self.noise_scheduler).float().contiguous()
                finally:
                    self.device = dev_saved
            goal2d = latent_goal_host.reshape(latent_goal_host.shape[0], -1)
            if goal2d.shape[1] != inner.goal_dim or state_images_host.shape[1:] != (inner.n_state_tokens, inner.obs_dim) \
                    or x_T_host.shape != (goal2d.shape[0], inner.action_seq_len, inner.action_dim):
                raise ValueError("denoise_actions_host: input shapes do not match the model configuration")
            if out_host is None:
                out_host = torch.empty(x_T_host.shape, dtype=torch.float32, pin_memory=True)
            out_host.copy_(x_T_host)
            modality_lang = inner.use_modality_encoder and modality == "lang"
            return inner.sample_host(state_images_host, out_host, goal2d, cache[key], self.sampler_type, modality_lang, dev)
        state = {"state_images": state_images_host.to(dev, non_blocking=True), "modality": modality}
        goal = latent_goal_host.to(dev, non_blocking=True)
        x_T = x_T_host.to(dev, non_blocking=True)
        actions = self.denoise_actions(None, state, goal, inference=True, x_T=x_T)
        if out_host is None:
            out_host = torch.empty(actions.shape, dtype=actions.dtype, pin_memory=True)
        out_host.copy_(actions, non_blocking=True)
        torch.cuda.current_stream(dev).synchronize()
        return out_host

    # mdtv_agent.py:552-591 (densities the shipped configs use)
    def make_sample_density(self):
        kind = self.sigma_sample_density_type
        if kind == 'loglogistic':
            return partial(utils.rand_log_logistic, loc=math.log(self.sigma_data), scale=0.5,
                           min_value=self.sigma_min, max_value=self.sigma_max)
        if kind == 'lognormal':
            return partial(utils.rand_log_normal, loc=getattr(self, 'sigma_sample_density_mean', -1.2),
                           scale=getattr(self, 'sigma_sample_density_std', 1.2))
        if kind == 'loguniform':
            return partial(utils.rand_log_uniform, min_value=self.sigma_min, max_value=self.sigma_max)
        if kind == 'uniform':
            return partial(utils.rand_uniform, min_value=self.sigma_min, max_value=self.sigma_max)
        if kind == 'v-diffusion':
            return partial(utils.rand_v_diffusion, sigma_data=self.sigma_data, min_value=self.sigma_min, max_value=self.sigma_max)
        # 'discrete' and 'split-lognormal' cannot be reached through the reference's make_sample_density either (float step count /
        # list-indexed config, mdtv_agent.py:582-589); the underlying utils.rand_discrete / rand_split_log_normal are provided
        raise ValueError('Unknown sample density type')

    # mdtv_agent.py:508-521 (forward value; see GCDenoiser.loss about the backward pass)
    def diffusion_loss(self, perceptual_emb, latent_goal, actions):
        self.model.train()               # as the reference (:517): sampling / validation may have left the model in eval mode
        sigmas = self.make_sample_density()(shape=(len(actions),), device=self.device).to(self.device)
        noise = torch.randn_like(actions)
        loss, _ = self.model.loss(perceptual_emb, actions, latent_goal, noise, sigmas)
        return loss, sigmas, noise
