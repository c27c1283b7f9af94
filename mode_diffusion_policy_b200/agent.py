"""Host-side restatement of the three MoDEAgent methods that bracket the hot path, for callers without Lightning:
`denoise_actions` (reference mode_agent.py:733-760), `sample_loop` (:771-840) and `get_noise_schedule` (:842-860),
plus `diffusion_loss` (:659-672). `MoDEAgent` itself is untouched by the drop-in (it instantiates GCDenoiser / MoDeDiT
through Hydra `_target_` strings); this class exists so that tests and bench.py can make the same calls a rollout
makes without the encoders, wandb and Lightning around them.
"""
from __future__ import annotations

import torch

from . import gc_sampling as S
from . import utils
from .score_wrappers import GCDenoiser


class DenoisingPolicy:
    def __init__(self, model: GCDenoiser, sampler_type="ddim", num_sampling_steps=10, sigma_data=0.5, sigma_min=0.001,
                 sigma_max=80.0, noise_scheduler="exponential", act_window_size=10, action_dim=7,
                 sigma_sample_density_type="loglogistic", device="cuda", sigma_sample_density_mean=-1.2,
                 sigma_sample_density_std=1.2):
        self.model, self.sampler_type, self.num_sampling_steps = model, sampler_type, num_sampling_steps
        self.sigma_data, self.sigma_min, self.sigma_max = sigma_data, sigma_min, sigma_max
        self.noise_scheduler, self.act_window_size, self.action_dim = noise_scheduler, act_window_size, action_dim
        self.sigma_sample_density_type = sigma_sample_density_type
        self.sigma_sample_density_mean, self.sigma_sample_density_std = sigma_sample_density_mean, sigma_sample_density_std
        self.device = device

    def load_pretrained_parameters(self, ckpt_path, strict: bool = False):
        """Denoiser part of MoDEAgent.load_pretrained_parameters (reference mode_agent.py:134-265): reads
        model_cleaned.safetensors / .pt, keeps the `model.inner_model.*` tensors. Returns a checkpoint.LoadReport."""
        from . import checkpoint

        return checkpoint.load_pretrained_parameters(self.model.inner_model, ckpt_path, strict=strict)

    def get_noise_schedule(self, n_sampling_steps, noise_schedule_type):
        t = noise_schedule_type
        if t == "karras":
            return S.get_sigmas_karras(n_sampling_steps, self.sigma_min, self.sigma_max, 7, self.device)
        if t == "exponential":
            return S.get_sigmas_exponential(n_sampling_steps, self.sigma_min, self.sigma_max, self.device)
        if t == "vp":
            return S.get_sigmas_vp(n_sampling_steps, device=self.device)
        if t == "linear":
            return S.get_sigmas_linear(n_sampling_steps, self.sigma_min, self.sigma_max, device=self.device)
        if t == "cosine_beta":
            return S.cosine_beta_schedule(n_sampling_steps, device=self.device)
        if t == "ve":
            return S.get_sigmas_ve(n_sampling_steps, self.sigma_min, self.sigma_max, device=self.device)
        if t == "iddpm":
            return S.get_iddpm_sigmas(n_sampling_steps, self.sigma_min, self.sigma_max, device=self.device)
        raise ValueError("Unknown noise schedule type")

    def sample_loop(self, sigmas, x_t, state, goal, latent_plan=None, sampler_type="ddim", extra_args={}):
        if sampler_type not in S.SAMPLERS:
            raise ValueError("desired sampler type not found!")
        fn = S.SAMPLERS[sampler_type]
        if sampler_type == "heun":
            return fn(self.model, state, x_t, goal, sigmas, s_churn=extra_args.get("s_churn", 0),
                      s_tmin=extra_args.get("s_min", 0), disable=True)
        return fn(self.model, state, x_t, goal, sigmas, disable=True)

    @torch.no_grad()
    def denoise_actions(self, latent_plan, perceptual_emb, latent_goal, inference=False, extra_args={}, x=None):
        steps = self.num_sampling_steps if inference else 10
        self.model.eval()
        state = perceptual_emb
        ref = state["state_images"] if isinstance(state, dict) else state
        if latent_goal.dim() < ref.dim():
            latent_goal = latent_goal.unsqueeze(1)
        sigmas = self.get_noise_schedule(steps, self.noise_scheduler)
        if x is None:  # the caller owns the RNG (mode_agent.py:756)
            x = torch.randn((len(latent_goal), self.act_window_size, self.action_dim), device=self.device) * self.sigma_max
        return self.sample_loop(sigmas, x, state, latent_goal, latent_plan, self.sampler_type, extra_args)

    def make_sample_density(self):
        """Training noise-level density by `sigma_sample_density_type` (reference mode_agent.py:692-731; the shipped
        config uses 'loglogistic'). 'split-lognormal' reads its parameters from an empty config dict in the reference
        and therefore raises KeyError there; it does here too."""
        import math
        from functools import partial

        kind = self.sigma_sample_density_type
        if kind == "lognormal":
            return partial(utils.rand_log_normal, loc=self.sigma_sample_density_mean, scale=self.sigma_sample_density_std)
        if kind == "loglogistic":
            return partial(utils.rand_log_logistic, loc=math.log(self.sigma_data), scale=0.5,
                           min_value=self.sigma_min, max_value=self.sigma_max)
        if kind == "loguniform":
            return partial(utils.rand_log_uniform, min_value=self.sigma_min, max_value=self.sigma_max)
        if kind == "uniform":
            return partial(utils.rand_uniform, min_value=self.sigma_min, max_value=self.sigma_max)
        if kind == "v-diffusion":
            return partial(utils.rand_v_diffusion, sigma_data=self.sigma_data, min_value=self.sigma_min,
                           max_value=self.sigma_max)
        if kind == "discrete":
            # the reference passes the float `num_sampling_steps * 1e5` on to torch.linspace (a TypeError); int() here
            sigmas = self.get_noise_schedule(int(self.num_sampling_steps * 1e5), "exponential")
            return partial(utils.rand_discrete, values=sigmas)
        if kind == "split-lognormal":
            raise KeyError("mean")  # the reference indexes an empty sd_config here (mode_agent.py:725-729)
        raise ValueError("Unknown sample density type")

    def diffusion_loss(self, perceptual_emb, latent_goal, actions):
        """Score-matching loss of one batch (reference mode_agent.py:659-672): train mode, per-sample sigma from the
        training density, caller-side Gaussian noise, `GCDenoiser.loss`."""
        self.model.train()
        sigmas = self.make_sample_density()(shape=(len(actions),), device=self.device).to(self.device)
        noise = torch.randn_like(actions)
        loss, _ = self.model.loss(perceptual_emb, actions, latent_goal, noise, sigmas)
        return loss
