"""Drop-in for `mode.models.edm_diffusion.score_wrappers.GCDenoiser` (reference score_wrappers.py:18-99).

Swap it in with Hydra:  model._target_: mode_diffusion_policy_b200.score_wrappers.GCDenoiser
"""
from __future__ import annotations

import torch
from torch import nn

from .modedit import MoDeDiT


def append_dims(x, target_dims):
    """reference mode/models/edm_diffusion/utils.py:146-151"""
    dims_to_append = target_dims - x.ndim
    if dims_to_append < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * dims_to_append]


def _instantiate(inner_model):
    if isinstance(inner_model, nn.Module):
        return inner_model
    try:  # the reference instantiates a Hydra config here (score_wrappers.py:28)
        import hydra

        return hydra.utils.instantiate(inner_model)
    except ImportError:
        cfg = dict(inner_model)
        cfg.pop("_target_", None)
        return MoDeDiT(**cfg)


class _EngineLoss(torch.autograd.Function):
    """GCDenoiser.loss through the engine's fused forward + hand-written backward (`mode_train_step`).

    The engine computes the gradient of the (scalar, mean) loss w.r.t. every parameter during `forward`; `backward`
    hands them to autograd scaled by the incoming gradient, so optimisers, Lightning and DDP hooks see ordinary
    `.grad` tensors. Parameters are passed as inputs only to be registered in the graph."""

    @staticmethod
    def forward(ctx, inner, state_images, action, goal, noise, sigma, *params):
        # A torch optimizer may have updated the masters since the last step without touching their version counters
        # (fused / foreach kernels do): re-pack unconditionally unless optim.EngineAdamW keeps the copies in sync itself.
        eng = inner._ensure_engine(action.shape[0], force_repack=not getattr(inner, "_engine_keeps_sync", False))
        stoch = inner._stochastic_args()
        eng.set_stochastic(**stoch)
        inner._last_step_multinomial = stoch["multinomial"]
        loss, out = eng.train_step(state_images, action, goal, noise, sigma)
        ctx.generation = eng.train_generation  # the flat gradient buffer now holds THIS call's gradients
        inner._advance_train_rng()
        ctx.eng, ctx.names, ctx.shapes = eng, inner._param_names, [tuple(p.shape) for p in params]
        ctx.inner = inner
        ctx.needs = [p.requires_grad for p in params]
        ctx.in_shapes = (action.shape[0], tuple(state_images.shape), tuple(goal.shape), state_images.dtype, goal.dtype)
        ctx.mark_non_differentiable(out)
        return loss, out

    @staticmethod
    def backward(ctx, g_loss, _g_out):
        grads = []
        inner = ctx.inner
        if ctx.eng.train_generation != ctx.generation:
            raise RuntimeError(
                "MoDE engine: GCDenoiser.loss was called again before this loss was back-propagated; the engine keeps ONE "
                "set of gradients (its flat buffer), so each training-mode loss() must be followed by its backward() "
                "before the next loss() (sum micro-batch losses into one call, or call backward per micro-batch).")
        if getattr(inner, "_skip_param_grads", False):
            # optim.EngineAdamW reads the engine's flat gradient buffer directly: no per-parameter copies
            inner._loss_grad_scale = g_loss.detach()
            grads = [None] * len(ctx.names)
        for name, shape, need in ([] if grads else zip(ctx.names, ctx.shapes, ctx.needs)):
            if not need or name == "gripper_embed.weight":  # unused unless use_proprio (reference modedit.py:684)
                grads.append(None)
            else:
                grads.append(ctx.eng.grad(name, shape) * g_loss)  # the product is a fresh tensor: the buffer is reused
        # gradients w.r.t. the observation tokens and the goal flow on into the caller's encoders (mode_agent.py:405-411)
        B, s_shape, g_shape, s_dtype, g_dtype = ctx.in_shapes
        d_state = d_goal = None
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[3]:
            d_state, d_goal = ctx.eng.input_grads(B, s_shape, g_shape, ctx.needs_input_grad[1], ctx.needs_input_grad[3])
            d_state = None if d_state is None else (d_state * g_loss).to(s_dtype)
            d_goal = None if d_goal is None else (d_goal * g_loss).to(g_dtype)
        return (None, d_state, None, d_goal, None, None, *grads)


class GCDenoiser(nn.Module):
    """Karras et al. preconditioner around the MoDE network; forward and loss run fused inside the CUDA engine
    (c_in scaling in the embedding kernel, c_out/c_skip combine in the head kernel)."""

    def __init__(self, inner_model, sigma_data=1.0):
        super().__init__()
        self.inner_model = _instantiate(inner_model)
        self.sigma_data = sigma_data
        if isinstance(self.inner_model, MoDeDiT):
            self.inner_model.set_sigma_data(sigma_data)

    def get_scalings(self, sigma):
        c_skip = self.sigma_data ** 2 / (sigma ** 2 + self.sigma_data ** 2)
        c_out = sigma * self.sigma_data / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        c_in = 1 / (sigma ** 2 + self.sigma_data ** 2) ** 0.5
        return c_skip, c_out, c_in

    def _engine(self, batch):
        m = self.inner_model
        if m.training:
            raise NotImplementedError("MoDE engine: in train mode use GCDenoiser.loss (fused forward + backward); call "
                                      ".eval() for denoising / sampling")
        return m._ensure_engine(batch)

    def forward(self, state, action, goal, sigma, uncond=False, **kwargs):
        m = self.inner_model
        goal = m._goals(goal, uncond)
        sigma = torch.as_tensor(sigma, device=action.device)
        return self._engine(action.shape[0]).denoise(state["state_images"], action, goal, sigma).to(action.dtype)

    def loss(self, state, action, goal, noise, sigma, **kwargs):
        """EDM loss (reference score_wrappers.py:45-63). In eval mode: the forward value. In train mode: forward and
        the hand-written backward in one engine call, wired into autograd, with the reference's train-mode
        regularisation (attention / expert dropout, goal masking, per-token multinomial routing) drawn from the
        engine's counter-based stream (MoDeDiT.set_train_rng); `inner_model.deterministic_training = True` turns it off."""
        m = self.inner_model
        goal = m._goals(goal, False)
        m._last_sigma = sigma.detach()  # for the auxiliary router losses (MoDeDiT.load_balancing_loss / z-loss)
        if m.training and torch.is_grad_enabled():
            m.check_trainable()
            params = [p for _, p in m.named_parameters()]
            return _EngineLoss.apply(m, state["state_images"], action, goal, noise, sigma, *params)
        loss, out = m._ensure_engine(action.shape[0]).loss(state["state_images"], action, goal, noise, sigma)
        return loss, out

    def sample_ddim(self, state, action, goal, sigmas):
        """Fused sample_ddim (reference gc_sampling.py:922-951): the whole loop is one CUDA-graph launch."""
        m = self.inner_model
        goal = m._goals(goal, False)
        return self._engine(action.shape[0]).sample_ddim(state["state_images"], action, goal, sigmas).to(action.dtype)

    def sample_fused(self, sampler, state, action, goal, sigmas):
        """Fused "ddim" / "euler" / "dpmpp_2m" sampler loops (engine `mode_sample`): one CUDA-graph launch each."""
        m = self.inner_model
        goal = m._goals(goal, False)
        return self._engine(action.shape[0]).sample(sampler, state["state_images"], action, goal, sigmas).to(action.dtype)

    def sample_program(self, state, action, goal, sigma_eval, reads_probe, prog, noise=None):
        """A sampler program (gc_sampling.SamplerProgram) as one CUDA-graph launch (engine `mode_sample_program`)."""
        m = self.inner_model
        goal = m._goals(goal, False)
        return self._engine(action.shape[0]).sample_program(state["state_images"], action, goal, sigma_eval, reads_probe,
                                                            prog, noise).to(action.dtype)

    def get_params(self):
        return self.inner_model.parameters()
