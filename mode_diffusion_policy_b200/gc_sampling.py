"""Noise schedules and k-diffusion samplers with the reference's names and signatures
(`mode.models.edm_diffusion.gc_sampling`, reference gc_sampling.py:26-994), written around one shared step helper.

`sample_ddim` — the reference's default sampler (conf/model/mode_agent.yaml:9) — and `sample_euler` (without churn) /
`sample_dpmpp_2m` run as ONE fused engine call (the whole sigma loop is a CUDA graph, the update is the epilogue of the
output-head kernel) whenever the model is the engine-backed GCDenoiser and no callback / scaler / extra_args
intervene; every other sampler calls `model(state, action, goal, sigma)` = the engine's fused denoiser once
per network evaluation and does its (tiny) update arithmetic in torch.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import utils


# ---------------------------------------------------------------------------------------------------- schedules
def append_zero(action):
    return torch.cat([action, action.new_zeros([1])])


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0, device="cpu"):
    """Karras et al. (2022) schedule (reference gc_sampling.py:26-32)."""
    ramp = torch.linspace(0, 1, n)
    lo, hi = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    return append_zero((hi + ramp * (lo - hi)) ** rho).to(device)


def get_sigmas_exponential(n, sigma_min, sigma_max, device="cpu"):
    """Exponential schedule, the reference default (gc_sampling.py:35-38)."""
    return append_zero(torch.linspace(math.log(sigma_max), math.log(sigma_min), n, device=device).exp())


def get_sigmas_linear(n, sigma_min, sigma_max, device="cpu"):
    return append_zero(torch.linspace(sigma_max, sigma_min, n, device=device))


def cosine_beta_schedule(n, s=0.008, device="cpu"):
    """reference gc_sampling.py:47-58"""
    steps = n + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = np.clip(1 - (ac[1:] / ac[:-1]), a_min=0, a_max=0.999)
    return append_zero(torch.tensor(np.flip(betas).copy(), device=device, dtype=torch.float32))


def get_sigmas_ve(n, sigma_min=0.02, sigma_max=100, device="cpu"):
    """reference gc_sampling.py:61-68 (note: t runs over [0, n+1] as written there)."""
    t = torch.linspace(0, n + 1, n, device=device)
    return append_zero(torch.sqrt((sigma_max ** 2) * ((sigma_min ** 2 / sigma_max ** 2) ** (t / (n - 1)))))


def get_iddpm_sigmas(n, sigma_min=0.02, sigma_max=100, M=1000, j_0=0, C_1=0.001, C_2=0.008, device="cpu"):
    """reference gc_sampling.py:71-81"""
    idx = torch.arange(n, dtype=torch.float64, device=device)
    u = torch.zeros(M + 1, dtype=torch.float64, device=device)
    abar = lambda j: (0.5 * np.pi * j / M / (C_2 + 1)).sin() ** 2  # noqa: E731
    for j in torch.arange(M, j_0, -1, device=device):
        u[j - 1] = ((u[j] ** 2 + 1) / (abar(j - 1) / abar(j)).clip(min=C_1) - 1).sqrt()
    uf = u[torch.logical_and(u >= sigma_min, u <= sigma_max)]
    return append_zero(uf[((len(uf) - 1) / (n - 1) * idx).round().to(torch.int64)]).to(torch.float32)


def get_sigmas_vp(n, beta_d=19.9, beta_min=0.1, eps_s=1e-3, device="cpu"):
    t = torch.linspace(1, eps_s, n, device=device)
    return append_zero(torch.sqrt(torch.exp(beta_d * t ** 2 / 2 + beta_min * t) - 1))


# ---------------------------------------------------------------------------------------------------- helpers
def to_d(action, sigma, denoised):
    """Karras ODE derivative (reference gc_sampling.py:91-93)."""
    return (action - denoised) / utils.append_dims(sigma, action.ndim)


def default_noise_sampler(x):
    return lambda sigma, sigma_next: torch.randn_like(x)


def get_ancestral_step(sigma_from, sigma_to, eta=1.0):
    """reference gc_sampling.py:102-109"""
    if not eta:
        return sigma_to, 0.0
    sigma_up = min(sigma_to, eta * (sigma_to ** 2 * (sigma_from ** 2 - sigma_to ** 2) / sigma_from ** 2) ** 0.5)
    return (sigma_to ** 2 - sigma_up ** 2) ** 0.5, sigma_up


def _t(sigma):  # t = -ln sigma
    return sigma.log().neg()


def _sig(t):
    return t.neg().exp()


class _Loop:
    """Shared plumbing of every sampler: evaluates the denoiser at a scalar sigma, reports to the callback, clips."""

    def __init__(self, model, state, goal, scaler, extra_args, callback, key="action"):
        self.model, self.state, self.goal, self.scaler = model, state, goal, scaler
        self.extra = {} if extra_args is None else extra_args
        self.callback, self.key = callback, key

    def denoise(self, x, sigma):
        return self.model(self.state, x, self.goal, sigma * x.new_ones([x.shape[0]]), **self.extra)

    def report(self, x, i, sigma, sigma_hat, denoised):
        if self.callback is not None:
            self.callback({self.key: x, "i": i, "sigma": sigma, "sigma_hat": sigma_hat, "denoised": denoised})

    def clip(self, x):
        return x if self.scaler is None else self.scaler.clip_output(x)


def _churn(x, sigmas, i, s_churn, s_tmin, s_tmax, s_noise):
    """Karras 'churn': raise the noise level to sigma_hat before the step (reference gc_sampling.py:196-201)."""
    gamma = min(s_churn / (len(sigmas) - 1), 2 ** 0.5 - 1) if s_tmin <= sigmas[i] <= s_tmax else 0.0
    eps = torch.randn_like(x) * s_noise
    sigma_hat = sigmas[i] * (gamma + 1)
    if gamma > 0:
        x = x + eps * (sigma_hat ** 2 - sigmas[i] ** 2) ** 0.5
    return x, sigma_hat


def _fusable(model, attr, sigmas, scaler, extra_args, callback) -> bool:
    """The engine's one-launch samplers take at most 64 evaluations at strictly positive sigmas (only the trailing
    sigma may be 0); every other schedule the reference samplers accept runs through the host loop."""
    if scaler is not None or callback is not None or extra_args or not hasattr(model, attr):
        return False
    s = torch.as_tensor(sigmas)
    return 2 <= s.numel() <= 65 and bool((s[:-1] > 0).all())


class SamplerProgram:
    """A sampler written as one row of coefficients per network evaluation, for the engine's one-launch execution
    (`GCDenoiser.sample_program` -> `mode_sample_program`). Every update in this file is linear in the step's base
    sample X, the probe P given to a second evaluation, the denoised output D, up to four history tensors H and a noise
    draw:   dst <- cX*X + cP*P + cD*D + sum_j cH[j]*H[j] + cN*noise,   H[slot] <- hX*x_in + hD*D.
    The coefficients are evaluated here in float64 from the schedule (the reference evaluates the same expressions on
    0-dim fp32 tensors; the difference is an fp32 ulp on numbers that multiply bf16-accurate network outputs)."""

    def __init__(self):
        self.sigma, self.reads_probe, self.rows, self.noise_index = [], [], [], []

    def eval(self, sigma, on_probe=False, cX=0.0, cP=0.0, cD=0.0, cH=(0.0, 0.0, 0.0, 0.0), cN=0.0, hX=0.0, hD=0.0,
             slot=-1, to_probe=False, noise=None):
        """Append one evaluation at `sigma` on X (or P) followed by its update row; `noise`: index of the caller's draw."""
        self.sigma.append(float(sigma))
        self.reads_probe.append(1 if on_probe else 0)
        self.rows.append([cX, cP, cD, *cH, cN if noise is not None else 0.0, hX, hD, float(slot), 1.0 if to_probe else 0.0,
                          0.0, 0.0, 0.0, 0.0])
        self.noise_index.append(noise)

    def run(self, model, state, action, goal, draws):
        """draws: the noise tensors the host loop would have drawn, in order (may be empty)."""
        noise = None
        if any(i is not None for i in self.noise_index):
            zero = torch.zeros_like(action)
            noise = torch.stack([zero if i is None else draws[i] for i in self.noise_index])
        return model.sample_program(state, action, goal, np.asarray(self.sigma, np.float32),
                                    np.asarray(self.reads_probe, np.int32), np.asarray(self.rows, np.float32), noise)


def _program_ok(model, sigmas, scaler, extra_args, callback, evals_per_step=2) -> bool:
    if scaler is not None or callback is not None or extra_args or not hasattr(model, "sample_program"):
        return False
    s = torch.as_tensor(sigmas)
    return s.numel() >= 2 and evals_per_step * (s.numel() - 1) <= 64 and bool((s[:-1] > 0).all())


def _floats(sigmas):
    return [float(v) for v in torch.as_tensor(sigmas).detach().double().cpu()]


def _exp_step(x, denoised, t, t_next):
    """x <- (sigma(t')/sigma(t)) x - expm1(-(t'-t)) D : the DPM-Solver-1 / DDIM update (reference gc_sampling.py:950)."""
    return (_sig(t_next) / _sig(t)) * x - (-(t_next - t)).expm1() * denoised


# ---------------------------------------------------------------------------------------------------- samplers
@torch.no_grad()
def sample_ddim(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None, eta=1.0):
    """DPM-Solver-1 / DDIM (reference gc_sampling.py:922-951)."""
    if _fusable(model, "sample_ddim", sigmas, scaler, extra_args, callback):
        return model.sample_ddim(state, action, goal, sigmas)  # fused: one CUDA-graph launch for the whole loop
    lp = _Loop(model, state, goal, scaler, extra_args, callback)
    for i in range(len(sigmas) - 1):
        denoised = lp.denoise(action, sigmas[i])
        lp.report(action, i, sigmas[i], sigmas[i], denoised)
        action = _exp_step(action, denoised, _t(sigmas[i]), _t(sigmas[i + 1]))
    return action


@torch.no_grad()
def sample_euler(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None,
                 s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0):
    """Algorithm 2 of Karras et al. without the 2nd-order correction (reference gc_sampling.py:164-211). Without churn,
    callback, scaler or extra_args the whole loop is one fused engine call."""
    if s_churn == 0 and _fusable(model, "sample_fused", sigmas, scaler, extra_args, callback):
        for _ in range(len(sigmas) - 1):  # the reference draws eps every step even when gamma = 0 (gc_sampling.py:196):
            torch.randn_like(action)      # leave the caller's RNG where the reference would
        return model.sample_fused("euler", state, action, goal, sigmas)
    lp = _Loop(model, state, goal, scaler, extra_args, callback, key="x")
    for i in range(len(sigmas) - 1):
        action, sigma_hat = _churn(action, sigmas, i, s_churn, s_tmin, s_tmax, s_noise)
        denoised = lp.denoise(action, sigma_hat)
        d = to_d(action, sigma_hat, denoised)
        lp.report(action, i, sigmas[i], sigma_hat, denoised)
        action = lp.clip(action + d * (sigmas[i + 1] - sigma_hat))
    return action


@torch.no_grad()
def sample_euler_ancestral(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None,
                           disable=None, eta=1.0):
    """reference gc_sampling.py:213-254. One engine launch; the noise is drawn here, by the caller's RNG, in the order the
    loop would draw it."""
    if _program_ok(model, sigmas, scaler, extra_args, callback, evals_per_step=1):
        sg, prog, draws = _floats(sigmas), SamplerProgram(), []
        for i in range(len(sg) - 1):
            s0 = sg[i]
            down, up = (float(v) for v in get_ancestral_step(s0, sg[i + 1], eta=eta))
            k = None
            if down > 0:
                draws.append(torch.randn_like(action))
                k = len(draws) - 1
            prog.eval(s0, cX=1 + (down - s0) / s0, cD=-(down - s0) / s0, cN=up, noise=k)
        return prog.run(model, state, action, goal, draws)
    lp = _Loop(model, state, goal, scaler, extra_args, callback, key="x")
    for i in range(len(sigmas) - 1):
        denoised = lp.denoise(action, sigmas[i])
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=eta)
        lp.report(action, i, sigmas[i], sigmas[i], denoised)
        action = action + to_d(action, sigmas[i], denoised) * (sigma_down - sigmas[i])
        if sigma_down > 0:
            action = action + torch.randn_like(action) * sigma_up
        action = lp.clip(action)
    return action


@torch.no_grad()
def sample_heun(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None,
                s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0):
    """Algorithm 2 of Karras et al. (Heun) (reference gc_sampling.py:256-312). Without churn the loop is one engine launch."""
    if s_churn == 0 and _program_ok(model, sigmas, scaler, extra_args, callback):
        sg, prog = _floats(sigmas), SamplerProgram()
        for i in range(len(sg) - 1):
            torch.randn_like(action)  # the reference draws eps every step even when gamma = 0 (:287)
            s0, s1 = sg[i], sg[i + 1]
            dt = s1 - s0
            if s1 == 0:
                prog.eval(s0, cX=1 + dt / s0, cD=-dt / s0)                                       # Euler step to sigma = 0
            else:
                prog.eval(s0, cX=1 + dt / s0, cD=-dt / s0, hX=1 / s0, hD=-1 / s0, slot=0, to_probe=True)  # H0 = d, P = X + d dt
                prog.eval(s1, on_probe=True, cX=1.0, cH=(dt / 2, 0, 0, 0), cP=dt / (2 * s1), cD=-dt / (2 * s1))  # X += (d + d2)/2 dt
        return prog.run(model, state, action, goal, [])
    lp = _Loop(model, state, goal, scaler, extra_args, callback, key="x")
    for i in range(len(sigmas) - 1):
        action, sigma_hat = _churn(action, sigmas, i, s_churn, s_tmin, s_tmax, s_noise)
        denoised = lp.denoise(action, sigma_hat)
        d = to_d(action, sigma_hat, denoised)
        lp.report(action, i, sigmas[i], sigma_hat, denoised)
        dt = sigmas[i + 1] - sigma_hat
        if sigmas[i + 1] == 0:
            action = action + d * dt
        else:
            probe = action + d * dt
            d_2 = to_d(probe, sigmas[i + 1], lp.denoise(probe, sigmas[i + 1]))
            action = action + (d + d_2) / 2 * dt
        action = lp.clip(action)
    return action


@torch.no_grad()
def sample_dpm_2(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None,
                 s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0):
    """DPM-Solver-2 flavoured midpoint steps (reference gc_sampling.py:314-373). Without churn: one engine launch."""
    if s_churn == 0 and _program_ok(model, sigmas, scaler, extra_args, callback):
        sg, prog = _floats(sigmas), SamplerProgram()
        for i in range(len(sg) - 1):
            torch.randn_like(action)  # unused eps draw of the reference (:344)
            s0, s1 = sg[i], sg[i + 1]
            if s1 == 0:
                prog.eval(s0, cX=1 + (s1 - s0) / s0, cD=-(s1 - s0) / s0)
            else:
                mid = math.exp(0.5 * (math.log(s0) + math.log(s1)))
                prog.eval(s0, cX=1 + (mid - s0) / s0, cD=-(mid - s0) / s0, to_probe=True)          # P = X + d (mid - s0)
                prog.eval(mid, on_probe=True, cX=1.0, cP=(s1 - s0) / mid, cD=-(s1 - s0) / mid)     # X += d2 (s1 - s0)
        return prog.run(model, state, action, goal, [])
    lp = _Loop(model, state, goal, scaler, extra_args, callback)
    for i in range(len(sigmas) - 1):
        action, sigma_hat = _churn(action, sigmas, i, s_churn, s_tmin, s_tmax, s_noise)
        denoised = lp.denoise(action, sigma_hat)
        d = to_d(action, sigma_hat, denoised)
        lp.report(action, i, sigmas[i], sigma_hat, denoised)
        if sigmas[i + 1] == 0:
            action = action + d * (sigmas[i + 1] - sigma_hat)
        else:
            sigma_mid = sigma_hat.log().lerp(sigmas[i + 1].log(), 0.5).exp()
            probe = action + d * (sigma_mid - sigma_hat)
            d_2 = to_d(probe, sigma_mid, lp.denoise(probe, sigma_mid))
            action = action + d_2 * (sigmas[i + 1] - sigma_hat)
        action = lp.clip(action)
    return action


@torch.no_grad()
def sample_dpm_2_ancestral(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None,
                           disable=None, eta=1.0):
    """reference gc_sampling.py:375-410. One engine launch (see sample_euler_ancestral for the noise)."""
    if _program_ok(model, sigmas, scaler, extra_args, callback):
        sg, prog, draws = _floats(sigmas), SamplerProgram(), []
        for i in range(len(sg) - 1):
            s0 = sg[i]
            down, up = (float(v) for v in get_ancestral_step(s0, sg[i + 1], eta=eta))
            if down == 0:
                prog.eval(s0, cX=1 + (down - s0) / s0, cD=-(down - s0) / s0)
            else:
                mid = math.exp(0.5 * (math.log(s0) + math.log(down)))
                draws.append(torch.randn_like(action))
                prog.eval(s0, cX=1 + (mid - s0) / s0, cD=-(mid - s0) / s0, to_probe=True)
                prog.eval(mid, on_probe=True, cX=1.0, cP=(down - s0) / mid, cD=-(down - s0) / mid, cN=up, noise=len(draws) - 1)
        return prog.run(model, state, action, goal, draws)
    lp = _Loop(model, state, goal, scaler, extra_args, callback, key="x")
    for i in range(len(sigmas) - 1):
        denoised = lp.denoise(action, sigmas[i])
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=eta)
        lp.report(action, i, sigmas[i], sigmas[i], denoised)
        d = to_d(action, sigmas[i], denoised)
        if sigma_down == 0:
            action = action + d * (sigma_down - sigmas[i])
        else:
            sigma_mid = sigmas[i].log().lerp(sigma_down.log(), 0.5).exp()
            probe = action + d * (sigma_mid - sigmas[i])
            d_2 = to_d(probe, sigma_mid, lp.denoise(probe, sigma_mid))
            action = action + d_2 * (sigma_down - sigmas[i])
            action = action + torch.randn_like(action) * sigma_up
        action = lp.clip(action)
    return action


def linear_multistep_coeff(order, t, i, j):
    """Integral of the j-th Lagrange basis polynomial over [t_i, t_{i+1}] (reference gc_sampling.py:413-427)."""
    if order - 1 > i:
        raise ValueError(f"Order {order} too high for step {i}")
    from scipy import integrate

    def basis(tau):
        prod = 1.0
        for k in range(order):
            if k != j:
                prod *= (tau - t[i - k]) / (t[i - j] - t[i - k])
        return prod

    return integrate.quad(basis, t[i], t[i + 1], epsrel=1e-4)[0]


@torch.no_grad()
def sample_lms(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None, order=4):
    """Linear multistep sampler (reference gc_sampling.py:429-466). Up to order 4 the loop is one engine launch: the
    derivative history lives in the engine's four history tensors."""
    if order <= 4 and _program_ok(model, sigmas, scaler, extra_args, callback, evals_per_step=1):
        sg, prog = _floats(sigmas), SamplerProgram()
        sig32 = torch.as_tensor(sigmas).detach().cpu().numpy()
        for i in range(len(sg) - 1):
            cur = min(i + 1, order)
            co = [linear_multistep_coeff(cur, sig32, i, j) for j in range(cur)]  # co[j] multiplies d_{i-j}
            cH = [0.0] * 4
            for j in range(1, cur):
                cH[(i - j) % 4] = co[j]
            # d_i = (X - D)/s_i goes to slot i % 4 (it replaces d_{i-4}, which order <= 4 no longer needs)
            prog.eval(sg[i], cX=1 + co[0] / sg[i], cD=-co[0] / sg[i], cH=tuple(cH), hX=1 / sg[i], hD=-1 / sg[i], slot=i % 4)
        return prog.run(model, state, action, goal, [])
    lp = _Loop(model, state, goal, scaler, extra_args, callback, key="x")
    sig = sigmas.detach().cpu().numpy()
    history = []
    for i in range(len(sigmas) - 1):
        denoised = lp.denoise(action, sigmas[i])
        history.append(to_d(action, sigmas[i], denoised))
        history = history[-order:]
        lp.report(action, i, sigmas[i], sigmas[i], denoised)
        cur = min(i + 1, order)
        coeffs = [linear_multistep_coeff(cur, sig, i, j) for j in range(cur)]
        action = lp.clip(action + sum(c * d for c, d in zip(coeffs, reversed(history))))
    return action


@torch.no_grad()
def sample_dpmpp_2m(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None):
    """DPM-Solver++(2M) (reference gc_sampling.py:699-734); fused into one engine call when nothing intervenes."""
    if _fusable(model, "sample_fused", sigmas, scaler, extra_args, callback):
        return model.sample_fused("dpmpp_2m", state, action, goal, sigmas)
    lp = _Loop(model, state, goal, scaler, extra_args, callback)
    old = None
    for i in range(len(sigmas) - 1):
        denoised = lp.denoise(action, sigmas[i])
        lp.report(action, i, sigmas[i], sigmas[i], denoised)
        t, t_next = _t(sigmas[i]), _t(sigmas[i + 1])
        target = denoised
        if old is not None and sigmas[i + 1] != 0:
            r = (t - _t(sigmas[i - 1])) / (t_next - t)
            target = (1 + 1 / (2 * r)) * denoised - (1 / (2 * r)) * old
        action = _exp_step(action, target, t, t_next)
        old = denoised
    return action


sample_dpmpp_2_with_lms = sample_dpmpp_2m  # identical bodies in the reference (gc_sampling.py:797-831)


@torch.no_grad()
def sample_dpmpp_2s(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None, disable=None,
                    eta=1.0):
    """DPM-Solver++(2S) (reference gc_sampling.py:955-994). One engine launch."""
    if _program_ok(model, sigmas, scaler, extra_args, callback):
        sg, prog = _floats(sigmas), SamplerProgram()
        for i in range(len(sg) - 1):
            s0, s1 = sg[i], sg[i + 1]
            if s1 == 0:
                prog.eval(s0, cX=1 + (s1 - s0) / s0, cD=-(s1 - s0) / s0)
            else:
                t, tn = -math.log(s0), -math.log(s1)
                sm = t + 0.5 * (tn - t)
                prog.eval(s0, cX=math.exp(-sm) / s0, cD=-math.expm1(-(sm - t)), to_probe=True)
                prog.eval(math.exp(-sm), on_probe=True, cX=s1 / s0, cD=-math.expm1(-(tn - t)))
        return prog.run(model, state, action, goal, [])
    lp = _Loop(model, state, goal, scaler, extra_args, callback)
    for i in range(len(sigmas) - 1):
        denoised = lp.denoise(action, sigmas[i])
        lp.report(action, i, sigmas[i], sigmas[i], denoised)
        if sigmas[i + 1] == 0:
            action = action + to_d(action, sigmas[i], denoised) * (sigmas[i + 1] - sigmas[i])
        else:
            t, t_next = _t(sigmas[i]), _t(sigmas[i + 1])
            s = t + 0.5 * (t_next - t)
            probe = _exp_step(action, denoised, t, s)
            action = _exp_step(action, lp.denoise(probe, _sig(s)), t, t_next)
        action = lp.clip(action)
    return action


@torch.no_grad()
def sample_dpmpp_2s_ancestral(model, state, action, goal, sigmas, scaler=None, extra_args=None, callback=None,
                              disable=None, eta=1.0, s_noise=1.0, noise_sampler=None):
    """reference gc_sampling.py:873-919. One engine launch; `noise_sampler` is called here once per step, in order."""
    noise_sampler = default_noise_sampler(action) if noise_sampler is None else noise_sampler
    if _program_ok(model, sigmas, scaler, extra_args, callback):
        sg, prog, draws = _floats(sigmas), SamplerProgram(), []
        for i in range(len(sg) - 1):
            s0 = sg[i]
            down, up = (float(v) for v in get_ancestral_step(s0, sg[i + 1], eta=eta))
            draws.append(noise_sampler(torch.as_tensor(sigmas)[i], torch.as_tensor(sigmas)[i + 1]))
            k = len(draws) - 1
            if down == 0:
                prog.eval(s0, cX=1 + (down - s0) / s0, cD=-(down - s0) / s0, cN=s_noise * up, noise=k)
            else:
                t, tn = -math.log(s0), -math.log(down)
                sm = t + 0.5 * (tn - t)
                prog.eval(s0, cX=math.exp(-sm) / s0, cD=-math.expm1(-(sm - t)), to_probe=True)
                prog.eval(math.exp(-sm), on_probe=True, cX=down / s0, cD=-math.expm1(-(tn - t)), cN=s_noise * up, noise=k)
        return prog.run(model, state, action, goal, draws)
    lp = _Loop(model, state, goal, scaler, extra_args, callback)
    for i in range(len(sigmas) - 1):
        denoised = lp.denoise(action, sigmas[i])
        sigma_down, sigma_up = get_ancestral_step(sigmas[i], sigmas[i + 1], eta=eta)
        lp.report(action, i, sigmas[i], sigmas[i], denoised)
        if sigma_down == 0:
            action = action + to_d(action, sigmas[i], denoised) * (sigma_down - sigmas[i])
        else:
            t, t_next = _t(sigmas[i]), _t(sigma_down)
            s = t + 0.5 * (t_next - t)
            probe = _exp_step(action, denoised, t, s)
            action = _exp_step(action, lp.denoise(probe, _sig(s)), t, t_next)
        action = lp.clip(action + noise_sampler(sigmas[i], sigmas[i + 1]) * s_noise * sigma_up)
    return action


class BrownianNoiseSampler:
    """Noise sampler of the stochastic DPM-Solver++ (reference BrownianTreeNoiseSampler, gc_sampling.py:139-162): returns
    the increment of ONE Brownian motion per sample element over [sigma, sigma_next], normalised to unit variance, so
    overlapping intervals are correlated the way the SDE solver expects. The reference builds it on torchsde's
    BrownianTree; this is a self-contained Brownian-bridge construction with the same interface and statistics (the
    draws themselves are a different — seedable — stream). W is pinned at sigma_min (0) and sigma_max; a new time point is
    sampled conditionally on its nearest known neighbours."""

    def __init__(self, x, sigma_min, sigma_max, seed=None, transform=lambda v: v):
        self.transform = transform
        t0, t1 = float(transform(torch.as_tensor(sigma_min))), float(transform(torch.as_tensor(sigma_max)))
        self.t0, self.t1 = min(t0, t1), max(t0, t1)
        self.gen = None
        if seed is not None:
            self.gen = torch.Generator(device=x.device)
            self.gen.manual_seed(int(seed))
        self.like = x
        self.times = [self.t0, self.t1]
        self.values = [torch.zeros_like(x), self._randn() * math.sqrt(self.t1 - self.t0)]

    def _randn(self):
        return torch.randn(self.like.shape, dtype=self.like.dtype, device=self.like.device, generator=self.gen)

    def _w(self, t):
        import bisect
        t = min(max(t, self.t0), self.t1)
        k = bisect.bisect_left(self.times, t)
        if self.times[k] == t:
            return self.values[k]
        ta, tb, wa, wb = self.times[k - 1], self.times[k], self.values[k - 1], self.values[k]
        mean = wa + (wb - wa) * ((t - ta) / (tb - ta))                 # Brownian bridge between the known neighbours
        w = mean + self._randn() * math.sqrt((t - ta) * (tb - t) / (tb - ta))
        self.times.insert(k, t)
        self.values.insert(k, w)
        return w

    def __call__(self, sigma, sigma_next):
        a, b = float(self.transform(torch.as_tensor(sigma))), float(self.transform(torch.as_tensor(sigma_next)))
        sign = 1.0 if a <= b else -1.0
        lo, hi = min(a, b), max(a, b)
        return (self._w(hi) - self._w(lo)) * (sign / math.sqrt(hi - lo))


@torch.no_grad()
def sample_dpmpp_sde(model, state, action, goal, sigmas, extra_args=None, callback=None, disable=None, eta=1.0, s_noise=1.0,
                     scaler=None, noise_sampler=None, r=1 / 2):
    """DPM-Solver++ (stochastic), reference gc_sampling.py:736-793 (`sampler_type='dpmpp_2m_sde'`, mode_agent.py:827): two
    network evaluations per step, the second on a probe at the intermediate time s = t + r h, both followed by a Brownian
    increment from `noise_sampler(sigma, sigma_next)` (default: BrownianNoiseSampler). One engine launch; the increments
    are requested here in the order the loop requests them."""
    sig = torch.as_tensor(sigmas)
    if noise_sampler is None:
        noise_sampler = BrownianNoiseSampler(action, sig[sig > 0].min(), sig.max())
    fac = 1 / (2 * r)
    if _program_ok(model, sigmas, scaler, extra_args, callback):
        sg, prog, draws, ok = _floats(sigmas), SamplerProgram(), [], True
        for i in range(len(sg) - 1):
            s0, s1 = sg[i], sg[i + 1]
            if s1 == 0:
                prog.eval(s0, cX=0.0, cD=1.0)  # Euler step to sigma = 0: x + (x - D) / s0 * (0 - s0) = D
                continue
            t, tn = -math.log(s0), -math.log(s1)
            sm = math.exp(-(t + (tn - t) * r))                                     # sigma at the intermediate time
            d1, u1 = (float(v) for v in get_ancestral_step(s0, sm, eta))
            d2, u2 = (float(v) for v in get_ancestral_step(s0, s1, eta))
            if d1 <= 0 or d2 <= 0:
                ok = False
                break
            draws.append(noise_sampler(sig[i].new_tensor(s0), sig[i].new_tensor(sm)))
            k1 = len(draws) - 1
            draws.append(noise_sampler(sig[i].new_tensor(s0), sig[i].new_tensor(s1)))
            k2 = len(draws) - 1
            e1, e2 = -math.expm1(t + math.log(d1)), -math.expm1(t + math.log(d2))  # -(t - t_fn(sd)).expm1()
            prog.eval(s0, cX=d1 / s0, cD=e1, cN=s_noise * u1, noise=k1, hD=1.0, slot=0, to_probe=True)   # P = x_2, H0 = D
            prog.eval(sm, on_probe=True, cX=d2 / s0, cH=(e2 * (1 - fac), 0, 0, 0), cD=e2 * fac, cN=s_noise * u2, noise=k2)
        if ok:
            return prog.run(model, state, action, goal, draws)
        raise ValueError("sample_dpmpp_sde: eta too large for this schedule (sigma_down = 0 inside the loop)")
    lp = _Loop(model, state, goal, scaler, extra_args, callback, key="x")
    x = action
    for i in range(len(sigmas) - 1):
        denoised = lp.denoise(x, sigmas[i])
        lp.report(x, i, sigmas[i], sigmas[i], denoised)
        if sigmas[i + 1] == 0:
            x = x + to_d(x, sigmas[i], denoised) * (sigmas[i + 1] - sigmas[i])
        else:
            t, t_next = _t(sigmas[i]), _t(sigmas[i + 1])
            s = t + (t_next - t) * r
            sd, su = get_ancestral_step(_sig(t), _sig(s), eta)
            s_ = _t(sd)
            x_2 = (_sig(s_) / _sig(t)) * x - (t - s_).expm1() * denoised
            x_2 = x_2 + noise_sampler(_sig(t), _sig(s)) * s_noise * su
            denoised_2 = lp.denoise(x_2, _sig(s))
            sd, su = get_ancestral_step(_sig(t), _sig(t_next), eta)
            t_next_ = _t(sd)
            denoised_d = (1 - fac) * denoised + fac * denoised_2
            x = (_sig(t_next_) / _sig(t)) * x - (t - t_next_).expm1() * denoised_d
            x = lp.clip(x + noise_sampler(_sig(t), _sig(t_next)) * s_noise * su)
    return x


SAMPLERS = {
    # sampler_type keys of MoDEAgent.sample_loop (reference mode_agent.py:771-840). 'dpm_fast', 'dpm_adaptive' raise in
    # the reference (undefined names) and are not provided.
    "lms": sample_lms, "heun": sample_heun, "euler": sample_euler, "ancestral": sample_dpm_2_ancestral,
    "euler_ancestral": sample_euler_ancestral, "dpm": sample_dpm_2, "dpmpp_2s_ancestral": sample_dpmpp_2s_ancestral,
    "dpmpp_2m": sample_dpmpp_2m, "ddim": sample_ddim, "dpmpp_2s": sample_dpmpp_2s,
    "debugging": sample_dpmpp_2_with_lms, "dpmpp_2_with_lms": sample_dpmpp_2_with_lms, "dpmpp_2m_sde": sample_dpmpp_sde,
}
