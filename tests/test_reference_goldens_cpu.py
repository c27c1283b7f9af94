"""CPU tests against the round-2 reference goldens (tests/golden/make_full_goldens.py):

* the numpy oracle at BASELINE.json's full size (12 layers, d=1024, 4 experts) against the REFERENCE's own fp32 run — network
  F, denoiser D, the 10-step DDIM sample and the router's top-k indices of every (step, layer), for the reference's
  effective init (`rg1`) and the wide-margin variant (`rg30`);
* all seven sigma schedules against the reference's values;
* every sampler's host loop (`mode_diffusion_policy_b200.gc_sampling`) against the reference's sampler run, with the
  oracle (fp32) standing in for the denoiser and the reference's recorded noise replayed — this pins the loop logic
  itself (update formulas, churn, ancestral noise, multistep history) independently of the GPU.
"""
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import mode_oracle as O
from mode_diffusion_policy_b200 import gc_sampling as S

GOLD = Path(__file__).resolve().parent / "golden"
FULL = O.ModeConfig()
TINY = O.ModeConfig(obs_dim=128, goal_dim=64, action_dim=7, embed_dim=256, n_layers=3, n_heads=4, n_state_tokens=2,
                    action_seq_len=10, num_experts=4, top_k=2)


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def _digest(sd):
    import hashlib

    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k]).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("gain", [30, 1])
def test_oracle_matches_reference_at_full_depth(gain):
    g = np.load(GOLD / f"model_full_d1024_l12_e4_rg{gain}.npz")
    B = int(g["B"])
    sd = O.make_weights(FULL, seed=1234, router_gain=float(gain))
    assert _digest(sd) == str(g["weights_sha256"])
    state, goal, x0 = O.make_inputs(FULL, B, seed=4321)
    sig = g["sigma_het"]
    acts = (x0 / np.float32(80.0)).astype(np.float32)
    F, routing = O.modedit_forward(sd, FULL, state, acts, goal, sig, "fp32", return_routing=True)
    assert rel_l2(F, g["forward_F"]) < 5e-6
    for l in range(FULL.n_layers):
        assert np.array_equal(routing[l]["idx"], g["forward_idx"][l]), l
        np.testing.assert_allclose(routing[l]["probs"], g["forward_probs"][l], atol=2e-6)
    D = O.denoiser_forward(sd, FULL, state, g["denoise_x"], goal, sig, "fp32")
    assert rel_l2(D, g["denoise_D"]) < 5e-6
    # the DDIM sample, one evaluation at a time so that every step's routing and denoiser output is compared
    x = x0.copy()
    sigmas = g["sigmas"]
    for i, (ratio, em1) in enumerate(O.ddim_coefficients(sigmas)):
        s_i = np.full(B, sigmas[i], np.float32)
        den, r = O.denoiser_forward(sd, FULL, state, x, goal, s_i, "fp32", return_routing=True)
        for l in range(FULL.n_layers):
            assert np.array_equal(r[l]["idx"], g["ddim_idx"][i, l]), (i, l)
        assert rel_l2(den, g["ddim_denoised"][i]) < 2e-5, i
        x = (ratio * x - em1 * den).astype(np.float32)
    assert rel_l2(x, g["ddim_actions"]) < 2e-5


def test_all_sigma_schedules_match_reference_values():
    g = np.load(GOLD / "schedules.npz")
    for n in (10, 25):
        got = {
            "karras": S.get_sigmas_karras(n, 1e-3, 80.0, 7, "cpu"),
            "exponential": S.get_sigmas_exponential(n, 1e-3, 80.0, "cpu"),
            "linear": S.get_sigmas_linear(n, 1e-3, 80.0, device="cpu"),
            "vp": S.get_sigmas_vp(n, device="cpu"),
            "cosine_beta": S.cosine_beta_schedule(n, device="cpu"),
            "ve": S.get_sigmas_ve(n, 1e-3, 80.0, device="cpu"),
            "iddpm": S.get_iddpm_sigmas(n, 1e-3, 80.0, device="cpu"),
        }
        for name, v in got.items():
            want = g[f"{name}_{n}"]
            assert v.dtype == torch.float32 and tuple(v.shape) == want.shape, name
            np.testing.assert_allclose(v.numpy(), want, rtol=1e-6, atol=0, err_msg=f"{name}_{n}")


class _OracleDenoiser:
    """GCDenoiser stand-in on the CPU: the fp32 numpy oracle (pinned to the reference by the goldens)."""

    def __init__(self, sd, cfg):
        self.sd, self.cfg = sd, cfg

    def __call__(self, state, x, goal, sigma, **kw):
        d = O.denoiser_forward(self.sd, self.cfg, state["state_images"].numpy(), x.numpy(), goal.numpy(),
                               sigma.numpy().astype(np.float32), "fp32")
        return torch.from_numpy(d)


class NoiseTape:
    """Replays the reference run's torch.randn_like draws (one tensor per call, in call order)."""

    def __init__(self, noise, device="cpu"):
        self.noise, self.i, self.device = noise, 0, device

    def __call__(self, like, *a, **k):
        z = torch.from_numpy(self.noise[self.i]).to(device=like.device, dtype=like.dtype)
        self.i += 1
        return z


SAMPLER_CALLS = {  # golden key -> (function name, kwargs) exactly as MoDEAgent.sample_loop calls them (mode_agent.py:796-838)
    "lms": ("sample_lms", {}), "heun": ("sample_heun", dict(s_churn=0, s_tmin=0)),
    "heun_churn": ("sample_heun", dict(s_churn=4.0, s_tmin=0)), "euler": ("sample_euler", {}),
    "euler_churn": ("sample_euler", dict(s_churn=4.0)), "ancestral": ("sample_dpm_2_ancestral", {}),
    "euler_ancestral": ("sample_euler_ancestral", {}), "dpm": ("sample_dpm_2", {}),
    "dpmpp_2s_ancestral": ("sample_dpmpp_2s_ancestral", {}), "dpmpp_2m": ("sample_dpmpp_2m", {}),
    "ddim": ("sample_ddim", {}), "dpmpp_2s": ("sample_dpmpp_2s", {}), "dpmpp_2_with_lms": ("sample_dpmpp_2_with_lms", {}),
    # the reference's default Brownian tree needs torchsde; the golden was generated with a noise_sampler that draws from
    # the recorded tape (tests/golden/make_full_goldens.py), which pins the update arithmetic and the order of the requests
    "dpmpp_2m_sde": ("sample_dpmpp_sde", dict(noise_sampler="tape")),
}


def sampler_call(key, like):
    """(function name, kwargs) of a golden key; `like`: the action tensor the tape's draws are shaped after."""
    name, kw = SAMPLER_CALLS[key]
    if kw.get("noise_sampler") == "tape":
        kw = dict(kw, noise_sampler=lambda sigma, sigma_next: torch.randn_like(like))
    return name, kw


@pytest.mark.parametrize("key", list(SAMPLER_CALLS))
def test_sampler_host_loops_match_reference_samplers(key, monkeypatch):
    g = np.load(GOLD / "samplers_tiny_d256_l3_e4.npz")
    sd = O.make_weights(TINY, seed=1234, router_gain=30.0)
    assert _digest(sd) == str(g["weights_sha256"])
    state, goal, x0 = O.make_inputs(TINY, 5, seed=4321)
    model = _OracleDenoiser(sd, TINY)
    tape = NoiseTape(g["noise_tape"])
    monkeypatch.setattr(torch, "randn_like", tape)
    name, kw = sampler_call(key, torch.from_numpy(x0))
    if key == "dpmpp_2m_sde":
        kw = dict(kw, callback=lambda d: None)  # the host loop (without a callback the sampler is one program, tested below)
    out = getattr(S, name)(model, {"state_images": torch.from_numpy(state)}, torch.from_numpy(x0), torch.from_numpy(goal),
                           torch.from_numpy(g["sigmas"]), disable=True, **kw)
    assert tape.i == int(g[key + "_draws"]), "the loop must consume the caller's RNG exactly like the reference"
    assert rel_l2(out.numpy(), g[key]) < 5e-5, rel_l2(out.numpy(), g[key])


class _ProgramOracle(_OracleDenoiser):
    """CPU interpreter of sampler programs (the semantics of the engine's `mode_sample_program`, include/mode_engine.h)
    over the fp32 oracle denoiser: checks the coefficient rows gc_sampling.py builds against the reference samplers
    without a GPU."""

    def sample_program(self, state, action, goal, sigma_eval, reads_probe, prog, noise=None):
        X, P = action.clone(), action.clone()
        H = [torch.zeros_like(action) for _ in range(4)]
        self.evals = len(sigma_eval)
        for i in range(len(sigma_eval)):
            xin = P if reads_probe[i] else X
            D = self(state, xin, goal, torch.full((action.shape[0],), float(sigma_eval[i])))
            c = [float(v) for v in prog[i]]
            v = c[0] * X + c[1] * P + c[2] * D + sum(c[3 + j] * H[j] for j in range(4))
            if noise is not None:
                v = v + c[7] * noise[i]
            if int(c[10]) >= 0:
                H[int(c[10])] = c[8] * xin + c[9] * D
            if c[11]:
                P = v
            else:
                X = v
        return X


PROGRAM_SAMPLERS = {"lms": 10, "heun": 19, "ancestral": 19, "euler_ancestral": 10, "dpm": 19, "dpmpp_2s_ancestral": 19,
                    "dpmpp_2s": 19, "dpmpp_2m_sde": 19}  # golden key -> network evaluations of the 10-step schedule


@pytest.mark.parametrize("key", list(PROGRAM_SAMPLERS))
def test_sampler_programs_match_reference_samplers(key, monkeypatch):
    """The one-launch form of Heun / DPM-2 / LMS / DPM++(2S) and the ancestral samplers: coefficient rows interpreted on
    the CPU against the reference's own sampler runs, same noise draws in the same order."""
    g = np.load(GOLD / "samplers_tiny_d256_l3_e4.npz")
    sd = O.make_weights(TINY, seed=1234, router_gain=30.0)
    state, goal, x0 = O.make_inputs(TINY, 5, seed=4321)
    model = _ProgramOracle(sd, TINY)
    tape = NoiseTape(g["noise_tape"])
    monkeypatch.setattr(torch, "randn_like", tape)
    name, kw = sampler_call(key, torch.from_numpy(x0))
    out = getattr(S, name)(model, {"state_images": torch.from_numpy(state)}, torch.from_numpy(x0), torch.from_numpy(goal),
                           torch.from_numpy(g["sigmas"]), disable=True, **kw)
    assert model.evals == PROGRAM_SAMPLERS[key]  # the program path ran (not the host loop)
    assert tape.i == int(g[key + "_draws"])
    assert rel_l2(out.numpy(), g[key]) < 5e-5, rel_l2(out.numpy(), g[key])


def test_brownian_noise_sampler_is_a_consistent_unit_variance_brownian_motion():
    """Default noise source of `sample_dpmpp_sde` (the reference's BrownianTreeNoiseSampler needs torchsde): increments
    over nested intervals add up, are normalised to unit variance, flip sign with the direction, and repeat for a seed."""
    x = torch.zeros(64, 10, 7)
    b = S.BrownianNoiseSampler(x, 1e-3, 80.0, seed=3)
    s = [torch.tensor(v) for v in (80.0, 30.0, 5.0, 0.5)]
    w01, w12, w02 = b(s[0], s[1]), b(s[1], s[2]), b(s[0], s[2])
    assert torch.allclose(w02 * (75.0 ** 0.5), w01 * (50.0 ** 0.5) + w12 * (25.0 ** 0.5), atol=1e-4)
    assert torch.allclose(b(s[1], s[0]), -w01) and torch.equal(b(s[0], s[1]), w01)
    for w in (w01, w12, w02, b(s[2], s[3])):
        assert abs(float(w.std()) - 1.0) < 0.08 and abs(float(w.mean())) < 0.08
    b2 = S.BrownianNoiseSampler(x, 1e-3, 80.0, seed=3)
    assert torch.equal(b2(s[0], s[1]), w01)
    # independent increments over disjoint intervals
    assert abs(float((w01 * w12).mean())) < 0.08
