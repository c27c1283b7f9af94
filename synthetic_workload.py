"""Synthetic CALVIN-shaped workload for bench.py and the measurement scripts: model shape, random-init weights with the
reference's parameter names and shapes, inputs, and the exponential noise schedule.

These are data generators, not a restatement of the algorithm: bench.py's engine arm must not touch `oracle/` (the
checker), so they live here. `tests/test_host_cpu.py::test_synthetic_workload_equals_the_oracles_generators` keeps them
bit-identical to the generators the parity tests use (oracle/mode_oracle.py), so the benchmark runs on the same numbers
the tests check."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

F32 = np.float32


@dataclass
class ModeConfig:
    """MoDeDiT constructor arguments that shape the path (modedit.py:643-674; conf/model/mode_agent.yaml:46-76)."""

    obs_dim: int = 2048
    goal_dim: int = 512
    action_dim: int = 7
    embed_dim: int = 1024
    n_layers: int = 12
    n_heads: int = 8
    n_state_tokens: int = 2
    action_seq_len: int = 10
    num_experts: int = 4
    top_k: int = 2
    router_normalize: bool = True
    sigma_data: float = 0.5
    rms_eps: float = 1e-6

    @property
    def seq_len(self) -> int:
        return 2 + self.n_state_tokens + self.action_seq_len


def state_dict_spec(cfg: ModeConfig):
    """Names and shapes of MoDeDiT.state_dict() (SURVEY.md §8b 'Weights contract'), in module order."""
    d, E = cfg.embed_dim, cfg.num_experts
    dh = d // cfg.n_heads
    spec = [
        ("pos_emb", (1, 1 + cfg.action_seq_len, d)),
        ("sigma_emb.weight", (d, 1)),
        ("sigma_emb.bias", (d,)),
        ("sigma_linear.weight", (d, d)),
        ("tok_emb.weight", (d, cfg.obs_dim)),
        ("gripper_embed.weight", (d, cfg.obs_dim)),
        ("goal_emb.weight", (d, cfg.goal_dim)),
        ("action_emb.weight", (d, cfg.action_dim)),
    ]
    for l in range(cfg.n_layers):
        b = f"blocks.{l}."
        spec += [
            (b + "ln_1.g", (d,)),
            (b + "attn.key.weight", (d, d)),
            (b + "attn.key.bias", (d,)),
            (b + "attn.query.weight", (d, d)),
            (b + "attn.query.bias", (d,)),
            (b + "attn.value.weight", (d, d)),
            (b + "attn.value.bias", (d,)),
            (b + "attn.c_proj.weight", (d, d)),
            (b + "attn.q_norm.g", (dh,)),
            (b + "attn.k_norm.g", (dh,)),
            (b + "ln_2.g", (d,)),
            (b + "router.router.mlp.0.weight", (2 * d, d)),
            (b + "router.router.mlp.0.bias", (2 * d,)),
            (b + "router.router.mlp.3.weight", (E, 2 * d)),
            (b + "router.router.mlp.3.bias", (E,)),
        ]
        for e in range(E):
            eb = b + f"experts.expert_{e}.mlp."
            spec += [
                (eb + "0.project.weight", (8 * d, d)),
                (eb + "0.project.bias", (8 * d,)),
                (eb + "2.weight", (d, 4 * d)),
            ]
    spec += [("ln.g", (d,)), ("out.weight", (cfg.action_dim, d)), ("out.bias", (cfg.action_dim,))]
    return spec


def make_weights_fast(cfg: ModeConfig, seed: int = 1234, router_gain: float = 30.0) -> dict:
    """Same distributions as make_weights but drawn as float32 uniforms directly (seconds instead of tens of seconds
    for the 686 M-parameter model). For timing runs, where only shapes and scales matter — never for goldens."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in state_dict_spec(cfg):
        u = rng.random(shape, dtype=F32)
        u -= F32(0.5)  # U(-0.5, 0.5)
        if name.endswith(".g"):
            w = F32(1.0) + F32(0.1) * u
        elif name == "pos_emb":
            w = F32(0.07) * u
        elif "router.router.mlp" in name:
            w = F32(0.07) * u if name.endswith("weight") else np.zeros(shape, dtype=F32)
            if name.endswith("mlp.3.weight"):
                w = w * F32(router_gain)
        else:
            fan_in = shape[1] if len(shape) == 2 else {"sigma_emb.bias": 1}.get(name, cfg.embed_dim)
            w = u * F32(2.0 / math.sqrt(fan_in))
        sd[name] = w
    return sd


def make_inputs(cfg: ModeConfig, B: int, seed: int = 4321, sigma_max: float = 80.0):
    """CALVIN-shaped synthetic inputs: state tokens, goal, initial noise x0 = randn * sigma_max."""
    rng = np.random.default_rng(seed)
    state = rng.standard_normal((B, cfg.n_state_tokens, cfg.obs_dim)).astype(F32)
    goal = rng.standard_normal((B, 1, cfg.goal_dim)).astype(F32)
    x0 = (rng.standard_normal((B, cfg.action_seq_len, cfg.action_dim)) * sigma_max).astype(F32)
    return state, goal, x0


def get_sigmas_exponential(n, sigma_min, sigma_max):
    """gc_sampling.py:35-38: exp(linspace(ln smax, ln smin, n)) ++ [0], fp32."""
    s = np.exp(np.linspace(math.log(sigma_max), math.log(sigma_min), n, dtype=F32), dtype=F32)
    return np.concatenate([s, np.zeros(1, dtype=F32)]).astype(F32)
