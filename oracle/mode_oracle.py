"""CPU oracle for the MoDE denoising hot path — TEST INFRASTRUCTURE, NOT A PRODUCT PATH.

A numpy restatement of the reference algorithm (intuitive-robots/MoDE_Diffusion_Policy @ 72cf8a2). Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may import this module, and only as
the checker or the timed CPU baseline; nothing under `mode_diffusion_policy_b200/` imports it.

Parity pin: the reference ships no golden vectors for this path, so the oracle is pinned against
outputs of the reference itself: `tests/golden/make_goldens.py` imports the reference modules from /root/reference in
the build container (CPU, fp32 and bf16-autocast) and commits the tensors under tests/golden/; tests/test_oracle.py
checks every function below against them.

Two precisions:
  prec="fp32"  — the reference's standalone-eval arithmetic (everything fp32), op order as written in the reference.
  prec="bf16"  — the engine's arithmetic contract: bf16 tensor-core operands with fp32 accumulation, rounding points
                 listed in DESIGN.md §"Rounding points" (they follow the reference under torch.autocast(bfloat16),
                 SURVEY.md A.3, with the output head, embeddings of sigma/actions and the router kept in fp32).

Each function cites the reference lines it restates (paths relative to the reference root).
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

try:  # erf for nn.GELU(); scipy is present in the image, math.erf is the fallback
    from scipy.special import erf as _erf
except Exception:  # pragma: no cover
    _erf = np.vectorize(math.erf, otypes=[np.float64])

F32 = np.float32


# ----------------------------------------------------------------------------------------------------------------
# numerics helpers
def bf16_round(x: np.ndarray) -> np.ndarray:
    """Round fp32 to the nearest bfloat16 (ties to even), returned as fp32 — what `.bfloat16().float()` does."""
    x = np.ascontiguousarray(x, dtype=F32)
    u = x.view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    out = rounded.astype(np.uint32).view(F32).reshape(x.shape)
    return np.where(np.isfinite(x), out, x)


def _q(x, prec):
    return bf16_round(x) if prec == "bf16" else np.asarray(x, dtype=F32)


def rmsnorm(x, g, eps=1e-6):
    """RMSNorm.forward, mode/models/networks/modedit.py:72-80."""
    x = np.asarray(x, dtype=F32)
    scale = F32(x.shape[-1] ** -0.5)
    norm = np.sqrt(np.sum(x * x, axis=-1, keepdims=True, dtype=F32)).astype(F32) * scale
    return (x / np.maximum(norm, F32(eps)) * g.astype(F32)).astype(F32)


def silu(x):
    x = np.asarray(x, dtype=F32)
    return (x / (F32(1) + np.exp(-x, dtype=F32))).astype(F32)


def gelu_erf(x):
    """nn.GELU() default (erf form), modedit.py:196."""
    x = np.asarray(x, dtype=F32)
    return (F32(0.5) * x * (F32(1) + _erf(x.astype(np.float64) * 0.7071067811865476).astype(F32))).astype(F32)


def linear(x, w, b=None, exact=False):
    """x @ w.T (+ b). fp32 like ATen by default; `exact=True` accumulates in float64 and rounds once to fp32 — the
    bf16 contract's definition of a tensor-core GEMM with fp32 accumulation (the exact sum, rounded), so that the
    checker adds no accumulation-order noise of its own."""
    x = np.asarray(x)
    lead = x.shape[:-1]
    dt = np.float64 if exact else F32
    x2 = np.ascontiguousarray(x.reshape(-1, x.shape[-1]), dtype=dt)  # one big GEMM instead of a batch of small ones
    y = np.matmul(x2, np.asarray(w, dtype=dt).T)
    if b is not None:
        y = y + b.astype(dt)
    return y.astype(F32).reshape(*lead, -1)


# ----------------------------------------------------------------------------------------------------------------
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


def make_weights(cfg: ModeConfig, seed: int = 1234, router_gain: float = 1.0) -> dict:
    """Synthetic weights with the reference's *effective* init: nn.Linear U(+-1/sqrt(fan_in)),
    router MLP N(0, 0.02) with zero bias, norm gains near 1, pos_emb N(0, 0.02). `router_gain` scales the last router
    layer so that top-k margins are far above rounding noise. Counter-based numpy generator: both the oracle and the
    engine load exactly these arrays."""
    rng = np.random.default_rng(seed)
    sd = {}
    for name, shape in state_dict_spec(cfg):
        if name.endswith(".g"):
            w = 1.0 + 0.05 * rng.standard_normal(shape)
        elif name == "pos_emb":
            w = 0.02 * rng.standard_normal(shape)
        elif "router.router.mlp" in name:
            w = 0.02 * rng.standard_normal(shape) if name.endswith("weight") else np.zeros(shape)
            if name.endswith("mlp.3.weight"):
                w = w * router_gain
        else:
            fan_in = shape[1] if len(shape) == 2 else {
                "sigma_emb.bias": 1, "out.bias": cfg.embed_dim}.get(name, cfg.embed_dim)
            bound = 1.0 / math.sqrt(fan_in)
            w = rng.uniform(-bound, bound, size=shape)
        sd[name] = np.ascontiguousarray(w, dtype=F32)
    return sd


def make_inputs(cfg: ModeConfig, B: int, seed: int = 4321, sigma_max: float = 80.0):
    """CALVIN-shaped synthetic inputs: state tokens, goal, initial noise x0 = randn * sigma_max."""
    rng = np.random.default_rng(seed)
    state = rng.standard_normal((B, cfg.n_state_tokens, cfg.obs_dim)).astype(F32)
    goal = rng.standard_normal((B, 1, cfg.goal_dim)).astype(F32)
    x0 = (rng.standard_normal((B, cfg.action_seq_len, cfg.action_dim)) * sigma_max).astype(F32)
    return state, goal, x0


# ----------------------------------------------------------------------------------------------------------------
# router
def sigma_embedding(sd, sigma, prec="fp32"):
    """MoDeDiT.process_sigma_embeddings, modedit.py:823-832: sigma_linear(sigma_emb(ln(sigma)/4)) -> (B, d).

    bf16 mode restates the engine's affine collapse emb_t = s*u + v with u = W2 w1, v = W2 b1 (fp64 -> fp32)."""
    sigma = np.asarray(sigma, dtype=F32).reshape(-1)
    s = (np.log(sigma, dtype=F32) / F32(4)).astype(F32)
    w1, b1, w2 = sd["sigma_emb.weight"], sd["sigma_emb.bias"], sd["sigma_linear.weight"]
    if prec == "fp32":
        e1 = (s[:, None] * w1[:, 0][None, :] + b1[None, :]).astype(F32)
        return linear(e1, w2)
    u = (w2.astype(np.float64) @ w1[:, 0].astype(np.float64)).astype(F32)
    v = (w2.astype(np.float64) @ b1.astype(np.float64)).astype(F32)
    return (s[:, None].astype(np.float64) * u[None, :] + v[None, :]).astype(F32)


def router_forward(sd, layer, c, cfg: ModeConfig, prec="fp32", sigma=None):
    """RouterCond.forward on cond only (modedit.py:312-421; CondRouterMLP :170-217), eval mode.

    c: (B, d). Returns dict(idx (B,k) int64 in torch.topk order, w (B,k) renormalised, probs (B,E), logits (B,E)).
    Ties: lowest expert index first (the engine's documented rule; torch.topk's tie order is backend-defined)."""
    p = f"blocks.{layer}.router.router.mlp."
    W1, b1, W2, b2 = sd[p + "0.weight"], sd[p + "0.bias"], sd[p + "3.weight"], sd[p + "3.bias"]
    if prec == "bf16" and sigma is not None:
        sigma = np.asarray(sigma, dtype=F32).reshape(-1)
        s = (np.log(sigma, dtype=F32) / F32(4)).astype(np.float64)
        w1s, b1s, w2s = sd["sigma_emb.weight"], sd["sigma_emb.bias"], sd["sigma_linear.weight"]
        u = (w2s.astype(np.float64) @ w1s[:, 0].astype(np.float64)).astype(F32).astype(np.float64)
        v = (w2s.astype(np.float64) @ b1s.astype(np.float64)).astype(F32).astype(np.float64)
        ra = (W1.astype(np.float64) @ u).astype(F32).astype(np.float64)
        rb = (W1.astype(np.float64) @ v + b1.astype(np.float64)).astype(F32).astype(np.float64)
        z = (s[:, None] * ra[None, :] + rb[None, :]).astype(F32)
    else:
        z = linear(c, W1, b1)
    hdn = gelu_erf(z)
    logits = linear(hdn, W2, b2)
    logits = (logits - logits.max(axis=-1, keepdims=True)).astype(F32)  # / temperature (1.0), modedit.py:345
    ex = np.exp(logits, dtype=F32)
    probs = (ex / ex.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)
    probs = np.clip(probs, F32(1e-9), F32(1 - 1e-9)).astype(F32)  # modedit.py:349
    k = cfg.top_k
    order = np.argsort(-probs, axis=-1, kind="stable")[:, :k]  # stable: equal probs keep ascending index
    top = np.take_along_axis(probs, order, axis=-1)
    w = top / top.sum(axis=-1, keepdims=True, dtype=F32) if cfg.router_normalize else top  # modedit.py:418-419
    return {"idx": order.astype(np.int64), "w": w.astype(F32), "probs": probs, "logits": logits}


def topk_margin(probs: np.ndarray, k: int) -> float:
    """Smallest gap between the k-th and (k+1)-th probability over the batch (inf if k == E)."""
    if probs.shape[-1] <= k:
        return float("inf")
    s = -np.sort(-probs, axis=-1)
    return float((s[:, k - 1] - s[:, k]).min())


# ----------------------------------------------------------------------------------------------------------------
# block
def attention(h, sd, layer, cfg: ModeConfig, prec="fp32"):
    """Attention.forward, modedit.py:133-167 (causal SDPA, per-head RMSNorm on q and k); h = ln_1(x) + c, (B,T,d)."""
    p = f"blocks.{layer}.attn."
    B, T, d = h.shape
    H = cfg.n_heads
    dh = d // H
    hq = _q(h, prec)
    ex = prec == "bf16"
    wq = lambda n: _q(sd[p + n + ".weight"], prec)  # noqa: E731
    q = _q(linear(hq, wq("query"), sd[p + "query.bias"], ex), prec)
    k = _q(linear(hq, wq("key"), sd[p + "key.bias"], ex), prec)
    v = _q(linear(hq, wq("value"), sd[p + "value.bias"], ex), prec)
    split = lambda t: t.reshape(B, T, H, dh).transpose(0, 2, 1, 3)  # noqa: E731
    q, k, v = split(q), split(k), split(v)
    q = _q(rmsnorm(q, sd[p + "q_norm.g"], cfg.rms_eps), prec)
    k = _q(rmsnorm(k, sd[p + "k_norm.g"], cfg.rms_eps), prec)
    if ex:
        s = np.matmul(q.astype(np.float64), k.transpose(0, 1, 3, 2).astype(np.float64)).astype(F32)
    else:
        s = np.matmul(q, k.transpose(0, 1, 3, 2)).astype(F32)
    s = s * F32(dh ** -0.5)
    mask = np.tril(np.ones((T, T), dtype=bool))
    s = np.where(mask, s, F32(-np.inf))
    m = s.max(axis=-1, keepdims=True)
    pexp = np.exp(s - m, dtype=F32)
    den = pexp.sum(axis=-1, keepdims=True, dtype=F32)
    if prec == "bf16":  # flash-style: unnormalised P rounded to bf16 for PV, fp32 row sum
        o = np.matmul(bf16_round(pexp).astype(np.float64), v.astype(np.float64)).astype(F32) / den
    else:
        o = np.matmul((pexp / den).astype(F32), v).astype(F32)
    o = _q(o.transpose(0, 2, 1, 3).reshape(B, T, d), prec)
    return linear(o, _q(sd[p + "c_proj.weight"], prec), None, ex)  # c_proj has no bias; resid_pdrop = 0


def expert_mlp(x, sd, layer, e, prec="fp32"):
    """Mlp.forward with SwishGLU (modedit.py:83-90, :246-255): Linear(d,8d)+b -> proj*silu(gate) -> Linear(4d,d)."""
    p = f"blocks.{layer}.experts.expert_{e}.mlp."
    ex = prec == "bf16"
    z = linear(_q(x, prec), _q(sd[p + "0.project.weight"], prec), sd[p + "0.project.bias"], ex)
    half = z.shape[-1] // 2
    hdn = _q(z[..., :half] * silu(z[..., half:]), prec)
    return _q(linear(hdn, _q(sd[p + "2.weight"], prec), None, ex), prec)


def block_forward(x, c, sd, layer, cfg: ModeConfig, prec="fp32", sigma=None, return_routing=False):
    """NoiseBlockMoE.forward(x, c), eval mode, modedit.py:530-595. x (B,T,d), c (B,d)."""
    b = f"blocks.{layer}."
    x = np.asarray(x, dtype=F32)
    B, T, d = x.shape
    h = (rmsnorm(x, sd[b + "ln_1.g"], cfg.rms_eps) + c[:, None, :]).astype(F32)
    x1 = (x + attention(h, sd, layer, cfg, prec)).astype(F32)
    xn = rmsnorm(x1, sd[b + "ln_2.g"], cfg.rms_eps)  # residual stream replaced by its norm, modedit.py:539
    r = router_forward(sd, layer, c, cfg, prec, sigma)
    nxt = np.zeros_like(xn)
    for e in range(cfg.num_experts):  # ascending expert order, modedit.py:561-566
        hit = r["idx"] == e  # (B, k)
        rows = np.nonzero(hit.any(axis=-1))[0]
        if rows.size == 0:
            continue
        pw = (r["w"] * hit).sum(axis=-1)[rows].astype(F32)  # routing prob of expert e for those samples
        y = expert_mlp(xn[rows], sd, layer, e, prec)
        nxt[rows] = (nxt[rows] + pw[:, None, None] * y).astype(F32)
    out = (xn + nxt).astype(F32)
    return (out, r) if return_routing else out


# ----------------------------------------------------------------------------------------------------------------
# network, preconditioner, samplers
def modedit_forward(sd, cfg: ModeConfig, state, actions, goal, sigma, prec="fp32", return_routing=False):
    """MoDeDiT.forward, eval mode (modedit.py:741-821, build_input_seq :847-860).

    state (B,S,obs), actions (B,A,adim), goal (B,1,G) or (B,G), sigma (B,) -> (B,A,adim)."""
    state = np.asarray(state, dtype=F32)
    actions = np.asarray(actions, dtype=F32)
    goal = np.asarray(goal, dtype=F32)
    if goal.ndim == 2:
        goal = goal[:, None, :]
    B = actions.shape[0]
    sigma = np.broadcast_to(np.asarray(sigma, dtype=F32).reshape(-1), (B,)).copy()
    A = cfg.action_seq_len
    pos = sd["pos_emb"]
    emb_t = sigma_embedding(sd, sigma, prec)  # (B, d)
    ex = prec == "bf16"
    goal_x = linear(_q(goal, prec), _q(sd["goal_emb.weight"], prec), None, ex) + pos[:, 0:1]
    state_x = linear(_q(state, prec), _q(sd["tok_emb.weight"], prec), None, ex) + pos[:, 1:2]
    act_x = linear(actions, sd["action_emb.weight"]) + pos[:, 1:1 + A]
    x = np.concatenate([emb_t[:, None, :], goal_x, state_x, act_x], axis=1).astype(F32)
    routing = []
    for l in range(cfg.n_layers):
        x, r = block_forward(x, emb_t, sd, l, cfg, prec, sigma, return_routing=True)
        routing.append(r)
    x = rmsnorm(x, sd["ln.g"], cfg.rms_eps)
    out = linear(x[:, -A:, :], sd["out.weight"], sd["out.bias"])
    return (out, routing) if return_routing else out


def get_scalings(sigma, sigma_data):
    """GCDenoiser.get_scalings, mode/models/edm_diffusion/score_wrappers.py:31-43."""
    sigma = np.asarray(sigma, dtype=F32)
    sd2 = F32(sigma_data) * F32(sigma_data)
    s2 = (sigma * sigma + sd2).astype(F32)
    c_skip = (sd2 / s2).astype(F32)
    c_out = (sigma * F32(sigma_data) / np.sqrt(s2)).astype(F32)
    c_in = (F32(1) / np.sqrt(s2)).astype(F32)
    return c_skip, c_out, c_in


def denoiser_forward(sd, cfg, state, actions, goal, sigma, prec="fp32", return_routing=False):
    """GCDenoiser.forward, score_wrappers.py:65-80: inner(c_in * x) * c_out + x * c_skip."""
    B = actions.shape[0]
    sigma = np.broadcast_to(np.asarray(sigma, dtype=F32).reshape(-1), (B,)).copy()
    c_skip, c_out, c_in = (t[:, None, None] for t in get_scalings(sigma, cfg.sigma_data))
    f = modedit_forward(sd, cfg, state, (actions * c_in).astype(F32), goal, sigma, prec, return_routing=return_routing)
    routing = None
    if return_routing:
        f, routing = f
    d = ((f * c_out).astype(F32) + (actions * c_skip).astype(F32)).astype(F32)
    return (d, routing) if return_routing else d


def denoiser_loss(sd, cfg, state, action, goal, noise, sigma, prec="fp32"):
    """GCDenoiser.loss, score_wrappers.py:45-63 (eval-mode network: no dropout, top-k routing)."""
    sigma = np.asarray(sigma, dtype=F32).reshape(-1)
    c_skip, c_out, c_in = (t[:, None, None] for t in get_scalings(sigma, cfg.sigma_data))
    noised = (action + (noise * sigma[:, None, None]).astype(F32)).astype(F32)
    f = modedit_forward(sd, cfg, state, (noised * c_in).astype(F32), goal, sigma, prec)
    target = ((action - c_skip * noised) / c_out).astype(F32)
    return F32(np.mean(((f - target).astype(F32) ** 2).reshape(len(sigma), -1), dtype=np.float64)), f


def get_sigmas_exponential(n, sigma_min, sigma_max):
    """gc_sampling.py:35-38: exp(linspace(ln smax, ln smin, n)) ++ [0], fp32."""
    s = np.exp(np.linspace(math.log(sigma_max), math.log(sigma_min), n, dtype=F32), dtype=F32)
    return np.concatenate([s, np.zeros(1, dtype=F32)]).astype(F32)


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0):
    """gc_sampling.py:26-32."""
    ramp = np.linspace(0, 1, n, dtype=F32)
    mn, mx = F32(sigma_min ** (1 / rho)), F32(sigma_max ** (1 / rho))
    s = ((mx + ramp * (mn - mx)) ** F32(rho)).astype(F32)
    return np.concatenate([s, np.zeros(1, dtype=F32)]).astype(F32)


def get_sigmas_linear(n, sigma_min, sigma_max):
    """gc_sampling.py:41-44."""
    return np.concatenate([np.linspace(sigma_max, sigma_min, n, dtype=F32), np.zeros(1, dtype=F32)]).astype(F32)


def ddim_coefficients(sigmas):
    """Per-step (sigma_next/sigma, expm1(-h)) of sample_ddim, gc_sampling.py:946-950, in fp32 op order."""
    out = []
    with np.errstate(divide="ignore"):
        for i in range(len(sigmas) - 1):
            t, tn = -np.log(F32(sigmas[i])), -np.log(F32(sigmas[i + 1]))
            h = F32(tn - t)
            out.append((F32(np.exp(-tn) / np.exp(-t)), F32(np.expm1(-h))))
    return out


def sample_ddim(sd, cfg, state, x, goal, sigmas, prec="fp32", trace=None):
    """sample_ddim (DPM-Solver-1), gc_sampling.py:922-951, over GCDenoiser.forward."""
    x = np.asarray(x, dtype=F32)
    B = x.shape[0]
    for i, (ratio, em1) in enumerate(ddim_coefficients(sigmas)):
        den = denoiser_forward(sd, cfg, state, x, goal, np.full((B,), sigmas[i], dtype=F32), prec)
        x = ((ratio * x).astype(F32) - (em1 * den).astype(F32)).astype(F32)
        if trace is not None:
            trace.append(x.copy())
    return x


def sample_euler(sd, cfg, state, x, goal, sigmas, prec="fp32"):
    """sample_euler with s_churn = 0, gc_sampling.py:164-205: d = (x - D)/sigma; x += d * (sigma_next - sigma)."""
    x = np.asarray(x, dtype=F32)
    B = x.shape[0]
    for i in range(len(sigmas) - 1):
        s = F32(sigmas[i])
        den = denoiser_forward(sd, cfg, state, x, goal, np.full((B,), s, dtype=F32), prec)
        d = ((x - den) / s).astype(F32)
        x = (x + d * F32(sigmas[i + 1] - s)).astype(F32)
    return x
