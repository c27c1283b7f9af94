"""Drop-in for the reference's `mode.models.networks.modedit.MoDeDiT` (reference modedit.py:641-1090).

Same constructor keyword arguments, same `state_dict()` names / shapes / order (EMA zips `state_dict().values()`
positionally, reference mode/callbacks/ema.py:96), same `forward(states, actions, goals, sigma, uncond=False)`.
The module holds the fp32 master parameters; every forward runs on the CUDA engine (libmode_engine.so) — there is no
PyTorch compute path behind it.

Swap it in with Hydra:  model.inner_model._target_: mode_diffusion_policy_b200.modedit.MoDeDiT
"""
from __future__ import annotations

import logging
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .engine import EngineConfig, ModeEngine

logger = logging.getLogger(__name__)


class RMSNorm(nn.Module):
    """Parameter holder with the reference's name (`g`); reference modedit.py:72-80."""

    def __init__(self, dim: int, eps: float = 1e-8):
        super().__init__()
        self.scale, self.eps = dim ** -0.5, eps
        self.g = nn.Parameter(torch.ones(dim))


class SwishGLU(nn.Module):
    """Holder for `project` = Linear(in, 2*out) (reference modedit.py:83-90)."""

    def __init__(self, in_dim: int, out_dim: int):
        super().__init__()
        self.act, self.project = nn.SiLU(), nn.Linear(in_dim, 2 * out_dim)


class Attention(nn.Module):
    """Parameter layout of reference modedit.py:94-131 (key/query/value with bias, c_proj without, q/k RMSNorm)."""

    def __init__(self, n_embd: int, n_head: int):
        super().__init__()
        self.key = nn.Linear(n_embd, n_embd)
        self.query = nn.Linear(n_embd, n_embd)
        self.value = nn.Linear(n_embd, n_embd)
        self.c_proj = nn.Linear(n_embd, n_embd, bias=False)
        self.n_head = n_head
        self.q_norm = RMSNorm(n_embd // n_head, eps=1e-6)
        self.k_norm = RMSNorm(n_embd // n_head, eps=1e-6)


class CondRouterMLP(nn.Module):
    """Linear(d,2d) -> GELU -> Dropout(0) -> Linear(2d,E); N(0,0.02) weights, zero bias (reference modedit.py:170-217)."""

    def __init__(self, n_embd: int, num_experts: int):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(n_embd, 2 * n_embd), nn.GELU(), nn.Dropout(0), nn.Linear(2 * n_embd, num_experts))
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.normal_(m.weight, mean=0.0, std=0.02)
                nn.init.zeros_(m.bias)


class RouterCond(nn.Module):
    def __init__(self, n_embd: int, num_experts: int, top_k: int, normalize: bool = True, use_argmax: bool = False):
        super().__init__()
        self.num_experts, self.top_k, self.normalize, self.use_argmax = num_experts, top_k, normalize, use_argmax
        self.router = CondRouterMLP(n_embd, num_experts)
        self.logits = None
        self.expert_cache = {}


class Mlp(nn.Module):
    """Expert: SwishGLU(d, 4d) -> Dropout -> Linear(4d, d, bias=False) (reference modedit.py:220-265)."""

    def __init__(self, n_embd: int, dropout: float = 0.0):
        super().__init__()
        self.mlp = nn.Sequential(SwishGLU(n_embd, 4 * n_embd), nn.Dropout(dropout), nn.Linear(4 * n_embd, n_embd, bias=False))


class NoiseBlockMoE(nn.Module):
    """Parameter layout and bookkeeping API of reference modedit.py:424-638. Compute happens in the engine; the
    usage counters live on the device and are read through the engine."""

    def __init__(self, n_embd, n_heads, mlp_pdrop, num_experts, top_k, router_normalize, use_argmax, layer_idx, owner):
        super().__init__()
        self.ln_1 = RMSNorm(n_embd, eps=1e-6)
        self.n_embd = n_embd
        self.attn = Attention(n_embd, n_heads)
        self.ln_2 = RMSNorm(n_embd, eps=1e-6)
        self.router = RouterCond(n_embd, num_experts, top_k, normalize=router_normalize, use_argmax=use_argmax)
        self.experts = nn.ModuleDict({f"expert_{i}": Mlp(n_embd, dropout=mlp_pdrop) for i in range(num_experts)})
        self.num_experts = num_experts
        self.logits = None
        self.probs = None
        self._layer_idx = layer_idx
        self._owner = [owner]  # list: keep the parent out of the module tree

    # --- reference API used by MoDEAgent's expert-usage logging (mode_agent.py:466-511)
    @property
    def total_tokens_processed(self) -> int:
        eng = self._owner[0]._engine
        return 0 if eng is None else eng.expert_usage(self._layer_idx)[1]

    @property
    def inference_expert_usage(self) -> torch.Tensor:
        return self.get_expert_usage()

    def get_expert_usage(self) -> torch.Tensor:
        eng = self._owner[0]._engine
        if eng is None:
            return torch.zeros(self.num_experts)
        return torch.from_numpy(eng.expert_usage(self._layer_idx)[0]).to(torch.float32)

    def reset_expert_usage(self) -> None:
        eng = self._owner[0]._engine
        if eng is not None:
            eng.reset_expert_usage()  # device counters are shared: resets every layer, as the caller loops over all

    def reset_expert_cache(self) -> None:
        pass  # the engine keeps no per-sigma expert cache (the reference's is keyed by float(c.mean()), modedit.py:543)

    def forward(self, x, c, context=None, custom_attn_mask=None):
        """NoiseBlockMoE.forward(x, c) through the engine's block-level entry (BASELINE.json configs[0])."""
        if self.training:
            raise NotImplementedError("MoDE engine: the block-level entry is eval-only; train-mode regularisation runs "
                                      "inside GCDenoiser.loss")
        return self._owner[0]._ensure_engine(x.shape[0]).block_forward(self._layer_idx, x, c)


class MoDeDiT(nn.Module):
    def __init__(
        self,
        obs_dim: int,
        goal_dim: int,
        device: str,
        goal_conditioned: bool,
        action_dim: int,
        embed_dim: int,
        embed_pdrob: float,
        attn_pdrop: float,
        n_layers: int,
        n_heads: int,
        goal_seq_len: int,
        obs_seq_len: int,
        action_seq_len: int,
        state_dim,
        mlp_pdrop: float = 0.1,
        goal_drop: float = 0.1,
        linear_output: bool = True,
        use_proprio: bool = False,
        cond_router: bool = True,
        num_experts: int = 4,
        top_k: int = 2,
        router_normalize: bool = True,
        use_goal_in_routing: bool = False,
        use_argmax: bool = False,
        causal: bool = True,
        use_shared_expert: bool = False,
        use_noise_token_as_input: bool = True,
        use_custom_attn_mask: bool = False,
        init_style: str = "default",
        n_state_tokens: int = 2,
        max_batch: int = 256,
        sigma_data: float = 0.5,
    ):
        super().__init__()
        unsupported = {
            "use_proprio": use_proprio, "use_shared_expert": use_shared_expert,
            "use_custom_attn_mask": use_custom_attn_mask, "use_goal_in_routing": use_goal_in_routing,
            "not goal_conditioned": not goal_conditioned, "not use_noise_token_as_input": not use_noise_token_as_input,
            "not linear_output": not linear_output, "not cond_router": not cond_router, "not causal": not causal,
            "goal_seq_len != 1": goal_seq_len != 1, "obs_seq_len != 1": obs_seq_len != 1,
        }
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f"MoDE engine: unsupported reference options {bad} (dead/disabled in conf/)")
        self.device = device
        self.obs_dim, self.embed_dim, self.num_layers = obs_dim, embed_dim, n_layers
        self.goal_seq_len, self.action_seq_len = goal_seq_len, action_seq_len
        self.num_experts, self.top_k = num_experts, top_k
        self.use_proprio = use_proprio
        self.cond_mask_prob = goal_drop
        self._train_cfg = dict(attn_pdrop=attn_pdrop, mlp_pdrop=mlp_pdrop, goal_drop=goal_drop, embed_pdrob=embed_pdrob,
                               use_argmax=use_argmax)
        self._warned_deterministic = False
        # parameters in the reference's registration order (modedit.py:680-722)
        self.sigma_emb = nn.Linear(1, embed_dim)
        self.sigma_linear = nn.Linear(embed_dim, embed_dim, bias=False)
        seq_size = goal_seq_len + obs_seq_len - 1 + action_seq_len
        self.tok_emb = nn.Linear(obs_dim, embed_dim, bias=False)
        self.gripper_embed = nn.Linear(obs_dim, embed_dim, bias=False)
        self.goal_emb = nn.Linear(goal_dim, embed_dim, bias=False)
        self.action_emb = nn.Linear(action_dim, embed_dim, bias=False)
        self.pos_emb = nn.Parameter(torch.zeros(1, seq_size, embed_dim))
        self.blocks = nn.ModuleList(
            NoiseBlockMoE(embed_dim, n_heads, mlp_pdrop, num_experts, top_k, router_normalize, use_argmax, i, self)
            for i in range(n_layers)
        )
        self.ln = RMSNorm(embed_dim, eps=1e-6)
        self.out = nn.Linear(embed_dim, action_dim)
        self.logits_per_layer = None
        self.probs_per_layer = None
        self._engine_cfg = EngineConfig(
            obs_dim=obs_dim, goal_dim=goal_dim, action_dim=action_dim, embed_dim=embed_dim, n_layers=n_layers,
            n_heads=n_heads, n_state_tokens=n_state_tokens, action_seq_len=action_seq_len, num_experts=num_experts,
            top_k=top_k, router_normalize=router_normalize, max_batch=max_batch, sigma_data=sigma_data, rms_eps=1e-6)
        self._engine: Optional[ModeEngine] = None
        self._weights_key = None

    # ------------------------------------------------------------------ engine management
    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        # reference order: pos_emb is registered as a Parameter of the root module and therefore comes first
        return sd

    def _weights_fingerprint(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _ensure_engine(self, batch: int, force_repack: bool = False) -> ModeEngine:
        """The engine for this module with its packed weights up to date. Changes are detected through the parameters'
        (data_ptr, _version) pairs; `force_repack` is for callers that cannot rely on version counters (a training step
        after a fused/foreach optimizer, which updates parameters without bumping them)."""
        cfg = self._engine_cfg
        if self._engine is None or batch > self._engine.cfg.max_batch:
            if self._engine is not None:
                self._refuse_if_train_state(f"batch {batch} exceeds max_batch={self._engine.cfg.max_batch}")
                self._engine.close()
            cfg.max_batch = max(cfg.max_batch, batch)
            self._engine = ModeEngine(cfg)
            self._weights_key = None
        key = self._weights_fingerprint()
        if key != self._weights_key or force_repack:
            self._engine.load_state_dict({k: v for k, v in self.state_dict().items()})
            self._weights_key = key
        return self._engine

    def _refuse_if_train_state(self, why: str) -> None:
        """The engine owns the Adam moments, the EMA buffer and the flat gradient buffer (and torch holds zero-copy views
        of them): re-creating it mid-training would silently reset the optimizer and leave those views dangling."""
        if self._engine is not None and self._engine.has_train_state:
            raise RuntimeError(
                f"MoDE engine: {why}, which needs a new engine, but this one holds training state (optimizer moments / "
                "EMA / gradient views). Construct MoDeDiT(max_batch=...) for the largest batch you will use (validation "
                "included) and set sigma_data before the first training step.")

    @property
    def _param_names(self):
        return [n for n, _ in self.named_parameters()]

    def check_trainable(self) -> None:
        """Everything the reference's train mode does is built into the engine (attention / expert / embedding dropout,
        goal masking, per-token multinomial routing); kept for callers of the earlier interface."""

    # ---- stochastic regularisation of the reference's train mode (modedit.py:149, :254, :389-390, :882-893)
    def set_train_rng(self, seed: int, step: int = 0) -> None:
        """Position of the engine's counter-based random stream: the masks and expert draws of training step n are a
        pure function of (seed, step + n). Default seed: torch.initial_seed() mixed with the data-parallel rank."""
        self._train_seed, self._train_step = int(seed), int(step)

    def _train_rng(self):
        if getattr(self, "_train_seed", None) is None:
            rank = torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
            self.set_train_rng((torch.initial_seed() * 1000003 + rank) & (2 ** 64 - 1), 0)
        return self._train_seed, self._train_step

    def _stochastic_args(self):
        """Arguments of ModeEngine.set_stochastic for the next training step (all off when `deterministic_training`)."""
        c = self._train_cfg
        if getattr(self, "deterministic_training", False):
            return dict(attn_pdrop=0.0, mlp_pdrop=0.0, goal_drop=0.0, embed_pdrop=0.0, multinomial=False, seed=0, step=0)
        seed, step = self._train_rng()
        return dict(attn_pdrop=c["attn_pdrop"], mlp_pdrop=c["mlp_pdrop"], goal_drop=c["goal_drop"],
                    embed_pdrop=c["embed_pdrob"], multinomial=not c["use_argmax"], seed=seed, step=step)

    def _advance_train_rng(self):
        if getattr(self, "_train_seed", None) is not None:
            self._train_step += 1

    def set_sigma_data(self, sigma_data: float) -> None:
        if float(sigma_data) != self._engine_cfg.sigma_data:
            if self._engine is not None:
                self._refuse_if_train_state("sigma_data changed")
                self._engine.close()
                self._engine = None
            self._engine_cfg.sigma_data = float(sigma_data)

    # ------------------------------------------------------------------ reference surface
    def _goals(self, goals, uncond):
        """preprocess_goals (reference modedit.py:862-880), eval mode."""
        if goals.dim() == 2:
            goals = goals.unsqueeze(1)
        if goals.shape[-1] == 2 * self.obs_dim:
            goals = goals[:, :, : self.obs_dim]
        if uncond:
            goals = torch.zeros_like(goals)
        return goals

    def forward(self, states, actions, goals, sigma, uncond: Optional[bool] = False):
        if self.training:
            raise NotImplementedError("MoDE engine: in train mode the network runs inside GCDenoiser.loss (fused forward + "
                                      "backward with dropout / multinomial routing); call .eval() for a plain forward")
        eng = self._ensure_engine(actions.shape[0])
        goals = self._goals(goals, uncond)
        sigma = torch.as_tensor(sigma, device=actions.device)
        return eng.forward(states["state_images"], actions, goals, sigma).to(actions.dtype)

    def get_params(self):
        return self.parameters()

    def precompute_experts_for_inference(self, sigma, goal=None):
        """No-op: the engine routes on the device every step (reference modedit.py:971-992 builds a per-sigma cache)."""

    def reset_all_caches(self):
        for b in self.blocks:
            b.reset_expert_cache()

    def freeze_router(self):
        for layer in self.blocks:
            layer.router.eval()
            for p in layer.router.parameters():
                p.requires_grad = False

    def unfreeze_router(self):
        for layer in self.blocks:
            layer.router.train()
            for p in layer.router.parameters():
                p.requires_grad = True

    def get_router_states(self):
        return [{"layer": i, "frozen": not any(p.requires_grad for p in l.router.parameters()),
                 "eval_mode": not l.router.training, "cache_size": 0} for i, l in enumerate(self.blocks)]

    def prepare_for_finetuning(self, freeze_routers: bool = True, freeze_expert_weights: float = 0.3,
                               reset_expert_stats: bool = True):
        if freeze_routers:
            self.freeze_router()

    # ---- auxiliary router losses (reference modedit.py:584-593, :898-969; weights are 0.0 in conf/model/mode_agent.yaml)
    def _router_aux(self):
        """Per-layer (shifted logits, clamped softmax) of the last `loss` call's sigma rows, recomputed with torch ops
        from the fp32 masters so they carry autograd history to the router / sigma-embedding parameters. The router
        sees only the sigma embedding (reference modedit.py:330-331), so this is B rows x a d->2d->E MLP per layer:
        not hot-path work. The top-k mask itself comes from the engine (`routing`), so both agree on the selection."""
        sigma = getattr(self, "_last_sigma", None)
        if sigma is None:
            raise RuntimeError("MoDE engine: auxiliary router losses need a preceding GCDenoiser.loss call")
        s = (sigma.float().log() / 4).reshape(-1, 1)
        c = F.linear(F.linear(s, self.sigma_emb.weight, self.sigma_emb.bias), self.sigma_linear.weight)
        outs = []
        for blk in self.blocks:
            mlp = blk.router.router.mlp
            logits = F.linear(F.gelu(F.linear(c, mlp[0].weight, mlp[0].bias)), mlp[3].weight, mlp[3].bias)
            logits = logits - logits.max(dim=-1, keepdim=True).values  # / temperature (1.0), reference :345
            probs = torch.clamp(torch.softmax(logits, dim=-1), min=1e-9, max=1 - 1e-9)
            outs.append((logits, probs))
        return outs

    def load_balancing_loss(self):
        """Mean over layers of E * sum_e mean(router_probs_e) * mean(mask_e) over all tokens (reference modedit.py:584-593,
        :898-928). With arg-max routing all T tokens of a sample share one routing row; under multinomial routing every
        token has its own draws (the engine's token-level table of the last training step)."""
        total = 0.0
        aux = self._router_aux()
        B = aux[0][0].shape[0]
        per_token = bool(getattr(self, "_last_step_multinomial", False))
        for layer, (_, probs) in enumerate(aux):
            if per_token:
                T = self._engine.cfg.seq_len
                idx = torch.from_numpy(self._engine.token_routing(layer, B)[0]).long().to(probs.device)
                probs = probs.repeat_interleave(T, dim=0)  # cond is repeated to every token (reference :328-330)
            else:
                idx = torch.from_numpy(self._engine.routing(layer, B)[0]).long().to(probs.device)
            mask = torch.zeros_like(probs).scatter_(1, idx, 1.0)
            rp = probs * mask
            if self.blocks[layer].router.normalize:
                rp = rp / rp.sum(dim=-1, keepdim=True)
            total = total + probs.shape[-1] * (rp.mean(dim=0) * mask.mean(dim=0)).sum()
        return total / len(aux)

    def compute_router_z_loss(self, eps=1e-6):
        """Mean over layers and rows of log(sum_e exp(logit_e) + eps)^2 on the max-shifted logits (reference :930-969)."""
        aux = self._router_aux()
        total = 0.0
        for logits, _ in aux:
            total = total + torch.log(torch.exp(logits).sum(dim=-1) + eps).pow(2).mean()
        return total / len(aux)

    def grad_norms(self):
        """Gradient monitoring of MoDEAgent.on_before_zero_grad (reference mode_agent.py:304-359) — total norm, the
        norm of the non-block ("input") layers and per-block per-layer norms of the last training step — from the
        engine's flat gradient buffer with two kernel launches and ONE device-to-host copy (the reference calls
        `.item()` several times per parameter). Works with `EngineAdamW` (no `.grad` tensors needed)."""
        eng = self._engine
        if eng is None:
            raise RuntimeError("MoDE engine: grad_norms() needs a preceding training-mode GCDenoiser.loss call")
        names = [n for n, p in self.named_parameters() if n != "gripper_embed.weight" and p.requires_grad]
        sumsq = eng.grad_sumsq([eng.grad_range(n) for n in names]).double().cpu().numpy()
        scale = getattr(self, "_loss_grad_scale", None)
        s2 = 1.0 if scale is None else float(scale) ** 2
        out = {"total": 0.0, "input_layers": 0.0, "blocks": {}}
        for n, v in zip(names, sumsq):
            v = float(v) * s2
            out["total"] += v
            if "blocks" in n:
                parts = n.split(".")
                blk = out["blocks"].setdefault(parts[1], {})
                layer = ".".join(parts[2:])
                blk[layer] = blk.get(layer, 0.0) + v
            else:
                out["input_layers"] += v
        out["total"], out["input_layers"] = out["total"] ** 0.5, out["input_layers"] ** 0.5
        for blk in out["blocks"].values():
            for k in blk:
                blk[k] = blk[k] ** 0.5
        return out

    def routing(self, layer: int, batch: int):
        """(top_k_indices, renormalised probs, clamped softmax) of the most recent call for `layer`."""
        return self._engine.routing(layer, batch)
