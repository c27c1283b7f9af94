"""Gradient goldens for the training path, from the reference's own autograd (build container only).

Deterministic training-parity mode of SURVEY.md A.5: attn_pdrop = mlp_pdrop = goal_drop = 0, use_argmax=True,
model.train(); loss = GCDenoiser.loss(state, actions, goal, noise, sigma)[0]; loss.backward() on CPU fp32.
Full gradients of even the small models are tens of MB, so each tensor is stored as (L2 norm, sum, 256 sampled
entries at indices drawn from a generator seeded by the tensor's position) — tests compare the engine's gradients at
the same indices and the norms.

    python tests/golden/make_train_goldens.py
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_goldens as MG  # noqa: E402  (stubs + reference imports)
from make_goldens import O, MoDeDiT, GCDenoiser  # noqa: E402

OUT = Path(__file__).resolve().parent
N_SAMPLES = 256


def sample_indices(numel, tensor_pos):
    rng = np.random.default_rng(10_000 + tensor_pos)
    return rng.integers(0, numel, size=min(N_SAMPLES, numel))


def golden_train(tag, cfg, B, router_gain=30.0):
    sd = O.make_weights(cfg, seed=1234, router_gain=router_gain)
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321)
    inner = MoDeDiT(obs_dim=cfg.obs_dim, goal_dim=cfg.goal_dim, device="cpu", goal_conditioned=True,
                    action_dim=cfg.action_dim, embed_dim=cfg.embed_dim, embed_pdrob=0, attn_pdrop=0.0,
                    n_layers=cfg.n_layers, n_heads=cfg.n_heads, goal_seq_len=1, obs_seq_len=1,
                    action_seq_len=cfg.action_seq_len, state_dim=7, mlp_pdrop=0.0, goal_drop=0.0,
                    num_experts=cfg.num_experts, top_k=cfg.top_k, use_argmax=True, init_style="olmoe")
    inner.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()})
    model = GCDenoiser(inner, sigma_data=cfg.sigma_data).train()
    g = np.load(OUT / f"{tag}.npz")
    acts = (x0 / np.float32(80.0)).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    st_in, goal_in = t(state).requires_grad_(True), t(goal).requires_grad_(True)  # the reference trains its encoders through these
    with torch.enable_grad():
        loss, f_out = model.loss({"state_images": st_in}, t(acts), goal_in, t(g["loss_noise"]), t(g["sigma_het"]))
        # auxiliary router losses of the same forward (modedit.py:898-969) and their gradients
        lb, zl = inner.load_balancing_loss(), inner.compute_router_z_loss()
        aux_names = [n for n, _ in inner.named_parameters() if "router" in n or n.startswith("sigma_")]
        aux_params = [dict(inner.named_parameters())[n] for n in aux_names]
        aux_grads = torch.autograd.grad(lb + zl, aux_params, retain_graph=True, allow_unused=True)
        loss.backward()
    out = {"loss": np.float32(loss.item()), "F": f_out.detach().numpy(), "aux_lb": np.float32(lb.item()),
           "aux_z": np.float32(zl.item())}
    out["d_state"] = st_in.grad.numpy().copy()
    out["d_goal"] = goal_in.grad.numpy().copy()
    for n, gr in zip(aux_names, aux_grads):
        out[f"auxnorm/{n}"] = np.float32(0.0 if gr is None else np.linalg.norm(gr.numpy().astype(np.float64)))
    assert abs(float(loss) - float(g["loss_value"])) < 1e-5 * abs(float(loss)), "train-mode loss differs from eval-mode"
    names = [n for n, _ in O.state_dict_spec(cfg)]
    params = dict(inner.named_parameters())
    for pos, name in enumerate(names):
        p = params[name]
        grad = np.zeros(p.shape, np.float32) if p.grad is None else p.grad.numpy()
        flat = grad.reshape(-1)
        idx = sample_indices(flat.size, pos)
        out[f"norm/{name}"] = np.float32(np.linalg.norm(flat.astype(np.float64)))
        out[f"sum/{name}"] = np.float32(flat.astype(np.float64).sum())
        out[f"val/{name}"] = flat[idx].astype(np.float32)
    np.savez_compressed(OUT / f"train_{tag}.npz", **out)
    unused = [n for n in names if params[n].grad is None]
    print(f"train_{tag}: loss {float(loss):.6f}; {len(names)} tensors; no-grad tensors: {len(unused)}")


# ------------------------------------------------------------------------------------------------------------------
# Stochastic training mode (the reference's default: attn_pdrop 0.3, mlp_pdrop 0.1, goal_drop 0.1, use_argmax=False).
# torch's generator cannot be reproduced by the engine, so the REFERENCE is run with the engine's masks instead: the
# four random sources are patched to read oracle/mode_rng.py (the numpy restatement of csrc/rng.cuh) —
#   torch.bernoulli (MoDeDiT.mask_cond, modedit.py:888), F.scaled_dot_product_attention's dropout_p (:149),
#   nn.Dropout inside every expert Mlp (:254), torch.multinomial (RouterCond, :389-390)
# — everything else (modules, autograd) is the reference's own code.
def golden_train_stochastic(tag, cfg, B, seed, step, p_attn=0.3, p_mlp=0.1, p_goal=0.1, router_gain=4.0, p_embed=0.0,
                            out_prefix="train_stoch"):
    import math

    from oracle import mode_rng as R

    sd = O.make_weights(cfg, seed=1234, router_gain=router_gain)
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321)
    inner = MoDeDiT(obs_dim=cfg.obs_dim, goal_dim=cfg.goal_dim, device="cpu", goal_conditioned=True,
                    action_dim=cfg.action_dim, embed_dim=cfg.embed_dim, embed_pdrob=p_embed, attn_pdrop=p_attn,
                    n_layers=cfg.n_layers, n_heads=cfg.n_heads, goal_seq_len=1, obs_seq_len=1,
                    action_seq_len=cfg.action_seq_len, state_dim=7, mlp_pdrop=p_mlp, goal_drop=p_goal,
                    num_experts=cfg.num_experts, top_k=cfg.top_k, use_argmax=False, init_style="olmoe")
    inner.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd.items()})
    model = GCDenoiser(inner, sigma_data=cfg.sigma_data).train()
    T, E, K, H, F = cfg.seq_len, cfg.num_experts, cfg.top_k, cfg.n_heads, 4 * cfg.embed_dim
    ctx = {"attn_calls": 0, "route_calls": 0, "embed_calls": 0, "masks": {}, "routing": {}}

    class EmbedDropout(torch.nn.Module):
        """MoDeDiT.drop (modedit.py:779-784): called on the goal token, the image tokens and the action tokens, in this
        order; rows t = 1, 2..1+S, 2+S..T-1 of the engine's [B, T, d] mask."""

        def forward(self, x):
            call = ctx["embed_calls"] % 3
            ctx["embed_calls"] += 1
            S = cfg.n_state_tokens
            lo, hi = [(1, 2), (2, 2 + S), (2 + S, T)][call]
            assert x.shape[1] == hi - lo
            keep = R.embed_keep_mask(seed, step, x.shape[0], T, cfg.embed_dim, p_embed)[:, lo:hi, :]
            return x * torch.from_numpy(keep.astype(np.float32)) / (1.0 - p_embed)

    if p_embed > 0:
        inner.drop = EmbedDropout()

    def fake_bernoulli(pt, *a, **k):
        bs, t, dd = pt.shape
        keep = R.goal_keep_mask(seed, step, bs, dd, p_goal).reshape(bs, t, dd)
        return torch.from_numpy((~keep).astype(np.float32))

    def fake_sdpa(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, **kw):
        layer = ctx["attn_calls"] % cfg.n_layers
        ctx["attn_calls"] += 1
        assert is_causal and attn_mask is None
        Bq, Hq, Tq, Dq = q.shape
        att = (q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(Dq))
        causal = torch.tril(torch.ones(Tq, Tq)).view(1, 1, Tq, Tq)
        att = torch.softmax(att.masked_fill(causal == 0, float("-inf")), dim=-1)
        if dropout_p > 0:
            keep = torch.from_numpy(R.attn_keep_mask(seed, step, layer, Bq, Hq, Tq, dropout_p).astype(np.float32))
            att = att * keep / (1.0 - dropout_p)
        return att @ v

    def fake_multinomial(probs, k, replacement=False):
        layer = ctx["route_calls"] % cfg.n_layers
        ctx["route_calls"] += 1
        assert not replacement and probs.shape[0] == B * T
        idx = R.multinomial_draws(seed, step, layer, probs.detach().numpy()[::T], T, k)
        ctx["routing"][layer] = idx
        return torch.from_numpy(idx)

    class MaskedDropout(torch.nn.Module):
        def __init__(self, layer, expert, p):
            super().__init__()
            self.layer, self.expert, self.p = layer, expert, p

        def forward(self, h):
            tokens = ctx["masks"][self.layer][:, self.expert].nonzero()[0]
            assert len(tokens) == h.shape[0]
            keep = R.mlp_keep_mask(seed, step, self.layer, tokens, self.expert, E, F, self.p)
            return h * torch.from_numpy(keep.astype(np.float32)) / (1.0 - self.p)

    def router_hook(layer):
        def hook(mod, args, out):
            ctx["masks"][layer] = out[0].detach().reshape(B * T, E).numpy() > 0
        return hook

    for li, blk in enumerate(inner.blocks):
        blk.router.register_forward_hook(router_hook(li))
        for e in range(E):
            mlp = blk.experts[f"expert_{e}"].mlp
            assert isinstance(mlp[1], torch.nn.Dropout)
            mlp[1] = MaskedDropout(li, e, p_mlp)

    g = np.load(OUT / f"{tag}.npz")
    acts = (x0 / np.float32(80.0)).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    st_in, goal_in = t(state).requires_grad_(True), t(goal).requires_grad_(True)
    saved = (torch.bernoulli, torch.nn.functional.scaled_dot_product_attention, torch.multinomial)
    torch.bernoulli, torch.nn.functional.scaled_dot_product_attention, torch.multinomial = fake_bernoulli, fake_sdpa, fake_multinomial
    try:
        with torch.enable_grad():
            loss, f_out = model.loss({"state_images": st_in}, t(acts), goal_in, t(g["loss_noise"]), t(g["sigma_het"]))
            loss.backward()
    finally:
        torch.bernoulli, torch.nn.functional.scaled_dot_product_attention, torch.multinomial = saved
    assert ctx["attn_calls"] == cfg.n_layers and ctx["route_calls"] == cfg.n_layers
    out = {"loss": np.float32(loss.item()), "F": f_out.detach().numpy(), "seed": np.int64(seed), "step": np.int64(step),
           "p": np.array([p_attn, p_mlp, p_goal], np.float32), "p_embed": np.float32(p_embed),
           "router_gain": np.float32(router_gain)}
    out["d_state"] = st_in.grad.numpy().copy()
    out["d_goal"] = goal_in.grad.numpy().copy()
    for li in range(cfg.n_layers):
        out[f"routing/{li}"] = ctx["routing"][li].astype(np.int32)
    names = [n for n, _ in O.state_dict_spec(cfg)]
    params = dict(inner.named_parameters())
    for pos, name in enumerate(names):
        p = params[name]
        grad = np.zeros(p.shape, np.float32) if p.grad is None else p.grad.numpy()
        flat = grad.reshape(-1)
        idx = sample_indices(flat.size, pos)
        out[f"norm/{name}"] = np.float32(np.linalg.norm(flat.astype(np.float64)))
        out[f"sum/{name}"] = np.float32(flat.astype(np.float64).sum())
        out[f"val/{name}"] = flat[idx].astype(np.float32)
    np.savez_compressed(OUT / f"{out_prefix}_{tag}.npz", **out)
    usage = [np.bincount(ctx["routing"][li].reshape(-1), minlength=E).tolist() for li in range(cfg.n_layers)]
    print(f"{out_prefix}_{tag}: loss {float(loss):.6f} (deterministic {float(g['loss_value']):.6f}); expert usage per layer {usage}")


TINY = MG.O.ModeConfig(obs_dim=128, goal_dim=64, action_dim=7, embed_dim=256, n_layers=3, n_heads=4, n_state_tokens=2,
                       action_seq_len=10, num_experts=4, top_k=2)
WIDE = MG.O.ModeConfig(obs_dim=64, goal_dim=64, action_dim=7, embed_dim=512, n_layers=2, n_heads=4, n_state_tokens=2,
                       action_seq_len=10, num_experts=8, top_k=2)

if __name__ == "__main__":
    torch.set_grad_enabled(True)
    golden_train("model_tiny_d256_l3_e4", TINY, 5)
    golden_train("model_wide_d512_l2_e8", WIDE, 4)
    golden_train_stochastic("model_tiny_d256_l3_e4", TINY, 5, seed=20261017, step=3)
    golden_train_stochastic("model_wide_d512_l2_e8", WIDE, 4, seed=77, step=0)
    golden_train_stochastic("model_tiny_d256_l3_e4", TINY, 5, seed=99, step=7, p_embed=0.2, out_prefix="train_stoch_embed")
