"""TEST INFRASTRUCTURE ONLY — numpy restatement of the reference's TRAIN-mode forward with the engine's random bits.

Eval-mode semantics live in oracle/mode_oracle.py. In train mode the reference adds: `mask_cond` on the
goal (modedit.py:882-893), `nn.Dropout(embed_pdrob)` on the goal / image / action token embeddings (:779-784), dropout on
the attention probabilities inside SDPA (:149), `nn.Dropout(mlp_pdrop)` between SwishGLU and each expert's down
projection (:254), and per-token `torch.multinomial(probs, k, replacement=False)` routing on the repeated conditioning
(:328-330, :389-390). The masks and draws come from oracle/mode_rng.py, the numpy restatement of csrc/rng.cuh, so this
forward can be compared on the CPU with the goldens the REFERENCE produced under the same masks
(tests/golden/make_train_goldens.py::golden_train_stochastic) — pinning the mask conventions independently of any GPU.
fp32 only (the goldens are the reference's fp32 CPU run)."""
from __future__ import annotations

import numpy as np

from . import mode_oracle as O
from . import mode_rng as R

F32 = np.float32


def _attention_train(h, sd, layer, cfg, keep, p_attn):
    """Attention.forward with dropout on the softmax probabilities (keep / (1 - p)), fp32."""
    p = f"blocks.{layer}.attn."
    B, T, d = h.shape
    H = cfg.n_heads
    dh = d // H
    q = O.linear(h, sd[p + "query.weight"], sd[p + "query.bias"])
    k = O.linear(h, sd[p + "key.weight"], sd[p + "key.bias"])
    v = O.linear(h, sd[p + "value.weight"], sd[p + "value.bias"])
    split = lambda t: t.reshape(B, T, H, dh).transpose(0, 2, 1, 3)  # noqa: E731
    q, k, v = split(q), split(k), split(v)
    q = O.rmsnorm(q, sd[p + "q_norm.g"], cfg.rms_eps)
    k = O.rmsnorm(k, sd[p + "k_norm.g"], cfg.rms_eps)
    s = np.matmul(q, k.transpose(0, 1, 3, 2)).astype(F32) * F32(dh ** -0.5)
    s = np.where(np.tril(np.ones((T, T), dtype=bool)), s, F32(-np.inf))
    pexp = np.exp(s - s.max(axis=-1, keepdims=True), dtype=F32)
    att = (pexp / pexp.sum(axis=-1, keepdims=True, dtype=F32)).astype(F32)
    if p_attn > 0:
        att = (att * keep.astype(F32) / F32(1.0 - p_attn)).astype(F32)
    o = np.matmul(att, v).astype(F32).transpose(0, 2, 1, 3).reshape(B, T, d)
    return O.linear(o, sd[p + "c_proj.weight"])


def modedit_forward_train(sd, cfg: O.ModeConfig, state, actions, goal, sigma, seed, step, p_attn=0.0, p_mlp=0.0,
                          p_goal=0.0, p_embed=0.0, multinomial=True):
    """MoDeDiT.forward in train mode under the engine's masks. Returns (F (B, A, adim), [per-layer draws (B*T, k)])."""
    state, actions = np.asarray(state, dtype=F32), np.asarray(actions, dtype=F32)
    goal = np.asarray(goal, dtype=F32)
    if goal.ndim == 2:
        goal = goal[:, None, :]
    B, A, T, E, K, d = actions.shape[0], cfg.action_seq_len, cfg.seq_len, cfg.num_experts, cfg.top_k, cfg.embed_dim
    Fh = 4 * d
    sigma = np.broadcast_to(np.asarray(sigma, dtype=F32).reshape(-1), (B,)).copy()
    if p_goal > 0:
        goal = (goal * R.goal_keep_mask(seed, step, B, goal.shape[-1], p_goal)[:, None, :]).astype(F32)
    pos = sd["pos_emb"]
    emb_t = O.sigma_embedding(sd, sigma, "fp32")
    goal_x = O.linear(goal, sd["goal_emb.weight"]) + pos[:, 0:1]
    state_x = O.linear(state, sd["tok_emb.weight"]) + pos[:, 1:2]
    act_x = O.linear(actions, sd["action_emb.weight"]) + pos[:, 1:1 + A]
    x = np.concatenate([emb_t[:, None, :], goal_x, state_x, act_x], axis=1).astype(F32)
    if p_embed > 0:
        x = (x * R.embed_keep_mask(seed, step, B, T, d, p_embed)).astype(F32)
        x[:, 1:, :] /= F32(1.0 - p_embed)  # the sigma token is not dropped (its mask row is all True)
    draws = []
    for layer in range(cfg.n_layers):
        b = f"blocks.{layer}."
        h = (O.rmsnorm(x, sd[b + "ln_1.g"], cfg.rms_eps) + emb_t[:, None, :]).astype(F32)
        keep = R.attn_keep_mask(seed, step, layer, B, cfg.n_heads, T, p_attn) if p_attn > 0 else None
        x1 = (x + _attention_train(h, sd, layer, cfg, keep, p_attn)).astype(F32)
        xn = O.rmsnorm(x1, sd[b + "ln_2.g"], cfg.rms_eps)
        r = O.router_forward(sd, layer, emb_t, cfg, "fp32")
        probs = r["probs"]  # (B, E): every token of a sample shares its row
        if multinomial:
            idx = R.multinomial_draws(seed, step, layer, probs, T, K)            # (B*T, K) draw order
        else:
            idx = np.repeat(r["idx"], T, axis=0)
        draws.append(idx)
        pt = np.repeat(probs, T, axis=0)                                          # (B*T, E)
        sel = np.take_along_axis(pt, idx, axis=1)
        w = (sel / sel.sum(axis=1, keepdims=True, dtype=F32)).astype(F32) if cfg.router_normalize else sel
        flat = xn.reshape(B * T, d)
        nxt = np.zeros_like(flat)
        for e in range(E):  # ascending expert order (modedit.py:557-566)
            hit = idx == e
            rows = np.nonzero(hit.any(axis=1))[0]
            if rows.size == 0:
                continue
            pw = (w * hit).sum(axis=1)[rows].astype(F32)
            p = b + f"experts.expert_{e}.mlp."
            z = O.linear(flat[rows], sd[p + "0.project.weight"], sd[p + "0.project.bias"])
            hdn = (z[:, :Fh] * O.silu(z[:, Fh:])).astype(F32)
            if p_mlp > 0:
                hdn = (hdn * R.mlp_keep_mask(seed, step, layer, rows, e, E, Fh, p_mlp) / F32(1.0 - p_mlp)).astype(F32)
            nxt[rows] = (nxt[rows] + pw[:, None] * O.linear(hdn, sd[p + "2.weight"])).astype(F32)
        x = (xn + nxt.reshape(B, T, d)).astype(F32)
    x = O.rmsnorm(x, sd["ln.g"], cfg.rms_eps)
    return O.linear(x[:, -A:, :], sd["out.weight"], sd["out.bias"]), draws


def denoiser_loss_train(sd, cfg, state, action, goal, noise, sigma, **kw):
    """GCDenoiser.loss (score_wrappers.py:45-63) over the train-mode network. Returns (loss, F, draws)."""
    sigma = np.asarray(sigma, dtype=F32).reshape(-1)
    c_skip, c_out, c_in = O.get_scalings(sigma, cfg.sigma_data)
    noised = (np.asarray(action, F32) + np.asarray(noise, F32) * sigma[:, None, None]).astype(F32)
    Fo, draws = modedit_forward_train(sd, cfg, state, noised * c_in[:, None, None], goal, sigma, **kw)
    target = (np.asarray(action, F32) - c_skip[:, None, None] * noised) / c_out[:, None, None]
    return float(np.mean((Fo - target).astype(np.float64) ** 2)), Fo, draws
