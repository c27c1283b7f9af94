"""TEST / BASELINE INFRASTRUCTURE ONLY — torch-CPU restatement of the reference's denoising path, op for op.

`oracle/mode_oracle.py` (numpy) is the parity checker: it restates the MATH and already exploits what the reference
does not (the router runs on B rows, not on all B*T). This module restates the reference's EXECUTION instead, so that
`bench.py --impl reference` / `cpu_baseline` time what the reference's own CPU path would spend on the host cores:
the same ATen kernels in the same structure — every `nn.Linear` as `F.linear`, the observation / goal embeddings
recomputed on every call (modedit.py:760-766), the router MLP evaluated on all B*T rows of the repeated conditioning
(modedit.py:326-336), `torch.topk` + `scatter_` masks (:391-400), one boolean-mask gather, expert MLP and `+=` scatter
per expert (:557-566), `F.scaled_dot_product_attention(is_causal=True)` (:149). The reference itself needs
/root/reference (absent on the GPU box), hence a restatement; it is pinned to the reference by the same goldens as the
numpy oracle (tests/test_oracle.py). Nothing under mode_diffusion_policy_b200/ imports this module.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def to_torch(sd_np: dict) -> dict:
    """state_dict of numpy arrays (oracle.make_weights*) -> torch CPU tensors (shared memory, no copy)."""
    return {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd_np.items()}


def rms_norm(x, g, eps=1e-6):
    """RMSNorm.forward, modedit.py:72-80: x / clamp(||x|| * dim^-0.5, eps) * g."""
    norm = torch.norm(x, dim=-1, keepdim=True) * (x.shape[-1] ** -0.5)
    return x / norm.clamp(min=eps) * g


def attention(x, sd, b, n_heads):
    """Attention.forward, modedit.py:133-167 (qk_norm, causal SDPA, c_proj without bias)."""
    B, T, C = x.shape
    hd = C // n_heads
    k = F.linear(x, sd[b + "attn.key.weight"], sd[b + "attn.key.bias"]).view(B, T, n_heads, hd).transpose(1, 2)
    q = F.linear(x, sd[b + "attn.query.weight"], sd[b + "attn.query.bias"]).view(B, T, n_heads, hd).transpose(1, 2)
    v = F.linear(x, sd[b + "attn.value.weight"], sd[b + "attn.value.bias"]).view(B, T, n_heads, hd).transpose(1, 2)
    q = rms_norm(q, sd[b + "attn.q_norm.g"])
    k = rms_norm(k, sd[b + "attn.k_norm.g"])
    y = F.scaled_dot_product_attention(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=True)
    y = y.transpose(1, 2).contiguous().view(B, T, C)
    return F.linear(y, sd[b + "attn.c_proj.weight"])


def router(x, c, sd, b, top_k, normalize=True):
    """RouterCond.forward with cond_router=True, router_context_cond_only=True (modedit.py:312-421), eval mode: the
    conditioning is repeated to every token and the MLP runs on all B*T rows."""
    B, T, _ = x.shape
    cond = c.repeat_interleave(T, dim=1) if c.shape[1] != T else c      # 'b t d -> b (t n) d'
    inp = cond.reshape(-1, cond.shape[-1])
    h = F.gelu(F.linear(inp, sd[b + "router.router.mlp.0.weight"], sd[b + "router.router.mlp.0.bias"]))
    logits = F.linear(h, sd[b + "router.router.mlp.3.weight"], sd[b + "router.router.mlp.3.bias"])
    logits = (logits - logits.max(dim=-1, keepdim=True).values) / 1.0
    probs = torch.clamp(torch.softmax(logits, dim=-1), min=1e-9, max=1 - 1e-9)
    idx = probs.topk(top_k, dim=-1).indices
    mask = torch.zeros_like(probs).scatter_(1, idx, 1)
    rprobs = torch.zeros_like(probs).scatter_(1, idx, probs.gather(1, idx))
    E = probs.shape[-1]
    mask, rprobs, idx = mask.view(B, T, E), rprobs.view(B, T, E), idx.view(B, T, top_k)
    if normalize:
        rprobs = rprobs / rprobs.sum(dim=-1, keepdim=True)
    return mask, idx, rprobs, probs.view(B, T, E)


def expert_mlp(x, sd, b, e):
    """Mlp with SwishGLU (modedit.py:83-90, :220-265): Linear(d, 8d)+bias -> projected * silu(gate) -> Linear(4d, d)."""
    p = b + f"experts.expert_{e}.mlp."
    projected, gate = F.linear(x, sd[p + "0.project.weight"], sd[p + "0.project.bias"]).tensor_split(2, dim=-1)
    return F.linear(projected * F.silu(gate), sd[p + "2.weight"])


def block(x, c, sd, layer, n_heads, num_experts, top_k):
    """NoiseBlockMoE.forward, eval mode without the fused-expert cache (modedit.py:530-595)."""
    b = f"blocks.{layer}."
    x = x + attention(rms_norm(x, sd[b + "ln_1.g"]) + c, sd, b, n_heads)
    x = rms_norm(x, sd[b + "ln_2.g"])
    mask, idx, rprobs, _ = router(x, c, sd, b, top_k)
    nxt = torch.zeros_like(x)
    for e in range(num_experts):
        tok = mask[:, :, e].bool()
        if tok.any():
            pw = rprobs[:, :, e][tok].unsqueeze(-1)
            nxt[tok] += pw * expert_mlp(x[tok], sd, b, e)
    return x + nxt, idx


def modedit_forward(sd, cfg, state, actions, goal, sigma, return_routing=False):
    """MoDeDiT.forward (modedit.py:741-821). state (B, S, obs), actions (B, A, adim), goal (B, 1, G), sigma (B,)."""
    emb_t = F.linear(F.linear((sigma.log() / 4)[:, None], sd["sigma_emb.weight"], sd["sigma_emb.bias"]),
                     sd["sigma_linear.weight"])[:, None, :]
    pos = sd["pos_emb"]
    goal_x = F.linear(goal, sd["goal_emb.weight"]) + pos[:, :1]
    state_x = F.linear(state, sd["tok_emb.weight"]) + pos[:, 1:2]
    action_x = F.linear(actions, sd["action_emb.weight"]) + pos[:, 1:]
    x = torch.cat([emb_t, goal_x, state_x, action_x], dim=1)
    routing = []
    for layer in range(cfg.n_layers):
        x, idx = block(x, emb_t, sd, layer, cfg.n_heads, cfg.num_experts, cfg.top_k)
        routing.append(idx)
    x = rms_norm(x, sd["ln.g"])
    out = F.linear(x[:, -cfg.action_seq_len:, :], sd["out.weight"], sd["out.bias"])
    return (out, routing) if return_routing else out


def denoiser_forward(sd, cfg, state, actions, goal, sigma):
    """GCDenoiser.forward (score_wrappers.py:31-43, :65-80)."""
    sd2 = cfg.sigma_data ** 2
    s = sigma.view(-1, 1, 1)
    c_skip = sd2 / (s ** 2 + sd2)
    c_out = s * cfg.sigma_data / (s ** 2 + sd2) ** 0.5
    c_in = 1 / (s ** 2 + sd2) ** 0.5
    return modedit_forward(sd, cfg, state, actions * c_in, goal, sigma) * c_out + actions * c_skip


def sample_ddim(sd, cfg, state, x, goal, sigmas):
    """sample_ddim (gc_sampling.py:922-951)."""
    s_in = x.new_ones([x.shape[0]])
    for i in range(len(sigmas) - 1):
        denoised = denoiser_forward(sd, cfg, state, x, goal, sigmas[i] * s_in)
        t, t_next = -sigmas[i].log(), -sigmas[i + 1].log()
        h = t_next - t
        x = ((-t_next).exp() / (-t).exp()) * x - (-h).expm1() * denoised
    return x


def flops_note() -> str:
    return ("router MLP on all B*T rows, tok_emb/goal_emb every call, per-expert boolean gather/scatter: the reference's "
            "execution structure (not the engine's)")

