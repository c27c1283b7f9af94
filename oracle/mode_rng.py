"""TEST INFRASTRUCTURE ONLY — numpy restatement of the engine's counter-based random bits (csrc/rng.cuh).

The reference draws its training-mode randomness from torch's global generator (nn.Dropout, SDPA dropout_p,
torch.bernoulli in MoDeDiT.mask_cond, torch.multinomial in RouterCond: reference modedit.py:149, :254, :389-390, :888).
Those streams are not reproducible outside torch, so the engine defines its own stateless bits; this file restates them
bit for bit so that tests/golden/make_train_goldens.py can run the REFERENCE with exactly the masks the engine will
draw, and tests can check the masks' statistics on CPU. Nothing under mode_diffusion_policy_b200/ imports this module.
"""
from __future__ import annotations

import numpy as np

RNG_GOAL, RNG_ATTN, RNG_MLP, RNG_ROUTE, RNG_EMBED = 1, 2, 3, 4, 5
_U32 = np.uint32


def lowbias32(x):
    """csrc/rng.cuh lowbias32: a 32-bit bijection (xorshift-multiply), vectorised over uint32 arrays."""
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x.astype(_U32)


def rng_key(seed: int, step: int, stream: int, layer: int) -> int:
    k = int(lowbias32((seed & 0xFFFFFFFF) ^ 0x9E3779B9))
    k = int(lowbias32((k + ((seed >> 32) & 0xFFFFFFFF)) & 0xFFFFFFFF))
    k = int(lowbias32((k + step) & 0xFFFFFFFF))
    k = int(lowbias32((k + stream * 0x85EBCA6B + layer * 0xC2B2AE35) & 0xFFFFFFFF))
    return k


def rng_bits(key: int, idx):
    return lowbias32((np.asarray(idx, dtype=np.uint64) & 0xFFFFFFFF) ^ np.uint64(key))


def drop_threshold(p: float) -> int:
    t = np.float32(p) * np.float32(65536.0) + np.float32(0.5)
    return 0 if t <= 0 else (65535 if t >= 65535 else int(t))


def _keep_from_elements(key: int, elem_idx, p: float):
    """keep mask for flat element indices: element i uses half (i & 1) of word i >> 1."""
    elem_idx = np.asarray(elem_idx, dtype=np.uint64)
    bits = rng_bits(key, elem_idx >> np.uint64(1)).astype(np.uint64)
    half = np.where(elem_idx & np.uint64(1), bits >> np.uint64(16), bits & np.uint64(0xFFFF))
    return half >= drop_threshold(p)


def goal_keep_mask(seed, step, B, G, p):
    """[B, G] bool: True = the goal feature is kept (MoDeDiT.mask_cond zeroes the others, no rescale)."""
    key = rng_key(seed, step, RNG_GOAL, 0)
    return _keep_from_elements(key, np.arange(B * G, dtype=np.uint64), p).reshape(B, G)


def embed_keep_mask(seed, step, B, T, d, p):
    """[B, T, d] bool keep mask of the token embeddings (+pos): element (row, col) uses half (col & 1) of word
    (row*d + col) // 2; the sigma token (t = 0) is never dropped (reference modedit.py:779-784 applies self.drop to the
    goal, image and action tokens only)."""
    key = rng_key(seed, step, RNG_EMBED, 0)
    keep = _keep_from_elements(key, np.arange(B * T * d, dtype=np.uint64), p).reshape(B, T, d)
    keep[:, 0, :] = True
    return keep


def attn_keep_mask(seed, step, layer, B, H, T, p):
    """[B, H, T, T] bool keep mask of the attention probabilities: element (b, h, i, j) uses half (j & 1) of word
    ((b*H + h)*T + i) * ceil(T/2) + j // 2."""
    key = rng_key(seed, step, RNG_ATTN, layer)
    thalf = (T + 1) // 2
    bh = np.arange(B * H, dtype=np.uint64)[:, None, None]
    i = np.arange(T, dtype=np.uint64)[None, :, None]
    j = np.arange(T, dtype=np.uint64)[None, None, :]
    word = (bh * T + i) * thalf + (j >> np.uint64(1))
    bits = rng_bits(key, word).astype(np.uint64)
    half = np.where(j & np.uint64(1), bits >> np.uint64(16), bits & np.uint64(0xFFFF))
    return (half >= drop_threshold(p)).reshape(B, H, T, T)


def mlp_keep_mask(seed, step, layer, token_ids, expert, E, F, p):
    """[len(token_ids), F] bool keep mask of expert `expert`'s hidden activations h for the given token rows: hidden
    unit j of token m uses half (j & 1) of word (m*E + expert) * (F/2) + j // 2."""
    key = rng_key(seed, step, RNG_MLP, layer)
    m = np.asarray(token_ids, dtype=np.uint64)[:, None]
    j = np.arange(F, dtype=np.uint64)[None, :]
    word = (m * E + expert) * (F // 2) + (j >> np.uint64(1))
    bits = rng_bits(key, word).astype(np.uint64)
    half = np.where(j & np.uint64(1), bits >> np.uint64(16), bits & np.uint64(0xFFFF))
    return half >= drop_threshold(p)


def rng_uniform(bits):
    """23 random bits -> (0, 1): (bits >> 9 + 0.5) / 2^23, exact in fp32."""
    return ((np.asarray(bits, dtype=np.uint32) >> 9).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)


def multinomial_draws(seed, step, layer, probs, T, K):
    """torch.multinomial(probs, K, replacement=False) semantics with the engine's bits: token m = b*T + t draws K experts
    one after the other, draw k inverting the CDF of the REMAINING clamped probabilities (ascending expert index) at
    u * S with u = rng_uniform(bits(m*K + k)) and S their fp32 sum (sequential, ascending). probs: [B, E] fp32 (every
    token of a sample shares its row). Returns int64 [B*T, K] in draw order."""
    probs = np.asarray(probs, dtype=np.float32)
    B, E = probs.shape
    key = rng_key(seed, step, RNG_ROUTE, layer)
    out = np.zeros((B * T, K), dtype=np.int64)
    f32 = np.float32
    for b in range(B):
        for t in range(T):
            m = b * T + t
            avail = [True] * E
            for k in range(K):
                S = f32(0.0)
                for e in range(E):
                    if avail[e]:
                        S = f32(S + probs[b, e])
                u = rng_uniform(rng_bits(key, m * K + k))
                target = f32(f32(u) * S)
                cum, chosen, last = f32(0.0), -1, 0
                for e in range(E):
                    if avail[e]:
                        last = e
                        cum = f32(cum + probs[b, e])
                        if chosen < 0 and target < cum:
                            chosen = e
                if chosen < 0:
                    chosen = last
                avail[chosen] = False
                out[m, k] = chosen
    return out
