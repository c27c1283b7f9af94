"""CPU tests: pin oracle/mode_oracle.py against golden vectors produced by the reference itself
(tests/golden/make_goldens.py; SURVEY.md §8c — the reference has no tests or fixtures of its own for this path)."""
from pathlib import Path

import numpy as np
import pytest

from oracle import mode_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"


def rel_l2(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b.astype(np.float64)) / (np.linalg.norm(b.astype(np.float64)) + 1e-30))


TINY = O.ModeConfig(obs_dim=128, goal_dim=64, action_dim=7, embed_dim=256, n_layers=3, n_heads=4, n_state_tokens=2,
                    action_seq_len=10, num_experts=4, top_k=2)
WIDE = O.ModeConfig(obs_dim=64, goal_dim=64, action_dim=7, embed_dim=512, n_layers=2, n_heads=4, n_state_tokens=2,
                    action_seq_len=10, num_experts=8, top_k=2)
MODELS = {"model_tiny_d256_l3_e4": (TINY, 5), "model_wide_d512_l2_e8": (WIDE, 4)}


def _digest(sd):
    import hashlib

    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k]).tobytes())
    return h.hexdigest()


def test_bf16_round_matches_torch():
    torch = pytest.importorskip("torch")
    x = np.random.default_rng(0).standard_normal(100000).astype(np.float32) * 37.0
    x[:5] = [0.0, -0.0, 1.0, 3.3895314e38, 1e-40]
    want = torch.from_numpy(x).bfloat16().float().numpy()
    assert np.array_equal(O.bf16_round(x), want)


@pytest.mark.parametrize("tag,d,H,E,T", [("block_b2_t32_d512_e2", 512, 8, 2, 32), ("block_b3_t14_d256_e4", 256, 4, 4, 14)])
def test_block_matches_reference_block(tag, d, H, E, T):
    """BASELINE.json configs[0]: NoiseBlockMoE.forward (modedit.py:530-595)."""
    g = np.load(GOLD / f"{tag}.npz")
    cfg = O.ModeConfig(obs_dim=64, goal_dim=64, embed_dim=d, n_layers=1, n_heads=H, n_state_tokens=2,
                       action_seq_len=T - 4, num_experts=E, top_k=2)
    sd = O.make_weights(cfg, seed=2024, router_gain=30.0)
    assert _digest(sd) == str(g["weights_sha256"])
    y, r = O.block_forward(g["x"], g["c"][:, 0, :], sd, 0, cfg, "fp32", return_routing=True)
    assert np.array_equal(np.sort(r["idx"], -1), np.sort(g["idx"], -1))
    assert np.array_equal(r["idx"], g["idx"])  # torch.topk order
    np.testing.assert_allclose(r["probs"], g["probs"], rtol=0, atol=2e-6)
    assert rel_l2(y, g["y"]) < 2e-6
    assert np.abs(y - g["y"]).max() < 2e-5


@pytest.mark.parametrize("tag", list(MODELS))
def test_network_denoiser_loss_match_reference(tag):
    cfg, B = MODELS[tag]
    g = np.load(GOLD / f"{tag}.npz")
    sd = O.make_weights(cfg, seed=1234, router_gain=30.0)
    assert _digest(sd) == str(g["weights_sha256"])
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321)
    sig = g["sigma_het"]
    acts = (x0 / np.float32(80.0)).astype(np.float32)
    F, routing = O.modedit_forward(sd, cfg, state, acts, goal, sig, "fp32", return_routing=True)
    for l in range(cfg.n_layers):
        assert np.array_equal(routing[l]["idx"], g["forward_idx"][l])
        np.testing.assert_allclose(routing[l]["w"], g["forward_w"][l], atol=2e-6)
        np.testing.assert_allclose(routing[l]["probs"], g["forward_probs"][l], atol=2e-6)
    assert rel_l2(F, g["forward_F"]) < 5e-6
    D = O.denoiser_forward(sd, cfg, state, g["denoise_x"], goal, sig, "fp32")
    assert rel_l2(D, g["denoise_D"]) < 5e-6
    loss, f = O.denoiser_loss(sd, cfg, state, acts, goal, g["loss_noise"], sig, "fp32")
    assert rel_l2(f, g["loss_F"]) < 5e-6
    assert abs(float(loss) - float(g["loss_value"])) <= 1e-5 * abs(float(g["loss_value"]))


@pytest.mark.parametrize("tag", list(MODELS))
def test_samplers_match_reference(tag):
    cfg, B = MODELS[tag]
    g = np.load(GOLD / f"{tag}.npz")
    sd = O.make_weights(cfg, seed=1234, router_gain=30.0)
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321)
    sigmas = O.get_sigmas_exponential(10, 1e-3, 80.0)
    np.testing.assert_allclose(sigmas, g["sigmas"], rtol=2e-6)
    a = O.sample_ddim(sd, cfg, state, x0, goal, g["sigmas"], "fp32")
    assert rel_l2(a, g["ddim_actions"]) < 2e-5
    e = O.sample_euler(sd, cfg, state, x0, goal, g["sigmas"], "fp32")
    assert rel_l2(e, g["euler_actions"]) < 2e-5


@pytest.mark.parametrize("tag", list(MODELS))
def test_bf16_contract_is_within_the_references_own_bf16_gap(tag):
    """The engine's arithmetic contract (prec='bf16') is not the reference's autocast bit for bit — the reference's own
    bf16 path differs from its fp32 path by ~1e-2. Pin it loosely: the contract must sit as close to the
    fp32 reference as the reference's CPU-autocast run does (x2 slack), on the denoiser and on the 10-step sample."""
    cfg, B = MODELS[tag]
    g = np.load(GOLD / f"{tag}.npz")
    sd = O.make_weights(cfg, seed=1234, router_gain=30.0)
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321)
    D = O.denoiser_forward(sd, cfg, state, g["denoise_x"], goal, g["sigma_het"], "bf16")
    ref_gap = rel_l2(g["denoise_D_autocast_bf16"], g["denoise_D"])
    assert rel_l2(D, g["denoise_D"]) < max(2 * ref_gap, 1e-3), (rel_l2(D, g["denoise_D"]), ref_gap)
    a = O.sample_ddim(sd, cfg, state, x0, goal, g["sigmas"], "bf16")
    ref_gap = rel_l2(g["ddim_actions_autocast_bf16"], g["ddim_actions"])
    assert rel_l2(a, g["ddim_actions"]) < max(2 * ref_gap, 1e-3), (rel_l2(a, g["ddim_actions"]), ref_gap)


def test_router_ties_take_lowest_index():
    cfg = O.ModeConfig(embed_dim=256, n_layers=1, n_heads=4, num_experts=4, top_k=2)
    sd = O.make_weights(cfg, seed=1)
    sd["blocks.0.router.router.mlp.3.weight"][:] = 0  # all logits equal -> all probs equal
    r = O.router_forward(sd, 0, np.zeros((3, 256), np.float32), cfg)
    assert np.array_equal(r["idx"], np.tile([0, 1], (3, 1)))
    np.testing.assert_allclose(r["w"], 0.5)


@pytest.mark.parametrize("tag", list(MODELS))
def test_torch_cpu_restatement_matches_reference(tag):
    """oracle/mode_ref_torch.py — the op-for-op torch-CPU restatement timed as the CPU baseline (`bench.py --impl
    reference`) — against the same reference-generated goldens: network output, top-k indices of every layer (the
    reference routes every token: all T rows of a sample must agree), denoiser output and the 10-step DDIM sample."""
    import torch

    from oracle import mode_ref_torch as RT

    cfg, B = MODELS[tag]
    g = np.load(GOLD / f"{tag}.npz")
    sd = RT.to_torch(O.make_weights(cfg, seed=1234, router_gain=30.0))
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    with torch.no_grad():
        acts = t((x0 / np.float32(80.0)).astype(np.float32))
        Fo, routing = RT.modedit_forward(sd, cfg, t(state), acts, t(goal), t(g["sigma_het"]), return_routing=True)
        D = RT.denoiser_forward(sd, cfg, t(state), t(g["denoise_x"]), t(goal), t(g["sigma_het"]))
        smp = RT.sample_ddim(sd, cfg, t(state), t(x0), t(goal), t(g["sigmas"]))
    assert rel_l2(Fo.numpy(), g["forward_F"]) < 5e-6
    for l in range(cfg.n_layers):
        idx = routing[l].numpy()  # (B, T, k)
        assert (idx == idx[:, :1, :]).all()
        assert np.array_equal(idx[:, 0, :], g["forward_idx"][l])
    assert rel_l2(D.numpy(), g["denoise_D"]) < 5e-6
    assert rel_l2(smp.numpy(), g["ddim_actions"]) < 2e-5


@pytest.mark.parametrize("tag,prefix", [(t, "train_stoch") for t in MODELS] + [("model_tiny_d256_l3_e4", "train_stoch_embed")])
def test_train_mode_oracle_matches_reference_under_the_engines_masks(tag, prefix):
    """oracle/mode_oracle_train.py (numpy train-mode forward with the masks of oracle/mode_rng.py) against the goldens the
    REFERENCE produced with its random sources patched to the same bits: expert draws identical for every token and
    layer, network output and loss to fp32 accuracy. This pins the mask conventions (which bit drops which element,
    the 1/(1-p) scalings, the draw algorithm) on the CPU, independently of the CUDA kernels that implement them."""
    from oracle import mode_oracle_train as OT

    cfg, B = MODELS[tag]
    g = np.load(GOLD / f"{tag}.npz")
    gs = np.load(GOLD / f"{prefix}_{tag}.npz")
    sd = O.make_weights(cfg, seed=1234, router_gain=float(gs["router_gain"]))
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321)
    acts = (x0 / np.float32(80.0)).astype(np.float32)
    p_attn, p_mlp, p_goal = (float(v) for v in gs["p"])
    loss, Fo, draws = OT.denoiser_loss_train(sd, cfg, state, acts, goal, g["loss_noise"], g["sigma_het"], seed=int(gs["seed"]),
                                             step=int(gs["step"]), p_attn=p_attn, p_mlp=p_mlp, p_goal=p_goal,
                                             p_embed=float(gs["p_embed"]), multinomial=True)
    for layer in range(cfg.n_layers):
        assert np.array_equal(draws[layer], gs[f"routing/{layer}"]), layer
    assert rel_l2(Fo, gs["F"]) < 1e-5, rel_l2(Fo, gs["F"])
    assert abs(loss - float(gs["loss"])) <= 1e-5 * abs(float(gs["loss"]))
