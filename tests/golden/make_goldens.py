"""Generates the golden vectors that pin oracle/mode_oracle.py to the reference.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_goldens.py

It imports the reference's own modules (mode/models/networks/modedit.py, mode/models/edm_diffusion/score_wrappers.py,
gc_sampling.py) on CPU with stub modules for hydra / torchsde / torchdiffeq / matplotlib, loads the
counter-based synthetic weights of oracle.make_weights into the reference state_dict, runs the reference in fp32 (and
under torch.autocast(cpu, bfloat16) where noted) and stores inputs' seeds + outputs in tests/golden/*.npz.
"""
import hashlib
import os
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = os.environ.get("MODE_REF", "/root/reference")
sys.path.insert(0, REF)
for n in ["hydra", "hydra.utils", "torchsde", "torchdiffeq", "matplotlib", "matplotlib.pyplot"]:
    sys.modules[n] = types.ModuleType(n)
sys.modules["hydra"].utils = sys.modules["hydra.utils"]
sys.modules["hydra.utils"].instantiate = lambda cfg, *a, **k: cfg
sys.modules["torchdiffeq"].odeint = None
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

from mode.models.edm_diffusion.gc_sampling import get_sigmas_exponential, sample_ddim, sample_euler  # noqa: E402
from mode.models.edm_diffusion.score_wrappers import GCDenoiser  # noqa: E402
from mode.models.networks.modedit import MoDeDiT, NoiseBlockMoE  # noqa: E402

from oracle import mode_oracle as O  # noqa: E402

OUT = Path(__file__).resolve().parent
torch.manual_seed(0)
torch.set_grad_enabled(False)


def weights_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k]).tobytes())
    return h.hexdigest()


def build_reference(cfg: O.ModeConfig, sd_np):
    inner = MoDeDiT(obs_dim=cfg.obs_dim, goal_dim=cfg.goal_dim, device="cpu", goal_conditioned=True,
                    action_dim=cfg.action_dim, embed_dim=cfg.embed_dim, embed_pdrob=0, attn_pdrop=0.3,
                    n_layers=cfg.n_layers, n_heads=cfg.n_heads, goal_seq_len=1, obs_seq_len=1,
                    action_seq_len=cfg.action_seq_len, state_dim=7, num_experts=cfg.num_experts, top_k=cfg.top_k,
                    init_style="olmoe")
    ref_sd = inner.state_dict()
    assert list(ref_sd.keys()) == [n for n, _ in O.state_dict_spec(cfg)], "state_dict names/order drifted"
    for k, v in ref_sd.items():
        assert tuple(v.shape) == sd_np[k].shape, (k, v.shape, sd_np[k].shape)
    inner.load_state_dict({k: torch.from_numpy(v.copy()) for k, v in sd_np.items()})
    model = GCDenoiser(inner, sigma_data=cfg.sigma_data).eval()
    return inner, model


def routing_hook(inner):
    """Records RouterCond.forward outputs per layer for the most recent network call."""
    rec = {}

    def mk(i):
        def hook(mod, args, out):
            mask, idx, rprobs, true_probs = out
            rec[i] = (idx[:, 0, :].clone(), rprobs[:, 0, :].clone(), true_probs[:, 0, :].clone())
        return hook

    hs = [blk.router.register_forward_hook(mk(i)) for i, blk in enumerate(inner.blocks)]
    return rec, hs


def golden_model(tag, cfg, B, router_gain, sigma_max=80.0, sigma_min=1e-3, n_steps=10):
    sd = O.make_weights(cfg, seed=1234, router_gain=router_gain)
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321, sigma_max=sigma_max)
    inner, model = build_reference(cfg, sd)
    rng = np.random.default_rng(777)
    sig_het = np.exp(rng.uniform(np.log(sigma_min), np.log(sigma_max), size=B)).astype(np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    st = {"state_images": t(state)}
    rec, hooks = routing_hook(inner)
    out = {"weights_sha256": weights_digest(sd), "sigma_het": sig_het}
    # (1) raw network, per-sample sigma, on x0 / 80 (order-one actions)
    acts = (x0 / np.float32(sigma_max)).astype(np.float32)
    F = inner(st, t(acts), t(goal), t(sig_het))
    out["forward_F"] = F.numpy()
    out["forward_idx"] = np.stack([rec[i][0].numpy() for i in range(cfg.n_layers)])
    out["forward_w"] = np.stack([np.take_along_axis(rec[i][1].numpy(), rec[i][0].numpy(), -1)
                                 for i in range(cfg.n_layers)])
    out["forward_probs"] = np.stack([rec[i][2].numpy() for i in range(cfg.n_layers)])
    # (2) preconditioned denoiser, per-sample sigma, on x0 scaled to each sample's sigma
    xs = (x0 / np.float32(sigma_max) * sig_het[:, None, None]).astype(np.float32)
    out["denoise_x"] = xs
    out["denoise_D"] = model(st, t(xs), t(goal), t(sig_het)).numpy()
    # (3) loss forward in eval mode (dropout off, top-k routing): GCDenoiser.loss
    noise = np.random.default_rng(99).standard_normal(acts.shape).astype(np.float32)
    loss, f_out = model.loss(st, t(acts), t(goal), t(noise), t(sig_het))
    out["loss_noise"] = noise
    out["loss_value"] = np.float32(loss.item())
    out["loss_F"] = f_out.numpy()
    # (4) DDIM / Euler samples with the exponential schedule
    sigmas = get_sigmas_exponential(n_steps, sigma_min, sigma_max, "cpu")
    out["sigmas"] = sigmas.numpy()
    trace_idx = []

    def cb(d):
        trace_idx.append(np.stack([rec[i][0].numpy() for i in range(cfg.n_layers)]))

    out["ddim_actions"] = sample_ddim(model, st, t(x0), t(goal), sigmas, disable=True, callback=cb).numpy()
    out["ddim_idx"] = np.stack(trace_idx)  # (steps, L, B, k)
    out["euler_actions"] = sample_euler(model, st, t(x0), t(goal), sigmas, disable=True).numpy()
    # (5) the reference's own bf16 path on CPU (autocast): context for the engine's bf16 contract, loosely pinned
    with torch.autocast("cpu", dtype=torch.bfloat16):
        out["ddim_actions_autocast_bf16"] = sample_ddim(model, st, t(x0), t(goal), sigmas, disable=True).float().numpy()
        out["denoise_D_autocast_bf16"] = model(st, t(xs), t(goal), t(sig_het)).float().numpy()
    for h in hooks:
        h.remove()
    np.savez_compressed(OUT / f"{tag}.npz", **out)
    margins = [O.topk_margin(out["forward_probs"][l], cfg.top_k) for l in range(cfg.n_layers)]
    print(f"{tag}: saved; min top-k margin (per-sample sigma) {min(margins):.3e}; "
          f"ddim |a| max {np.abs(out['ddim_actions']).max():.3f}")


def golden_block(tag, d, H, E, k, B, T, router_gain):
    """BASELINE.json configs[0]: one NoiseBlockMoE forward on CPU via the reference module."""
    cfg = O.ModeConfig(obs_dim=64, goal_dim=64, embed_dim=d, n_layers=1, n_heads=H, n_state_tokens=2,
                       action_seq_len=T - 4, num_experts=E, top_k=k)
    sd = O.make_weights(cfg, seed=2024, router_gain=router_gain)
    blk = NoiseBlockMoE(d, H, 0.3, 0.1, 0.1, cond_router=True, num_experts=E, top_k=k).eval()
    blk.load_state_dict({n[len("blocks.0."):]: torch.from_numpy(v.copy()) for n, v in sd.items()
                         if n.startswith("blocks.0.")})
    rng = np.random.default_rng(5)
    x = rng.standard_normal((B, T, d)).astype(np.float32)
    c = (0.5 * rng.standard_normal((B, 1, d))).astype(np.float32)
    rec = {}
    h = blk.router.register_forward_hook(lambda m, a, o: rec.update(idx=o[1][:, 0, :].clone(), probs=o[3][:, 0, :].clone()))
    y = blk(torch.from_numpy(x), torch.from_numpy(c))
    h.remove()
    np.savez_compressed(OUT / f"{tag}.npz", weights_sha256=weights_digest(sd), x=x, c=c, y=y.numpy(),
                        idx=rec["idx"].numpy(), probs=rec["probs"].numpy())
    print(f"{tag}: saved; min margin {O.topk_margin(rec['probs'].numpy(), k):.3e}")


if __name__ == "__main__":
    # configs[0] of BASELINE.json (B=2, seq=32, d=512, 2 experts) and a routed E=4 variant
    golden_block("block_b2_t32_d512_e2", d=512, H=8, E=2, k=2, B=2, T=32, router_gain=30.0)
    golden_block("block_b3_t14_d256_e4", d=256, H=4, E=4, k=2, B=3, T=14, router_gain=30.0)
    # a small full model: 3 layers, d=256, 4 experts — full 10-step sample, loss, routing
    tiny = O.ModeConfig(obs_dim=128, goal_dim=64, action_dim=7, embed_dim=256, n_layers=3, n_heads=4,
                        n_state_tokens=2, action_seq_len=10, num_experts=4, top_k=2)
    golden_model("model_tiny_d256_l3_e4", tiny, B=5, router_gain=30.0)
    # head dim 128 / 8 experts variant (Dh = 128 like the full model)
    wide = O.ModeConfig(obs_dim=64, goal_dim=64, action_dim=7, embed_dim=512, n_layers=2, n_heads=4,
                        n_state_tokens=2, action_seq_len=10, num_experts=8, top_k=2)
    golden_model("model_wide_d512_l2_e8", wide, B=4, router_gain=30.0)
