"""Where a training step's time goes (1 GPU, B=128): engine fwd+bwd, autograd hand-off, AdamW, weight re-pack."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import mode_oracle as O  # noqa: E402
from mode_diffusion_policy_b200.modedit import MoDeDiT  # noqa: E402
from mode_diffusion_policy_b200.score_wrappers import GCDenoiser  # noqa: E402

B = 128
cfg = O.ModeConfig()
inner = MoDeDiT(obs_dim=2048, goal_dim=512, device="cuda", goal_conditioned=True, action_dim=7, embed_dim=1024, embed_pdrob=0,
                attn_pdrop=0.0, n_layers=12, n_heads=8, goal_seq_len=1, obs_seq_len=1, action_seq_len=10, state_dim=7,
                mlp_pdrop=0.0, goal_drop=0.0, num_experts=4, top_k=2, use_argmax=True, max_batch=B)
inner.load_state_dict({k: torch.from_numpy(v) for k, v in O.make_weights_fast(cfg, seed=1234).items()})
model = GCDenoiser(inner, sigma_data=0.5).cuda().train()
opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05, fused=True)
state, goal, x0 = O.make_inputs(cfg, B, seed=4321)
rng = np.random.default_rng(7)
S, G = torch.from_numpy(state).cuda(), torch.from_numpy(goal).cuda()
A_ = torch.from_numpy((x0 / np.float32(80.0)).astype(np.float32)).cuda()
noise = torch.from_numpy(rng.standard_normal(x0.shape).astype(np.float32)).cuda()
sigma = torch.from_numpy(np.exp(rng.uniform(np.log(1e-3), np.log(80.0), B)).astype(np.float32)).cuda()


def timed(fn, n=5):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return 1e3 * (time.perf_counter() - t0) / n


eng = inner._ensure_engine(B)
print("engine train_step (fwd+bwd)      ms:", round(timed(lambda: eng.train_step(S, A_, G, noise, sigma)), 3))
print("engine loss (fwd only)           ms:", round(timed(lambda: eng.loss(S, A_, G, noise, sigma)), 3))
loss, _ = model.loss({"state_images": S}, A_, G, noise, sigma)
loss.backward()
print("AdamW fused step                 ms:", round(timed(lambda: opt.step()), 3))
print("re-pack weights (load_state_dict) ms:", round(timed(lambda: eng.load_state_dict(dict(inner.state_dict()))), 3))


def full():
    opt.zero_grad(set_to_none=True)
    l, _ = model.loss({"state_images": S}, A_, G, noise, sigma)
    l.backward()
    opt.step()


print("full python step                 ms:", round(timed(full), 3))
import os
os.environ["MODE_TRAIN_SYNC"] = "0"
