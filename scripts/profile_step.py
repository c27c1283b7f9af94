"""Workload for ncu: loads the full-size model (fast synthetic weights) and runs a few denoiser evaluations at B=256
(uniform sigma, sampler semantics). Usage under gpurun:
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm_tcgen05|attention_kernel|router_kernel|plan_kernel|embed_kernel|ln2_permute|combine_kernel|head_kernel' \
      --csv --log-file gpurun_out/launches.csv python scripts/profile_step.py --evals 2
"""
import argparse
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import mode_oracle as O  # noqa: E402  (synthetic weights / inputs only)
from mode_diffusion_policy_b200.engine import EngineConfig, ModeEngine  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--evals", type=int, default=2)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--layers", type=int, default=12)
ap.add_argument("--per-sample-sigma", action="store_true")
a = ap.parse_args()
cfg = O.ModeConfig(n_layers=a.layers)
eng = ModeEngine(EngineConfig(n_layers=a.layers, max_batch=a.batch))
eng.load_state_dict(O.make_weights_fast(cfg, seed=1234))
state, goal, x0 = O.make_inputs(cfg, a.batch, seed=4321)
S, G, X = (torch.from_numpy(t).cuda() for t in (state, goal, x0 / np.float32(80.0)))
if a.per_sample_sigma:
    sig = torch.from_numpy(np.exp(np.random.default_rng(3).uniform(np.log(1e-3), np.log(80.0), a.batch)).astype(np.float32)).cuda()
else:
    sig = torch.full((1,), 0.5, device="cuda")
for _ in range(a.evals):
    out = eng.denoise(S, X, G, sig)
torch.cuda.synchronize()
print("done", float(out.abs().mean()))
