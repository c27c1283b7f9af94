"""Prints engine-vs-oracle parity numbers (rel-L2) as a function of depth and sigma at d=1024 — run under gpurun.
The oracle is used here only as the checker."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import mode_oracle as O  # noqa: E402
from test_engine_gpu import cu, engine_for, rel_l2  # noqa: E402

rows = []
for L in (1, 2, 4, 8, 12):
    cfg = O.ModeConfig(n_layers=L)
    B = 8
    sd = O.make_weights(cfg, seed=1234, router_gain=30.0)
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321)
    eng = engine_for(cfg, sd, B)
    for sigma in (80.0, 0.5, 0.01):
        sig = np.full(B, sigma, np.float32)
        xs = (x0 / np.float32(80.0) * np.float32(max(sigma, 0.5))).astype(np.float32)
        F = eng.forward(cu(state), cu(xs), cu(goal), cu(sig)).cpu().numpy()
        D = eng.denoise(cu(state), cu(xs), cu(goal), cu(sig)).cpu().numpy()
        wF = O.modedit_forward(sd, cfg, state, xs, goal, sig, "bf16")
        wD = O.denoiser_forward(sd, cfg, state, xs, goal, sig, "bf16")
        fF = O.modedit_forward(sd, cfg, state, xs, goal, sig, "fp32")
        rows.append({"layers": L, "sigma": sigma, "F_vs_contract": rel_l2(F, wF), "D_vs_contract": rel_l2(D, wD),
                     "F_vs_fp32": rel_l2(F, fF), "contract_vs_fp32": rel_l2(wF, fF)})
        print(json.dumps(rows[-1]), flush=True)
    if L == 12:
        sigmas = O.get_sigmas_exponential(10, 1e-3, 80.0)
        a = eng.sample_ddim(cu(state), cu(x0), cu(goal), sigmas).cpu().numpy()
        w = O.sample_ddim(sd, cfg, state, x0, goal, sigmas, "bf16")
        f = O.sample_ddim(sd, cfg, state, x0, goal, sigmas, "fp32")
        print(json.dumps({"layers": L, "ddim10_vs_contract": rel_l2(a, w), "ddim10_vs_fp32": rel_l2(a, f),
                          "contract_vs_fp32": rel_l2(w, f)}), flush=True)
    eng.close()
    del eng, sd
