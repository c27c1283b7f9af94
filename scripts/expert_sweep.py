"""BASELINE.json configs[4]: expert-count sweep E in {2,4,8,16} at d=1024, B=256 on one B200, in both routing modes —
uniform sigma (sampler semantics: the whole batch uses the same top-2 experts) and per-sample sigma (ragged groups).
Prints one JSON line per (E, mode) with denoising-steps/s and the per-kernel-class device times."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from oracle import mode_oracle as O  # noqa: E402  (synthetic weights / inputs only)
from mode_diffusion_policy_b200.engine import EngineConfig, ModeEngine  # noqa: E402

B = 256
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for E in (2, 4, 8, 16):
    cfg = O.ModeConfig(num_experts=E, n_layers=layers)
    eng = ModeEngine(EngineConfig(num_experts=E, n_layers=layers, max_batch=B))
    eng.load_state_dict(O.make_weights_fast(cfg, seed=1234))
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321)
    S, G, X = (torch.from_numpy(t).cuda() for t in (state, goal, x0 / np.float32(80.0)))
    sig_u = torch.full((1,), 0.5, device="cuda")
    sig_p = torch.from_numpy(np.exp(np.random.default_rng(3).uniform(np.log(1e-3), np.log(80.0), B)).astype(np.float32)).cuda()
    for mode, sig in (("uniform_sigma", sig_u), ("per_sample_sigma", sig_p)):
        for _ in range(3):
            eng.denoise(S, X, G, sig)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20
        e0.record()
        for _ in range(n):
            eng.denoise(S, X, G, sig)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        prof = eng.profile_eval(S, X, G, sig, reps=2)
        idx = np.concatenate([eng.routing(l, B)[0].reshape(-1) for l in range(layers)])
        print(json.dumps({"experts": E, "mode": mode, "layers": layers, "batch": B, "ms_per_denoising_step": round(ms, 4),
                          "denoising_steps_per_s": round(1e3 / ms, 1),
                          "experts_used": int(len(np.unique(idx))),
                          "kernels_ms": {k: round(v[0], 4) for k, v in prof.items()}}), flush=True)
    eng.close()
    del eng
    torch.cuda.empty_cache()
