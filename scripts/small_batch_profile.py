"""Where a rollout-sized sample's time goes: per-kernel-class device time of one evaluation (event pairs around every
launch, `mode_profile_eval`) and the 10-step sample latency, for B in argv (default 1 2 4 8 32)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import synthetic_workload as O  # noqa: E402
from mode_diffusion_policy_b200.engine import EngineConfig, ModeEngine  # noqa: E402

cfg = O.ModeConfig()
weights = O.make_weights_fast(cfg, seed=1234)
batches = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 32]
eng = ModeEngine(EngineConfig(max_batch=max(batches)))
eng.load_state_dict(weights)
sigmas = O.get_sigmas_exponential(10, 1e-3, 80.0)
for B in batches:
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321)
    S, G, X = (torch.from_numpy(a).cuda() for a in (state, goal, x0))
    for _ in range(5):
        eng.sample_ddim(S, X, G, sigmas)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        eng.sample_ddim(S, X, G, sigmas)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    prof = eng.profile_eval(S, X / 80.0, G, torch.full((1,), 0.5, device="cuda"), reps=5)
    print(json.dumps({"batch": B, "ms_per_10_step_sample": round(ms, 3), "launches": eng.last_launch_count(),
                      "us_per_launch_class": {k: round(1e3 * m / max(n, 1), 2) for k, (m, n) in prof.items()}}), flush=True)
