"""Timing of the fused expert-MLP kernel variants (MODE_MLP_FLAGS) against the two-launch path."""
import os
import sys
import time

import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import mode_oracle as O  # noqa: E402
from test_engine_gpu import engine_for, cu  # noqa: E402

cfg = O.ModeConfig()
sd = O.make_weights_fast(cfg, seed=1234)
sig = O.get_sigmas_exponential(10, 1e-3, 80.0)
B = 256
state, goal, x0 = O.make_inputs(cfg, B, seed=4321)
ref = None
variants = [(a.split(":")[0], int(a.split(":")[1])) for a in sys.argv[1:]] or [("0", 0), ("1", 0)]
for fused, flags in variants:
    os.environ["MODE_MLP_FUSED"] = fused
    os.environ["MODE_MLP_FLAGS"] = str(flags)
    eng = engine_for(cfg, sd, B)
    x = eng.sample_ddim(cu(state), cu(x0), cu(goal), sig)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        eng.sample_ddim(cu(state), cu(x0), cu(goal), sig)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 5 * 1e3
    if ref is None:
        ref = x.clone()
    prof = eng.profile_eval(cu(state), cu(x0), cu(goal), torch.full((B,), float(sig[3]), device="cuda"))
    print(f"fused={fused} flags={flags:2d}: {ms:.2f} ms/sample -> {10 / ms * 1e3:.1f} steps/s; identical={torch.equal(ref, x)}; "
          f"up={prof.get('up_gemm_swiglu', prof)}", flush=True)
    del eng
