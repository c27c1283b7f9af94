"""Experiment: does running two half-batches on two streams (two engines) beat one full batch on one stream?
Persistent GEMM tails and the latency-bound row kernels leave SMs idle; a second independent stream can fill them."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import mode_oracle as O  # noqa: E402
from test_engine_gpu import engine_for, cu  # noqa: E402

cfg = O.ModeConfig()
sd = O.make_weights_fast(cfg, seed=1234)
sig = O.get_sigmas_exponential(10, 1e-3, 80.0)


def run(n_streams, B_each, reps=8):
    engs = [engine_for(cfg, sd, B_each) for _ in range(n_streams)]
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    ins = []
    for i in range(n_streams):
        state, goal, x0 = O.make_inputs(cfg, B_each, seed=4321 + i)
        ins.append((cu(state), cu(x0), cu(goal)))

    def once():
        for e, s, (st, x, g) in zip(engs, streams, ins):
            with torch.cuda.stream(s):
                e.sample_ddim(st, x, g, sig)

    for _ in range(3):
        once()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    total_B = n_streams * B_each
    print(f"{n_streams} stream(s) x B={B_each}: {dt * 1e3:.2f} ms per 10-step sample of {total_B} trajectories -> "
          f"{10 / dt * total_B / 256:.1f} denoising-steps/s (B=256 equivalents)", flush=True)
    del engs


run(1, 256)
run(2, 128)
run(2, 256)
run(4, 64)
run(1, 512)
