"""Persistent small-batch kernel (MODE_SMALL_FUSED=1) against the CUDA-graph path: max |diff| per sampler and batch,
and run-to-run determinism of the persistent kernel (scripts/README.md)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import mode_oracle as O  # noqa: E402
from test_engine_gpu import MODELS, cu  # noqa: E402
from test_reference_full_gpu import _modules  # noqa: E402
from mode_diffusion_policy_b200 import gc_sampling as S  # noqa: E402

for tag in MODELS:
    cfg, _ = MODELS[tag]
    sd = O.make_weights(cfg, seed=1234, router_gain=30.0)
    state, goal, x0 = O.make_inputs(cfg, 2, seed=4321)
    sigmas = O.get_sigmas_exponential(10, 1e-3, 80.0)
    outs = {}
    for flag in ("0", "1", "1b"):
        os.environ["MODE_SMALL_FUSED"] = flag[0]
        inner, model = _modules(cfg, sd, max_batch=2)
        res = {}
        for B in (1, 2):
            st = {"state_images": cu(state[:B])}
            for name, fn in (("ddim", S.sample_ddim), ("dpmpp_2m", S.sample_dpmpp_2m), ("heun", S.sample_heun), ("euler", S.sample_euler)):
                res[(name, B)] = fn(model, st, cu(x0[:B]), cu(goal[:B]), cu(sigmas), disable=True).clone()
        outs[flag] = res
        del inner, model
    for k in outs["0"]:
        a, b, c = outs["0"][k], outs["1"][k], outs["1b"][k]
        print(tag, k, "graph-vs-persistent max|d|", float((a - b).abs().max()), "persistent rerun max|d|", float((b - c).abs().max()),
              "scale", float(a.abs().max()), flush=True)
