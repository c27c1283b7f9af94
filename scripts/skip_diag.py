"""In-situ cost of each kernel class inside the captured 10-step DDIM graph: time the sampler with one class of
launches removed (MODE_DEBUG_SKIP bit mask, outputs are garbage) and report the step-time delta. Complements the
event-bracketed `mode_profile_eval` (includes launch gaps) and the ncu launch list (cold cache, serialised)."""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from oracle import mode_oracle as O  # noqa: E402  (input/weight generator only)
from mode_diffusion_policy_b200.engine import EngineConfig, ModeEngine  # noqa: E402

CLASSES = ["router+plan", "embed", "qkv_gemm", "attention", "proj_gemm", "ln2_permute", "up_gemm_swiglu", "down_gemm",
           "combine_ln1"]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    cfg = O.ModeConfig()
    eng = ModeEngine(EngineConfig(max_batch=B))
    eng.load_state_dict(O.make_weights_fast(cfg, seed=1234))
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321)
    S, G, X = (torch.from_numpy(a).cuda() for a in (state, goal, x0))
    sig = O.get_sigmas_exponential(10, 1e-3, 80.0)

    def timed(mask, reps=10):
        os.environ["MODE_DEBUG_SKIP"] = str(mask)
        eng.sample_ddim(S[:2], X[:2], G[:2], sig)  # different batch: drops the cached graphs
        for _ in range(3):
            eng.sample_ddim(S, X, G, sig)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            eng.sample_ddim(S, X, G, sig)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps / 10  # ms per denoising step

    base = timed(0)
    print(f"B={B} full step: {base:.4f} ms/denoising step ({1e3 / base:.1f} steps/s)")
    for i in range(2, 9):
        t = timed(1 << i)
        print(f"  without {CLASSES[i]:16s}: {t:.4f} ms  -> in-situ cost {base - t:.4f} ms/step ({(base - t) / 12 * 1e3:.1f} us/layer)")
    t = timed(sum(1 << i for i in (3, 5, 8)))
    print(f"  GEMMs only: {t:.4f} ms")
    t = timed(sum(1 << i for i in (2, 4, 6, 7)))
    print(f"  row kernels + attention only: {t:.4f} ms")
    base2 = timed(0)
    print(f"full step again: {base2:.4f} ms")


if __name__ == "__main__":
    main()
