#!/usr/bin/env python
"""Headline benchmark: denoising-steps/sec of the MoDE EDM/DDIM sampling loop (BASELINE.json `metric`).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A bench "step" is one pass of the hot path over one batch: a full 10-step DDIM sample of B=256 CALVIN-shaped
trajectories per GPU (12 layers, d=1024, 4 experts top-2, T=14 tokens) = 10 denoising steps (full-batch network
evaluations). value = N_gpus * K * 10 / time. Data-parallel inference shards trajectories over GPUs with no
collective -> weak scaling. One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "denoising-steps/sec"
UNIT = "denoising-steps/s"
B_PER_GPU = 256
N_SAMPLING_STEPS = 10
SIGMA_MIN, SIGMA_MAX = 1e-3, 80.0
WORKLOAD = ("CALVIN-shaped DDIM sampling, BASELINE.json configs[2] per-GPU shard: B=256 trajectories/GPU, "
            "10 EDM steps, MoDE 12L d=1024 H=8, 4 experts top-2, T=14 tokens (sigma|goal|2 img|10 act), "
            "obs 2x2048, goal 512, actions 10x7")


def algorithmic_flops_per_denoising_step(cfg, B):
    """per layer QKV 6Md^2 + proj 2Md^2 + experts k*24*M*d^2 + attention 2BT(T+1)d; router on R=1
    distinct sigma (sampler mode); sigma-embed, action-embed and head per step."""
    d, T, L, k, E = cfg.embed_dim, cfg.seq_len, cfg.n_layers, cfg.top_k, cfg.num_experts
    M = B * T
    per_layer = 6 * M * d * d + 2 * M * d * d + k * 24 * M * d * d + 2 * B * T * (T + 1) * d + (4 * d * d + 4 * d * E)
    extra = (2 * d + 2 * d * d) + 2 * B * cfg.action_seq_len * cfg.action_dim * d * 2
    return L * per_layer + extra


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max = index, threading.Event(), [], set(), None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks_setting": 0x2}
        while not self.stop_flag.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max,
                "reasons": sorted(self.reasons)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_leg(cfg, sample_B=B_PER_GPU, repeats=4, sd=None):
    """Times the reference's CPU path on the host cores on a bounded sample of the workload: `repeats` full-network
    denoiser calls at B=sample_B, scaled linearly to B=256. The reference's own modules need /root/reference, which does
    not exist on the GPU box, so the timed code is oracle/mode_ref_torch.py: an op-for-op torch-CPU restatement (same
    ATen kernels, same structure: router MLP on all B*T rows, embeddings recomputed per call, per-expert boolean
    gather/scatter), pinned to the reference by the goldens (tests/test_oracle.py). `sd`: torch state_dict to reuse."""
    import torch

    import synthetic_workload as O
    from oracle import mode_ref_torch as RT  # the only use of oracle/ in this file: the CPU arm being timed

    torch.set_num_threads(host_cores())
    if sd is None:
        sd = RT.to_torch(O.make_weights_fast(cfg, seed=1234))
    state, goal, x0 = O.make_inputs(cfg, sample_B, seed=4321)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))  # noqa: E731
    S, G = t(state), t(goal)
    sig = torch.full((sample_B,), 0.5)
    x = t((x0 / np.float32(SIGMA_MAX)).astype(np.float32))
    with torch.no_grad():
        RT.denoiser_forward(sd, cfg, S[:2], x[:2], G[:2], sig[:2])  # warm the thread pool / page in the weights
        t0 = time.perf_counter()
        for _ in range(repeats):
            RT.denoiser_forward(sd, cfg, S, x, G, sig)
        dt = (time.perf_counter() - t0) / repeats
    steps_per_s = (sample_B / B_PER_GPU) / dt
    return {"value": steps_per_s, "unit": UNIT, "cores": host_cores(), "kind": "port",
            "implementation": "torch-CPU restatement of the reference modules, op for op (oracle/mode_ref_torch.py), fp32",
            "sample": f"{repeats} denoiser call(s) at B={sample_B} ({dt:.2f} s each)" +
                      ("" if sample_B == B_PER_GPU else f", scaled linearly to B={B_PER_GPU}")}


def run_reference(args, rank):
    import synthetic_workload as O

    if rank != 0:
        return
    cfg = O.ModeConfig()
    vals = []
    base = None
    from oracle import mode_ref_torch as RT

    sd = RT.to_torch(O.make_weights_fast(cfg, seed=1234))  # once: every step times the same network
    for i in range(args.warmup + args.steps):
        base = cpu_reference_leg(cfg, sample_B=B_PER_GPU, repeats=1, sd=sd)  # one full-batch network evaluation per step
        if i >= args.warmup:
            vals.append(base["value"])
    v = float(np.mean(vals))
    base["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * N_SAMPLING_STEPS / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference path on host cores (torch-CPU restatement of the reference's PyTorch modules, op for op)"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="trajectories per GPU (default: the headline 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak (default, what the driver runs): --batch trajectories per GPU; strong: --batch is the GLOBAL "
                         "batch, split evenly over the ranks (SURVEY.md §8d asks for both, labelled)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    import synthetic_workload as O  # data generators only; oracle/ (the checker) is touched by the cpu_baseline leg alone
    from mode_diffusion_policy_b200 import parallel
    from mode_diffusion_policy_b200.engine import EngineConfig, ModeEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the MoDE engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    cfg = O.ModeConfig()
    B = args.batch
    if args.scaling == "strong":
        from mode_diffusion_policy_b200.parallel import shard_bounds
        b0, b1 = shard_bounds(args.batch, rank, world)
        B = b1 - b0
    # a "denoising step" is one evaluation of a FULL batch of args.batch trajectories: N ranks each evaluating their own
    # full batch do N of them per step (weak), N ranks sharing one batch do one (strong)
    units_per_step = world if args.scaling == "weak" else 1
    # The headline numbers evaluate EVERY token row of every block, like the reference. The engine's default additionally
    # drops the 4 rows of the last block that cannot reach the output (bit-identical results); that
    # variant is measured afterwards and reported separately as `dead_row_elimination`.
    trim_env_given = "MODE_TRIM_LAST" in os.environ
    if not trim_env_given:
        os.environ["MODE_TRIM_LAST"] = "0"
    weights = O.make_weights_fast(cfg, seed=1234)
    eng = ModeEngine(EngineConfig(max_batch=B))
    eng.load_state_dict(weights)
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321 + rank)
    sigmas = O.get_sigmas_exponential(N_SAMPLING_STEPS, SIGMA_MIN, SIGMA_MAX)
    S, G, X = (torch.from_numpy(a).cuda() for a in (state, goal, x0))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident throughput (inputs already in HBM)
    for _ in range(args.warmup):
        eng.sample_ddim(S, X, G, sigmas)
    launches_per_step = eng.last_launch_count()
    clocks = ClockSampler(local_rank)
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = eng.sample_ddim(S, X, G, sigmas)
    e1.record()
    barrier()
    clocks.stop_flag.set()
    ms = e0.elapsed_time(e1)
    assert torch.isfinite(out).all()

    # ---------------- end to end through the C ABI with HOST buffers (pinned): H2D + sample + D2H every step
    hs, hg = torch.from_numpy(state).pin_memory(), torch.from_numpy(np.ascontiguousarray(goal[:, 0, :])).pin_memory()
    hx0 = torch.from_numpy(x0).pin_memory()
    hx = torch.empty_like(hx0).pin_memory()
    for _ in range(2):
        hx.copy_(hx0)
        eng.sample_ddim_host(hs, hx, hg, sigmas)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hx.copy_(hx0)
        eng.sample_ddim_host(hs, hx, hg, sigmas)  # synchronises the stream before returning
    barrier()
    e2e_s = time.perf_counter() - t0
    assert np.array_equal(hx.numpy(), out.cpu().numpy()), "host entry and device entry disagree"

    # ---------------- the engine's default configuration (dead rows of the last block eliminated), device-resident loop
    trimmed_ms = None
    if not trim_env_given:
        os.environ["MODE_TRIM_LAST"] = "1"
        eng2 = ModeEngine(EngineConfig(max_batch=B))
        eng2.load_state_dict(weights)
        for _ in range(args.warmup):
            out2 = eng2.sample_ddim(S, X, G, sigmas)
        assert torch.equal(out2, out), "dead-row elimination changed the result"
        barrier()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for _ in range(args.steps):
            eng2.sample_ddim(S, X, G, sigmas)
        t1e.record()
        barrier()
        trimmed_ms = t0e.elapsed_time(t1e)
        eng2.close()
        del eng2
        os.environ["MODE_TRIM_LAST"] = "0"
    del weights

    # the job is as slow as its slowest rank: MAX over ranks of the device time, whole-job units / that time
    if trimmed_ms is not None:
        (trimmed_ms,) = parallel.max_over_ranks([trimmed_ms], "cuda")
    ms, e2e_ms = parallel.max_over_ranks([ms, e2e_s * 1e3], "cuda")
    value = units_per_step * args.steps * N_SAMPLING_STEPS / (ms * 1e-3)
    e2e_value = units_per_step * args.steps * N_SAMPLING_STEPS / (e2e_ms * 1e-3)

    if rank == 0:
        # ---------------- roofline of the dominant kernel: grouped expert up-projection GEMM (tcgen05, SwiGLU epilogue)
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "measured (MEASURED_PEAKS.json bf16_tflops_sustained)" if pk.exists() else "fallback (B200_PROFILING.md sustained)"
        one = torch.full((1,), 0.5, device="cuda")
        prof = eng.profile_eval(S, X / SIGMA_MAX, G, one, reps=3)
        d, k = cfg.embed_dim, cfg.top_k
        M = B * cfg.seq_len
        up_ms, up_n = prof["up_gemm_swiglu"]
        # per launch: k*M routed rows x 8d outputs x d; the last block only routes its action rows (dead-row
        # elimination), so the AVERAGE launch does ((L-1) + A/T)/L of that
        trim = os.environ.get("MODE_TRIM_LAST", "1") != "0"
        launch_frac = ((cfg.n_layers - 1) + cfg.action_seq_len / cfg.seq_len) / cfg.n_layers if trim else 1.0
        up_flops = 2.0 * (k * M) * (8 * d) * d * launch_frac
        achieved = up_flops / (up_ms / up_n * 1e-3) / 1e12
        flops = {"qkv_gemm": 6.0 * M * d * d, "proj_gemm": 2.0 * M * d * d, "up_gemm_swiglu": up_flops,
                 "down_gemm": 2.0 * (k * M) * d * (4 * d) * launch_frac}
        # algorithmic HBM bytes per launch of the HBM-bound row kernels
        hbm_gbs = peaks.get("hbm_gbs", 6650.0)
        T_ = cfg.seq_len
        nbytes = {"attention": M * 3 * d * 2 + M * d * 2,                       # read qkv bf16, write out bf16
                  "ln2_permute": M * d * 4 * 2 + k * M * d * 2,                 # r/w fp32 residual, write k bf16 copies
                  "combine_ln1": M * d * 4 * 2 + k * M * d * 2 + M * d * 2,     # r/w fp32 residual, read k bf16 rows, write hA
                  "embed": M * d * 4 + M * d * 2 + B * (1 + cfg.n_state_tokens) * d * 4}
        kernels = {}
        for name, (kms, n) in prof.items():
            kernels[name] = {"ms_per_denoising_step": round(kms, 4), "launches": n}
            if name in flops and kms > 0:
                kernels[name]["tflops"] = round(flops[name] * n / (kms * 1e-3) / 1e12, 1)
                kernels[name]["frac_of_bf16_peak"] = round(kernels[name]["tflops"] / peak_tf, 3)
            if name in nbytes and kms > 0:
                gbs = nbytes[name] * n / (kms * 1e-3) / 1e9
                kernels[name]["hbm_gbs"] = round(gbs, 1)
                kernels[name]["frac_of_hbm_peak"] = round(gbs / hbm_gbs, 3)
        step_flops = algorithmic_flops_per_denoising_step(cfg, B)
        traffic = None  # dram__bytes_read+write per launch of the dominant kernel, from the committed ncu --set full capture
        tr = ROOT / "profiles" / "ncu_traffic.json"
        if tr.exists() and B == B_PER_GPU:
            traffic = json.loads(tr.read_text()).get("up_gemm_swiglu_bytes_per_launch")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "denoising_steps_per_bench_step": N_SAMPLING_STEPS,
                       "parallelism": f"dp{world} (independent trajectory shards, no collective)",
                       "l2": "no explicit flush: each denoising step streams 705 MB of bf16 weights (> 126 MB L2)",
                       "weights": "random init, reference shapes (686 M params)",
                       "dead_rows": ("last block's experts run on the 10 action rows of each trajectory only (the other 4 "
                                     "never reach the head); step_roofline counts the reference's full FLOPs") if trim
                       else "none: every token row of every block is evaluated"},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(hs.numel() * 4 + hg.numel() * 4 + hx.numel() * 4),
                    "d2h_bytes_per_step": int(hx.numel() * 4)},
            "gpu_launches": int(launches_per_step * args.steps),
            "clocks": clocks.summary(),
            "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel<EPI_SWIGLU_BF16> (grouped expert up-projection)",
                         "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                         "traffic": traffic, "peak_source": peak_src,
                         "flops_per_launch": up_flops, "us_per_launch": 1e3 * up_ms / up_n},
            "step_roofline": {"algorithmic_tflop_per_denoising_step": step_flops / 1e12,  # this rank's batch
                              "achieved_tflops": step_flops * args.steps * N_SAMPLING_STEPS / (ms * 1e-3) / 1e12,
                              "frac_of_peak": step_flops * args.steps * N_SAMPLING_STEPS / (ms * 1e-3) / 1e12 / peak_tf},
            "kernels": kernels,
        }
        if trimmed_ms is not None:
            line["dead_row_elimination"] = {
                "value": units_per_step * args.steps * N_SAMPLING_STEPS / (trimmed_ms * 1e-3), "unit": UNIT,
                "note": "engine default: the last block's experts skip the 4 token rows per trajectory that cannot reach the "
                        "output head; results bit-identical to the headline run (asserted); NOT used for `value`/`e2e`"}
        if not args.no_cpu_baseline and world == 1:  # the CPU leg belongs to the N=1 line only
            line["cpu_baseline"] = cpu_reference_leg(cfg)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
