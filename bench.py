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


REF_SAMPLE_B = 64  # the CPU arms time a bounded sample of the workload: a quarter of the 256-trajectory batch


def engine_config_record(B, world, trim):
    """`config` of the JSON line; both arms print the same keys and values for the same command line."""
    return {"workload": WORKLOAD, "batch_per_gpu": B, "denoising_steps_per_bench_step": N_SAMPLING_STEPS,
            "parallelism": f"dp{world} (independent trajectory shards, no collective)",
            "l2": "no explicit flush: each denoising step streams 705 MB of bf16 weights (> 126 MB L2)",
            "weights": "random init, reference shapes (686 M params)",
            "dead_rows": ("last block's experts run on the 10 action rows of each trajectory only (the other 4 "
                          "never reach the head); step_roofline counts the reference's full FLOPs") if trim
            else "none: every token row of every block is evaluated"}


class CpuArm:
    """The reference's CPU path for the sampling loop on the host cores, fp32, all cores.

    kind "reference": the reference's OWN modules (MoDeDiT, GCDenoiser, sample_ddim), unmodified, when a checkout is
    reachable (MODE_REF, /root/reference, baseline/_ref/mode) — the case in the build container. kind "port": on the GPU
    box no checkout exists (a Python reference cannot travel; `pip install --target baseline/_ref` of its setup.py yields an
    empty distribution because mode/ has no __init__.py), so the timed code is oracle/mode_ref_torch.py, an
    op-for-op torch-CPU restatement with the reference's execution structure, pinned to it by the goldens."""

    def __init__(self, cfg):
        import torch

        import synthetic_workload as O

        self.torch, self.O, self.cfg = torch, O, cfg
        torch.set_num_threads(host_cores())
        torch.set_grad_enabled(False)
        sd_np = O.make_weights_fast(cfg, seed=1234)
        self.kind, self.impl = "port", "torch-CPU restatement of the reference modules, op for op (oracle/mode_ref_torch.py), fp32"
        self.model = None
        for root in (os.environ.get("MODE_REF"), "/root/reference", str(ROOT / "baseline" / "_ref")):
            if root and os.path.isfile(os.path.join(root, "mode", "models", "networks", "modedit.py")):
                try:
                    self._load_reference(root, sd_np)
                    self.kind, self.impl = "reference", f"the reference's own MoDeDiT / GCDenoiser / sample_ddim from {root}, fp32"
                    break
                except Exception as exc:  # an unusable checkout: fall through to the port, say why
                    print(f"bench: reference at {root} not usable ({exc!r}); timing the port", file=sys.stderr)
        if self.model is None:
            from oracle import mode_ref_torch as RT  # the only use of oracle/ in this file: the CPU arm being timed

            self.RT, self.sd = RT, RT.to_torch(sd_np)

    def _load_reference(self, root, sd_np):
        import types

        torch = self.torch
        for n in ["hydra", "hydra.utils", "torchsde", "torchdiffeq", "matplotlib", "matplotlib.pyplot"]:
            if n not in sys.modules:
                try:
                    __import__(n)
                except Exception:
                    sys.modules[n] = types.ModuleType(n)  # imported at module top by the reference, unused on this path
        if not hasattr(sys.modules["hydra"], "utils"):
            sys.modules["hydra"].utils = sys.modules["hydra.utils"]
        if not hasattr(sys.modules["hydra.utils"], "instantiate"):
            sys.modules["hydra.utils"].instantiate = lambda cfg, *a, **k: cfg
        if not hasattr(sys.modules["torchdiffeq"], "odeint"):
            sys.modules["torchdiffeq"].odeint = None
        if not hasattr(sys.modules["matplotlib"], "pyplot"):
            sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        sys.path.insert(0, root)
        from mode.models.edm_diffusion.gc_sampling import sample_ddim
        from mode.models.edm_diffusion.score_wrappers import GCDenoiser
        from mode.models.networks.modedit import MoDeDiT

        c = self.cfg
        inner = MoDeDiT(obs_dim=c.obs_dim, goal_dim=c.goal_dim, device="cpu", goal_conditioned=True, action_dim=c.action_dim,
                        embed_dim=c.embed_dim, embed_pdrob=0, attn_pdrop=0.3, n_layers=c.n_layers, n_heads=c.n_heads,
                        goal_seq_len=1, obs_seq_len=1, action_seq_len=c.action_seq_len, state_dim=7,
                        num_experts=c.num_experts, top_k=c.top_k, init_style="olmoe")
        inner.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd_np.items()})
        self.model, self.ref_sample_ddim = GCDenoiser(inner, sigma_data=c.sigma_data).eval(), sample_ddim

    def inputs(self, B):
        torch = self.torch
        state, goal, x0 = self.O.make_inputs(self.cfg, B, seed=4321)
        sig = torch.from_numpy(self.O.get_sigmas_exponential(N_SAMPLING_STEPS, SIGMA_MIN, SIGMA_MAX))
        return torch.from_numpy(state), torch.from_numpy(goal), torch.from_numpy(x0), sig

    def sample(self, S, G, X, sig):
        """One full 10-step DDIM sample (10 network evaluations) of the given trajectories; returns seconds."""
        t0 = time.perf_counter()
        if self.model is not None:
            out = self.ref_sample_ddim(self.model, {"state_images": S}, X, G, sig, disable=True)
        else:
            out = self.RT.sample_ddim(self.sd, self.cfg, S, X, G, sig)
        dt = time.perf_counter() - t0
        assert bool(self.torch.isfinite(out).all())
        return dt

    def record(self, value, seconds, n):
        return {"value": value, "unit": UNIT, "cores": host_cores(), "kind": self.kind, "implementation": self.impl,
                "sample": f"{n} full {N_SAMPLING_STEPS}-step DDIM sample(s) of {REF_SAMPLE_B} of the {B_PER_GPU} trajectories "
                          f"({seconds:.2f} s each, {N_SAMPLING_STEPS} network evaluations), scaled by {REF_SAMPLE_B}/{B_PER_GPU} to "
                          f"full-batch denoising steps (CPU time is linear in the batch at these sizes, SURVEY.md §8d)"}


def cpu_reference_leg(cfg, repeats=3):
    """cpu_baseline of the engine arm's N=1 line: a bounded sample (about 10-30 s) of the same workload on the host cores."""
    arm = CpuArm(cfg)
    S, G, X, sig = arm.inputs(REF_SAMPLE_B)
    arm.sample(S[:2], G[:2], X[:2], sig)  # warm the thread pool / page in the weights
    dts = [arm.sample(S, G, X, sig) for _ in range(repeats)]
    dt = float(np.mean(dts))
    return arm.record((REF_SAMPLE_B / B_PER_GPU) * N_SAMPLING_STEPS / dt, dt, repeats)


def run_reference(args, rank):
    """`--impl reference`: the reference arm. A bench step = one full 10-step DDIM sample, as in the engine arm, over a
    bounded sample of the batch (REF_SAMPLE_B of the 256 trajectories); `ms_per_step` is the measured time of such a step
    and `value` the resulting full-batch denoising-steps/s."""
    import synthetic_workload as O

    if rank != 0:
        return
    cfg = O.ModeConfig()
    arm = CpuArm(cfg)
    S, G, X, sig = arm.inputs(REF_SAMPLE_B)
    dts = []
    for i in range(args.warmup + args.steps):
        dt = arm.sample(S, G, X, sig)
        if i >= args.warmup:
            dts.append(dt)
    dt = float(np.mean(dts))
    v = (REF_SAMPLE_B / B_PER_GPU) * N_SAMPLING_STEPS / dt
    config = engine_config_record(B_PER_GPU, args.gpus, False)
    config["reference_sample"] = (f"each timed step samples {REF_SAMPLE_B} of the {B_PER_GPU} trajectories on the host cores; "
                                  f"value = ({REF_SAMPLE_B}/{B_PER_GPU}) * {N_SAMPLING_STEPS} / step time")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": arm.record(v, dt, args.steps),
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def train_leg(world, rank, steps, warmup, batch=128):
    """BASELINE.json configs[3] — the one path with a collective (reference mode/training_calvin.py:92-103, DDP): training
    step of the full MoDE (12 L, d=1024, 4 experts), noised-action MSE, per-GPU batch 128 (global 1024 on 8 GPUs), bf16
    tensor-core operands, fused forward+backward, optimizer state sharded over the ranks (NCCL reduce-scatter of the flat
    fp32 gradient buffer per block during the backward, 1/world of the fused AdamW per rank, all-gather of the new bf16
    weights during the next forward). Also times the same step WITHOUT the exchange on every rank: the difference is the
    communication that stays exposed. Returns the `train` sub-record (rank 0) or None."""
    import torch
    import torch.distributed as dist

    import synthetic_workload as O
    from mode_diffusion_policy_b200 import parallel
    from mode_diffusion_policy_b200.modedit import MoDeDiT
    from mode_diffusion_policy_b200.optim import EngineAdamW
    from mode_diffusion_policy_b200.score_wrappers import GCDenoiser

    cfg = O.ModeConfig()
    B = batch
    inner = MoDeDiT(obs_dim=cfg.obs_dim, goal_dim=cfg.goal_dim, device="cuda", goal_conditioned=True, action_dim=7,
                    embed_dim=cfg.embed_dim, embed_pdrob=0, attn_pdrop=0.3, n_layers=cfg.n_layers, n_heads=cfg.n_heads,
                    goal_seq_len=1, obs_seq_len=1, action_seq_len=10, state_dim=7, mlp_pdrop=0.1, goal_drop=0.1,
                    num_experts=cfg.num_experts, top_k=cfg.top_k, use_argmax=False, max_batch=B)  # conf/model/mode_agent.yaml recipe
    inner.load_state_dict({k: torch.from_numpy(v) for k, v in O.make_weights_fast(cfg, seed=1234).items()})
    model = GCDenoiser(inner, sigma_data=0.5).cuda().train()
    opt = EngineAdamW(inner, lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05)
    state, goal, x0 = O.make_inputs(cfg, B, seed=4321 + rank)
    rng = np.random.default_rng(7 + rank)
    S, G = torch.from_numpy(state).cuda(), torch.from_numpy(goal).cuda()
    A_ = torch.from_numpy((x0 / np.float32(SIGMA_MAX)).astype(np.float32)).cuda()
    noise = torch.from_numpy(rng.standard_normal(x0.shape).astype(np.float32)).cuda()
    sigma = torch.from_numpy(np.exp(rng.uniform(np.log(SIGMA_MIN), np.log(SIGMA_MAX), B)).astype(np.float32)).cuda()
    names = [n for n, _ in inner.named_parameters() if n != "gripper_embed.weight"]
    reducer = [None]

    def step(exchange, master_sync="lazy"):
        loss, _ = model.loss({"state_images": S}, A_, G, noise, sigma)
        loss.backward()
        if exchange and world > 1:
            if reducer[0] is None:
                reducer[0] = parallel.ShardedGradExchange(inner._engine, names, cfg.n_layers)
            opt.step_sharded(reducer[0], master_sync=master_sync)
        else:
            opt.step()
        return loss

    def timed(exchange, master_sync="lazy"):
        for _ in range(warmup):
            step(exchange, master_sync)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = step(exchange, master_sync)
        if getattr(opt, "_opt_stream", None) is not None:  # the last step's updates still in flight on the optimizer's stream
            torch.cuda.current_stream().wait_stream(opt._opt_stream)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        (ms,) = parallel.max_over_ranks([e0.elapsed_time(e1)], "cuda")
        return ms / steps, float(loss.detach())

    local_ms, _ = timed(False)                       # forward + backward + optimizer, no exchange (what one GPU does)
    full_ms, loss = timed(True) if world > 1 else (local_ms, _)
    sync_ms = timed(True, "step")[0] if world > 1 else local_ms   # fp32 masters re-gathered on every step
    if world > 1:
        opt.synchronize_parameters()                 # what a checkpoint does once: every rank sees every fp32 master
        torch.cuda.synchronize()
    inner._engine.close()
    if rank != 0:
        return None
    flat_bytes = 4 * sum(p.numel() for n, p in inner.named_parameters() if n != "gripper_embed.weight")
    return {"metric": "training-samples/sec", "value": world * B / (full_ms * 1e-3), "unit": "samples/s",
            "ms_per_step": full_ms, "ms_per_step_without_exchange": local_ms, "exposed_exchange_ms": full_ms - local_ms,
            "ms_per_step_masters_gathered_every_step": sync_ms,
            "steps": steps, "warmup": warmup, "global_batch": world * B, "batch_per_gpu": B, "scaling": "weak",
            "loss": loss, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "BASELINE.json configs[3]: MoDE 12L d=1024 4 experts top-2, noised-action MSE, fwd + bwd + "
                                   "gradient exchange + fused AdamW + weight re-pack, reference regularisation (attention "
                                   "dropout 0.3, expert dropout 0.1, goal masking 0.1, per-token multinomial routing)",
                       "collective": (f"sharded optimizer state over the {world} ranks (optim.EngineAdamW.step_sharded): per block, "
                                      f"NCCL reduce_scatter(AVG) of the fp32 gradients ({flat_bytes / 1e9:.2f} GB in all) overlapped "
                                      f"with the backward, 1/{world} of the fused AdamW per rank, all_gather of the new bf16 weights "
                                      "overlapped with the next forward; small tensors all-reduced and replicated. ms_per_step: "
                                      "fp32 masters gathered on demand (checkpoint time); ms_per_step_masters_gathered_every_step: "
                                      "also an fp32 all_gather of the masters every step (DDP's every-rank-holds-everything)") if world > 1
                       else "none (one GPU)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--batch", type=int, default=B_PER_GPU, help="trajectories per GPU (default: the headline 256)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the configs[3] training sub-record")
    ap.add_argument("--train-steps", type=int, default=8)
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
        # NCCL's CTAs must win SMs from the persistent GEMM grids of the training leg: high-priority communication stream
        opts = dist.ProcessGroupNCCL.Options()
        opts.is_high_priority_stream = True
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), pg_options=opts)
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

    # ---------------- configs[3]: the training step (the one path with a collective), all ranks take part
    train = None
    if not args.no_train and args.scaling == "weak" and B == B_PER_GPU:
        train = train_leg(world, rank, args.train_steps, 3)

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
            "config": engine_config_record(B, world, trim),
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
        if train is not None:
            line["train"] = train
        if not args.no_cpu_baseline and world == 1:  # the CPU leg belongs to the N=1 line only
            line["cpu_baseline"] = cpu_reference_leg(cfg)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
