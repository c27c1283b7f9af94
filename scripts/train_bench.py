"""BASELINE.json configs[3]: training step, noised-action MSE, full MoDE (12 L, d=1024, 4 experts), per-GPU batch 128
(global 1024 on 8 GPUs), bf16 tensor-core operands, ONE NCCL all-reduce of the flat gradient buffer per step, fused
AdamW on the fp32 masters, weights re-packed for the next step. Prints one JSON line (rank 0).

    python scripts/train_bench.py                                  # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/train_bench.py
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from oracle import mode_oracle as O  # noqa: E402  (synthetic weights / inputs only)
from mode_diffusion_policy_b200 import parallel  # noqa: E402
from mode_diffusion_policy_b200.modedit import MoDeDiT  # noqa: E402
from mode_diffusion_policy_b200.score_wrappers import GCDenoiser  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--layers", type=int, default=12)
ap.add_argument("--stochastic", action="store_true",
                help="the reference's default train-mode regularisation (conf/model/mode_agent.yaml: attn_pdrop 0.3, "
                     "mlp_pdrop 0.1, goal_drop 0.1, use_argmax False = per-token multinomial routing)")
ap.add_argument("--ema", type=float, default=None, help="keep an EMA of the weights inside the optimizer launch (decay)")
a = ap.parse_args()
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
if world > 1:
    # NCCL's kernels need 512+ threads on one SM at once; next to a small-CTA kernel with a deep queue (the optimizer)
    # such a hole never opens unless NCCL's stream has priority over it
    opts = dist.ProcessGroupNCCL.Options()
    opts.is_high_priority_stream = os.environ.get("MODE_NCCL_HIGH_PRIORITY", "1") == "1"
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), pg_options=opts)
cfg = O.ModeConfig(n_layers=a.layers)
B = a.batch
inner = MoDeDiT(obs_dim=2048, goal_dim=512, device="cuda", goal_conditioned=True, action_dim=7, embed_dim=1024, embed_pdrob=0,
                attn_pdrop=0.3 if a.stochastic else 0.0, n_layers=a.layers, n_heads=8, goal_seq_len=1, obs_seq_len=1,
                action_seq_len=10, state_dim=7, mlp_pdrop=0.1 if a.stochastic else 0.0, goal_drop=0.1 if a.stochastic else 0.0,
                num_experts=4, top_k=2, use_argmax=not a.stochastic, max_batch=B)
inner.load_state_dict({k: torch.from_numpy(v) for k, v in O.make_weights_fast(cfg, seed=1234).items()})
model = GCDenoiser(inner, sigma_data=0.5).cuda().train()
if os.environ.get("MODE_TRAIN_FUSED_OPT", "1") == "1":  # one engine launch: AdamW over the flat gradient buffer + re-pack
    from mode_diffusion_policy_b200.optim import EngineAdamW
    opt = EngineAdamW(inner, lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05, ema_decay=a.ema)
    opt_name = "engine fused AdamW+repack" + (f"+EMA({a.ema})" if a.ema is not None else "")
else:
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, betas=(0.9, 0.95), weight_decay=0.05, fused=True)
    opt_name = "torch fused AdamW + engine re-pack"
state, goal, x0 = O.make_inputs(cfg, B, seed=4321 + rank)
rng = np.random.default_rng(7 + rank)
S, G = torch.from_numpy(state).cuda(), torch.from_numpy(goal).cuda()
A_ = torch.from_numpy((x0 / np.float32(80.0)).astype(np.float32)).cuda()
noise = torch.from_numpy(rng.standard_normal(x0.shape).astype(np.float32)).cuda()
sigma = torch.from_numpy(np.exp(rng.uniform(np.log(1e-3), np.log(80.0), B)).astype(np.float32)).cuda()
ap_overlap = os.environ.get("MODE_TRAIN_OVERLAP", "1") == "1"
reducer = None


TIMELINE = [] if os.environ.get("MODE_TRAIN_TIMELINE") == "1" else None  # per-stream event marks of the last step
# MODE_TRAIN_PIPELINE: 1 = exchange + optimizer pipelined per layer behind the backward (default for N > 1);
# 2 = additionally let block l's update start while blocks l-1..0 are still in backward (also valid on one GPU);
# 3 = sharded optimizer (reduce-scatter -> 1/world AdamW -> bf16 all-gather, optim.EngineAdamW.step_sharded) behind the
# backward; 4 = sharded with updates starting during the backward. MODE_TRAIN_MASTER_SYNC=step|lazy (sharded modes)
pipe_mode = int(os.environ.get("MODE_TRAIN_PIPELINE", "1" if world > 1 else "0"))
pipelined = pipe_mode > 0 and hasattr(opt, "step_overlapped")


def step():
    global reducer
    opt.zero_grad(set_to_none=True)
    loss, _ = model.loss({"state_images": S}, A_, G, noise, sigma)
    if pipelined and pipe_mode >= 3 and world > 1:
        if reducer is None:
            reducer = parallel.ShardedGradExchange(inner._engine, [n for n, _ in inner.named_parameters()
                                                                   if n != "gripper_embed.weight"], a.layers)
        sync = os.environ.get("MODE_TRAIN_MASTER_SYNC", "step")
        ofw = os.environ.get("MODE_TRAIN_OVERLAP_FWD", "1") == "1"
        if pipe_mode >= 4:
            opt.step_sharded(reducer, loss_scale=1.0, master_sync=sync, overlap_forward=ofw, timeline=TIMELINE)
        else:
            loss.backward()
            opt.step_sharded(reducer, master_sync=sync, overlap_forward=ofw, timeline=TIMELINE)
        return loss
    if pipelined:  # exchange and optimizer pipelined per layer (optim.EngineAdamW.step_overlapped)
        if reducer is None:
            reducer = parallel.GradAllReduce(inner._engine, [n for n, _ in inner.named_parameters()
                                                             if n != "gripper_embed.weight"], a.layers)
        if pipe_mode >= 2:  # unscaled loss: no autograd round trip, updates start while earlier blocks are in backward
            opt.step_overlapped(reducer, timeline=TIMELINE, loss_scale=1.0)
        else:
            loss.backward()
            opt.step_overlapped(reducer, timeline=TIMELINE)
        return loss
    if world > 1:  # average the engine's flat gradient buffer before autograd hands out the views
        if ap_overlap:  # per-layer buckets on a side stream, overlapped with the rest of the backward
            if reducer is None:
                reducer = parallel.GradAllReduce(inner._engine, [n for n, _ in inner.named_parameters()
                                                                 if n != "gripper_embed.weight"], a.layers)
            reducer.run()
        else:  # one blocking all-reduce after the backward
            dist.all_reduce(inner._engine.flat_grads(), op=dist.ReduceOp.AVG)
    loss.backward()
    opt.step()
    return loss


for _ in range(a.warmup):
    step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    if TIMELINE is not None:
        TIMELINE.clear()
        t_start = torch.cuda.Event(enable_timing=True)
        t_start.record()
    loss = step()
if getattr(opt, "_opt_stream", None) is not None:  # the last step's updates still in flight on the optimizer's stream
    torch.cuda.current_stream().wait_stream(opt._opt_stream)
e1.record()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
(ms,) = parallel.max_over_ranks([e0.elapsed_time(e1)], "cuda")
if TIMELINE and rank == 0:
    print("timeline of the last step (ms after its start): " +
          ", ".join(f"{lab} {t_start.elapsed_time(ev):.2f}" for lab, ev in TIMELINE), file=sys.stderr, flush=True)
if rank == 0:
    fwd_flops = 2.526888e12 * B / 256 * a.layers / 12
    print(json.dumps({"metric": "training-samples/sec", "value": world * B * a.steps / (ms * 1e-3), "unit": "samples/s",
                      "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
                      "global_batch": world * B, "dtype": "bf16", "data": "synthetic", "loss": float(loss),
                      "approx_tflops": 3 * fwd_flops * a.steps / (ms * 1e-3) / 1e12,
                      "config": {"workload": f"MoDE {a.layers}L d=1024 E=4 top-2, B={B}/GPU, fwd+bwd+all-reduce+AdamW+repack",
                                 "regularisation": ("attn_pdrop 0.3, mlp_pdrop 0.1, goal_drop 0.1, per-token multinomial routing" if a.stochastic else "none (deterministic mode)"), "optimizer": opt_name,
                                 "grad_allreduce": (f"sharded optimizer (reduce-scatter, 1/{world} AdamW per rank, bf16 all-gather; mode {pipe_mode}, master sync {os.environ.get('MODE_TRAIN_MASTER_SYNC', 'step')})" if pipelined and pipe_mode >= 3 and world > 1 else "per-layer NCCL all-reduce buckets pipelined with per-layer fused AdamW launches (step_overlapped)" if pipelined else "per-layer NCCL all-reduce buckets of the flat fp32 gradient buffer, overlapped with backward" if ap_overlap else "one NCCL all-reduce over the flat fp32 gradient buffer")}}), flush=True)
if world > 1:
    dist.destroy_process_group()
