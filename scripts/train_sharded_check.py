"""2-GPU (or N-GPU) check of the sharded data-parallel step (optim.EngineAdamW.step_sharded: reduce-scatter -> 1/world of
the fused AdamW per rank -> all-gather of the bf16 weights) against the replicated one (step_overlapped: all-reduce, every
rank updates everything). After 3 steps the fp32 masters (after synchronize_parameters), the moments, the EMA and the
loss of a 4th forward (i.e. the packed bf16 weights) must agree between the schedules and between the ranks. On two ranks
the agreement is bit-exact (a two-term mean does not depend on the order); on more ranks NCCL's reduce-scatter and
all-reduce sum in different orders, and Adam turns an ulp of a near-zero gradient into an lr-sized difference of that
element: there the check is the relative L2 distance of the parameter MOVEMENT (<= 2e-2), the loss (<= 2e-3) and — always
exact — that every rank ends with identical parameters.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/train_sharded_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import synthetic_workload as O  # noqa: E402
from mode_diffusion_policy_b200 import parallel  # noqa: E402
from mode_diffusion_policy_b200.modedit import MoDeDiT  # noqa: E402
from mode_diffusion_policy_b200.optim import EngineAdamW  # noqa: E402
from mode_diffusion_policy_b200.score_wrappers import GCDenoiser  # noqa: E402

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = O.ModeConfig(obs_dim=64, goal_dim=64, action_dim=7, embed_dim=1024, n_layers=2, n_heads=8, n_state_tokens=2,
                   action_seq_len=10, num_experts=2, top_k=2)
B = 8
sd = O.make_weights_fast(cfg, seed=1234)
state, goal, x0 = O.make_inputs(cfg, B, seed=4321 + rank)  # every rank its own shard of the global batch
rng = np.random.default_rng(5 + rank)
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
st = {"state_images": cu(state)}
acts, goal_t = cu((x0 / np.float32(80.0)).astype(np.float32)), cu(goal)
noise = cu(rng.standard_normal(x0.shape).astype(np.float32))
sig = cu(np.exp(rng.uniform(np.log(1e-3), np.log(80.0), B)).astype(np.float32))


def build():
    inner = MoDeDiT(obs_dim=cfg.obs_dim, goal_dim=cfg.goal_dim, device="cuda", goal_conditioned=True, action_dim=7,
                    embed_dim=cfg.embed_dim, embed_pdrob=0, attn_pdrop=0.0, n_layers=cfg.n_layers, n_heads=cfg.n_heads,
                    goal_seq_len=1, obs_seq_len=1, action_seq_len=10, state_dim=7, mlp_pdrop=0.0, goal_drop=0.0,
                    num_experts=cfg.num_experts, top_k=cfg.top_k, use_argmax=True, max_batch=B)
    inner.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    model = GCDenoiser(inner, sigma_data=0.5).cuda().train()
    return inner, model, EngineAdamW(inner, lr=3e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05, ema_decay=0.99)


def names_of(inner):
    return [n for n, _ in inner.named_parameters() if n != "gripper_embed.weight"]


variants = {"replicated": build(), "sharded": build(), "sharded_early_lazy": build()}
exch = {}
losses = {k: [] for k in variants}
for it in range(4):
    for key, (inner, model, opt) in variants.items():
        loss, _ = model.loss(st, acts, goal_t, noise, sig)
        loss.backward()
        losses[key].append(float(loss.detach()))
        if it == 3:
            continue  # 4th forward only: its loss reads the packed weights of step 3
        if key not in exch:
            cls = parallel.GradAllReduce if key == "replicated" else parallel.ShardedGradExchange
            exch[key] = cls(inner._engine, names_of(inner), cfg.n_layers)
        if key == "replicated":
            opt.step_overlapped(exch[key])
        elif key == "sharded":
            opt.step_sharded(exch[key])
        else:
            opt.step_sharded(exch[key], loss_scale=1.0, master_sync="lazy")
torch.cuda.synchronize()
ref_inner, _, ref_opt = variants["replicated"]
ref_p = {n: p.detach().clone() for n, p in ref_inner.named_parameters()}
ref_sd, ref_ema = ref_opt.state_dict(), {k: v.clone() for k, v in ref_opt.ema_state_dict().items()}
ok = True
tol, loss_tol = (0.0, 0.0) if world == 2 else (2e-2, 2e-3)
init_p = {n: torch.from_numpy(v).cuda() for n, v in sd.items()}


def movement_distance(params):
    """|| (p - p0) - (p_ref - p0) || / || p_ref - p0 || over all parameters."""
    num = sum(float(((p.detach() - ref_p[n]).double() ** 2).sum()) for n, p in params)
    den = sum(float(((ref_p[n] - init_p[n]).double() ** 2).sum()) for n, p in params if n in init_p)
    return (num / max(den, 1e-300)) ** 0.5


for key in ("sharded", "sharded_early_lazy"):
    inner, _, opt = variants[key]
    n_sharded = sum(len(lay) for lay in exch[key].layers)
    opt.synchronize_parameters()
    torch.cuda.synchronize()
    dp = max(float((p.detach() - ref_p[n]).abs().max()) for n, p in inner.named_parameters())
    osd = opt.state_dict()
    dm = float((osd["exp_avg"] - ref_sd["exp_avg"]).abs().max())
    dv = float((osd["exp_avg_sq"] - ref_sd["exp_avg_sq"]).abs().max())
    de = max(float((v - ref_ema[k]).abs().max()) for k, v in opt.ema_state_dict().items())
    dl = max(abs(a - b) for a, b in zip(losses[key], losses["replicated"]))
    chk = torch.stack([p.detach().double().sum() for p in inner.parameters()]).sum().reshape(1)
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    rel = movement_distance(list(inner.named_parameters()))
    exact = max(dp, dm, dv, de) == 0.0
    fell_back = exch[key].fallback is not None  # world sizes that split no tensor evenly: the replicated step, exactly
    good = ((exact and dl == 0.0) if (world == 2 or fell_back) else rel <= tol) and dl <= loss_tol and float(hi - lo) == 0.0 \
        and (n_sharded > 0 or fell_back)
    ok &= good
    print(f"rank {rank} {key}: {n_sharded} sharded tensors{' (fallback: replicated step)' if fell_back else ''}, "
          f"{len(exch[key].tail)} replicated spans; vs replicated movement rel-L2 {rel:.3e} max |dp| {dp:.3e} "
          f"|dm| {dm:.3e} |dv| {dv:.3e} |dema| {de:.3e} |dloss| {dl:.3e}; rank checksum spread {float(hi - lo):.3e}; "
          f"losses {losses[key]}; {'OK' if good else 'FAIL'}", flush=True)
moved = max(float((p.detach().cpu() - torch.from_numpy(sd[n])).abs().max()) for n, p in ref_inner.named_parameters() if n in sd)
ok &= moved > 0
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
