"""2-GPU check of the pipelined data-parallel step (optim.EngineAdamW.step_overlapped): per-layer all-reduce buckets on
one stream, per-layer fused AdamW launches on another, against the plain sequence (one blocking all-reduce of the flat
gradient buffer, then one optimizer launch). After 3 steps the parameters must agree between the two schedules and
between the ranks.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/train_pipeline_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
from oracle import mode_oracle as O  # noqa: E402  (synthetic weights / inputs only)
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
state, goal, x0 = O.make_inputs(cfg, B, seed=4321 + rank)  # every rank its own shard
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
    return inner, model, EngineAdamW(inner, lr=3e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.05)


inner_a, model_a, opt_a = build()
inner_b, model_b, opt_b = build()
reducer = None
for _ in range(3):
    la, _ = model_a.loss(st, acts, goal_t, noise, sig)
    dist.all_reduce(inner_a._engine.flat_grads(), op=dist.ReduceOp.AVG)
    la.backward()
    opt_a.step()
    lb, _ = model_b.loss(st, acts, goal_t, noise, sig)
    lb.backward()
    if reducer is None:
        names = [n for n, _ in inner_b.named_parameters() if n != "gripper_embed.weight"]
        reducer = parallel.GradAllReduce(inner_b._engine, names, cfg.n_layers)
        assert all(len(b) > 0 for b in reducer.layer_buckets) and reducer.active()
    opt_b.step_overlapped(reducer)
torch.cuda.synchronize()
pa, pb = dict(inner_a.named_parameters()), dict(inner_b.named_parameters())
worst, moved = 0.0, 0.0
for n in pa:
    worst = max(worst, float((pa[n].detach() - pb[n].detach()).abs().max()))
    moved = max(moved, float((pb[n].detach().cpu() - torch.from_numpy(sd[n])).abs().max())) if n in sd else moved
# both ranks must hold the same parameters: compare a checksum across ranks
chk = torch.stack([p.detach().double().sum() for p in pb.values()]).sum().reshape(1)
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
ok = worst == 0.0 and float(hi - lo) == 0.0 and moved > 0
print(f"rank {rank}: pipelined vs sequential max |dp| {worst:.3e} (parameters moved by up to {moved:.2e}); "
      f"rank checksum spread {float(hi - lo):.3e}; losses {float(la):.6f} / {float(lb):.6f}; {'OK' if ok else 'FAIL'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
