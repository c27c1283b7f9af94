"""2-GPU check of the data-parallel training exchange: the NCCL all-reduce of the engine's flat gradient
buffer, divided by the world size, equals the mean of the per-shard gradients each rank can also compute locally, and
equals (to bf16-operand noise) the gradient of the whole global batch computed on one GPU.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/train_allreduce_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import mode_oracle as O  # noqa: E402  (synthetic weights / inputs only)
from test_engine_gpu import engine_for, cu  # noqa: E402

rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = O.ModeConfig(obs_dim=128, goal_dim=64, embed_dim=256, n_layers=3, n_heads=4, num_experts=4, top_k=2)
Bl = 8
sd = O.make_weights(cfg, seed=1234, router_gain=30.0)
state, goal, x0 = O.make_inputs(cfg, Bl * world, seed=4321)
rng = np.random.default_rng(5)
noise = rng.standard_normal(x0.shape).astype(np.float32)
sigma = np.exp(rng.uniform(np.log(1e-3), np.log(80.0), Bl * world)).astype(np.float32)
acts = (x0 / np.float32(80.0)).astype(np.float32)
eng = engine_for(cfg, sd, Bl * world)


def grads(lo, hi):
    sl = slice(lo, hi)
    loss, _ = eng.train_step(cu(state[sl]), cu(acts[sl]), cu(goal[sl]), cu(noise[sl]), cu(sigma[sl]))
    return float(loss), eng.flat_grads().clone()


per_shard = [grads(r * Bl, (r + 1) * Bl) for r in range(world)]
want = torch.stack([g for _, g in per_shard]).sum(0) / world  # fixed order: identical on every rank
mine_loss, _ = grads(rank * Bl, (rank + 1) * Bl)
from mode_diffusion_policy_b200 import parallel  # noqa: E402

names = [n for n, _ in O.state_dict_spec(cfg) if n != "gripper_embed.weight"]
reducer = parallel.GradAllReduce(eng, names, cfg.n_layers, min_bucket=1 << 12)  # small model: force per-layer buckets
assert all(len(b) > 0 for b in reducer.layer_buckets), reducer.layer_buckets
reducer.run()  # overlapped per-layer buckets + tail; the main stream waits for the side stream
flat = eng.flat_grads()
err = float((flat - want).abs().max() / want.abs().max())
_, whole = grads(0, Bl * world)
err_whole = float((want - whole).norm() / whole.norm())
ok = err < 1e-6 and err_whole < 2e-2
print(f"rank {rank}: all-reduce vs local mean rel-max-err {err:.2e}; mean-of-shards vs whole-batch rel-l2 {err_whole:.2e}; "
      f"{'OK' if ok else 'FAIL'}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
