#!/usr/bin/env python
"""Cross-GPU check of gym_macm.dist.PeerGather (torchrun, one rank per GPU, NCCL): what the step kernels stored into
the learner's buffers over NVLink must equal an NCCL all-gather of the shards' own output buffers.

    python -m torch.distributed.run --nproc-per-node 2 profiles/peer_gather_check.py"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
import gym_macm
from gym_macm.dist import PeerGather, all_gather_envs, shard_range

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
total, N = 512 * world, 64
start, count = shard_range(total, rank, world)
env = gym_macm.BatchedFlock(count, n_agents=[N], device=dev, seed=3, env_index_base=start, reward_mode="linear")
pg = PeerGather(env, total, learner=0, names=("obs", "nn_idx", "rewards", "collided", "done"))
g = torch.Generator(device=dev)
g.manual_seed(5)
ok = True
for k in range(20):
    a = torch.zeros((total, N, 4), dtype=torch.uint8, device=dev)
    a[..., :3] = torch.randint(0, 3, (total, N, 3), generator=g, device=dev, dtype=torch.uint8)
    pg.step(a[start:start + count].contiguous())
    pg.fence()
    ref = all_gather_envs({n: env.state[n] for n in pg.names}, total)
    if rank == 0:
        for n in pg.names:
            ok = ok and torch.equal(pg.gathered()[n], ref[n])
    dist.barrier()
if rank == 0:
    print("peer gather over %d GPUs, %d envs x %d agents, 20 steps: %s" % (world, total, N, "ok" if ok else "MISMATCH"))
env.close()
dist.barrier()
dist.destroy_process_group()
