#!/usr/bin/env python
"""One batch of the bench workload (4096 x 64 Flock, linear reward, random actions): SETTLE steps, then K more.
The thing to put under ncu:  ncu -k regex:macm_step -s <SETTLE> -c <K> ... python profiles/one_batch.py SETTLE K"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
import torch
import gym_macm

settle = int(sys.argv[1]) if len(sys.argv) > 1 else 64
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
E = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
N = 64
dev = torch.device("cuda", 0)
sim = gym_macm.BatchedFlock(E, n_agents=[N], reward_mode="linear", device=dev, seed=1234)
g = torch.Generator(device=dev)
g.manual_seed(99)
POOL = 61
acts = torch.zeros((POOL, E, N, 4), dtype=torch.uint8, device=dev)
acts[..., :3] = torch.randint(0, 3, (POOL, E, N, 3), generator=g, device=dev, dtype=torch.uint8)
for k in range(settle + K):
    sim.engine.step(acts[k % POOL])
torch.cuda.synchronize()
print("touching/env %.2f" % float(sim.state["env_state"][:, 2].float().mean()))
