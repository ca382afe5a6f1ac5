#!/usr/bin/env python
"""A small workload for compute-sanitizer (memcheck / racecheck / synccheck): dense 64-agent flock envs through the
single-step kernel and through macm_rollout, six-agent envs (four per warp), a team-deathmatch arena.

    compute-sanitizer --tool racecheck python profiles/sanitize_case.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
import gym_macm

dev = "cuda:0"
g = torch.Generator(device=dev)
g.manual_seed(1)


def acts(K, E, N, attack=False):
    a = torch.zeros((K, E, N, 4), dtype=torch.uint8, device=dev)
    a[..., :3] = torch.randint(0, 3, (K, E, N, 3), generator=g, device=dev, dtype=torch.uint8)
    if attack:
        a[..., 3] = torch.randint(0, 2, (K, E, N), generator=g, device=dev, dtype=torch.uint8)
    return a


env = gym_macm.BatchedFlock(40, n_agents=[64], device=dev, seed=1, start_spread=9.0, reward_mode="linear")
a = acts(8, 40, 64)
for k in range(8):
    env.step(a[k])
env.rollout(a[:6])
env.rollout(None, n_steps=4, policy="flock")
pile = gym_macm.BatchedFlock(4, n_agents=[64], device=dev, seed=2, start_spread=6.5, max_contacts=2016, max_touching=240)
a = acts(4, 4, 64)
for k in range(4):
    pile.step(a[k])
small = gym_macm.BatchedFlock(64, n_agents=[6], targets=[0, 0, 1, 1, 2, 2], device=dev, seed=3, start_spread=3.0)
a = acts(6, 64, 6)
for k in range(6):
    small.step(a[k])
small.rollout(a[:4])
tdm = gym_macm.BatchedTDM(16, n_agents=[15, 15, 15], device=dev, seed=4, world_width=8.0, world_height=8.0)
a = acts(6, 16, 45, attack=True)
for k in range(6):
    tdm.step(a[k])
tdm.rollout(a[:3], want=("rewards",))
torch.cuda.synchronize()
print("touching contacts per env:", float(env.state["env_state"][:, 2].float().mean()), float(pile.state["env_state"][:, 2].float().mean()))
print("done")
