#!/usr/bin/env python
"""Shapes of the contact islands in the bench workload (4096 envs x 64 agents after the settle steps): how many
touching contacts an island has, how many bodies, and the largest number of contacts on one body -- chains
(degree <= 2) allow a wavefront over Gauss-Seidel passes, stars and triangles do not.

    python profiles/island_shapes.py [settle steps=64]
"""
import os
import sys
from collections import Counter

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
import gym_macm

E, N, POOL = 4096, 64, 61
settle = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda", 0)
sim = gym_macm.BatchedFlock(E, n_agents=[N], reward_mode="linear", device=dev, seed=1234)
g = torch.Generator(device=dev)
g.manual_seed(99)
acts = torch.zeros((POOL, E, N, 4), dtype=torch.uint8, device=dev)
acts[..., :3] = torch.randint(0, 3, (POOL, E, N, 3), generator=g, device=dev, dtype=torch.uint8)
for k in range(settle):
    sim.engine.step(acts[k % POOL])
torch.cuda.synchronize()
cnt = sim.state["contact_count"].cpu().numpy()
ab = sim.state["contact_ab"].cpu().numpy().astype(np.uint32)
shapes, per_env_max = Counter(), Counter()
for e in range(E):
    rec = ab[e, :cnt[e]]
    t = rec[(rec >> 16) & 1 == 1]
    a, b = (t & 0xff).astype(int), ((t >> 8) & 0xff).astype(int)
    parent = {}

    def find(x):
        while parent.setdefault(x, x) != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x
    for x, y in zip(a, b):
        parent[find(x)] = find(y)
    isl = {}
    for x, y in zip(a, b):
        isl.setdefault(find(x), []).append((x, y))
    big = 0
    for cs in isl.values():
        deg = Counter()
        for x, y in cs:
            deg[x] += 1
            deg[y] += 1
        shapes[(len(cs), len(deg), max(deg.values()))] += 1
        big = max(big, len(cs))
    per_env_max[big] += 1
print("islands by (contacts, bodies, max contacts on one body): count")
for k, v in sorted(shapes.items()):
    kind = "pair" if k[0] == 1 else ("chain" if k[2] <= 2 and k[1] == k[0] + 1 else ("ring" if k[2] <= 2 else "branched"))
    print("  %s  %-9s %6d" % (k, kind, v))
print("envs by their largest island (contacts):", dict(sorted(per_env_max.items())))
