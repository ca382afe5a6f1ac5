#!/usr/bin/env python
"""Device-timed cost of a step inside macm_rollout against the steps per launch (bench workload: 4096 envs x 64
agents, linear reward, iid U{0,1,2}^3 actions, 64 settle steps, rotation over ROT independent batches).  Every
step writes its obs / nn_idx / rewards / collided / done to the per-step arrays.  An experiment driver.

    python profiles/rollout_probe.py [K ...]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
import gym_macm

dev = torch.device("cuda", 0)
E, N, SETTLE, ROT = int(os.environ.get("ENVS", 4096)), int(os.environ.get("AGENTS", 64)), 64, 16
Ks = [int(x) for x in sys.argv[1:]] or [1, 2, 4, 8, 16, 32, 64]
KMAX = max(Ks + [SETTLE])
sims = [gym_macm.BatchedFlock(E, n_agents=[N], reward_mode="linear", device=dev, seed=1234 + r) for r in range(ROT)]
g = torch.Generator(device=dev)
g.manual_seed(99)
acts = torch.zeros((KMAX + 16, E, N, 4), dtype=torch.uint8, device=dev)
acts[..., :3] = torch.randint(0, 3, (KMAX + 16, E, N, 3), generator=g, device=dev, dtype=torch.uint8)
for r, s in enumerate(sims):
    s.rollout(acts[r:r + SETTLE], want=())
torch.cuda.synchronize()
for K in Ks:
    want = tuple(w for w in os.environ.get("WANT", "obs,nn_idx,rewards,collided,done").split(",") if w)
    outs = [s.engine.rollout_buffers(K, want) for s in sims[:2]]   # two sets in alternation (the consumer double-buffers)
    launches = max(4 * ROT, (2048 // K) // ROT * ROT)
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for j in range(launches):
            r = j % ROT
            sims[r].engine.rollout(acts[(j % 16):(j % 16) + K], K, None, 0, outs[j & 1])
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (launches * K)
    print(json.dumps({"envs": E, "steps_per_launch": K, "want": list(want), "sync": os.environ.get("MACM_ROLLOUT_SYNC", "default"), "launches": launches, "us_per_step": 1e3 * ms,
                      "agent_steps_per_sec": E * N / (ms * 1e-3)}), flush=True)
    del outs
