#!/usr/bin/env python
"""Device-timed throughput of the BASELINE configs that are not the bench line (parity-test cases in
BASELINE.json: configs[2] multi-flock 6 agents x 65,536 envs, configs[3] TDM 3 x 15 agents x 16,384 envs) and
of the larger batches of configs[4] (64-agent flock envs, 16k-65k envs on one GPU).  Same method as bench.py:
CUDA events around K back-to-back launches after a settle phase, actions pre-generated on the device."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
import gym_macm

dev = torch.device("cuda", 0)
STEPS, SETTLE = 200, 64


def timed(env, acts, label, E, N, bytes_per_agent_step):
    for k in range(SETTLE):
        env.engine.step(acts[k % len(acts)])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(STEPS):
        env.engine.step(acts[k % len(acts)])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / STEPS
    info = env.engine.info
    print(json.dumps({"config": label, "envs": E, "agents_per_env": N, "ms_per_step": ms,
                      "agent_steps_per_sec": E * N / (ms * 1e-3),
                      "algorithmic_GBps": bytes_per_agent_step * E * N / (ms * 1e-3) / 1e9,
                      "launch": {"threads_per_block": info.threads_per_block, "blocks": info.blocks,
                                 "lanes_per_env": info.lanes_per_env, "smem_per_block": info.smem_bytes_per_block},
                      "contacts_per_agent": float(env.state["contact_count"].sum()) / (E * N),
                      "state": "L2-resident (one batch stepped repeatedly)"}))


def rand_actions(E, N, n=16, attack=False):
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    a = torch.zeros((n, E, N, 4), dtype=torch.uint8, device=dev)
    a[..., :3] = torch.randint(0, 3, (n, E, N, 3), generator=g, device=dev, dtype=torch.uint8)
    if attack:
        a[..., 3] = torch.randint(0, 2, (n, E, N), generator=g, device=dev, dtype=torch.uint8)
    return a


which = sys.argv[1:] or ["cfg3", "cfg4", "cfg5"]
if "cfg3" in which:
    E, N = 65536, 6
    env = gym_macm.BatchedFlock(E, n_agents=[N], targets=[0, 0, 1, 1, 2, 2], device=dev, seed=3)
    timed(env, rand_actions(E, N), "configs[2]: multi-flock 6 agents, targets=[0,0,1,1,2,2], binary reward", E, N, 109 + 32 * 0.2 + 25 / 6)
    env.close()
if "cfg4" in which:
    E, N = 16384, 45
    env = gym_macm.BatchedTDM(E, n_agents=[15, 15, 15], device=dev, seed=4)
    timed(env, rand_actions(E, N, attack=True), "configs[3]: TDM 3 teams x 15 agents (repaired semantics)", E, N, 828 + 32 * 0.4)
    env.close()
if "cfg5" in which:
    for E in (16384, 65536):
        N = 64
        env = gym_macm.BatchedFlock(E, n_agents=[N], reward_mode="linear", device=dev, seed=6)
        timed(env, rand_actions(E, N, n=4), "configs[4] shard: 64-agent flock envs, %d agents on one GPU" % (E * N), E, N, 121)
        env.close()
