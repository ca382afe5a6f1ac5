#!/usr/bin/env python
"""us per launch of get_obs alone (macm_observe: nearest-agent search + target node) on the bench batch, and of a
rollout without per-step observations: what a phase-split step (engine kernel + observation kernel) would have to beat."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
import torch
import gym_macm

dev = torch.device("cuda", 0)
E, N, R = 4096, 64, 16
sims = [gym_macm.BatchedFlock(E, n_agents=[N], reward_mode="linear", device=dev, seed=1234 + r) for r in range(R)]
g = torch.Generator(device=dev)
g.manual_seed(99)
acts = torch.zeros((61, E, N, 4), dtype=torch.uint8, device=dev)
acts[..., :3] = torch.randint(0, 3, (61, E, N, 3), generator=g, device=dev, dtype=torch.uint8)
for k in range(64 * R):
    sims[k % R].engine.step(acts[(k // R + 7 * (k % R)) % 61])


def timed(fn, n):
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for k in range(n):
            fn(k)
        e1.record()
        torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n


print("observe alone: %.2f us per launch" % timed(lambda k: sims[k % R].engine.observe(), 800))
print("step: %.2f us per launch" % timed(lambda k: sims[k % R].engine.step(acts[k % 61]), 800))
K = 8
print("rollout K=%d without per-step obs: %.2f us per step" % (K, timed(lambda k: sims[k % R].engine.rollout(acts[:K], K, None, 0, {}), 100) / K))
outs = sims[0].engine.rollout_buffers(K)
print("rollout K=%d with per-step outputs: %.2f us per step" % (K, timed(lambda k: sims[k % R].engine.rollout(acts[:K], K, None, 0, outs), 100) / K))
