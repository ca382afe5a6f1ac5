#!/usr/bin/env python
"""One batch of BASELINE config 3 (multi-flock 65536 x 6) or config 4 (TDM 16384 x 45): SETTLE steps, then K more --
the thing to put under ncu:  ncu -k regex:macm_step --launch-skip <SETTLE> -c <K> ... python profiles/one_batch_cfg.py 4 64 2"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
import torch
import gym_macm

cfg = int(sys.argv[1])
settle = int(sys.argv[2]) if len(sys.argv) > 2 else 64
K = int(sys.argv[3]) if len(sys.argv) > 3 else 2
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(99)
if cfg == 3:
    E, N = 65536, 6
    sim = gym_macm.BatchedFlock(E, n_agents=[N], targets=[0, 0, 1, 1, 2, 2], device=dev, seed=31)
    acts = torch.zeros((31, E, N, 4), dtype=torch.uint8, device=dev)
    acts[..., :3] = torch.randint(0, 3, (31, E, N, 3), generator=g, device=dev, dtype=torch.uint8)
else:
    E, N = 16384, 45
    sim = gym_macm.BatchedTDM(E, n_agents=[15, 15, 15], device=dev, seed=41)
    acts = torch.randint(0, 3, (17, E, N, 4), generator=g, device=dev, dtype=torch.uint8)
    acts[..., 3] = torch.randint(0, 2, (17, E, N), generator=g, device=dev, dtype=torch.uint8)
for k in range(settle + K):
    sim.engine.step(acts[k % acts.shape[0]])
torch.cuda.synchronize()
print("contacts/agent %.3f" % float(sim.state["contact_count"].float().sum() / (E * N)))
