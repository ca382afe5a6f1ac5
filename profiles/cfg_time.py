#!/usr/bin/env python
"""us per step of BASELINE config 3 (multi-flock 65536 x 6) or 4 (TDM 16384 x 45), two batches on two streams, 300 steps."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
import torch
import gym_macm

cfg = int(sys.argv[1])
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev)
g.manual_seed(99)
if cfg == 3:
    E, N, R = 65536, 6, 10
    sims = [gym_macm.BatchedFlock(E, n_agents=[N], targets=[0, 0, 1, 1, 2, 2], device=dev, seed=31 + r) for r in range(R)]
    acts = torch.zeros((31, E, N, 4), dtype=torch.uint8, device=dev)
    acts[..., :3] = torch.randint(0, 3, (31, E, N, 3), generator=g, device=dev, dtype=torch.uint8)
else:
    E, N, R = 16384, 45, 2
    sims = [gym_macm.BatchedTDM(E, n_agents=[15, 15, 15], device=dev, seed=41 + r) for r in range(R)]
    acts = torch.randint(0, 3, (17, E, N, 4), generator=g, device=dev, dtype=torch.uint8)
    acts[..., 3] = torch.randint(0, 2, (17, E, N), generator=g, device=dev, dtype=torch.uint8)
pool = gym_macm.BatchPool(sims)
K = 300
for rep in range(2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream(dev)
    e0.record()
    for st in pool.streams:
        st.wait_stream(main)
    for k in range(K):
        pool.step(acts[(k // R + 7 * (k % R)) % acts.shape[0]])
    for st in pool.streams:
        main.wait_stream(st)
    e1.record()
    torch.cuda.synchronize()
print("config %d: %.2f us per step, %.3g agent-steps/s" % (cfg, 1e3 * e0.elapsed_time(e1) / K, E * N * K / (e0.elapsed_time(e1) * 1e-3)))
