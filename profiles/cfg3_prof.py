import os, sys, torch
sys.path.insert(0, '/root/repo/gym-macm_b200')
import gym_macm
dev = torch.device('cuda', 0)
E, N = 65536, 6
env = gym_macm.BatchedFlock(E, n_agents=[N], targets=[0,0,1,1,2,2], device=dev, seed=3)
g = torch.Generator(device=dev); g.manual_seed(5)
a = torch.zeros((16, E, N, 4), dtype=torch.uint8, device=dev)
a[..., :3] = torch.randint(0, 3, (16, E, N, 3), generator=g, device=dev, dtype=torch.uint8)
for k in range(80):
    env.engine.step(a[k % 16])
torch.cuda.synchronize()
