#!/usr/bin/env python
"""Multi-wave view of one step launch (macm_set_trace): per SM, the sequence of blocks it ran -- how long each
block lived, how long the SM sat between two blocks, how evenly the blocks' warps ended.

    python profiles/wave_trace.py [envs=32768] [block threads]
"""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
E = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
if len(sys.argv) > 2:
    os.environ["MACM_BLOCK_THREADS"] = sys.argv[2]
import gym_macm
from gym_macm import _lib

N, SETTLE, POOL = 64, 64, 7
dev = torch.device("cuda", 0)
sim = gym_macm.BatchedFlock(E, n_agents=[N], reward_mode="linear", device=dev, seed=1234)
g = torch.Generator(device=dev)
g.manual_seed(99)
acts = torch.zeros((POOL, E, N, 4), dtype=torch.uint8, device=dev)
acts[..., :3] = torch.randint(0, 3, (POOL, E, N, 3), generator=g, device=dev, dtype=torch.uint8)
for k in range(SETTLE):
    sim.engine.step(acts[k % POOL])
trace = torch.zeros((E, 4), dtype=torch.int64, device=dev)
_lib.check(_lib.lib().macm_set_trace(sim.engine._h, C.c_void_p(trace.data_ptr())))
torch.cuda.synchronize()
sim.engine.step(acts[0])
torch.cuda.synchronize()
t = trace.cpu().numpy().astype(np.int64)
st, en = (t[:, 0] - t[:, 0].min()) / 1e3, (t[:, 1] - t[:, 0].min()) / 1e3
smid = t[:, 3] & 0xffff
multi = (t[:, 3] >> 48) & 1
wpb = sim.engine.info.threads_per_block // 32
blk = np.arange(E) // wpb
nb = blk.max() + 1
b_st = np.array([st[blk == b].min() for b in range(nb)])
b_en = np.array([en[blk == b].max() for b in range(nb)])
b_med = np.array([np.median(en[blk == b]) for b in range(nb)])
b_sm = np.array([smid[blk == b][0] for b in range(nb)])
print("envs %d  threads/block %d  blocks %d  kernel %.1f us  (%.2f us per 4096 envs)" % (E, wpb * 32, nb, en.max(), en.max() * 4096 / E))
print("per-warp duration us: mean %.2f median %.2f p90 %.2f max %.2f" % ((en - st).mean(), np.median(en - st), np.percentile(en - st, 90), (en - st).max()))
print("per-block: duration mean %.2f median %.2f p90 %.2f ; median-warp end after block start %.2f" % (
    (b_en - b_st).mean(), np.median(b_en - b_st), np.percentile(b_en - b_st, 90), (b_med - b_st).mean()))
gaps, busy, resident = [], [], []
for s in np.unique(b_sm):
    idx = np.where(b_sm == s)[0]
    idx = idx[np.argsort(b_st[idx])]
    # blocks resident together on this SM (narrow blocks): time-weighted count
    ev = sorted([(b_st[i], 1) for i in idx] + [(b_en[i], -1) for i in idx])
    cur, last_t, area, cover = 0, ev[0][0], 0.0, 0.0
    for tt, d in ev:
        area += cur * (tt - last_t)
        cover += (tt - last_t) if cur > 0 else 0.0
        cur += d
        last_t = tt
    busy.append(cover / en.max())
    resident.append(area / max(cover, 1e-9))
    if wpb > 4:
        for a, b in zip(idx[:-1], idx[1:]):
            gaps.append(b_st[b] - b_en[a])
print("per-SM: fraction of the kernel with a block resident %.3f ; mean blocks resident while busy %.2f" % (np.mean(busy), np.mean(resident)))
if gaps:
    gaps = np.array(gaps)
    print("gap between consecutive blocks on an SM (us): mean %.2f median %.2f p90 %.2f max %.2f" % (
        gaps.mean(), np.median(gaps), np.percentile(gaps, 90), gaps.max()))
first = b_st < 1.0
print("first-wave blocks: duration mean %.2f ; later blocks: duration mean %.2f" % ((b_en - b_st)[first].mean(), (b_en - b_st)[~first].mean()))
print("multi envs %.3f ; warp duration multi %.2f others %.2f" % (multi.mean(), (en - st)[multi == 1].mean(), (en - st)[multi == 0].mean()))
