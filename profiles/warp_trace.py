#!/usr/bin/env python
"""Per-env timing of one step launch (macm_set_trace): where the kernel's wall time goes.

Runs the bench workload (4096 envs x 64 agents, settled), switches the trace hook on for a few
back-to-back launches and prints: launch gap, ramp (first -> last warp start), per-warp duration
by contact structure, and the tail (last warp end vs the median warp end)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gym-macm_b200"))
import ctypes as C

import gym_macm
from gym_macm import _lib

N = 64
settle = int(sys.argv[1]) if len(sys.argv) > 1 else 64
E = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
dev = torch.device("cuda", 0)
sims = [gym_macm.BatchedFlock(E, n_agents=[N], reward_mode="linear", device=dev, seed=1234 + r) for r in range(4)]
g = torch.Generator(device=dev)
g.manual_seed(99)
POOL = 61
acts = torch.zeros((POOL, E, N, 4), dtype=torch.uint8, device=dev)
acts[..., :3] = torch.randint(0, 3, (POOL, E, N, 3), generator=g, device=dev, dtype=torch.uint8)
for k in range(settle):
    for r, s in enumerate(sims):
        s.engine.step(acts[(k + 7 * r) % POOL])
traces = [torch.zeros((E, 4), dtype=torch.int64, device=dev) for _ in sims]
for s, t in zip(sims, traces):
    _lib.check(_lib.lib().macm_set_trace(s.engine._h, C.c_void_p(t.data_ptr())))
torch.cuda.synchronize()
for rep in range(3):
    for r, s in enumerate(sims):
        s.engine.step(acts[(settle + rep + 7 * r) % POOL])
torch.cuda.synchronize()
tr = [t.cpu().numpy().astype(np.int64) for t in traces]
t00 = min(t[:, 0].min() for t in tr)
print("launch  start_first  start_last  end_median  end_p90  end_last   (us, relative to the first launch)")
prev_end = None
for i, t in enumerate(tr):
    st, en = (t[:, 0] - t00) / 1e3, (t[:, 1] - t00) / 1e3
    gap = "" if prev_end is None else "  gap after previous launch %.2f us" % (st.min() - prev_end)
    print("%d  %8.2f %8.2f %8.2f %8.2f %8.2f%s" % (i, st.min(), st.max(), np.median(en), np.percentile(en, 90), en.max(), gap))
    prev_end = en.max()
t = tr[-1]
cyc = t[:, 2]
smid = t[:, 3] & 0xffff
tc = (t[:, 3] >> 16) & 0xffff
nlev = (t[:, 3] >> 32) & 0xffff
multi = (t[:, 3] >> 48) & 1
slot = (t[:, 3] >> 49) & 0x7fff
dur = (t[:, 1] - t[:, 0]) / 1e3
print("per-env duration us: mean %.2f median %.2f p90 %.2f p99 %.2f max %.2f; cycles mean %.0f max %d" % (
    dur.mean(), np.median(dur), np.percentile(dur, 90), np.percentile(dur, 99), dur.max(), cyc.mean(), cyc.max()))
print("touching contacts/env %.2f, multi %.3f" % (tc.mean(), multi.mean()))
print("tc  multi  nlev   envs   mean_us   max_us")
for m in (0, 1):
    for k in sorted(set(tc[multi == m])):
        sel = (tc == k) & (multi == m)
        print("%2d    %d    %4.1f  %5d   %7.2f  %7.2f" % (k, m, nlev[sel].mean(), sel.sum(), dur[sel].mean(), dur[sel].max()))
# per-SM: when does the SM's last warp end, and which env was it
ends = (t[:, 1] - t[:, 0].min()) / 1e3
last = np.array([ends[smid == s].max() for s in np.unique(smid)])
print("per-SM last-warp end (us): min %.2f median %.2f max %.2f ; SMs used %d" % (last.min(), np.median(last), last.max(), len(last)))
for q in (50, 75, 90, 100):
    print("  %3d%% of all warps have ended by %.2f us" % (q, np.percentile(ends, q)))
# per-SM view: is an SM slow because of what it holds (multi envs) or for its own reasons?
print("SM  envs  multi  mean_us  max_us   mean_us(non-multi)")
rows = []
for s_ in np.unique(smid):
    sel = smid == s_
    rows.append((s_, sel.sum(), int(multi[sel].sum()), dur[sel].mean(), ends[sel].max(),
                 dur[sel & (multi == 0)].mean() if (sel & (multi == 0)).any() else 0.0, int(tc[sel].sum())))
rows.sort(key=lambda r: r[4])
for r in rows[:6] + rows[-10:]:
    print("%3d  %3d  %3d   %6.2f  %6.2f   %6.2f   tc_sum %d" % r)
a = np.array([(r[1], r[2], r[3], r[4], r[6]) for r in rows], dtype=np.float64)
print("corr(last end, multi envs on the SM) = %.2f ; corr(last end, touching contacts on the SM) = %.2f ; corr(last end, envs) = %.2f" % (
    np.corrcoef(a[:, 3], a[:, 1])[0, 1], np.corrcoef(a[:, 3], a[:, 4])[0, 1], np.corrcoef(a[:, 3], a[:, 0])[0, 1]))
print("envs per SM histogram:", np.bincount(a[:, 0].astype(int)))
# does the position of a warp among the SM's resident blocks (launch order) decide when it ends?
env_id = slot
wpb = sims[0].engine.info.threads_per_block // 32
blk = env_id // wpb
rank = np.zeros(E, np.int64)
if wpb > 4:     # one block per SM: position of the warp inside the block, by sub-partition round
    rank = (slot % wpb) // 4
else:
    for s_ in np.unique(smid):
        sel = np.where(smid == s_)[0]
        ub = np.unique(blk[sel])
        rank[sel] = np.searchsorted(ub, blk[sel])
print("block rank on its SM (0 = lowest block id)   envs   mean_us  (non-multi only)  p90")
for r in range(int(rank.max()) + 1):
    sel = (rank == r) & (multi == 0)
    if sel.any():
        print("   %d   %5d   %6.2f   %6.2f" % (r, sel.sum(), dur[sel].mean(), np.percentile(dur[sel], 90)))
print("warp in block   mean_us")
for w in range(4):
    sel = ((env_id % 4) == w) & (multi == 0)
    print("   %d   %6.2f" % (w, dur[sel].mean()))
print("threads per block %d" % (wpb * 32))
print("multi envs by block rank:", np.bincount(rank[multi == 1]).tolist(), "; mean end of multi envs %.2f us, of the others %.2f us" % (ends[multi == 1].mean(), ends[multi == 0].mean()))
print("the 16 warps that end last:  end_us  dur_us  rank  multi  tc  islands  SM  slot")
for k in np.argsort(-ends)[:16]:
    print("   %6.2f  %6.2f   %d     %d    %2d    %2d   %3d  %4d" % (ends[k], dur[k], rank[k], multi[k], tc[k], nlev[k], smid[k], slot[k]))
