"""Would a cell list pay for the nearest-agent search?  Candidate statistics of the bench workload (CPU, oracle).

64-agent flock envs from the reference's spawn distribution, 64 settle steps of random actions (the bench's state),
then an 8 x 8 grid over each env's bounding box (one agent per cell on average).  A warp searches the 3 x 3 cells
around each of its 64 agents; lanes run in lock step, so the warp pays for the LONGEST candidate list of its two
agent slots in every round, and an agent whose nearest neighbour is farther than one cell must fall back."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402

E, N, G = 256, 64, 8
rng = np.random.default_rng(0)
pos = 20.0 * (rng.random((E, N, 2)) - 0.5)
ang = rng.uniform(-1, 1, (E, N)) * np.pi
ref = oracle.OracleBatch(E, n_agents=N, n_targets=1, reward_mode=1)
ref.reset(pos, ang, targets=np.full((E, 1, 2), 30.0))
for k in range(64):
    ref.flock_step(rng.integers(0, 3, (E, N, 3)), 8)
p = ref.bodies()[..., 0:2].astype(np.float64)
mean_c, max_c, fallback, warp_cost = [], [], [], []
for e in range(E):
    lo, hi = p[e].min(0), p[e].max(0)
    h = (hi - lo).max() / G + 1e-9
    cell = np.minimum(((p[e] - lo) / h).astype(int), G - 1)
    d = np.linalg.norm(p[e][:, None] - p[e][None], axis=-1) + np.eye(N) * 1e9
    nn = d.min(1)
    near = (np.abs(cell[:, None, 0] - cell[None, :, 0]) <= 1) & (np.abs(cell[:, None, 1] - cell[None, :, 1]) <= 1)
    cnt = near.sum(1) - 1
    mean_c.append(cnt.mean()); max_c.append(cnt.max())
    # exact only if the nearest neighbour is closer than the distance to the edge of the 3 x 3 block (>= h)
    fallback.append(int((nn > h).sum()))
    warp_cost.append(max(cnt[:32].max(), 0) + max(cnt[32:].max(), 0))   # two agent slots per lane, lock step
print("candidates in the 3x3 cells: mean %.1f per agent, max over an env's agents %.1f (brute force: 63)" % (np.mean(mean_c), np.mean(max_c)))
print("lock-step cost of a warp: %.1f candidate rounds for its two slots (brute force: 63 packed rounds for both slots)" % np.mean(warp_cost))
print("agents per env whose nearest neighbour lies beyond one cell (exact fallback needed): %.1f" % np.mean(fallback))
