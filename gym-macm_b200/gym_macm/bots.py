"""Scripted actors with the call signature the env hosts expect (`actor(obs) -> action`), after
test_scripts/bots.py of the reference.  These are the HOST versions used by the dict API's
`actions=None` mode (mvmnt.py:86-92); `BatchedFlock.bot_actions` runs the same policies on the
device for closed-loop rollouts."""
import numpy as np


def idle(obs=None):
    return np.array([1, 1, 1, 0])


def forward(obs=None):
    return np.array([2, 1, 1, 0])


def rotate(obs=None):
    return np.array([1, 1, 2, 0])


def diag(obs=None):
    return np.array([2, 2, 1, 0])


def circle(obs=None):
    return np.array([2, 1, 2, 0]) if np.random.rand() < 0.5 else np.array([2, 1, 1, 0])


def flock(obs, coord="polar"):
    """Turn towards the target node and walk when it is within +-45 degrees; stop inside 1 m."""
    own = list(obs.values())[0]
    tnodes = [n for n in own["nodes"] if n["type"] == 1]
    if not tnodes:
        return idle()[:3]
    pos = tnodes[0]["position"]
    if pos[0] < 1:
        return idle()[:3]
    if len(pos) == 3:
        turn, ahead = np.sign(pos[2]) + 1, int(pos[1] > np.cos(np.pi / 4)) + 1
    else:
        turn, ahead = np.sign(pos[1]) + 1, int(np.abs(pos[1]) < (np.pi / 4)) + 1
    return np.array([ahead, 1, int(turn)])


def flock_cont(obs, coord="polar"):
    return np.array([1, 1])


def combat(obs):
    """Face and approach the nearest enemy, strike inside 3 m."""
    enemies = [a for a in obs["agents"] if a["type"] == 0]
    if not enemies:
        return idle()
    tgt = min(enemies, key=lambda a: a["position"][0])
    r, th = tgt["position"][0], tgt["position"][1]
    return np.array([int(np.abs(th) < (np.pi / 5)) + 1, 1, int(np.sign(th) + 1), int(r < 3)])
