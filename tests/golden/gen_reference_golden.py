#!/usr/bin/env python
"""Generate golden vectors by running the REFERENCE's own Flock host code.

    python tests/golden/gen_reference_golden.py       # writes tests/golden/flock_ref_*.npz

/root/reference/gym_macm (mvmnt.py, cm_framework.py, settings.py, backends/no_render.py) is
imported UNMODIFIED; `Box2D` and `gym`, which cannot be installed here, are replaced by the
stand-ins in tests/golden/shim/, whose b2World is backed by the CPU oracle.  So these vectors
pin the reference's host logic (action decode and force in float64, reward pass incl. the
"every listed contact" rule, observation pass, episode clock, draw order of the initial state)
executed by the reference's own Python; the engine underneath is our restatement (parity with
pybox2d itself stays unpinned).  Needs /root/reference, so it runs in the build container only;
its outputs are committed.
"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def main():
    assert os.path.isdir(REF), "the reference tree is needed to regenerate the golden vectors"
    sys.path.insert(0, ROOT)                       # for `oracle`
    sys.path.insert(0, os.path.join(HERE, "shim"))  # Box2D, gym stand-ins
    sys.path.insert(0, REF)                         # the reference's gym_macm
    import gym_macm.envs.mvmnt as mv               # noqa: E402  (reference code)
    assert mv.__file__.startswith(REF), mv.__file__

    cases = [
        # BASELINE config 1: n_agents=[4], binary reward, discrete, polar
        dict(name="cfg1_n4_binary", seed=0, n_agents=[4], targets=None, steps=400, kw={}),
        # crowded 4 agents (contacts, -1 rewards, position correction)
        dict(name="n4_crowded", seed=1, n_agents=[4], targets=None, steps=200, kw=dict(start_spread=2)),
        # BASELINE config 3 shape: per-agent target gather
        dict(name="cfg3_n6_targets", seed=2, n_agents=[6], targets=[0, 0, 1, 1, 2, 2], steps=200, kw=dict(start_spread=6)),
        # linear reward + cartesian coordinates, 16 agents, dense
        dict(name="n16_linear_cartesian", seed=3, n_agents=[16], targets=None, steps=150,
             kw=dict(reward_mode="linear", coord="cartesian", start_spread=6)),
        # continuous actions (bug-compatible normalisation), short episode: done flips inside the run
        dict(name="n8_continuous_done", seed=4, n_agents=[8], targets=None, steps=80,
             kw=dict(action_mode="continuous", start_spread=4, time_limit=1)),
        # 64 agents as in BASELINE config 2 (one env)
        dict(name="cfg2_n64_linear", seed=5, n_agents=[64], targets=None, steps=60, kw=dict(reward_mode="linear")),
    ]
    for c in cases:
        random.seed(c["seed"])
        rng = np.random.default_rng(c["seed"])
        env = mv.Flock(n_agents=c["n_agents"], targets=c["targets"], **c["kw"])
        N = sum(c["n_agents"])
        world = env.framework.world
        b0 = world._ow.bodies().copy()
        D = 2 if env.settings.coord == "polar" else 3
        rec = dict(pos0=b0[:, 0:2].astype(np.float64), angle0=b0[:, 4].astype(np.float64),
                   targets=np.array([[t.x, t.y] for t in env.targets], np.float64),
                   target_idx=np.array(env.targets_idx, np.uint8), N=N,
                   settings=np.array([env.settings.reward_mode, env.settings.action_mode, env.settings.coord,
                                      str(env.settings.time_limit), str(env.settings.start_spread)]))

        def obs_arrays(obs):
            nn_id = np.array([obs[i]["nodes"][0]["id"] for i in range(N)], np.int32)
            nn = np.array([obs[i]["nodes"][0]["position"] for i in range(N)], np.float64)
            tg = np.array([obs[i]["nodes"][1]["position"] for i in range(N)], np.float64)
            assert all(obs[i]["nodes"][1]["id"] == N and obs[i]["nodes"][0]["type"] == 0 for i in range(N))
            return nn_id, nn, tg

        nn_id, nn, tg = obs_arrays(env.obs)
        rec.update(obs0_nn_id=nn_id, obs0_nn=nn, obs0_tg=tg)
        A, R, NI, NN, TG, DONE, BODY, CNT = [], [], [], [], [], [], [], []
        for k in range(c["steps"]):
            if env.settings.action_mode == "discrete":
                act = rng.integers(0, 3, (N, 3))
                actions = {i: act[i] for i in range(N)}
            else:
                act = rng.uniform(-1, 1, (N, 2)).astype(np.float32).astype(np.float64)
                actions = {i: act[i] for i in range(N)}
            obs, rewards = env.step(actions)
            nn_id, nn, tg = obs_arrays(obs)
            A.append(act); NI.append(nn_id); NN.append(nn); TG.append(tg)
            R.append(np.array([rewards[i] for i in range(N)], np.float64))
            DONE.append(env.done)
            BODY.append(world._ow.bodies().copy())
            CNT.append(len(world.contacts))
        rec.update(actions=np.array(A), rewards=np.array(R), nn_id=np.array(NI), nn=np.array(NN), tg=np.array(TG),
                   done=np.array(DONE), bodies=np.array(BODY), n_contacts=np.array(CNT, np.int32))
        out = os.path.join(HERE, "flock_ref_%s.npz" % c["name"])
        np.savez_compressed(out, **rec)
        print("%-28s N=%2d steps=%3d contacts(max)=%3d collided agent-steps=%d done@%s -> %s" % (
            c["name"], N, c["steps"], max(CNT), int((np.array(R) == -1).sum()),
            (int(np.argmax(DONE)) + 1) if any(DONE) else None, os.path.relpath(out, ROOT)))


if __name__ == "__main__":
    main()
