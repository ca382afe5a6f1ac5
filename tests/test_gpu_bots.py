"""the scripted actors of test_scripts/bots.py on the device -- `combat` (bots.py:3-16) and `circle`
(bots.py:31-35) beside the ones test_gpu_api / test_gpu_rollout cover -- held to the host versions in gym_macm.bots
acting on the reference's observation dicts, and the in-rollout actors to the standalone bot kernel."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_combat_bot_matches_host_bot_on_the_obs_dicts():
    import torch
    import gym_macm
    from gym_macm import bots
    from gym_macm.envs.combat import tdm_obs_to_dict
    E, teams = 24, [4, 4, 3]
    env = gym_macm.BatchedTDM(E, n_agents=teams, device="cuda:0", seed=11, world_width=8.0, world_height=8.0)
    N = env.engine.N
    ids = list(range(N))
    struck = 0
    for k in range(260):
        a = env.bot_actions("combat")
        torch.cuda.synchronize()
        if k % 13 == 0 or k > 250:
            ob = env.state["obs"].cpu().numpy().reshape(E, N, N, 4)
            alive = env.alive.cpu().numpy()
            health = env.health.cpu().numpy()
            got = a.cpu().numpy()
            for e in range(E):
                d = tdm_obs_to_dict(ids, env.teams, health[e], alive[e], ob[e])
                for i in ids:
                    want = bots.combat(d[i]) if alive[e, i] else bots.idle()
                    assert np.array_equal(got[e, i], want), (k, e, i, got[e, i], want)
            struck += int(got[..., 3].sum())
        env.step(a)
    assert struck > 0                                   # somebody came within 3 m and struck
    assert int((~env.alive).sum()) > 0                  # and people died of it (melee range 2 m, four hits)


def test_combat_and_circle_inside_a_rollout_equal_bot_kernel_plus_step():
    import torch
    import gym_macm
    from test_gpu_rollout import _same_state
    E, K = 40, 200
    one, many = [gym_macm.BatchedTDM(E, n_agents=[5, 5, 5], device="cuda:0", seed=4, world_width=8.0, world_height=9.0)
                 for _ in range(2)]
    rew, done = [], []
    for k in range(K):
        one.step(one.bot_actions("combat"))
        rew.append(one.state["rewards"].clone())
        done.append(one.state["done"].clone())
    out = many.rollout(None, n_steps=K, policy="combat", want=("rewards", "done"))
    assert torch.equal(out["rewards"], torch.stack(rew)) and torch.equal(out["done"], torch.stack(done))
    _same_state(one, many, "combat actor")
    assert int((~one.alive).sum()) > 0
    # circle: forward with a coin-flip turn, same draws in the bot kernel and inside the rollout
    one, many = [gym_macm.BatchedFlock(E, n_agents=[12], device="cuda:0", seed=6, start_spread=6.0) for _ in range(2)]
    turns = 0
    for k in range(50):
        a = one.bot_actions("circle", seed=5)
        assert bool((a[..., 0] == 2).all()) and bool((a[..., 1] == 1).all()) and bool(((a[..., 2] == 1) | (a[..., 2] == 2)).all())
        turns += int((a[..., 2] == 2).sum())
        one.step(a)
    assert 0.45 < turns / (50 * E * 12) < 0.55          # np.random.rand() < 0.5 (bots.py:32)
    many.rollout(None, n_steps=50, policy="circle", seed=5)
    _same_state(one, many, "circle actor")


def test_bot_status_codes():
    import torch
    import gym_macm
    from gym_macm import _lib
    L = _lib.lib()
    flock = gym_macm.BatchedFlock(4, n_agents=[4], device="cuda:0", seed=0)
    out = torch.zeros((4, 4, 4), dtype=torch.uint8, device="cuda:0")
    # the combat actor reads a TDM observation row: refused on Flock, never silently idle
    assert L.macm_bot_actions(flock.engine._h, _lib.BOTS["combat"], 0, C.c_void_p(out.data_ptr()), None) == -6
    assert L.macm_rollout(flock.engine._h, None, 2, _lib.BOTS["combat"], 0, None, None) == -6
    assert L.macm_bot_actions(flock.engine._h, 8, 0, C.c_void_p(out.data_ptr()), None) == -1
    tdm = gym_macm.BatchedTDM(4, n_agents=[2, 2], device="cuda:0", seed=0)
    assert L.macm_bot_actions(tdm.engine._h, _lib.BOTS["flock"], 0, C.c_void_p(out.data_ptr()), None) == -6
    assert L.macm_bot_actions(tdm.engine._h, _lib.BOTS["combat"], 0, C.c_void_p(out.data_ptr()), None) == 0
    assert L.macm_rollout(tdm.engine._h, None, 2, _lib.BOTS["combat"], 0, None, None) == 0
    torch.cuda.synchronize()
