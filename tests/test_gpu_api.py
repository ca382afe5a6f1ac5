"""GPU tests of the C-ABI surface beyond the step itself: device sampler, shards, bots on the
device, host-buffer stepping, status codes."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_sampler_distributions_and_shard_invariance():
    import torch
    import gym_macm
    full = gym_macm.BatchedFlock(64, n_agents=[16], targets=[0] * 8 + [1] * 8, device="cuda:0", seed=5)
    pv = full.state["posvel"].cpu().numpy()
    ang = full.state["angsleep"][..., 0].cpu().numpy()
    tg = full.targets.cpu().numpy()
    # mvmnt.py:62-64: pos = 20 * (U - 0.5), angle = uniform(-1, 1) * pi; velocities zero
    assert pv[..., :2].min() >= -10 and pv[..., :2].max() <= 10 and np.all(pv[..., 2:] == 0)
    assert abs(pv[..., :2].mean()) < 0.5 and 5.0 < pv[..., :2].std() < 6.5
    assert np.abs(ang).max() <= np.pi + 1e-6 and abs(ang.mean()) < 0.3
    # mvmnt.py:50-52: target distance in [25, 60]
    d = np.hypot(tg[..., 0], tg[..., 1])
    assert d.min() >= 25 - 1e-4 and d.max() <= 60 + 1e-4
    # fresh worlds: fat = tight +- 0.1, no contacts, first-step flag
    fat = full.state["fat"].cpu().numpy()
    assert np.array_equal(fat[..., 0], (pv[..., 0] - np.float32(0.5)) - np.float32(0.1))
    assert int(full.state["contact_count"].sum()) == 0 and np.all(full.state["env_state"][:, 1].cpu().numpy() == 1)
    # the same batch cut into two shards keyed by the global env index
    parts = [gym_macm.BatchedFlock(32, n_agents=[16], targets=[0] * 8 + [1] * 8, device="cuda:0", seed=5,
                                   env_index_base=b) for b in (0, 32)]
    got = torch.cat([p.state["posvel"] for p in parts], 0).cpu().numpy()
    assert np.array_equal(got, pv)
    assert np.array_equal(torch.cat([p.targets for p in parts], 0).cpu().numpy(), tg)
    # and the shards step like the whole
    act = torch.randint(0, 3, (64, 16, 3), device="cuda:0")
    for _ in range(20):
        full.step(act)
        parts[0].step(act[:32])
        parts[1].step(act[32:])
    got = torch.cat([p.state["posvel"] for p in parts], 0)
    assert torch.equal(got, full.state["posvel"])
    assert torch.equal(torch.cat([p.state["obs"] for p in parts], 0), full.state["obs"])


def test_device_bots_match_host_bots():
    import torch
    import gym_macm
    from gym_macm import bots
    from gym_macm.envs.mvmnt import obs_to_dict
    for coord in ("polar", "cartesian"):
        env = gym_macm.BatchedFlock(16, n_agents=[8], device="cuda:0", seed=3, coord=coord)
        for k in range(30):
            a = env.bot_actions("flock")
            ob = env.state["obs"].cpu().numpy()
            nn = env.state["nn_idx"].cpu().numpy()
            for e in range(16):
                d = obs_to_dict(list(range(8)), nn[e], ob[e], 8)
                want = np.array([bots.flock({i: d[i]}) for i in range(8)])
                assert np.array_equal(a[e, :, :3].cpu().numpy(), want), (coord, k, e)
            env.step(a)
        r = env.bot_actions("random", seed=7).cpu().numpy()
        assert r[..., :3].max() <= 2 and r[..., 3].max() == 0 and len(np.unique(r[..., :3])) == 3
        assert np.array_equal(env.bot_actions("forward").cpu().numpy()[0, 0], [2, 1, 1, 0])
        assert np.array_equal(env.bot_actions("diag").cpu().numpy()[0, 0], [2, 2, 1, 0])


def test_flock_bot_closed_loop_reaches_target():
    import gym_macm
    env = gym_macm.BatchedFlock(32, n_agents=[6], device="cuda:0", seed=11)
    r0 = env.state["obs"][..., 2].mean().item()
    best = -1.0
    for k in range(1500):
        env.step(env.bot_actions("flock"))
        best = max(best, env.rewards.max().item())
    r1 = env.state["obs"][..., 2].mean().item()
    assert r0 > 25 and r1 < 4.0   # 25 s at ~4 m/s: every flock sits around its target
    assert best == 1.0            # inside the 7 m reward radius on the way in ...
    # ... and once piled up at the target everybody is in somebody's fat AABB: -1 (mvmnt.py:162-164)
    assert env.collided.float().mean().item() > 0.9


def test_step_host_equals_step_device():
    import torch
    import gym_macm
    a = gym_macm.BatchedFlock(128, n_agents=[10], device="cuda:0", seed=2)
    b = gym_macm.BatchedFlock(128, n_agents=[10], device="cuda:0", seed=2)
    g = torch.Generator().manual_seed(0)
    for k in range(25):
        act = torch.zeros((128, 10, 4), dtype=torch.uint8)
        act[..., :3] = torch.randint(0, 3, (128, 10, 3), generator=g, dtype=torch.uint8)
        obs_d, rew_d = a.step(act.cuda())
        obs_h, rew_h = b.step_host(act)
        torch.cuda.synchronize()
        assert torch.equal(obs_d["nodes"][0]["position"].cpu(), obs_h["nodes"][0]["position"])
        assert torch.equal(obs_d["nodes"][1]["position"].cpu(), obs_h["nodes"][1]["position"])
        assert torch.equal(obs_d["nodes"][0]["id"].cpu(), obs_h["nodes"][0]["id"])
        assert torch.equal(rew_d.cpu(), rew_h)
    assert a.engine.launch_count == b.engine.launch_count


def test_status_codes():
    import torch
    from gym_macm import _lib
    L = _lib.lib()
    p = _lib.default_params(_lib.FLOCK)
    p.n_envs, p.n_agents = 4, 4
    h = C.c_void_p()
    assert L.macm_create(C.byref(h), C.byref(p), 0) == 0
    dummy = torch.zeros(64, dtype=torch.uint8, device="cuda:0")
    assert L.macm_step(h, C.c_void_p(dummy.data_ptr()), None) == -3          # MACM_E_UNBOUND
    assert L.macm_reset(h, None) == -3
    b = _lib.MacmBuffers()
    assert L.macm_bind(h, C.byref(b)) == -3                                   # NULL buffers
    assert L.macm_step(h, None, None) == -1                                   # MACM_E_INVALID
    assert L.macm_destroy(h) == 0
    p.n_agents = 1
    assert L.macm_create(C.byref(h), C.byref(p), 0) == -1
    p.n_agents, p.n_targets = 4, 99
    assert L.macm_create(C.byref(h), C.byref(p), 0) == -1
    p.n_targets = 1
    assert L.macm_create(C.byref(h), C.byref(p), 1000) == -2                  # no such device
    assert b"invalid" in L.macm_strerror(-1)


def test_tdm_dict_api_runs():
    import random
    import gym_macm
    from gym_macm import bots
    random.seed(3)
    env = gym_macm.make("gym_macm:cm-tdm-v0", n_agents=[2, 2], actors=[[bots.combat] * 2] * 2, world_width=4, world_height=3)
    assert sorted(env.obs) == ["00", "01", "10", "11"] and env.obs["00"]["myTeam"] == 0
    assert len(env.obs["00"]["agents"]) == 3 and env.obs["00"]["myHealth"][0] == 1.0
    for k in range(600):
        obs, rewards = env.step()
        if env.done:
            break
    assert env.done and (env.winner in (0, 1) or sum(env.n_alive) == 0 or env.time_passed > 60)
    assert set(rewards.values()) <= {0, -1}
    env.close()


def test_batches_on_their_own_streams():
    """Independent batches stepped on two streams (they overlap on the device) end up where serial stepping does."""
    import torch
    import gym_macm
    E, N, K = 296, 64, 30
    ss = [torch.cuda.Stream(device="cuda:0") for _ in range(2)]
    serial = [gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=20 + r) for r in range(4)]
    torch.cuda.synchronize()
    par = [gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=20 + r, stream=ss[r & 1]) for r in range(4)]
    acts = torch.zeros((K, E, N, 4), dtype=torch.uint8, device="cuda:0")
    acts[..., :3] = torch.randint(0, 3, (K, E, N, 3), device="cuda:0", dtype=torch.uint8)
    torch.cuda.synchronize()
    for k in range(K):
        for r in range(4):
            serial[r].step(acts[(k + r) % K])
            par[r].step(acts[(k + r) % K])
    torch.cuda.synchronize()
    for a, b in zip(serial, par):
        for n in ("posvel", "angsleep", "fat", "obs", "nn_idx", "rewards", "contact_count"):
            assert torch.equal(a.state[n], b.state[n]), n
