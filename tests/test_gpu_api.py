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


def test_host_path_is_ordered_against_the_callers_stream():
    """reset / step on the caller's stream followed by step_host (the handle's own stream) and back --
    with the pinned buffers already allocated, so nothing synchronises by accident."""
    import torch
    import gym_macm
    E, N = 4096, 16
    a = gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=2)
    b = gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=2)
    b.engine.pinned()
    torch.cuda.synchronize()
    g = torch.Generator().manual_seed(0)
    acts = torch.zeros((12, E, N, 4), dtype=torch.uint8)
    acts[..., :3] = torch.randint(0, 3, (12, E, N, 3), generator=g, dtype=torch.uint8)
    acts_d = acts.cuda()
    hact = [x.pin_memory() for x in acts]
    side = torch.cuda.Stream(device="cuda:0")
    for rep in range(3):
        a.reset(seed=40 + rep)
        with torch.cuda.stream(side):        # b's device-side calls go to a stream of their own
            b.reset(seed=40 + rep)           # sample + reset + observe kernels, still running when ...
            for k in range(3):
                b.step(acts_d[k])
        for k in range(3):
            a.step(acts_d[k])
        for k in range(3, 6):                # ... the host path takes over
            a.step(acts_d[k])
            b.engine.step_host(hact[k], wait=False)
        with torch.cuda.stream(side):
            for k in range(6, 9):            # and back to the caller's stream without a host_sync in between
                b.step(acts_d[k])
        for k in range(6, 9):
            a.step(acts_d[k])
        torch.cuda.synchronize()
        for n in ("posvel", "angsleep", "fat", "obs", "nn_idx", "rewards", "contact_count", "env_state"):
            assert torch.equal(a.state[n], b.state[n]), (rep, n)


def test_one_d2h_slab_and_pack_kernel():
    import torch
    import gym_macm
    E, N = 64, 10
    env = gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=1)
    t, lay = env.engine.t, env.engine._out_layout
    base = env.engine._out_slab.data_ptr()
    order = sorted(lay, key=lambda n: lay[n][0])
    assert order == ["obs", "rewards", "done", "nn_idx", "collided"]
    for n in order:      # the five outputs lie back to back in one device slab, 16-byte aligned
        assert t[n].data_ptr() == base + lay[n][0] and lay[n][0] % 16 == 0
    p = env.engine.pinned()
    assert all(p[n].data_ptr() - p["_slab"].data_ptr() == lay[n][0] for n in order)
    # any integer dtype / width-3 actions go through the library's own pack kernel
    a64 = torch.randint(0, 3, (E, N, 3), device="cuda:0")
    env.step(a64)
    want = torch.zeros((E, N, 4), dtype=torch.uint8, device="cuda:0")
    want[..., :3] = a64.to(torch.uint8)
    assert torch.equal(env._act4, want)
    env2 = gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=1)
    env2.step(want)
    env2.step_host(want.cpu())
    env.step(a64.to(torch.int32))
    torch.cuda.synchronize()
    for n in ("posvel", "obs", "rewards"):
        assert torch.equal(env.state[n], env2.state[n])
    hp = env2.engine.pinned()
    for n in order:
        assert torch.equal(hp[n], env2.state[n].cpu()), n


def test_overflow_is_reported():
    """A pile denser than the contact capacity: the flags are sticky, counted on the device, visible as
    `overflowed`, and step() raises with check_overflow=True."""
    import torch
    import gym_macm
    from gym_macm import _lib
    E, N = 8, 32
    rng = np.random.default_rng(0)
    pos = rng.uniform(-1.2, 1.2, (E, N, 2))
    pos[4:] = rng.uniform(-40, 40, (E - 4, N, 2))                 # envs 4.. are sparse
    ang = np.zeros((E, N))
    act = torch.ones((E, N, 4), dtype=torch.uint8, device="cuda:0")
    env = gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=None, max_contacts=40, max_touching=16)
    env.load_state(pos, ang)
    env.step(act)
    assert env.overflowed[:4].all() and not env.overflowed[4:].any()
    c, t = env.overflow_count()
    assert c == 4 and 1 <= t <= 4
    for _ in range(3):
        env.step(act)
    assert env.overflowed[:4].all()                                # sticky
    dbg = gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=None, max_contacts=40, max_touching=16,
                                check_overflow=True)
    dbg.load_state(pos, ang)
    with pytest.raises(_lib.MacmError, match="overflow"):
        dbg.step(act)
    ok = gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=3, check_overflow=True)
    ok.step(act)
    assert ok.overflow_count() == (0, 0)


def test_render_state_of_one_env():
    import gym_macm
    env = gym_macm.BatchedFlock(16, n_agents=[5], targets=[0, 0, 1, 1, 1], device="cuda:0", seed=4, start_spread=2.0)
    import torch
    env.step(torch.ones((16, 5, 4), dtype=torch.uint8, device="cuda:0"))
    fr = env.render_state(3)
    pv = env.state["posvel"][3].cpu().numpy()
    col = env.state["collided"][3].cpu().numpy()
    assert len(fr.world.bodies) == 5 and set(fr.gui_objects) == {"target0", "target1"}
    for i, b in enumerate(fr.world.bodies):
        assert b.transform.position == (float(pv[i, 0]), float(pv[i, 1])) and b.fixtures[0].shape.radius == 0.5
        assert tuple(b.userData.color) == ((1.0, 0.2, 0.2) if col[i] else (0.4, 0.4, 0.6))
    assert col.any()     # 5 agents in a 2 m square touch
    tdm = gym_macm.BatchedTDM(4, n_agents=[2, 2], device="cuda:0", seed=0)
    fr = tdm.render_state(1)
    assert [tuple(b.userData.color) for b in fr.world.bodies] == [(0.2, 0.2, 1.0)] * 2 + [(1.0, 0.2, 0.2)] * 2


def test_batch_pool_equals_serial_stepping():
    import torch
    import gym_macm
    E, N, K = 296, 64, 24
    serial = [gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=60 + r) for r in range(3)]
    pool = gym_macm.BatchPool([gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=60 + r) for r in range(3)])
    acts = torch.zeros((K, E, N, 4), dtype=torch.uint8, device="cuda:0")
    acts[..., :3] = torch.randint(0, 3, (K, E, N, 3), device="cuda:0", dtype=torch.uint8)
    torch.cuda.synchronize()
    for k in range(K):
        serial[k % 3].step(acts[k])
        assert pool.step(acts[k]) is pool.batches[k % 3]
    pool.synchronize()
    torch.cuda.synchronize()
    for a, b in zip(serial, pool.batches):
        for n in ("posvel", "angsleep", "fat", "obs", "nn_idx", "rewards", "contact_count"):
            assert torch.equal(a.state[n], b.state[n]), n
