"""per-env reset and auto-reset (macm_reset_masked, MACM_FLAG_AUTO_RESET) -- the reference's
env.reset() (mvmnt.py:224-233, combat.py:229-239; repaired semantics, App. B3) for the envs of a batch that are
done (mvmnt.py:134-136, combat.py:171-182), the others untouched.  The step after a reset is held bit-exact to the
oracle's fresh world at the same (device-drawn) state."""
import numpy as np
import pytest

from _parity import compare_step, gpu_bodies, make_pair

pytestmark = pytest.mark.gpu


def _reset_oracle_envs(env, ref, which):
    """Give the oracle's envs `which` the states the device just drew for them."""
    st = env.state
    pv = st["posvel"].cpu().numpy().astype(np.float64)
    ang = st["angsleep"][..., 0].cpu().numpy().astype(np.float64)
    tg = st["targets"].cpu().numpy().astype(np.float64) if "targets" in st else None
    for e in which:
        ref.reset_env(int(e), pv[e, :, 0:2], ang[e], None if tg is None else tg[e])


def test_masked_reset_flock_bit_exact_after_reset():
    import torch
    E, N = 48, 16
    env, ref, rng = make_pair(E, N, targets=[0] * 8 + [1] * 8, seed=21, spread=6.0, reward_mode="linear")
    for k in range(25):
        act = rng.integers(0, 3, (E, N, 3))
        env.step(torch.as_tensor(act, device="cuda:0"))
        o = ref.flock_step(act)
    compare_step(env, ref, o, 24, "polar", "linear")
    before = {n: env.state[n].clone() for n in ("posvel", "angsleep", "fat", "targets", "contact_count", "env_state", "obs")}
    mask = torch.zeros(E, dtype=torch.uint8, device="cuda:0")
    which = [1, 5, 6, 30, 47]
    mask[which] = 1
    env.reset_done(mask, seed=99)
    torch.cuda.synchronize()
    keep = np.setdiff1d(np.arange(E), which)
    for n, b in before.items():   # the other envs are untouched
        assert torch.equal(env.state[n][keep], b[keep]), n
    st = env.state
    pv = st["posvel"].cpu().numpy()
    # fresh bodies from the reference's distributions (mvmnt.py:62-64), new targets (mvmnt.py:48-52), fresh world
    assert np.all(np.abs(pv[which][..., :2]) <= 10) and np.all(pv[which][..., 2:] == 0)
    assert not np.array_equal(pv[which][..., :2], before["posvel"][which][..., :2].cpu().numpy())
    d = np.hypot(*np.moveaxis(st["targets"][which].cpu().numpy(), -1, 0))
    assert d.min() >= 25 - 1e-4 and d.max() <= 60 + 1e-4
    es = st["env_state"].cpu().numpy()
    assert np.all(es[which, 0] == 0) and np.all((es[which, 1] & 7) == 1) and np.all(es[which, 1] >> 8 == 1)
    assert np.all(es[keep, 1] >> 8 == 0) and np.all(st["contact_count"][which].cpu().numpy() == 0)
    assert np.array_equal(env.episode.cpu().numpy(), (np.isin(np.arange(E), which)).astype(np.int32))
    fat = st["fat"].cpu().numpy()
    assert np.array_equal(fat[which][..., 0], (pv[which][..., 0] - np.float32(0.5)) - np.float32(0.1))
    # the oracle gets the same new episodes; observations of the fresh worlds, then 30 more steps bit-exact
    _reset_oracle_envs(env, ref, which)
    oo = ref.flock_observe()
    assert np.array_equal(st["nn_idx"].cpu().numpy(), oo["nn_idx"])
    for k in range(30):
        act = rng.integers(0, 3, (E, N, 3))
        env.step(torch.as_tensor(act, device="cuda:0"))
        o = ref.flock_step(act)
        if k in (0, 1, 29):
            compare_step(env, ref, o, 25 + k, "polar", "linear")
    # a second reset of the same envs draws another episode
    first = env.state["posvel"][which].clone()
    env.reset_done(mask, seed=99)
    pv2 = env.state["posvel"][which]
    assert not torch.equal(pv2, first) and int(env.episode[1]) == 2
    # same (seed, env, episode) -> same draw, whatever happened in between
    other, _, _ = make_pair(E, N, targets=[0] * 8 + [1] * 8, seed=3, spread=6.0, reward_mode="linear")
    other.reset_done(mask, seed=99)
    other.reset_done(mask, seed=99)
    assert torch.equal(other.state["posvel"][which], pv2) and torch.equal(other.state["targets"][which], env.state["targets"][which])


def test_auto_reset_on_time_limit():
    """done after the time limit (mvmnt.py:134-136): with auto_reset the env starts over by itself."""
    import torch
    import gym_macm
    E, N = 40, 8
    env = gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=2, time_limit=0.2, auto_reset=True)
    done_step = env.engine.info.done_step      # 13 for 0.2 s at 60 Hz
    assert done_step == 13
    act = torch.ones((E, N, 4), dtype=torch.uint8, device="cuda:0")
    for k in range(1, 2 * done_step + 1):
        pv_before = env.state["posvel"].clone()
        env.step(act)
        torch.cuda.synchronize()
        if k % done_step == 0:
            assert bool(env.done.all())                      # the learner still sees the flag ...
            assert int(env.step_count.max()) == 0            # ... and the envs are already in their next episode
            assert torch.equal(env.episode, torch.full((E,), k // done_step, dtype=torch.int32, device="cuda:0"))
            assert not torch.equal(env.state["posvel"], pv_before)
        else:
            assert not bool(env.done.any()) and int(env.step_count.min()) == k % done_step
    # without the flag nothing resets
    env2 = gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=2, time_limit=0.2)
    for k in range(done_step + 2):
        env2.step(act)
    assert bool(env2.done.all()) and int(env2.step_count.min()) == done_step + 2 and int(env2.episode.max()) == 0


def test_tdm_envs_end_at_different_steps_and_reset_alone():
    """TDM matches end when one team is left (combat.py:171-182), every env at its own step; reset_done() restarts
    exactly those, and the restarted envs step bit-exactly like fresh oracle worlds."""
    import torch
    import gym_macm
    from oracle import oracle
    E, teams = 64, [3, 3]
    env = gym_macm.BatchedTDM(E, n_agents=teams, device="cuda:0", seed=7, world_width=5.0, world_height=4.0,
                              max_contacts=15, max_touching=16)
    N = env.engine.N
    team = np.array(env.teams, np.uint8)
    ref = oracle.OracleBatch(E, env_kind=oracle.TDM, n_agents=N, n_targets=0)
    torch.cuda.synchronize()
    ref.reset(env.state["posvel"][..., :2].cpu().numpy().astype(np.float64),
              env.state["angsleep"][..., 0].cpu().numpy().astype(np.float64), team=team)
    rng = np.random.default_rng(5)
    ended_at = np.full(E, -1)
    n_resets = 0
    for k in range(900):
        # the combat actor (bots.py:3-16) closes in and strikes, so matches do end; every 5th step is random
        if k % 5:
            act = env.bot_actions("combat").cpu().numpy().astype(np.int64)
        else:
            act = np.concatenate([rng.integers(0, 3, (E, N, 3)), (rng.random((E, N, 1)) < 0.9).astype(np.int64)], -1)
        env.step(torch.as_tensor(act, device="cuda:0"))
        o = ref.tdm_step(act)
        done = env.state["done"].cpu().numpy()
        assert np.array_equal(done, o["done"]), k
        assert np.array_equal(gpu_bodies(env), ref.bodies()), "step %d" % k
        new = np.flatnonzero(done)
        if len(new) and k % 7 == 0:      # the learner collects finished matches every few steps
            ended_at[new] = k
            env.reset_done(seed=1234)    # mask = done
            torch.cuda.synchronize()
            assert np.all(env.step_count.cpu().numpy()[new] == 0) and bool(env.alive[new].all())
            assert np.all(env.state["env_state"][:, 3].cpu().numpy()[new] == -1)
            _reset_oracle_envs(env, ref, new)
            n_resets += len(new)
    assert n_resets >= 8 and len(set(ended_at[ended_at >= 0])) >= 3      # matches ended at different steps
    assert int(env.episode.max()) >= 1


def test_auto_reset_inside_a_rollout_equals_single_steps():
    """With auto_reset a K-step macm_rollout restarts finished envs between its steps exactly as K single steps with
    the flag do (same draws, same order of events): per-step outputs and final state bit-identical."""
    import torch
    import gym_macm
    from test_gpu_rollout import _same_state
    # Flock: every env hits the time limit on step 13 (mvmnt.py:134-136), three times in 40 steps
    E, N, K = 48, 64, 40
    one, many = [gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=8, time_limit=0.2, auto_reset=True,
                                       start_spread=10.0, targets=[i % 2 for i in range(N)]) for _ in range(2)]
    for e in (one, many):
        e.engine.set_auto_reset_seed(77)
    acts = torch.zeros((K, E, N, 4), dtype=torch.uint8, device="cuda:0")
    acts[..., :3] = torch.randint(0, 3, (K, E, N, 3), device="cuda:0", dtype=torch.uint8)
    per = {n: [] for n in ("obs", "nn_idx", "rewards", "collided", "done")}
    for k in range(K):
        one.step(acts[k])
        for n in per:
            per[n].append(one.state[n].clone())
    out = many.rollout(acts)
    torch.cuda.synchronize()
    for n in per:
        assert torch.equal(out[n], torch.stack(per[n])), n
    _same_state(one, many, "flock auto-reset")
    assert torch.equal(one.state["targets"], many.state["targets"])
    assert int(one.episode.min()) == 3 and int(out["done"].sum()) == 3 * E
    # TDM with the combat actor: matches end at different steps (combat.py:171-182)
    E, K = 40, 400
    one, many = [gym_macm.BatchedTDM(E, n_agents=[3, 3], device="cuda:0", seed=9, world_width=5.0, world_height=4.0,
                                     auto_reset=True) for _ in range(2)]
    rew, done = [], []
    for k in range(K):
        one.step(one.bot_actions("combat"))
        rew.append(one.state["rewards"].clone())
        done.append(one.state["done"].clone())
    out = many.rollout(None, n_steps=K, policy="combat", want=("rewards", "done"))
    assert torch.equal(out["rewards"], torch.stack(rew)) and torch.equal(out["done"], torch.stack(done))
    _same_state(one, many, "tdm auto-reset")
    d = torch.stack(done).cpu().numpy()                      # [K, E]
    first = np.where(d.any(0), d.argmax(0), -1)
    assert int(one.episode.max()) >= 1 and len(set(first[first >= 0].tolist())) >= 3      # matches ended at different steps
