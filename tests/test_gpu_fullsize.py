"""BASELINE.json's configurations at their FULL sizes.

The oracle cannot step 65,536 worlds in test time, and does not have to: environments never interact
(nothing in mvmnt.py / combat.py couples two worlds), so any subset of a batch must evolve exactly as the
same worlds stepped alone.  Each test runs the whole batch on the GPU from device-sampled initial states,
hands the initial state and the action stream of a random subset of envs to the CPU oracle and holds the
subset to the usual bars (bodies, contact lists, nearest-agent ids, collision flags, rewards bit-exact;
observation floats within the fp32 tolerance of tests/_parity.py).  Two size-independent properties ride
along: the batch is deterministic (same seed, same bits) and shard-invariant (two half batches with
`env_index_base` are the full batch).
"""
import numpy as np
import pytest

from _parity import LIN_REWARD_ATOL, compare_obs

from _parity import OBS_ATOL, OBS_RTOL, ang_diff
from test_gpu_tdm import steps_left

pytestmark = pytest.mark.gpu


def _subset_vs_oracle(E, N, steps, targets, reward_mode, n_sub, seed, check_every):
    import torch
    import gym_macm
    from oracle import oracle
    dev = "cuda:0"
    T = 1 if targets is None else len(set(targets))
    env = gym_macm.BatchedFlock(E, n_agents=[N], targets=targets, reward_mode=reward_mode, device=dev, seed=seed)
    twin = gym_macm.BatchedFlock(E, n_agents=[N], targets=targets, reward_mode=reward_mode, device=dev, seed=seed)
    rng = np.random.default_rng(seed)
    sub = np.sort(rng.choice(E, n_sub, replace=False))
    subt = torch.as_tensor(sub, device=dev)
    st = env.state
    pos = st["posvel"][subt][..., 0:2].cpu().numpy().astype(np.float64)
    ang = st["angsleep"][subt][..., 0].cpu().numpy().astype(np.float64)
    tg = st["targets"][subt].cpu().numpy().astype(np.float64)
    ref = oracle.OracleBatch(n_sub, n_agents=N, n_targets=T, reward_mode={"binary": 0, "linear": 1}[reward_mode])
    ref.reset(pos, ang, targets=tg, target_idx=targets)
    g = torch.Generator(device=dev)
    g.manual_seed(seed + 1)
    for k in range(steps):
        act = torch.randint(0, 3, (E, N, 3), generator=g, device=dev, dtype=torch.uint8)
        env.step(act)
        twin.step(act)
        o = ref.flock_step(act[subt].cpu().numpy().astype(np.int64), 4)
        if (k + 1) % check_every and k != steps - 1:
            continue
        torch.cuda.synchronize()
        st = env.state
        assert not (st["env_state"][:, 1] & 6).any(), "step %d: contact capacity overflow" % k
        body = torch.cat([st["posvel"][subt], st["angsleep"][subt], st["fat"][subt]], -1).cpu().numpy()
        assert np.array_equal(body, ref.bodies()), "step %d: body state of the subset" % k
        assert np.array_equal(st["nn_idx"][subt].cpu().numpy(), o["nn_idx"]), "step %d: nearest-agent ids" % k
        assert np.array_equal(st["collided"][subt].cpu().numpy(), o["collided"]), "step %d: collision flags" % k
        rew = st["rewards"][subt].cpu().numpy().astype(np.float64)
        if reward_mode == "binary":
            assert np.array_equal(rew, o["rewards"]), "step %d: rewards" % k
        else:
            assert np.array_equal(rew == -1, o["rewards"] == -1)
            assert np.allclose(rew, o["rewards"], rtol=0, atol=LIN_REWARD_ATOL), "step %d: linear rewards" % k
        compare_obs(st["obs"][subt].cpu().numpy(), o, "polar", k)
        cnt = st["contact_count"][subt].cpu().numpy()
        for j in range(0, n_sub, max(1, n_sub // 8)):
            ab, fl, imp = env.contacts(int(sub[j]))
            rab, rfl, rimp = ref.contacts(j)
            assert cnt[j] == len(rab) and np.array_equal(ab, rab) and np.array_equal(fl, rfl), "step %d: contacts" % k
            assert np.array_equal(imp[rfl.astype(bool)], rimp[rfl.astype(bool)]), "step %d: impulses" % k
        # determinism: the twin batch holds the same bits everywhere
        for name in ("posvel", "angsleep", "fat", "contact_count", "rewards", "nn_idx", "obs"):
            assert torch.equal(st[name], twin.state[name]), "step %d: twin batch differs in %s" % (k, name)
    env.close()
    twin.close()


def test_cfg2_full_size_long_run_subset():
    # BASELINE configs[1]: 64 agents x 4096 envs, linear reward; 240 steps = 4 s of simulated time
    _subset_vs_oracle(4096, 64, 240, None, "linear", n_sub=48, seed=101, check_every=40)


def test_cfg3_full_size_subset():
    # BASELINE configs[2]: 6 agents, targets=[0,0,1,1,2,2], 65,536 envs (per-agent target gather)
    _subset_vs_oracle(65536, 6, 120, [0, 0, 1, 1, 2, 2], "binary", n_sub=256, seed=102, check_every=30)


def test_cfg5_one_gpu_shard_subset():
    # BASELINE configs[4]: 1M agents of 64-agent flock envs on one GPU (16,384 envs, multi-wave launch)
    _subset_vs_oracle(16384, 64, 40, None, "linear", n_sub=32, seed=103, check_every=20)


def test_cfg3_shard_invariance_at_full_size():
    # two ranks' shards (env_index_base) stepped separately are the full batch, bit for bit
    import torch
    import gym_macm
    dev = "cuda:0"
    E, N, tg = 65536, 6, [0, 0, 1, 1, 2, 2]
    full = gym_macm.BatchedFlock(E, n_agents=[N], targets=tg, device=dev, seed=7)
    halves = [gym_macm.BatchedFlock(E // 2, n_agents=[N], targets=tg, device=dev, seed=7, env_index_base=b)
              for b in (0, E // 2)]
    g = torch.Generator(device=dev)
    g.manual_seed(8)
    for k in range(30):
        act = torch.randint(0, 3, (E, N, 3), generator=g, device=dev, dtype=torch.uint8)
        full.step(act)
        halves[0].step(act[: E // 2].contiguous())
        halves[1].step(act[E // 2:].contiguous())
    torch.cuda.synchronize()
    for name in ("posvel", "angsleep", "fat", "contact_count", "rewards", "nn_idx", "obs", "collided"):
        both = torch.cat([h.state[name] for h in halves], 0)
        assert torch.equal(full.state[name], both), name


def test_cfg4_full_size_tdm_subset():
    # BASELINE configs[3]: 3 teams x 15 agents x 16,384 envs (repaired TDM semantics)
    import torch
    import gym_macm
    from oracle import oracle
    dev = "cuda:0"
    E, teams, steps, n_sub = 16384, [15, 15, 15], 90, 24
    team = np.array([t for t, n in enumerate(teams) for _ in range(n)], np.uint8)
    N = len(team)
    env = gym_macm.BatchedTDM(E, n_agents=teams, device=dev, seed=55)
    rng = np.random.default_rng(55)
    sub = np.sort(rng.choice(E, n_sub, replace=False))
    subt = torch.as_tensor(sub, device=dev)
    st = env.state
    pos = st["posvel"][subt][..., 0:2].cpu().numpy().astype(np.float64)
    ang = st["angsleep"][subt][..., 0].cpu().numpy().astype(np.float64)
    ref = oracle.OracleBatch(n_sub, env_kind=oracle.TDM, n_agents=N, n_targets=0)
    ref.reset(pos, ang, team=team)
    g = torch.Generator(device=dev)
    g.manual_seed(56)
    for k in range(steps):
        act = torch.cat([torch.randint(0, 3, (E, N, 3), generator=g, device=dev, dtype=torch.uint8),
                         torch.randint(0, 2, (E, N, 1), generator=g, device=dev, dtype=torch.uint8)], -1)
        env.step(act)
        o = ref.tdm_step(act[subt].cpu().numpy().astype(np.int64), 4)
        if (k + 1) % 30:
            continue
        torch.cuda.synchronize()
        st = env.state
        assert not (st["env_state"][:, 1] & 6).any(), "step %d: contact capacity overflow" % k
        body = torch.cat([st["posvel"][subt], st["angsleep"][subt], st["fat"][subt]], -1).cpu().numpy()
        assert np.array_equal(body, ref.bodies()), "step %d: body state of the subset" % k
        ts = ref.tdm_state()
        assert np.array_equal(st["tdm_state"][subt][..., 0].cpu().numpy().astype(np.float64), ts[..., 0]), "health"
        assert np.array_equal(st["collided"][subt].cpu().numpy(), o["collided"]), "step %d: collision flags" % k
        assert np.array_equal(st["rewards"][subt].cpu().numpy().astype(np.float64), o["rewards"]), "step %d: rewards" % k
        assert np.array_equal(st["done"][subt].cpu().numpy(), o["done"]), "step %d: done" % k
        obs = st["obs"][subt].cpu().numpy().reshape(n_sub, N, N, 4)
        assert np.array_equal(obs[..., 3].astype(np.int8), o["type"]), "step %d: ally/enemy/none flags" % k
        # observation floats (combat.py:206-227; float64 in the reference, fp32 here: the bars of tests/_parity.py)
        m = o["type"] >= 0
        assert np.allclose(obs[..., 0][m], o["obs"][..., 0][m], rtol=OBS_RTOL, atol=OBS_ATOL), "step %d: obs r" % k
        assert ang_diff(obs[..., 1][m].astype(np.float64), o["obs"][..., 1][m]).max() <= 3e-6, "step %d: obs theta" % k
        assert ang_diff(obs[..., 2][m].astype(np.float64), o["obs"][..., 2][m]).max() <= 3e-6, "step %d: obs phi" % k
        # cool-downs (whole steps left; combat.py:142,155) and alive flags
        ti = st["tdm_state"][subt].cpu().view(torch.int32).numpy()
        alive = ts[..., 3] > 0
        assert np.array_equal(ti[..., 3] & 1, ts[..., 3].astype(np.int32)), "step %d: alive" % k
        assert np.array_equal(ti[..., 1][alive], np.vectorize(steps_left)(ts[..., 1])[alive]), "step %d: attack cool-down" % k
        assert np.array_equal(ti[..., 2][alive], np.vectorize(steps_left)(ts[..., 2])[alive]), "step %d: movement cool-down" % k
        assert np.array_equal(st["env_state"][subt][:, 3].cpu().numpy(), o["winner"]), "step %d: winner" % k
        # contact lists of the subset: pairs in birth order, touching flags, warm-start impulses
        for j, e in enumerate(sub):
            ab, fl, imp = env.contacts(int(e))
            rab, rfl, rimp = ref.contacts(j)
            assert np.array_equal(ab, rab) and np.array_equal(fl, rfl), "step %d env %d: contact list" % (k, e)
            assert np.array_equal(imp[rfl.astype(bool)], rimp[rfl.astype(bool)]), "step %d env %d: impulses" % (k, e)
    env.close()
