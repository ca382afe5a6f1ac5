"""CUDA path (through the C ABI) against the CPU oracle on identical seeded states."""
import numpy as np
import pytest

from _parity import run_parity

pytestmark = pytest.mark.gpu


def test_cfg1_single_env_4_agents_binary():
    # BASELINE config 1: cm-flock-v0 n_agents=[4], one env, binary reward
    run_parity(1, 4, 300, seed=0)


def test_cfg1_crowded():
    run_parity(32, 4, 200, seed=1, spread=2.0)


def test_cfg2_small_linear():
    # BASELINE config 2 at a size the oracle steps in milliseconds
    run_parity(64, 64, 120, seed=1234, reward_mode="linear")


def test_cfg2_full_size_linear_steps():
    # 4096 envs x 64 agents, linear reward: full size, a few steps from spawn (densest contacts)
    run_parity(4096, 64, 6, seed=1234, reward_mode="linear", check_every=3)


def test_cfg3_multi_flock_targets():
    # BASELINE config 3: 6 agents, targets=[0,0,1,1,2,2] (per-agent target gather)
    run_parity(2048, 6, 60, targets=[0, 0, 1, 1, 2, 2], seed=3, check_every=10)


def test_dense_pile_multi_island_ordering():
    # 64 agents spawned inside 5 m x 5 m: hundreds of touching contacts, bodies with many
    # contacts each -> exercises the island DFS order and the level schedule
    st = run_parity(16, 64, 40, seed=7, spread=6.5, max_contacts=2016, max_touching=240)
    assert st["max_touching"] > 64


def test_mid_density_n32_and_n16():
    run_parity(64, 32, 80, seed=8, spread=6.0)
    run_parity(64, 16, 80, seed=9, spread=4.0)
    run_parity(64, 45, 60, seed=10, spread=8.0)


def test_flock_bot_congregation():
    # bots.flock drives every agent to the target: the flock piles up, contact density rises
    st = run_parity(8, 24, 900, seed=11, policy="flock", check_every=50, _reward_radius=7)
    assert st["max_touching"] >= 8


def test_idle_sleep():
    # idle agents fall asleep (velocity snapped to zero) island by island
    run_parity(8, 8, 80, seed=12, spread=3.0, policy="idle", check_every=5)


def test_continuous_and_cartesian():
    run_parity(64, 8, 100, seed=13, spread=5.0, action_mode="continuous", coord="cartesian")


def test_pade_damping_and_odd_iterations():
    run_parity(32, 12, 60, seed=14, spread=4.0, damping_model="taylor", velocityIterations=5, positionIterations=2)
    run_parity(32, 12, 30, seed=15, spread=4.0, positionIterations=0, enableWarmStarting=False)


def test_episode_done_step():
    run_parity(2, 4, 40, seed=16, time_limit=0.5, check_every=1)


@pytest.mark.parametrize("N,cols", [(64, 8), (40, 8), (33, 11)])
def test_nn_exact_draws_on_a_lattice(N, cols):
    # agents on a square lattice (spacing 2 m: no contacts): every agent has two to four nearest
    # neighbours at exactly the same squared distance, so the lowest index has to win
    # (mvmnt.py:194, strict '<' in ascending order) -- the block-minimum search's draw handling,
    # its rotated half and its padding lanes (N < 64) are all on this path
    import torch
    import gym_macm
    from oracle import oracle
    from _parity import compare_obs
    E = 4
    rng = np.random.default_rng(21)
    idx = np.arange(N)
    base = np.stack([2.0 * (idx % cols), 2.0 * (idx // cols)], -1).astype(np.float64)
    pos = np.stack([base[rng.permutation(N)] + off for off in ((0, 0), (-7, 3), (100, -50), (0.5, 0.25))])
    ang = rng.uniform(-3, 3, (E, N))
    tg = rng.uniform(-40, 40, (E, 1, 2))
    env = gym_macm.BatchedFlock(E, n_agents=[N], device="cuda:0", seed=None)
    env.load_state(pos, ang, targets=tg)
    ref = oracle.OracleBatch(E, n_agents=N, n_targets=1)
    ref.reset(pos, ang, targets=tg)
    torch.cuda.synchronize()
    o = ref.flock_observe()
    assert np.array_equal(env.state["nn_idx"].cpu().numpy(), o["nn_idx"])
    compare_obs(env.state["obs"].cpu().numpy(), o, "polar")
    idle = np.ones((E, N, 3), np.int64)
    for k in range(3):   # the step kernel's own observation pass; idle agents stay on the lattice
        env.step(torch.as_tensor(idle, device="cuda:0"))
        o = ref.flock_step(idle)
        torch.cuda.synchronize()
        assert np.array_equal(env.state["nn_idx"].cpu().numpy(), o["nn_idx"]), "step %d" % k
        compare_obs(env.state["obs"].cpu().numpy(), o, "polar", k)
    env.close()


def test_more_than_64_agents_per_env():
    """The reference has no cap on n_agents (mvmnt.py:61); envs of 65..128 agents run four agents per lane with
    128-bit contact adjacency rows.  Same bars as everywhere: engine state, contact lists, ids bit-exact."""
    run_parity(24, 96, 60, seed=31, reward_mode="linear")                                  # the reference's 20 m spawn square
    run_parity(12, 128, 50, seed=32, spread=14.0, targets=[i % 3 for i in range(128)])      # full width, three targets, denser
    st = run_parity(8, 80, 40, seed=33, spread=8.0, max_contacts=80 * 79 // 2, max_touching=240)   # a pile: multi-contact islands
    assert st["max_touching"] > 32       # the level-scheduled solver path of the wide shape ran
    run_parity(6, 70, 300, seed=34, policy="flock", check_every=25, max_contacts=70 * 69 // 2, max_touching=240)


def test_piles_beyond_the_shared_memory_stage():
    """More than 240 touching contacts in one env (an overlapping spawn pile): with max_touching > 240 the global-memory
    solver stage takes such envs, bit-exact like everything else; without it they are flagged, not silently wrong."""
    import torch
    import gym_macm
    from _parity import make_pair, compare_step
    for N, spread, seed in ((64, 3.6, 41), (128, 5.5, 42), (33, 2.2, 43)):
        E = 6
        env, ref, rng = make_pair(E, N, seed=seed, spread=spread, max_contacts=N * (N - 1) // 2, max_touching=N * (N - 1) // 2)
        worst = 0
        for k in range(25):
            act = rng.integers(0, 3, (E, N, 3))
            env.step(torch.as_tensor(act, device="cuda:0"))
            o = ref.flock_step(act)
            worst = max(worst, max(ref.env_info(e)["touching"] for e in range(E)))
            compare_step(env, ref, o, k)
        assert worst > 240, worst          # the global-memory stage was exercised
        assert env.overflow_count() == (0, 0)
        env.close()
    # the same pile on a sim without the stage: reported
    pos = np.random.default_rng(41).uniform(-1.8, 1.8, (2, 64, 2))
    flagged = gym_macm.BatchedFlock(2, n_agents=[64], device="cuda:0", seed=None, max_contacts=2016, max_touching=240)
    flagged.load_state(pos, np.zeros((2, 64)))
    flagged.step(torch.ones((2, 64, 4), dtype=torch.uint8, device="cuda:0"))
    assert bool(flagged.overflowed.all())
