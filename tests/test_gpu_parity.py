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
    run_parity(32, 12, 60, seed=14, spread=4.0, damping_model="pade", velocityIterations=5, positionIterations=2)
    run_parity(32, 12, 30, seed=15, spread=4.0, positionIterations=0, enableWarmStarting=False)


def test_episode_done_step():
    run_parity(2, 4, 40, seed=16, time_limit=0.5, check_every=1)
